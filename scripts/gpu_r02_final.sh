#!/bin/bash
# final single-GPU evidence of the round: bench line (with extras + cpu baseline), reference arm, ncu launch list
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r02_final_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_ref.json 2> gpurun_out/r02_final_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/r02_final_ref.json
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 900 python bench.py --mode blind --steps 3 --warmup 1 > gpurun_out/r02_final_blind.json 2>/dev/null; echo "blind rc=$?"
timeout 1200 python bench.py --mode long --micro-batch 8 --steps 2 --warmup 1 > gpurun_out/r02_final_long.json 2>/dev/null; echo "long rc=$?"
python - <<'PY'
import json
for f in ("r02_final_bench", "r02_final_blind", "r02_final_long"):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, round(d['value'],2), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'frac', d['roofline']['frac'], d['kernel_time_share'], d['clocks'])
PY
du -sh gpurun_out
