#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c3_bench.json 2> gpurun_out/r02_c3_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c3_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['kernel_time_share'], d['clocks'])
PY
timeout 600 python scripts/layer_table.py 16 mixed > gpurun_out/r02_c3_layers.txt 2>&1
head -45 gpurun_out/r02_c3_layers.txt
SHAPESET=n128 timeout 300 python scripts/ncu_conv.py 16 5
timeout 300 python scripts/ncu_conv.py 16 5 | tail -3
