#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_network.py tests/test_gpu_elementwise.py -x -q -m gpu > gpurun_out/r02_c4_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02_c4_pytest.log
SHAPESET=n128 timeout 300 python scripts/ncu_conv.py 16 5
SHAPESET=n128 SINGLE=1 timeout 300 python scripts/ncu_conv.py 16 5 | head -2
timeout 300 python scripts/ncu_conv.py 16 5
for mb in 0 48 80 110; do BUDDY_GN_L2_MB=$mb timeout 300 python scripts/bench_gn_l2.py 16; done
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c4_bench.json 2> gpurun_out/r02_c4_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c4_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['kernel_time_share'], d['clocks'])
PY
BUDDY_GN_L2_MB=0 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('L2 off:', d['value'], d['ms_per_step'], d['roofline']['frac'], d['kernel_time_share'])"
timeout 600 python scripts/layer_table.py 16 mixed > gpurun_out/r02_c4_layers.txt 2>&1
head -30 gpurun_out/r02_c4_layers.txt
