#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_network.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/ncu_conv.py 16 5 | tail -3
SHAPESET=n128 timeout 300 python scripts/ncu_conv.py 16 5 | sed -n 3p
for i in 1 2; do timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4), d['kernel_time_share'], d['clocks']['sm_mhz'])"; done
