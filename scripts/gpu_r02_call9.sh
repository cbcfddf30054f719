#!/bin/bash
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],2), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],4), d['kernel_time_share'], d['clocks']['sm_mhz'])"; }
EXTRA="" run A=1
EXTRA="--streams 2 --micro-batch 16" run A=1
EXTRA="--streams 2 --micro-batch 16" run BUDDY_CONV_SMEM_KB=200
EXTRA="--streams 2 --micro-batch 16" run BUDDY_CONV_SMEM_KB=180
EXTRA="--streams 2 --micro-batch 8" run BUDDY_CONV_SMEM_KB=200
EXTRA="--micro-batch 16" run A=1
