#!/bin/bash
for pol in "" F G E; do
  echo "===== BUDDY_X1_BWD=$pol"
  BUDDY_X1_BWD=$pol timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_sampler.py -q -m gpu -s -k "mixed or other_lengths or informed or uncond or 30s" 2>&1 | grep -a "^\[net\|^\[informed\|^\[uncond\|passed\|failed"
  BUDDY_X1_BWD=$pol timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench', round(d['value'],2), round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4), 'passes', round(d['roofline']['fp16_pass_equivalents_per_product'],3), d['clocks']['sm_mhz'])"
done
