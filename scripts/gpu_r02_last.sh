#!/bin/bash
# last GPU calls of round 2 (a few minutes of box time left): the operator / helper surface test against the
# unmodified reference and the default bench line of the final tree
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_reference_integration.py -m gpu -x -q -s -k "api_operator" > gpurun_out/last_api.log 2>&1
echo "api rc=$?"; tail -12 gpurun_out/last_api.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/last_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/last_bench.json').read().strip().splitlines()[-1])
print(round(d['value'], 2), round(d['ms_per_step'], 1), 'e2e', round(d['e2e']['value'], 2), 'frac', round(d['roofline']['frac'], 3),
      d['kernel_time_share'], d['clocks'], 'blind', round(d['extra']['blind']['value'], 1), 'b1', round(d['extra']['b1_latency']['ms_per_step'], 2),
      'full', round(d['extra']['e2e_full']['seconds'], 2), 'cpu', d['cpu_baseline']['value'])
PY
