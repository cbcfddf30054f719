#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu -s > gpurun_out/r02_c6_pytest.log 2>&1
echo "pytest rc=$?"; grep -a "^\[\|passed\|failed\|Error\|error" gpurun_out/r02_c6_pytest.log | tail -40
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r02_c6_bench.json 2> gpurun_out/r02_c6_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_c6_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c6_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['kernel_time_share'], d['clocks'], d.get('cpu_baseline'))
PY
