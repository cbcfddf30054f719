#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/r02_c8_pytest.log 2>&1
echo "pytest rc=$?"; grep -a "^\[\|passed\|failed\|^FAILED\|^ERROR" gpurun_out/r02_c8_pytest.log | tail -30
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/r02_c8_bench.json 2> gpurun_out/r02_c8_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_c8_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c8_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['kernel_time_share'], d['clocks'])
print(json.dumps(d.get('extra'), indent=1)[:1500]); print(d.get('roofline_secondary'))
PY
python -c "import __graft_entry__ as g; g.smoke()"
