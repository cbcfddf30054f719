#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_spectral.py tests/test_reference_integration.py tests/test_gpu_network.py -q -m gpu -s > gpurun_out/r02_c7_pytest.log 2>&1
echo "pytest rc=$?"; grep -a "^\[\|passed\|failed\|Error\|error\|assert" gpurun_out/r02_c7_pytest.log | tail -40
timeout 1200 python bench.py --mode long --micro-batch 8 --steps 2 --warmup 1 > gpurun_out/r02_c7_long.json 2> gpurun_out/r02_c7_long.err
echo "long rc=$?"; tail -3 gpurun_out/r02_c7_long.err; cut -c1-400 gpurun_out/r02_c7_long.json
