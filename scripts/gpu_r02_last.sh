#!/bin/bash
# last GPU call of round 2 (≈ 4 minutes of box time left): the new operator-surface test against the unmodified
# reference, smoke(), the tests around the radix-4 FFT kernels, and a blind bench record with those kernels
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_reference_integration.py -m gpu -x -q -s -k "api_operator" > gpurun_out/last_api.log 2>&1
echo "api rc=$?"; tail -15 gpurun_out/last_api.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/last_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/last_smoke.log
timeout 100 python bench.py --mode blind --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/last_blind.json 2> gpurun_out/last_blind.err
echo "blind rc=$?"; cut -c1-400 gpurun_out/last_blind.json; tail -2 gpurun_out/last_blind.err
timeout 80 python -m pytest tests/test_gpu_spectral.py -m gpu -x -q > gpurun_out/last_spectral.log 2>&1
echo "spectral rc=$?"; tail -3 gpurun_out/last_spectral.log
