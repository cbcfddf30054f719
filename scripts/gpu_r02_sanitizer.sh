#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_elementwise.py tests/test_gpu_spectral.py -q -m gpu -k "not 256" > gpurun_out/r02_memcheck1.log 2>&1
echo "memcheck1 rc=$?"; grep -a "passed\|failed\|ERROR SUMMARY" gpurun_out/r02_memcheck1.log | tail -3
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_network.py tests/test_gpu_blind.py -q -m gpu -k "fixture and not full or blocked" > gpurun_out/r02_memcheck2.log 2>&1
echo "memcheck2 rc=$?"; grep -a "passed\|failed\|ERROR SUMMARY" gpurun_out/r02_memcheck2.log | tail -3
