#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_network.py -x -q -m gpu > gpurun_out/r02_c2_pytest_conv.log 2>&1
echo "conv/network rc=$?" | tee -a gpurun_out/r02_c2_pytest_conv.log
tail -5 gpurun_out/r02_c2_pytest_conv.log
echo "== new kernel" > gpurun_out/r02_c2_shapes.txt
timeout 300 python scripts/ncu_conv.py 16 5 >> gpurun_out/r02_c2_shapes.txt 2>&1
SHAPESET=n128 timeout 300 python scripts/ncu_conv.py 16 5 >> gpurun_out/r02_c2_shapes.txt 2>&1
for v in "DBG=1" "NOSTATS=1" "DBG=4"; do
  echo "== n128 $v" >> gpurun_out/r02_c2_shapes.txt
  env SHAPESET=n128 $v timeout 300 python scripts/ncu_conv.py 16 5 >> gpurun_out/r02_c2_shapes.txt 2>&1
done
cat gpurun_out/r02_c2_shapes.txt
timeout 1500 python -m pytest tests/test_gpu_blind.py tests/test_gpu_sampler.py -x -q -m gpu -s > gpurun_out/r02_c2_pytest_blind.log 2>&1
echo "blind/sampler rc=$?" | tee -a gpurun_out/r02_c2_pytest_blind.log
grep -a "\[blind\|\[informed\|passed\|failed\|Error\|assert" gpurun_out/r02_c2_pytest_blind.log | tail -40
