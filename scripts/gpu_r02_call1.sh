#!/bin/bash
# round-2 GPU call 1: SW128 shifted-descriptor probe + ablations of the N=128 conv shapes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt
timeout 60 scripts/probes/umma_sw128_shift_probe > gpurun_out/r02_sw128_probe.txt 2>&1
echo "probe rc=$?" >> gpurun_out/r02_sw128_probe.txt
export SHAPESET=n128
for v in "BASE=1" "DBG=1" "DBG=2" "DBG=4" "NOSTATS=1" "ONETAP=1" "NOPAIR=1"; do
  echo "== $v" >> gpurun_out/r02_ablate.txt
  env $v timeout 300 python scripts/ncu_conv.py 16 5 >> gpurun_out/r02_ablate.txt 2>&1
done
tail -5 gpurun_out/r02_sw128_probe.txt
