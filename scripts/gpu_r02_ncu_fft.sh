#!/bin/bash
# ncu --set full of the radix-4 / radix-16 1024-point STFT kernels (the capture in misc_kernels_r02_ncu_full.txt predates them)
mkdir -p gpurun_out /tmp/ncu
timeout 150 ncu --set full --clock-control none -k 'regex:fft_analysis|fft_synthesis' -c 4 -o /tmp/ncu/fft -f python scripts/ncu_misc.py 4 > gpurun_out/r02_ncu_fft.log 2>&1
echo "ncu fft rc=$?"; tail -3 gpurun_out/r02_ncu_fft.log
ncu -i /tmp/ncu/fft.ncu-rep --page raw --csv > gpurun_out/r02_fft_full_raw.csv 2>/dev/null
wc -c gpurun_out/r02_fft_full_raw.csv
