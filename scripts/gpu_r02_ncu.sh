#!/bin/bash
# ncu evidence of the round-2 kernels (one GPU; numbers printed under ncu are never bench values).
# The .ncu-rep files are exported to CSV on the box and deleted (gpurun_out/ may not exceed 64 MiB).
mkdir -p gpurun_out /tmp/ncu
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -o /tmp/ncu/conv -f python scripts/ncu_conv.py 16 1 > gpurun_out/r02_ncu_conv.log 2>&1
echo "ncu conv rc=$?"
ncu -i /tmp/ncu/conv.ncu-rep --page raw --csv > gpurun_out/r02_conv_full_raw.csv 2>/dev/null
ncu -i /tmp/ncu/conv.ncu-rep --page source --csv -k regex:conv_gemm --launch-skip 1 --launch-count 1 > gpurun_out/r02_conv_source.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k 'regex:gn_apply|gn_bwd|dft_analysis|dft_synthesis|fft_analysis|fft_synthesis|fftconv|subband_fir|comp_loss|wpe_kernel|upfirdn' -o /tmp/ncu/misc -f python scripts/ncu_misc.py 4 > gpurun_out/r02_ncu_misc.log 2>&1
echo "ncu misc rc=$?"
ncu -i /tmp/ncu/misc.ncu-rep --page raw --csv > gpurun_out/r02_misc_full_raw.csv 2>/dev/null
ls -la /tmp/ncu gpurun_out | head -30
du -sh gpurun_out
