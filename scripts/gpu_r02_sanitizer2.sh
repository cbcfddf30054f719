#!/bin/bash
# memcheck over the kernels added late in round 2 (cast_operand, gn_act32, upfirdn2d inside the variant graphs)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_elementwise.py tests/test_reference_integration.py -q -m gpu -k "cast_operand or gn_act32 or upfirdn or ddpm_resblock or fir_variant or progressive_variants" > gpurun_out/r02_memcheck3.log 2>&1
echo "memcheck3 rc=$?"; grep -a "passed\|failed\|ERROR SUMMARY" gpurun_out/r02_memcheck3.log | tail -3
