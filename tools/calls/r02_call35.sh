#!/bin/bash
# round 2, GPU call 35: NaN-poison run of the GPU suite; compute-sanitizer memcheck over the conv + kernel tests
mkdir -p gpurun_out
MI_B200_POISON=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t35_poison.log 2>&1
echo "poison rc=$?"; tail -4 gpurun_out/r02_t35_poison.log | cut -c1-300
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_conv_tc_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 1200 -x > gpurun_out/r02_t35_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r02_t35_memcheck.log | cut -c1-300
