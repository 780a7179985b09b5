#!/bin/bash
# round 2, GPU call 37: CAIN: tiny-map conv kernel + 16-byte interior reduce
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_flow_kernels_gpu.py -m gpu -q --timeout 300 -x > gpurun_out/r02_t37_kernels.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/r02_t37_kernels.log | cut -c1-300
timeout 900 python -m pytest tests/test_system_gpu.py -m gpu -q --timeout 600 -x -k "cain" > gpurun_out/r02_t37_cain.log 2>&1
echo "cain system rc=$?"; tail -3 gpurun_out/r02_t37_cain.log | cut -c1-300
timeout 600 python tools/bench_backbones.py cain 2>&1 | tail -1 | cut -c1-300
