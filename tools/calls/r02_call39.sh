#!/bin/bash
# round 2, GPU call 39: bias gradient inside the weight-gradient kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --timeout 120 -x -k "wgrad" > gpurun_out/r02_t39_wgrad.log 2>&1
echo "wgrad rc=$?"; tail -4 gpurun_out/r02_t39_wgrad.log | cut -c1-400
