#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t13_all.log 2>&1
echo "all rc=$?"; tail -8 gpurun_out/r02_t13_all.log | cut -c1-300
MI_B200_KXS=2 timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --timeout 120 > gpurun_out/r02_t13_conv.log 2>&1
echo "conv rc=$?"; tail -3 gpurun_out/r02_t13_conv.log | cut -c1-300
