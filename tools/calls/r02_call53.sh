#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t53_all.log 2>&1
echo "all rc=$?"; tail -2 gpurun_out/r02_t53_all.log | cut -c1-300
MI_B200_POISON=1 timeout 1500 python -m pytest tests/test_conv_tc_gpu.py tests/test_system_gpu.py -m gpu -q --timeout 900 > gpurun_out/r02_t53_poison.log 2>&1
echo "poison rc=$?"; tail -2 gpurun_out/r02_t53_poison.log | cut -c1-300
timeout 600 python tools/bench_backbones.py superslomo cain rrin 2>&1 | grep tasks_per_s | cut -c1-150
