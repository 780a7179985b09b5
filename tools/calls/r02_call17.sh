#!/bin/bash
# round 2, GPU call 17: trimmed MMA-lane stage overhead; epilogue leader waits; RRIN lanes A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --timeout 120 -x > gpurun_out/r02_t17_conv.log 2>&1
echo "conv rc=$?"; tail -2 gpurun_out/r02_t17_conv.log | cut -c1-300
timeout 300 python tools/bench_conv.py fprop > gpurun_out/r02_conv17_kxs.txt 2>&1; cat gpurun_out/r02_conv17_kxs.txt
for s in "2 258 450 51 51" "2 192 256 64 64" "2 384 512 32 32" "2 48 64 256 256"; do
MI_B200_DEBUG_TIMING=1 timeout 120 python tools/one_conv.py $s 2>&1 | tail -1
done
for l in 4 8; do
MI_B200_TASK_STREAMS=$l timeout 600 python tools/bench_backbones.py rrin 2>&1 | tail -2 | cut -c1-400
nvidia-smi --query-gpu=memory.used --format=csv,noheader
done
