#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 120 -x > gpurun_out/r02_t47_conv.log 2>&1
echo "conv rc=$?"; tail -2 gpurun_out/r02_t47_conv.log | cut -c1-300
timeout 300 python tools/bench_conv.py fprop 2>&1 | head -4
MI_B200_DEBUG_TIMING=1 timeout 120 python tools/one_conv.py 2 384 512 32 32 2>&1 | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'])"
timeout 900 python -m pytest tests/test_system_gpu.py -m gpu -q --timeout 600 -x -k "full_size" > gpurun_out/r02_t47_full.log 2>&1
echo "full-size rc=$?"; tail -2 gpurun_out/r02_t47_full.log | cut -c1-300
