#!/bin/bash
# round 2, GPU call 25: third pipeline stage for the 64-channel resident layers (one-box staging, plain epilogue)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 120 -x > gpurun_out/r02_t25_conv.log 2>&1
echo "conv+kernels rc=$?"; tail -3 gpurun_out/r02_t25_conv.log | cut -c1-300
MI_B200_KXS=4 timeout 300 python tools/bench_conv.py fprop > gpurun_out/r02_conv25_two_box.txt 2>&1
timeout 300 python tools/bench_conv.py fprop > gpurun_out/r02_conv25_one_box.txt 2>&1
paste -d'|' gpurun_out/r02_conv25_two_box.txt gpurun_out/r02_conv25_one_box.txt | cut -c1-170
for s in "2 258 450 51 51" "2 192 256 64 64"; do
MI_B200_DEBUG_TIMING=1 timeout 120 python tools/one_conv.py $s 2>&1 | tail -1
done
for m in 4 1; do
MI_B200_KXS=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('kxs mode $m', d['value'], d['e2e']['value'], d['roofline']['frac'])"
done
