#!/bin/bash
# round 2, GPU call 9: filter-column-stacked 3x3 kernel (N = 192 MMAs): parity, per-layer A/B, per-role counters
mkdir -p gpurun_out
MI_B200_KXS=2 timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --timeout 300 -x > gpurun_out/r02_t9_conv.log 2>&1
echo "conv(kxs forced) rc=$?"; tail -15 gpurun_out/r02_t9_conv.log | cut -c1-300
MI_B200_KXS=0 python tools/bench_conv.py fprop > gpurun_out/r02_conv9_halo.txt 2>&1
MI_B200_KXS=2 python tools/bench_conv.py fprop > gpurun_out/r02_conv9_kxs.txt 2>&1
paste -d'|' gpurun_out/r02_conv9_halo.txt gpurun_out/r02_conv9_kxs.txt | cut -c1-200
for s in "2 258 450 51 51" "2 192 256 64 64" "2 384 512 32 32" "2 48 64 256 256" "2 24 32 512 512"; do
MI_B200_KXS=2 MI_B200_DEBUG_TIMING=1 python tools/one_conv.py $s 2>&1 | tail -1
done
