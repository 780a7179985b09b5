#!/bin/bash
# round 2, GPU call 8: 16-byte-pixel warp kernels + 32-bit index arithmetic of the resampling kernels: parity, HBM GB/s, ncu
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 > gpurun_out/r02_t8_kernels.log 2>&1
echo "kernels rc=$?"; tail -4 gpurun_out/r02_t8_kernels.log | cut -c1-300
python tools/bench_warp.py 2 32 > gpurun_out/r02_warp8.txt 2>&1; cat gpurun_out/r02_warp8.txt
python tools/bench_conv.py > gpurun_out/r02_conv8.txt 2>&1; cat gpurun_out/r02_conv8.txt
ncu --set full --clock-control none --import-source on -k regex:warp_fwd_vec -s 6 -c 1 -o gpurun_out/r02_ncu_warp_fwd_vec -f python tools/bench_warp.py 32 > gpurun_out/r02_ncu8a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_bwd_vec -s 6 -c 1 -o gpurun_out/r02_ncu_warp_bwd_vec -f python tools/bench_warp.py 32 > gpurun_out/r02_ncu8b.log 2>&1
python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench8.json')); print(d['value'], d['e2e']['value'])"; tail -3 gpurun_out/r02_bench8.err
MI_B200_DEBUG_TIMING=1 python tools/one_conv.py 2 258 450 51 51 2>&1 | tail -2
MI_B200_DEBUG_TIMING=1 python tools/one_conv.py 2 48 64 256 256 2>&1 | tail -2
MI_B200_DEBUG_TIMING=1 python tools/one_conv.py 2 24 32 512 512 2>&1 | tail -2
