#!/bin/bash
# round 2, GPU call 18: store warp + pad-lane policy; RRIN regression A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 120 -x > gpurun_out/r02_t18_conv.log 2>&1
echo "conv+kernels rc=$?"; tail -3 gpurun_out/r02_t18_conv.log | cut -c1-300
timeout 300 python tools/bench_conv.py fprop > gpurun_out/r02_conv18_kxs.txt 2>&1; cat gpurun_out/r02_conv18_kxs.txt
for s in "2 258 450 51 51" "2 192 256 64 64" "2 384 512 32 32" "2 48 64 256 256"; do
MI_B200_DEBUG_TIMING=1 timeout 120 python tools/one_conv.py $s 2>&1 | tail -1
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench18.json 2> gpurun_out/r02_bench18.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench18.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"; tail -3 gpurun_out/r02_bench18.err
MI_B200_KXS=0 timeout 600 python tools/bench_backbones.py rrin 2>&1 | tail -1 | cut -c1-300
MI_B200_SM_BUDGET=147 timeout 600 python tools/bench_backbones.py rrin 2>&1 | tail -1 | cut -c1-300
MI_B200_TASK_STREAMS=4 MI_B200_SM_BUDGET=147 MI_B200_KXS=0 timeout 600 python tools/bench_backbones.py rrin 2>&1 | tail -1 | cut -c1-300
