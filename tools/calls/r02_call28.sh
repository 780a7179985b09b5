#!/bin/bash
# round 2, GPU call 28: separable convolution with packed FFMA2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_reference_gpu.py -m gpu -q --timeout 300 -k "sepconv or reference" > gpurun_out/r02_t28_sepconv.log 2>&1
echo "sepconv tests rc=$?"; tail -3 gpurun_out/r02_t28_sepconv.log | cut -c1-300
timeout 300 python tools/bench_sepconv.py > gpurun_out/r02_sepconv28.txt 2>&1; tail -6 gpurun_out/r02_sepconv28.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value']); print(d['roofline']['per_kernel']['sepconv_fwd'], d['roofline']['per_kernel']['sepconv_bwd'])"
ncu --set full --clock-control none --import-source on -k regex:sepconv_fwd_quad -s 3 -c 1 -o gpurun_out/r02b_ncu_sepconv_fwd_quad_ffma2 -f python tools/bench_sepconv.py > gpurun_out/r02_ncu28a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sepconv_bwd_quad -s 3 -c 1 -o gpurun_out/r02b_ncu_sepconv_bwd_quad_ffma2 -f python tools/bench_sepconv.py > gpurun_out/r02_ncu28b.log 2>&1
