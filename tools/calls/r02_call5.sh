#!/bin/bash
# round 2, GPU call 5: separable convolution with tap-planar filters (parity, micro-benchmark, bench)
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_reference_gpu.py -m gpu -q --timeout 900 -k "sepconv or reference" > gpurun_out/r02_t5_kernels.log 2>&1
echo "kernels rc=$?"; tail -8 gpurun_out/r02_t5_kernels.log | cut -c1-300
python tools/bench_sepconv.py > gpurun_out/r02_sepconv_planar.txt 2>&1; cat gpurun_out/r02_sepconv_planar.txt | tail -6
python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench5.json')); print(d['value'], d['e2e']['value'], d['roofline']['per_kernel']['sepconv_fwd'], d['roofline']['per_kernel']['sepconv_bwd'])"; tail -3 gpurun_out/r02_bench5.err
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t5_all.log 2>&1
echo "all rc=$?"; tail -8 gpurun_out/r02_t5_all.log | cut -c1-300
