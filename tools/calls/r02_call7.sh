#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_reference_gpu.py -m gpu -q --timeout 900 -k "sepconv or reference" > gpurun_out/r02_t7_kernels.log 2>&1
echo "kernels rc=$?"; tail -4 gpurun_out/r02_t7_kernels.log | cut -c1-300
python tools/bench_sepconv.py > gpurun_out/r02_sepconv_planar2.txt 2>&1; cat gpurun_out/r02_sepconv_planar2.txt | tail -6
python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench7.json')); print(d['value'], d['e2e']['value'], d['roofline']['per_kernel']['sepconv_fwd'], d['roofline']['per_kernel']['sepconv_bwd'])"; tail -3 gpurun_out/r02_bench7.err
MI_B200_SEPCONV_QUAD=0 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench7_gen1.json 2> gpurun_out/r02_bench7_gen1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench7_gen1.json')); print('gen1', d['value'], d['e2e']['value'], d['roofline']['per_kernel']['sepconv_fwd'], d['roofline']['per_kernel']['sepconv_bwd'])"
