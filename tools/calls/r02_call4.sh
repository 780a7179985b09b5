#!/bin/bash
# round 2, GPU call 4: y-quad separable-convolution kernels (parity + A/B), lane sweep
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_reference_gpu.py -m gpu -q --timeout 900 > gpurun_out/r02_t4_kernels.log 2>&1
echo "kernels rc=$?"; tail -8 gpurun_out/r02_t4_kernels.log | cut -c1-300
python tools/bench_sepconv.py > gpurun_out/r02_sepconv_quad.txt 2>&1; cat gpurun_out/r02_sepconv_quad.txt | tail -6
MI_B200_SEPCONV_QUAD=0 python tools/bench_sepconv.py > gpurun_out/r02_sepconv_gen1.txt 2>&1; cat gpurun_out/r02_sepconv_gen1.txt | tail -6
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t4_all.log 2>&1
echo "all rc=$?"; tail -8 gpurun_out/r02_t4_all.log | cut -c1-300
for lanes in 4 6 8; do
  MI_B200_TASK_STREAMS=$lanes python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench4_l$lanes.json 2> gpurun_out/r02_bench4_l$lanes.err
  echo "lanes $lanes rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench4_l$lanes.json')); print(d['value'], d['e2e']['value'], d['roofline']['per_kernel']['sepconv_fwd'], d['roofline']['per_kernel']['sepconv_bwd'])"
done
