#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_reference_gpu.py -m gpu -q --timeout 300 -k "sepconv or reference" > gpurun_out/r02_t30_sepconv.log 2>&1
echo "sepconv tests rc=$?"; tail -2 gpurun_out/r02_t30_sepconv.log | cut -c1-300
timeout 300 python tools/bench_sepconv.py 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value']); print(d['roofline']['per_kernel']['sepconv_fwd'], d['roofline']['per_kernel']['sepconv_bwd'])"
