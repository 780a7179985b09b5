#!/bin/bash
# round 2, GPU call 26: wgrad filter-column kernel with elect.sync roles + unrolled row loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --timeout 120 -x > gpurun_out/r02_t26_conv.log 2>&1
echo "conv rc=$?"; tail -3 gpurun_out/r02_t26_conv.log | cut -c1-300
timeout 300 python tools/bench_conv.py wgrad > gpurun_out/r02_conv26_wgrad.txt 2>&1; cat gpurun_out/r02_conv26_wgrad.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'], d['roofline']['frac']); print(d['roofline']['per_kernel']['wgrad_tc_kx'])"
