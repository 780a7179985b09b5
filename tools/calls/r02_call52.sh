#!/bin/bash
mkdir -p gpurun_out
MI_B200_WGRAD_MIN16=2 timeout 300 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --timeout 120 -x -k "wgrad" 2>&1 | tail -1
timeout 300 python tools/bench_conv.py wgrad > gpurun_out/r02_conv52_min3.txt 2>&1
MI_B200_WGRAD_MIN16=2 timeout 300 python tools/bench_conv.py wgrad > gpurun_out/r02_conv52_min2.txt 2>&1
paste -d'|' gpurun_out/r02_conv52_min3.txt gpurun_out/r02_conv52_min2.txt | cut -c1-150 | head -12
for v in 3 2; do
MI_B200_WGRAD_MIN16=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('min16=$v', d['value'], d['e2e']['value'])"
done
