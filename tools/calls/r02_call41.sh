#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --timeout 120 -x -k "wgrad" > gpurun_out/r02_t41_wgrad.log 2>&1
echo "wgrad rc=$?"; tail -2 gpurun_out/r02_t41_wgrad.log | cut -c1-400
MI_B200_WGRAD_BIAS_FUSED=0 timeout 300 python tools/bench_conv.py wgrad > gpurun_out/r02_conv41_wgrad_sep.txt 2>&1
timeout 300 python tools/bench_conv.py wgrad > gpurun_out/r02_conv41_wgrad_fused.txt 2>&1
paste -d'|' gpurun_out/r02_conv41_wgrad_sep.txt gpurun_out/r02_conv41_wgrad_fused.txt | cut -c1-150
for v in 0 1; do
MI_B200_WGRAD_BIAS_FUSED=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fused=$v', d['value'], d['e2e']['value'])"
done
MI_B200_WGRAD_BIAS_FUSED=0 timeout 600 python tools/bench_backbones.py cain rrin 2>&1 | grep tasks_per_s | cut -c1-200
timeout 600 python tools/bench_backbones.py cain rrin 2>&1 | grep tasks_per_s | cut -c1-200
