#!/bin/bash
# round 2, GPU call 51: paired filter columns in the weight-gradient kernel (Cout <= 64)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --timeout 120 -x -k "wgrad" > gpurun_out/r02_t51_wgrad.log 2>&1
echo "wgrad rc=$?"; tail -3 gpurun_out/r02_t51_wgrad.log | cut -c1-400
MI_B200_WGRAD_PAIR=0 timeout 300 python tools/bench_conv.py wgrad > gpurun_out/r02_conv51_wgrad_single.txt 2>&1
timeout 300 python tools/bench_conv.py wgrad > gpurun_out/r02_conv51_wgrad_pair.txt 2>&1
paste -d'|' gpurun_out/r02_conv51_wgrad_single.txt gpurun_out/r02_conv51_wgrad_pair.txt | cut -c1-150
for v in 0 1; do
MI_B200_WGRAD_PAIR=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pair=$v', d['value'], d['e2e']['value'])"
done
