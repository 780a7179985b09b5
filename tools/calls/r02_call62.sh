#!/bin/bash
mkdir -p gpurun_out
for nt in 256 128 64; do echo "TPOSE_NT=$nt"; MI_B200_SEPCONV_TPOSE_NT=$nt python tools/bench_sepconv.py 2>&1 | grep planar; done
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "sepconv" > gpurun_out/r02_t62_sepconv.log 2>&1
echo "sepconv tests rc=$?"; tail -2 gpurun_out/r02_t62_sepconv.log | cut -c1-300
for nt in 256 128 64; do MI_B200_SEPCONV_TPOSE_NT=$nt timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench nt=$nt', d['value'], d['e2e']['value'])"; done
