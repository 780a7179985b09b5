#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['e2e']['value'])"; }
run default
MI_B200_SEPCONV_QUAD=0 run sepconv_gen1
MI_B200_SM_BUDGET=30 run budget30
MI_B200_SM_BUDGET=24 run budget24
MI_B200_SM_BUDGET=49 run budget49
