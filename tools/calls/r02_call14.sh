#!/bin/bash
# round 2, GPU call 14: SM budget per persistent launch x number of task lanes
mkdir -p gpurun_out
python tools/bench_warp.py 2 32 > gpurun_out/r02_warp14.txt 2>&1; cat gpurun_out/r02_warp14.txt | cut -c1-160
for cfg in "148 4" "74 4" "74 8" "50 6" "37 8" "37 4" "100 4"; do
set -- $cfg
MI_B200_SM_BUDGET=$1 MI_B200_TASK_STREAMS=$2 timeout 300 python bench.py --steps 4 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench14_$1_$2.json 2> gpurun_out/r02_bench14_$1_$2.err
python -c "
import json,sys; d=json.load(open('gpurun_out/r02_bench14_$1_$2.json')); print('budget $1 lanes $2:', d['value'], d['e2e']['value'])" || tail -3 gpurun_out/r02_bench14_$1_$2.err
done
