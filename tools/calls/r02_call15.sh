#!/bin/bash
mkdir -p gpurun_out
for cfg in "50 8" "30 8" "24 8" "18 8" "12 8"; do
set -- $cfg
MI_B200_SM_BUDGET=$1 MI_B200_TASK_STREAMS=$2 timeout 300 python bench.py --steps 4 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench15_$1_$2.json 2> gpurun_out/r02_bench15_$1_$2.err
python -c "
import json,sys; d=json.load(open('gpurun_out/r02_bench15_$1_$2.json')); print('budget $1 lanes $2:', d['value'], d['e2e']['value'])" || tail -3 gpurun_out/r02_bench15_$1_$2.err
done
