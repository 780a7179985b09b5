#!/bin/bash
# round 2, GPU call 20: SM budget x lanes for the other BASELINE configurations
mkdir -p gpurun_out
for m in superslomo cain rrin; do
for cfg in "147 4" "147 8" "74 8" "37 8"; do
set -- $cfg
echo -n "$m budget $1 lanes $2: "
MI_B200_SM_BUDGET=$1 MI_B200_TASK_STREAMS=$2 timeout 600 python tools/bench_backbones.py $m 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['tasks_per_s'], d['ms_per_meta_batch'])
except Exception as e: print('failed', e)"
done
done 2>&1 | tee gpurun_out/r02_budget_sweep_other_configs.txt
