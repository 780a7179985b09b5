#!/bin/bash
# round 2, GPU call 21: MSL host-sync fix + share policy: bench with other configs; rrin budget sweep again
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench21.json 2> gpurun_out/r02_bench21.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench21.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac']); print({k:v['value'] for k,v in d['config']['other_configs'].items()})"; tail -3 gpurun_out/r02_bench21.err
for cfg in "147 8" "74 8" "37 8"; do
set -- $cfg
echo -n "rrin budget $1 lanes $2: "
MI_B200_SM_BUDGET=$1 MI_B200_TASK_STREAMS=$2 timeout 600 python tools/bench_backbones.py rrin 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['tasks_per_s'], d['ms_per_meta_batch'])
except Exception as e: print('failed', e)"
done 2>&1 | tee gpurun_out/r02_budget_sweep_rrin_after_sync_fix.txt
