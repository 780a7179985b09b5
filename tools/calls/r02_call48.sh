#!/bin/bash
# round 2, GPU call 48: final state: GPU suite, smoke, default bench
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t48_all.log 2>&1
echo "all rc=$?"; tail -2 gpurun_out/r02_t48_all.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/r02_bench48.json 2> gpurun_out/r02_bench48.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench48.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], r['frac'], r['achieved'], r['at_sm_share'], d['clocks'], d['gpu_launches']); print({k:v.get('value') for k,v in d['config']['other_configs'].items()}); g=d.get('gpu_reference',{}); print(g.get('allow_tf32',{}).get('tasks_per_s'), g.get('fp32',{}).get('tasks_per_s'), g.get('ratio_e2e_over_reference_tf32'), g.get('ratio_e2e_over_reference_fp32')); print(d['cpu_baseline']['value'])"; tail -3 gpurun_out/r02_bench48.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches48.csv python tools/one_task.py > gpurun_out/r02_one_task48.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches48.csv > gpurun_out/r02_launches48.txt; head -12 gpurun_out/r02_launches48.txt; tail -1 gpurun_out/r02_launches48.txt
