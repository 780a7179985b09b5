#!/bin/bash
# round 2, GPU call 43: the default bench line of the final state (+ reference arm), smoke
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/r02_bench43.json 2> gpurun_out/r02_bench43.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench43.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], r['frac'], r['at_sm_share'], d['clocks']); print({k:v.get('value') for k,v in d['config']['other_configs'].items()}); g=d.get('gpu_reference',{}); print(g.get('allow_tf32',{}).get('tasks_per_s'), g.get('fp32',{}).get('tasks_per_s'), g.get('ratio_e2e_over_reference_tf32'), g.get('ratio_e2e_over_reference_fp32')); print(d['cpu_baseline'])"; tail -3 gpurun_out/r02_bench43.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench43_ref.json 2> gpurun_out/r02_bench43_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/r02_bench43_ref.json
