#!/bin/bash
# round 2, GPU call 32: deferred (batched) finishing stage of the weight gradients
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_t32_all.log 2>&1
echo "all rc=$?"; tail -5 gpurun_out/r02_t32_all.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench32.json 2> gpurun_out/r02_bench32.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench32.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']); print({k:v['value'] for k,v in d['config']['other_configs'].items()})"; tail -3 gpurun_out/r02_bench32.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches32.csv python tools/one_task.py > gpurun_out/r02_one_task32.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches32.csv > gpurun_out/r02_launches32.txt; head -14 gpurun_out/r02_launches32.txt; tail -1 gpurun_out/r02_launches32.txt
