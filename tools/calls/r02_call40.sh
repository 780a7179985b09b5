#!/bin/bash
# round 2, GPU call 40: fused bias gradient: GPU suite (plain and NaN-poisoned), bench with other configs, launch list
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_t40_all.log 2>&1
echo "all rc=$?"; tail -3 gpurun_out/r02_t40_all.log | cut -c1-300
MI_B200_POISON=1 timeout 1500 python -m pytest tests/test_conv_tc_gpu.py tests/test_system_gpu.py -m gpu -q --timeout 900 -x > gpurun_out/r02_t40_poison.log 2>&1
echo "poison rc=$?"; tail -2 gpurun_out/r02_t40_poison.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench40.json 2> gpurun_out/r02_bench40.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench40.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']); print({k:v.get('value') for k,v in d['config']['other_configs'].items()})"; tail -3 gpurun_out/r02_bench40.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches40.csv python tools/one_task.py > gpurun_out/r02_one_task40.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches40.csv > gpurun_out/r02_launches40.txt; head -14 gpurun_out/r02_launches40.txt; tail -1 gpurun_out/r02_launches40.txt
