#!/bin/bash
# last verification of round 2: whole GPU suite, smoke, short bench, launch list
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t68_all.log 2>&1
echo "all rc=$?"; tail -2 gpurun_out/r02_t68_all.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null > gpurun_out/r02g_bench_1gpu_short_last.json; python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_1gpu_short_last.json')); print('bench', d['value'], d['e2e']['value'], d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches68.csv python tools/one_task.py > gpurun_out/r02_one_task68.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches68.csv > gpurun_out/r02_launches68.txt; grep "planar\|TOTAL" gpurun_out/r02_launches68.txt
