#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "sepconv" > gpurun_out/r02_t64_sepconv.log 2>&1
echo "sepconv tests rc=$?"; tail -2 gpurun_out/r02_t64_sepconv.log | cut -c1-300
python tools/bench_sepconv.py 2>&1 | grep planar
MI_B200_SEPCONV_TPOSE_VEC=0 python tools/bench_sepconv.py 2>&1 | grep planar
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches64.csv python tools/one_task.py > gpurun_out/r02_one_task64.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches64.csv > gpurun_out/r02_launches64.txt; grep "planar\|TOTAL" gpurun_out/r02_launches64.txt
