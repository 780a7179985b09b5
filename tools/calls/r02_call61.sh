#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_t61_all.log 2>&1
echo "all rc=$?"; tail -2 gpurun_out/r02_t61_all.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'])"
MI_B200_FINISH_VEC=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench scalar finish', d['value'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches61.csv python tools/one_task.py > gpurun_out/r02_one_task61.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches61.csv > gpurun_out/r02_launches61.txt; grep "finish\|TOTAL" gpurun_out/r02_launches61.txt
