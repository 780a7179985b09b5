#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_t67_all.log 2>&1
echo "all rc=$?"; tail -2 gpurun_out/r02_t67_all.log | cut -c1-300
MI_B200_POISON=1 timeout 900 python -m pytest tests/test_system_gpu.py -m gpu -q --timeout 900 -x -k "sepconv" > gpurun_out/r02_t67_poison.log 2>&1
echo "poison rc=$?"; tail -2 gpurun_out/r02_t67_poison.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches67.csv python tools/one_task.py > gpurun_out/r02_one_task67.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches67.csv > gpurun_out/r02_launches67.txt; grep "planar\|Fill\|TOTAL" gpurun_out/r02_launches67.txt
