#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_flow_kernels_gpu.py -m gpu -q --timeout 300 -x > gpurun_out/r02_t55_kernels.log 2>&1
echo "kernels rc=$?"; tail -2 gpurun_out/r02_t55_kernels.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'])"
MI_B200_UPSAMPLE_STRIP=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench per-pixel form', d['value'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches55.csv python tools/one_task.py > gpurun_out/r02_one_task55.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches55.csv > gpurun_out/r02_launches55.txt; grep "upsample\|TOTAL" gpurun_out/r02_launches55.txt
