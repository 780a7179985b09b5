#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_t58_all.log 2>&1
echo "all rc=$?"; tail -2 gpurun_out/r02_t58_all.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'])"
MI_B200_ROTATE_BATCH=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench rotate-per-layer', d['value'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches58.csv python tools/one_task.py > gpurun_out/r02_one_task58.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches58.csv > gpurun_out/r02_launches58.txt; grep "upsample\|dgrad\|TOTAL" gpurun_out/r02_launches58.txt
python tools/bench_upsample.py > gpurun_out/r02_upsample_58.jsonl 2>/dev/null
python - <<'PY'
import json
for l in open('gpurun_out/r02_upsample_58.jsonl'):
    d=json.loads(l); print(d['strip'],d['rows'],d['kernel'],d['kind'],d['c'],d['lo'],d['us'],d['frac_of_hbm'])
PY
