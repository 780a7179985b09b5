#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02_upsample_sweep57.jsonl
MI_B200_UPSAMPLE_STRIP=0 python tools/bench_upsample.py >> gpurun_out/r02_upsample_sweep57.jsonl 2>gpurun_out/r02_upsample_err57.log
python tools/bench_upsample.py >> gpurun_out/r02_upsample_sweep57.jsonl 2>>gpurun_out/r02_upsample_err57.log
for r in 1 2 4 8 16; do MI_B200_UPSAMPLE_ROWS=$r python tools/bench_upsample.py >> gpurun_out/r02_upsample_sweep57.jsonl 2>>gpurun_out/r02_upsample_err57.log; done
python - <<'PY'
import json
for l in open('gpurun_out/r02_upsample_sweep57.jsonl'):
    d=json.loads(l); print(d['strip'],d['rows'],d['kernel'],d['kind'],d['c'],d['lo'],d['us'],d['frac_of_hbm'])
PY
tail -3 gpurun_out/r02_upsample_err57.log
