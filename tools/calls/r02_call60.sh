#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_flow_kernels_gpu.py -m gpu -q --timeout 300 -x > gpurun_out/r02_t60_kernels.log 2>&1
echo "kernels rc=$?"; tail -2 gpurun_out/r02_t60_kernels.log | cut -c1-300
python tools/bench_upsample.py > gpurun_out/r02_upsample_60.jsonl 2>gpurun_out/r02_upsample_err60.log
python - <<'PY'
import json
for l in open('gpurun_out/r02_upsample_60.jsonl'):
    d=json.loads(l); print(d['strip'],d['rows'],d['kernel'],d['kind'],d['c'],d['lo'],d['us'],d['frac_of_hbm'])
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches60.csv python tools/one_task.py > gpurun_out/r02_one_task60.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches60.csv > gpurun_out/r02_launches60.txt; grep "upsample\|TOTAL" gpurun_out/r02_launches60.txt
ncu --set full --clock-control none --import-source on -k regex:upsample2_fwd_strip -s 6 -c 1 -o gpurun_out/r02g_ncu_upsample_fwd_strip -f python tools/bench_upsample.py > gpurun_out/r02_ncu60a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:upsample2_bwd_strip -s 6 -c 1 -o gpurun_out/r02g_ncu_upsample_bwd_strip -f python tools/bench_upsample.py > gpurun_out/r02_ncu60b.log 2>&1
python tools/ncu_extract.py gpurun_out/r02g_ncu_upsample_fwd_strip.ncu-rep | grep "duration\|dram\|warps_active\|sm__throughput"
python tools/ncu_extract.py gpurun_out/r02g_ncu_upsample_bwd_strip.ncu-rep | grep "duration\|dram\|warps_active\|sm__throughput"
