#!/bin/bash
# round 2, GPU call 12: kx-stacked kernel as the default 3x3 engine: whole GPU suite, bench, launch list, ncu
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_t12_all.log 2>&1
echo "all rc=$?"; tail -8 gpurun_out/r02_t12_all.log | cut -c1-300
python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench12.json 2> gpurun_out/r02_bench12.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench12.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac']); print(d['roofline']['per_kernel'])"; tail -3 gpurun_out/r02_bench12.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches12.csv python tools/one_task.py > gpurun_out/r02_one_task12.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches12.csv > gpurun_out/r02_launches12.txt; head -24 gpurun_out/r02_launches12.txt
ncu --set full --clock-control none --import-source on -k regex:kxs -s 2 -c 1 -o gpurun_out/r02_ncu_kxs_51x51_258x450 -f python tools/one_conv.py 2 258 450 51 51 > gpurun_out/r02_ncu12a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kxs -s 2 -c 1 -o gpurun_out/r02_ncu_kxs_64x64_192x256 -f python tools/one_conv.py 2 192 256 64 64 > gpurun_out/r02_ncu12b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kxs -s 2 -c 1 -o gpurun_out/r02_ncu_kxs_256x256_48x64 -f python tools/one_conv.py 2 48 64 256 256 > gpurun_out/r02_ncu12c.log 2>&1
ls gpurun_out/*.ncu-rep | tail -4
