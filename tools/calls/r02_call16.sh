#!/bin/bash
# round 2, GPU call 16: 8 lanes + per-call SM budget as defaults: GPU suite, bench with the other configurations
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t16_all.log 2>&1
echo "all rc=$?"; tail -4 gpurun_out/r02_t16_all.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench16.json 2> gpurun_out/r02_bench16.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench16.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac']); print({k:v['value'] for k,v in d['config']['other_configs'].items()})"; tail -3 gpurun_out/r02_bench16.err
for s in "2 258 450 51 51" "2 192 256 64 64" "2 384 512 32 32" "2 48 64 256 256"; do
MI_B200_DEBUG_TIMING=1 timeout 120 python tools/one_conv.py $s 2>&1 | tail -1
done
