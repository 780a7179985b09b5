#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench34.json 2> gpurun_out/r02_bench34.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench34.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], r['frac'], r['achieved'], r['at_sm_share'], r['share_of_step'])"; tail -3 gpurun_out/r02_bench34.err
