#!/bin/bash
# 2-GPU sanity run of the final state (the driver's scaling run launches bench.py exactly like this)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02g_bench_2gpu.json 2> gpurun_out/r02g_bench_2gpu.err
echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_2gpu.json')); print(d['value'], d['e2e']['value'], d['n_gpus'], d['config'].get('other_configs'))"; tail -3 gpurun_out/r02g_bench_2gpu.err
