#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench49_8gpu.json 2> gpurun_out/r02_bench49_8gpu.err
echo "8gpu rc=$?"; python -c "
import json
l=[x for x in open('gpurun_out/r02_bench49_8gpu.json') if x.startswith('{')][-1]
d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']); print({k:v.get('value') for k,v in d['config']['other_configs'].items()})"; tail -3 gpurun_out/r02_bench49_8gpu.err
