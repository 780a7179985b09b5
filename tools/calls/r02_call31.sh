#!/bin/bash
# round 2, GPU call 31 (2 GPUs): weak scaling sanity of the 8-lane / SM-budget path, reference arm under torchrun
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-other-configs > gpurun_out/r02_bench31_2gpu.json 2> gpurun_out/r02_bench31_2gpu.err
echo "2gpu rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench31_2gpu.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])"; tail -3 gpurun_out/r02_bench31_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_bench31_ref_2gpu.json 2> gpurun_out/r02_bench31_ref_2gpu.err
echo "ref rc=$?"; cat gpurun_out/r02_bench31_ref_2gpu.json | cut -c1-600; tail -2 gpurun_out/r02_bench31_ref_2gpu.err
