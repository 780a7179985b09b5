#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sepconv_fwd_quad -s 3 -c 1 -o gpurun_out/r02_ncu_sepconv_fwd_quad -f python tools/bench_sepconv.py > gpurun_out/r02_ncu6a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sepconv_bwd_quad -s 3 -c 1 -o gpurun_out/r02_ncu_sepconv_bwd_quad -f python tools/bench_sepconv.py > gpurun_out/r02_ncu6b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sepconv_fwd_kernel -s 3 -c 1 -o gpurun_out/r02_ncu_sepconv_fwd_gen1 -f python tools/bench_sepconv.py > gpurun_out/r02_ncu6c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
