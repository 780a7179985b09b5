#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:upsample2_fwd_strip -s 6 -c 1 -o gpurun_out/r02g_ncu_upsample_fwd_strip -f python tools/bench_upsample.py > gpurun_out/r02_ncu59a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:upsample2_bwd_strip -s 6 -c 1 -o gpurun_out/r02g_ncu_upsample_bwd_strip -f python tools/bench_upsample.py > gpurun_out/r02_ncu59b.log 2>&1
python tools/ncu_extract.py gpurun_out/r02g_ncu_upsample_fwd_strip.ncu-rep
python tools/ncu_extract.py gpurun_out/r02g_ncu_upsample_bwd_strip.ncu-rep
