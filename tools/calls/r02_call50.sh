#!/bin/bash
# round 2, GPU call 50: final-state ncu captures of the weight-gradient kernel
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kx -s 2 -c 1 -o gpurun_out/r02e_ncu_wgrad_kx_64x64_192x256 -f python tools/one_conv.py 2 192 256 64 64 wgrad > gpurun_out/r02_ncu50a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kx -s 2 -c 1 -o gpurun_out/r02e_ncu_wgrad_kx_51x51_258x450 -f python tools/one_conv.py 2 258 450 51 51 wgrad > gpurun_out/r02_ncu50b.log 2>&1
MI_B200_SM_BUDGET=37 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kx -s 2 -c 1 -o gpurun_out/r02e_ncu_wgrad_kx_64x64_192x256_37ctas -f python tools/one_conv.py 2 192 256 64 64 wgrad > gpurun_out/r02_ncu50c.log 2>&1
ls gpurun_out/r02e_*.ncu-rep
