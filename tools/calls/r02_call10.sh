#!/bin/bash
mkdir -p gpurun_out
MI_B200_KXS=2 ncu --set full --clock-control none --import-source on -k regex:kxs -s 2 -c 1 -o gpurun_out/r02_ncu_kxs_51x51_v0 -f python tools/one_conv.py 2 258 450 51 51 > gpurun_out/r02_ncu10a.log 2>&1
tail -3 gpurun_out/r02_ncu10a.log
MI_B200_KXS=0 ncu --set full --clock-control none --import-source on -k regex:halo_kernel -s 2 -c 1 -o gpurun_out/r02_ncu_halo_51x51_r02 -f python tools/one_conv.py 2 258 450 51 51 > gpurun_out/r02_ncu10b.log 2>&1
tail -3 gpurun_out/r02_ncu10b.log
