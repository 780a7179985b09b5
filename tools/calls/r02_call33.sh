#!/bin/bash
# round 2, GPU call 33: ncu of the 3x3 kernel at the SM share of the real run (37 CTAs); smoke()
mkdir -p gpurun_out
MI_B200_SM_BUDGET=37 ncu --set full --clock-control none --import-source on -k regex:kxs -s 2 -c 1 -o gpurun_out/r02c_ncu_kxs_51x51_258x450_37ctas -f python tools/one_conv.py 2 258 450 51 51 > gpurun_out/r02_ncu33a.log 2>&1
MI_B200_SM_BUDGET=37 ncu --set full --clock-control none --import-source on -k regex:kxs -s 2 -c 1 -o gpurun_out/r02c_ncu_kxs_64x64_192x256_37ctas -f python tools/one_conv.py 2 192 256 64 64 > gpurun_out/r02_ncu33b.log 2>&1
ls gpurun_out/r02c_*.ncu-rep
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
