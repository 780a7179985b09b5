#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"filters_to_planar|planar_to_filters" -s 10 -c 40 -o gpurun_out/r02g_ncu_sepconv_transposes_float4 -f python tools/bench_sepconv.py > gpurun_out/r02_ncu69.log 2>&1
python tools/ncu_extract.py gpurun_out/r02g_ncu_sepconv_transposes_float4.ncu-rep > gpurun_out/r02g_ncu_sepconv_transposes_float4.txt
grep "kernel:\|duration\|dram__bytes\|dram_throughput\|sm__throughput" gpurun_out/r02g_ncu_sepconv_transposes_float4.txt | tail -24
