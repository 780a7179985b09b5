#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"filters_to_planar|planar_to_filters" -s 4 -c 2 -o gpurun_out/r02g_ncu_sepconv_transposes -f python tools/bench_sepconv.py > gpurun_out/r02_ncu63.log 2>&1
python tools/ncu_extract.py gpurun_out/r02g_ncu_sepconv_transposes.ncu-rep
ncu -i gpurun_out/r02g_ncu_sepconv_transposes.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
  print('==',r[h.index('Kernel Name')][:40])
  for k,v in zip(h,r):
    if any(s in k for s in ['smsp__average_warps_issue_stalled','l1tex__t_sector_hit_rate','lts__t_sector_hit_rate.pct','smsp__issue_active.avg.pct','lts__t_sectors_srcunit_tex_op_write.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','dram__sectors','l1tex__data_bank_conflicts_pipe_lsu_mem_shared','smsp__inst_executed.sum ','launch__occupancy_limit','achieved_occupancy','launch__waves']):
      try:
        if float(v.replace(',',''))!=0: print('  ',k,v)
      except: pass
"
