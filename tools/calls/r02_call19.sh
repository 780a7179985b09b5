#!/bin/bash
# round 2, GPU call 19: where does RRIN lose under a 37-CTA budget? launch lists of one task under both budgets
mkdir -p gpurun_out
for b in 37 147; do
MI_B200_SM_BUDGET=$b timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches19_rrin_$b.csv python tools/one_task_model.py rrin > gpurun_out/r02_one_task19_$b.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches19_rrin_$b.csv > gpurun_out/r02_launches19_rrin_$b.txt; head -16 gpurun_out/r02_launches19_rrin_$b.txt; tail -1 gpurun_out/r02_launches19_rrin_$b.txt
done
