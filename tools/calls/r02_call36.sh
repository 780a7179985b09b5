#!/bin/bash
mkdir -p gpurun_out
for m in superslomo cain; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches36_$m.csv python tools/one_task_model.py $m > gpurun_out/r02_one_task36_$m.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches36_$m.csv > gpurun_out/r02_launches36_$m.txt; head -22 gpurun_out/r02_launches36_$m.txt; tail -1 gpurun_out/r02_launches36_$m.txt
done
