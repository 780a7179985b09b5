#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t42_all.log 2>&1
echo "all rc=$?"; tail -2 gpurun_out/r02_t42_all.log | cut -c1-300
MI_B200_POISON=1 timeout 1500 python -m pytest tests/test_conv_tc_gpu.py tests/test_system_gpu.py -m gpu -q --timeout 900 > gpurun_out/r02_t42_poison.log 2>&1
echo "poison rc=$?"; tail -2 gpurun_out/r02_t42_poison.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches42.csv python tools/one_task.py > gpurun_out/r02_one_task42.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches42.csv > gpurun_out/r02_launches42.txt; head -8 gpurun_out/r02_launches42.txt; tail -1 gpurun_out/r02_launches42.txt
