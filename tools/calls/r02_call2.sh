#!/bin/bash
# round 2, GPU call 2: TF32 round-to-nearest operand convention -- parity, A/B against truncation, cost
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t2_all.log 2>&1
echo "all rc=$?"; tail -30 gpurun_out/r02_t2_all.log | cut -c1-300
cp gpurun_out/parity_full_size.jsonl gpurun_out/r02_parity_full_size_rn.jsonl
rm -f gpurun_out/parity_full_size.jsonl
MI_B200_TF32_RN=0 python -m pytest tests/test_system_gpu.py -m gpu -q --timeout 900 -k "full_size_train_iter" > gpurun_out/r02_t2_trunc.log 2>&1
echo "trunc rc=$?"; tail -4 gpurun_out/r02_t2_trunc.log
cp gpurun_out/parity_full_size.jsonl gpurun_out/r02_parity_full_size_trunc.jsonl
python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
echo "bench rc=$?"; cat gpurun_out/r02_bench2.json; tail -5 gpurun_out/r02_bench2.err
MI_B200_TF32_RN=0 python bench.py --steps 5 --warmup 3 --no-other-configs --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench2_trunc.json 2> gpurun_out/r02_bench2_trunc.err
echo "bench trunc rc=$?"; cat gpurun_out/r02_bench2_trunc.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches2.csv python tools/one_task.py > gpurun_out/r02_one_task2.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches2.csv > gpurun_out/r02_launches2.txt; head -40 gpurun_out/r02_launches2.txt
