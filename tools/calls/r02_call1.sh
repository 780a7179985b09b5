#!/bin/bash
# round 2, GPU call 1: the new parity tests first, then the whole GPU suite, then bench.py (with the GPU reference leg)
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
python -m pytest tests/test_reference_gpu.py -m gpu -q -x --timeout 900 > gpurun_out/r02_t_refgpu.log 2>&1
echo "refgpu rc=$?"; tail -5 gpurun_out/r02_t_refgpu.log
python -m pytest tests/test_system_gpu.py -m gpu -q --timeout 900 -k "full_size_train_iter or run_test_iter or super_loss" > gpurun_out/r02_t_new.log 2>&1
echo "new rc=$?"; tail -15 gpurun_out/r02_t_new.log
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t_all.log 2>&1
echo "all rc=$?"; tail -8 gpurun_out/r02_t_all.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r02_clocks1.csv &
SMI=$!
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
echo "bench rc=$?"; cat gpurun_out/r02_bench1.json; tail -5 gpurun_out/r02_bench1.err
kill $SMI
