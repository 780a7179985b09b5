#!/bin/bash
# round 2, GPU call 22: state after the SM-sharing commit: GPU suite, default bench (incl. GPU reference + CPU baseline),
# reference GPU numbers of the other configurations, launch list, final ncu captures
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t22_all.log 2>&1
echo "all rc=$?"; tail -4 gpurun_out/r02_t22_all.log | cut -c1-300
timeout 1500 python bench.py > gpurun_out/r02_bench22.json 2> gpurun_out/r02_bench22.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench22.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel']); print({k:v['value'] for k,v in d['config']['other_configs'].items()}); print(d.get('gpu_reference',{}).get('ratio_e2e_over_reference_tf32'), d.get('gpu_reference',{}).get('ratio_e2e_over_reference_fp32'), d['cpu_baseline'])"; tail -3 gpurun_out/r02_bench22.err
for cfg in "superslomo 4" "cain 4 --weight-gain 0.4" "rrin 8"; do
set -- $cfg
CUDA_VISIBLE_DEVICES=0 timeout 600 python baseline/reference_gpu.py --model $1 --batch $2 --steps 3 --warmup 2 $3 $4 2>&1 | tail -1 | cut -c1-400
done | tee gpurun_out/r02_reference_gpu_other_configs.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches22.csv python tools/one_task.py > gpurun_out/r02_one_task22.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches22.csv > gpurun_out/r02_launches22.txt; head -12 gpurun_out/r02_launches22.txt; tail -1 gpurun_out/r02_launches22.txt
ncu --set full --clock-control none --import-source on -k regex:kxs -s 2 -c 1 -o gpurun_out/r02b_ncu_kxs_51x51_258x450 -f python tools/one_conv.py 2 258 450 51 51 > gpurun_out/r02_ncu22a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kxs -s 2 -c 1 -o gpurun_out/r02b_ncu_kxs_64x64_192x256 -f python tools/one_conv.py 2 192 256 64 64 > gpurun_out/r02_ncu22b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kx -s 2 -c 1 -o gpurun_out/r02b_ncu_wgrad_kx_64x64_192x256 -f python tools/one_conv.py 2 192 256 64 64 wgrad > gpurun_out/r02_ncu22c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_fwd_vec -s 6 -c 1 -o gpurun_out/r02b_ncu_warp_fwd_vec -f python tools/bench_warp.py 32 > gpurun_out/r02_ncu22d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_bwd_vec -s 6 -c 1 -o gpurun_out/r02b_ncu_warp_bwd_vec -f python tools/bench_warp.py 32 > gpurun_out/r02_ncu22e.log 2>&1
ls gpurun_out/r02b_*.ncu-rep
