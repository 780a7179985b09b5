#!/bin/bash
# final state of round 2: whole GPU suite, default bench line (reference on the same GPU, CPU arm, other configurations)
mkdir -p gpurun_out
rm -f gpurun_out/parity_full_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_t65_all.log 2>&1
echo "all rc=$?"; tail -2 gpurun_out/r02_t65_all.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02g_bench_1gpu_final.json 2> gpurun_out/r02g_bench_1gpu_final.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_1gpu_final.json')); print(d['value'], d['e2e']['value'], d.get('gpu_reference'), d['roofline'].get('frac'), d['config'].get('other_configs'), d.get('clocks'))"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
