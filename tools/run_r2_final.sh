#!/bin/bash
# final tree: full GPU suite, smoke, per-shape table (own vs torch), default bench (all legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/r2final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python tools/shape_table.py > gpurun_out/r2final_shape_table.log 2>&1; echo "shape table rc=$?"
timeout 900 python bench.py > gpurun_out/r2final_bench_n1.json 2> gpurun_out/r2final_bench_n1.log; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2final_bench_n1.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], {k:v['ms'] for k,v in d['kernel_shares'].items() if v['ms']>1})"
