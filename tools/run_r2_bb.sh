#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gpu_kernel_check.py upsample conv3x3 2>&1 | grep -E "PASS|FAIL|EXC"
timeout 900 python -m pytest tests -m gpu -q --timeout 200 -x 2>&1 | tail -4
B="--steps 2 --warmup 3 --no-full-pipeline --no-cpu-baseline --no-eager --quick-e2e"
timeout 600 python bench.py $B > gpurun_out/r2bb_bench.json 2> gpurun_out/r2bb_bench.log; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2bb_bench.json')); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k:v['ms'] for k,v in d['kernel_shares'].items()})"
