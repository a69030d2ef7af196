#!/bin/bash
mkdir -p gpurun_out
B="--steps 2 --warmup 3 --no-full-pipeline --no-cpu-baseline --no-eager --quick-e2e"
for v in 1 0 1 0; do
  TTVDM_UPSAMPLE_PARITY=$v timeout 600 python bench.py $B > gpurun_out/r2cc_bench_$v.json 2> gpurun_out/r2cc_bench_$v.log
  python -c "
import json; d=json.load(open('gpurun_out/r2cc_bench_$v.json')); print('parity=$v', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['kernel_shares']['gemm']['ms'])"
done
