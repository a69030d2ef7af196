#!/bin/bash
mkdir -p gpurun_out
B="--height 256 --width 384 --no-cpu-baseline --no-eager --no-full-pipeline"
timeout 300 python bench.py $B > gpurun_out/r2final_bench_256_vgl.json 2> gpurun_out/r2final_bench_256_vgl.log; echo "vgl rc=$?"
timeout 300 python bench.py --vl $B > gpurun_out/r2final_bench_256_vl.json 2> gpurun_out/r2final_bench_256_vl.log; echo "vl rc=$?"
python -c "
import json
for n in ('vgl','vl'):
    d=json.load(open(f'gpurun_out/r2final_bench_256_{n}.json')); print(n, d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['whole_step']['frac'])"
