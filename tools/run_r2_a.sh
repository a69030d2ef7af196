#!/bin/bash
# Round 2, first GPU call: sanity of the inherited tree, exp2 microbenchmark, own-vs-torch shape table, bench with the new legs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 120 tools/microbench/ex2_rate > gpurun_out/r2a_ex2_rate.txt 2>&1; echo "ex2_rate rc=$?"; cat gpurun_out/r2a_ex2_rate.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2a_pytest_gpu.log
timeout 900 python tools/shape_table.py > gpurun_out/r2a_shape_table.log 2>&1; echo "shape_table rc=$?"; grep -c ROW gpurun_out/r2a_shape_table.log
cp gpurun_out/shape_table.json gpurun_out/r2a_shape_table.json
timeout 900 python bench.py --steps 2 --warmup 3 --no-full-pipeline > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2a_bench_n1.json'))
print(d['value'], d['e2e'], d['clocks'])
print({k:(v['ms']) for k,v in d['kernel_shares'].items()})
print(d['eager_bf16_gpu']); print(d['cpu_baseline'])
PY
