#!/bin/bash
# Round 2, third GPU call: shared-memory GroupNorm accumulators + partial row sums (no per-tile atomics).
mkdir -p gpurun_out
timeout 300 python tools/gpu_kernel_check.py pack gemm_gnstats gemm_lnfold > gpurun_out/r2c_check_fusions.log 2>&1; echo "fusion cases rc=$?"
grep -E "^(PASS|FAIL|EXC)" gpurun_out/r2c_check_fusions.log
timeout 600 python tools/shape_table.py --fusions > gpurun_out/r2c_fusions.log 2>&1; echo "fusion table rc=$?"; grep ROW gpurun_out/r2c_fusions.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "vs reference|vs oracle|rel-L2 per kept|passed|failed|^FAILED|^ERROR" gpurun_out/r2c_pytest_gpu.log | tail -30
timeout 600 python bench.py --steps 2 --warmup 3 --no-full-pipeline --no-cpu-baseline --no-eager --quick-e2e > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench_n1.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])
print({k:(v['ms'],v['launches']) for k,v in d['kernel_shares'].items()})
for g in d['gemm_shapes']: print(g)
PY
TTVDM_GEMM_CONTIG=1 timeout 600 python tools/shape_table.py > gpurun_out/r2c_shape_table_contig1.log 2>&1; echo "shape_table contig=1 rc=$?"
cp gpurun_out/shape_table.json gpurun_out/r2c_shape_table_contig1.json
