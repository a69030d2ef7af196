#!/bin/bash
# Round 2, second GPU call: epilogue-fused norm statistics + LayerNorm fold — kernel cases, whole GPU suite, bench.
mkdir -p gpurun_out
timeout 300 python tools/gpu_kernel_check.py gemm_gnstats gemm_lnfold > gpurun_out/r2b_check_fusions.log 2>&1; echo "fusion cases rc=$?"
grep -E "^(PASS|FAIL|EXC)" gpurun_out/r2b_check_fusions.log
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "vs reference|vs oracle|rel-L2 per kept|passed|failed|Error|error" gpurun_out/r2b_pytest_gpu.log | tail -20
timeout 600 python bench.py --steps 2 --warmup 3 --no-full-pipeline --no-cpu-baseline --no-eager --quick-e2e > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_n1.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])
print({k:(v['ms'],v['launches']) for k,v in d['kernel_shares'].items()})
for g in d['gemm_shapes']: print(g)
PY
TTVDM_GEMM_PAIR=1 timeout 600 python tools/shape_table.py > gpurun_out/r2b_shape_table_pair1.log 2>&1; echo "shape_table pair=1 rc=$?"
cp gpurun_out/shape_table.json gpurun_out/r2b_shape_table_pair1.json
