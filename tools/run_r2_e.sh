#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_kernel_check.py gemm_ pack > gpurun_out/r2e_check.log 2>&1; echo "kernel cases rc=$?"; grep -E "^(FAIL|EXC)" gpurun_out/r2e_check.log; grep -c PASS gpurun_out/r2e_check.log
timeout 600 python tools/shape_table.py --fusions > gpurun_out/r2e_fusions.log 2>&1; echo "fusion table rc=$?"; grep ROW gpurun_out/r2e_fusions.log
TTVDM_GN_TMA_MAXK=2048 timeout 600 python tools/shape_table.py --fusions > gpurun_out/r2e_fusions_maxk2048.log 2>&1; echo "fusion table maxk=2048 rc=$?"; grep "tconv" gpurun_out/r2e_fusions_maxk2048.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "vs reference|vs oracle|rel-L2 per kept|passed|failed|^FAILED|^ERROR" gpurun_out/r2e_pytest_gpu.log | tail -30
timeout 600 python bench.py --steps 2 --warmup 3 --no-full-pipeline --no-cpu-baseline --no-eager --quick-e2e > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.log; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench_n1.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_n1.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])
print({k:(v['ms'],v['launches']) for k,v in d['kernel_shares'].items()})
for g in d['gemm_shapes']: print(g)
PY
