#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest_gpu.log
B="--steps 2 --warmup 3 --no-full-pipeline --no-cpu-baseline --no-eager --quick-e2e"
for g in 1 0; do
  TTVDM_STEP_GRAPH=$g timeout 600 python bench.py $B > gpurun_out/r2g_bench_576_graph$g.json 2> gpurun_out/r2g_bench_576_graph$g.log; echo "bench 576 graph=$g rc=$?"
  TTVDM_STEP_GRAPH=$g timeout 600 python bench.py $B --height 256 --width 384 > gpurun_out/r2g_bench_256_graph$g.json 2> gpurun_out/r2g_bench_256_graph$g.log; echo "bench 256 graph=$g rc=$?"
done
python - <<'PY'
import json
for n in ("576_graph1","576_graph0","256_graph1","256_graph0"):
    try:
        d=json.load(open(f'gpurun_out/r2g_bench_{n}.json'))
        print(n, d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], {k:v['ms'] for k,v in d['kernel_shares'].items() if v['ms']>1})
    except Exception as e: print(n, 'ERR', e)
PY
tail -5 gpurun_out/r2g_bench_576_graph1.log
