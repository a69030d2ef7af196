#!/bin/bash
# full GPU suite (per-test timeout), smoke, default bench (all legs), 256x384 VGL + VL lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2p_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2p_smoke.log
timeout 900 python bench.py > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.log; echo "bench rc=$?"
timeout 600 python bench.py --height 256 --width 384 --no-cpu-baseline --no-eager --no-full-pipeline > gpurun_out/r2p_bench_256_vgl.json 2> gpurun_out/r2p_bench_256_vgl.log; echo "bench256 rc=$?"
python - <<'PY'
import json
for n in ("n1","256_vgl"):
    try:
        d=json.load(open(f'gpurun_out/r2p_bench_{n}.json'))
        print(n, d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], {k:v['ms'] for k,v in d['kernel_shares'].items() if v['ms']>1}, d.get('eager_bf16_gpu'))
    except Exception as e: print(n, 'ERR', e)
PY
