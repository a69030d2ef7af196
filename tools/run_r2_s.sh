#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/gemm_knobs.py
B="--no-full-pipeline --no-cpu-baseline --no-eager"
timeout 600 python bench.py --vl --height 256 --width 384 $B > gpurun_out/r2s_bench_256_vl.json 2> gpurun_out/r2s_bench_256_vl.log; echo "vl256 rc=$?"
timeout 600 python bench.py --vl $B > gpurun_out/r2s_bench_576_vl.json 2> gpurun_out/r2s_bench_576_vl.log; echo "vl576 rc=$?"
python - <<'PY'
import json
for n in ("256_vl","576_vl"):
    try:
        d=json.load(open(f'gpurun_out/r2s_bench_{n}.json'))
        print(n, d['metric'], d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['whole_step'])
    except Exception as e: print(n, 'ERR', e); print(open(f'gpurun_out/r2s_bench_{n}.log').read()[-800:])
PY
