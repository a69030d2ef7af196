#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_vae_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout 200 2>&1 | tail -3
for f in 1 0; do
  TTVDM_UPSAMPLE_PARITY=$f timeout 300 python tools/vae_time.py --no-cpu --iters 3 --out gpurun_out/r2_vae_time_parity$f.json > gpurun_out/r2_vae_time_parity$f.log 2>&1; echo "vae_time parity=$f rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2_vae_time_parity$f.json')); print('parity=$f', {k:round(v,3) for k,v in d.items() if isinstance(v,(int,float))})" 2>&1 | cut -c1-500
done
