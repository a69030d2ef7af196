#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_vae_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout 200 2>&1 | tail -3
for f in 1 0; do
  TTVDM_VAE_FUSE_GN=$f timeout 300 python tools/vae_time.py --no-cpu --iters 3 --out gpurun_out/r2_vae_time_fuse$f.json > gpurun_out/r2_vae_time_fuse$f.log 2>&1; echo "vae_time fuse=$f rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2_vae_time_fuse$f.json')); print('fuse=$f', {k:v for k,v in d.items() if isinstance(v,(int,float))}, d.get('shares', d.get('decode_shares')))" 2>&1 | cut -c1-600
done
