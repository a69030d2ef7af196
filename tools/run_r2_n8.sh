#!/bin/bash
# 8 GPUs: K = 8 videos (whole CFG pairs, one per rank) and K = 4 (split pairs: every video on two ranks, eps exchange per step)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2n8_smi.txt
B="--steps 1 --warmup 3 --no-full-pipeline --no-cpu-baseline --no-eager --quick-e2e"
for K in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$K bench.py --gpus 8 --videos $K $B > gpurun_out/r2_sweep_n8_k$K.out 2> gpurun_out/r2_sweep_n8_k$K.log; echo "bench n8 k$K rc=$?"
  grep '^{' gpurun_out/r2_sweep_n8_k$K.out > gpurun_out/r2_sweep_n8_k$K.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_sweep_n8_k$K.json')); print('N=8 K=$K', d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['parallelism'][:60])"
done
