#!/bin/bash
# 2 GPUs: split-pair equality over NCCL (tiny + SVD at the headline size), bench K=1 (split pair) and K=2 (whole pairs).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2n2_smi.txt
timeout 900 python -m pytest tests/test_sharding_nccl.py -m gpu -q > gpurun_out/r2n2_pytest.log 2>&1; echo "nccl test rc=$?"; tail -2 gpurun_out/r2n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/split_pair_check.py > gpurun_out/r2n2_split_svd.log 2>&1; echo "split svd rc=$?"; grep SPLIT_PAIR gpurun_out/r2n2_split_svd.log
B="--steps 1 --warmup 3 --no-full-pipeline --no-cpu-baseline --no-eager --quick-e2e"
for K in 1 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$K bench.py --gpus 2 --videos $K $B > gpurun_out/r2_sweep_n2_k$K.json 2> gpurun_out/r2_sweep_n2_k$K.log; echo "bench n2 k$K rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2_sweep_n2_k$K.json')); print('N=2 K=$K', d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['parallelism'][:60])"
done
