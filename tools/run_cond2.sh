#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pipeline_gpu.py -q > gpurun_out/pipe_pytest_gpu.log 2>&1; echo "pipeline pytest rc=$?"; tail -15 gpurun_out/pipe_pytest_gpu.log
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/all_pytest_gpu.log 2>&1; echo "all pytest rc=$?"; tail -4 gpurun_out/all_pytest_gpu.log
timeout 400 ncu --set full --clock-control none -k regex:"gn_apply|gn_stats|gemm_kernel|softmax_rows" --launch-skip 190 -c 26 \
   -o gpurun_out/vae_full -f python tools/vae_time.py --no-cpu --iters 1 --out gpurun_out/vae_time_ncu_full.json > gpurun_out/vae_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/vae_full.ncu-rep
