#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "all pytest rc=$?"; tail -6 gpurun_out/final_pytest_gpu.log
timeout 120 python tools/vae_time.py --no-cpu --out gpurun_out/vae_time_final.json > gpurun_out/vae_time_final.log 2>&1; echo "vae_time rc=$?"; tail -1 gpurun_out/vae_time_final.log | cut -c1-1200
