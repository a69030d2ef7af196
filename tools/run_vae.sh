#!/bin/bash
# VAE path evidence run on one B200: GPU parity tests of the VAE, timing + full-size parity, ncu launch list of one
# decode + encode at 256x384, then the whole GPU suite.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/vae_smi.txt 2>&1
timeout 420 python -m pytest tests/test_vae_gpu.py -q > gpurun_out/vae_pytest_gpu.log 2>&1; echo "vae pytest rc=$?"; tail -25 gpurun_out/vae_pytest_gpu.log
timeout 420 python tools/vae_time.py --out gpurun_out/vae_time.json > gpurun_out/vae_time.log 2>&1; echo "vae_time rc=$?"; tail -3 gpurun_out/vae_time.log | cut -c1-3000
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/vae_launches.csv \
    python tools/vae_time.py --height 256 --width 384 --iters 1 --no-cpu --out gpurun_out/vae_time_ncu.json > gpurun_out/vae_ncu.log 2>&1
echo "ncu rc=$? lines=$(wc -l < gpurun_out/vae_launches.csv)"
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/all_pytest_gpu.log 2>&1; echo "all pytest rc=$?"; tail -4 gpurun_out/all_pytest_gpu.log
