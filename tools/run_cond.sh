#!/bin/bash
# Conditioning builder + whole drop-in pipeline evidence run on one B200.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_clip_gpu.py tests/test_pipeline_gpu.py -q > gpurun_out/cond_pytest_gpu.log 2>&1; echo "cond pytest rc=$?"; tail -30 gpurun_out/cond_pytest_gpu.log
timeout 300 python tools/cond_time.py --out gpurun_out/cond_time.json > gpurun_out/cond_time.log 2>&1; echo "cond_time rc=$?"; tail -2 gpurun_out/cond_time.log | cut -c1-2000
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/all_pytest_gpu.log 2>&1; echo "all pytest rc=$?"; tail -4 gpurun_out/all_pytest_gpu.log
