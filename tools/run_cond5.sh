#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_clip_gpu.py tests/test_pipeline_gpu.py -q > gpurun_out/cond_pytest_gpu.log 2>&1; echo "cond pytest rc=$?"; tail -12 gpurun_out/cond_pytest_gpu.log
timeout 120 python tools/cond_time.py --out gpurun_out/cond_time.json > gpurun_out/cond_time.log 2>&1; echo "cond_time rc=$?"; tail -1 gpurun_out/cond_time.log | cut -c1-1500
