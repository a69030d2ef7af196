#!/bin/bash
mkdir -p gpurun_out
tools/microbench/mma_chain > gpurun_out/r2h_mma_chain.txt 2>&1; echo "mma_chain rc=$?"
timeout 600 python tools/shape_table.py --attn > gpurun_out/r2h_attn_table.log 2>&1; echo "attn table rc=$?"; tail -6 gpurun_out/r2h_attn_table.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2h_pytest_gpu.log
