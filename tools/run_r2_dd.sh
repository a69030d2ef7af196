#!/bin/bash
mkdir -p gpurun_out
for wa in 0 1; do
  TTVDM_GEMM_WAVE_AWARE=$wa timeout 300 python tools/shape_table.py --gemm-only > gpurun_out/r2dd_shape_wa$wa.log 2>&1; echo "wa=$wa rc=$?"
  cp gpurun_out/shape_table.json gpurun_out/r2dd_shape_table_wa$wa.json
done
TTVDM_GEMM_WAVE_AWARE=1 timeout 200 python tools/gpu_kernel_check.py gemm conv tconv 2>&1 | grep -E "FAIL|EXC|PASS" | awk '{print $1}' | sort | uniq -c
