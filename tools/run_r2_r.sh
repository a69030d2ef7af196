#!/bin/bash
mkdir -p gpurun_out
for pf in 0 1 2; do
  TTVDM_GEMM_L2_PREFETCH=$pf timeout 300 python tools/shape_table.py --gemm-only > gpurun_out/r2r_shape_pf$pf.log 2>&1; echo "pf=$pf rc=$?"
  cp gpurun_out/shape_table.json gpurun_out/r2r_shape_table_pf$pf.json
done
timeout 200 python tools/gpu_kernel_check.py gemm conv tconv 2>&1 | grep -E "FAIL|EXC|PASS" | awk '{print $1}' | sort | uniq -c
