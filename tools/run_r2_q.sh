#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/shape_table.py --fusions 2>&1 | grep -i "groupnorm"
timeout 300 ncu --set full --clock-control none -k regex:gemm_kernel -c 8 -o gpurun_out/prof_gemm_l0 -f python tools/gemm_one.py > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summarize.py gpurun_out/prof_gemm_l0.ncu-rep gpurun_out/ncu_full_gemm_l0_shapes.csv
ncu -i gpurun_out/prof_gemm_l0.ncu-rep --page details --csv > gpurun_out/ncu_details_gemm_l0_shapes.csv 2>/dev/null
rm -f gpurun_out/prof_gemm_l0.ncu-rep
