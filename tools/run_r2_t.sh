#!/bin/bash
mkdir -p gpurun_out
for pair in unset 1; do
  if [ $pair = 1 ]; then export TTVDM_GEMM_PAIR=1; fi
  timeout 300 ncu --set full --clock-control none -k regex:gemm_kernel -c 8 -o gpurun_out/prof_l0_pair$pair -f python tools/gemm_one.py > gpurun_out/r2t_ncu_$pair.log 2>&1; echo "ncu pair=$pair rc=$?"
  python tools/ncu_summarize.py gpurun_out/prof_l0_pair$pair.ncu-rep gpurun_out/ncu_gemm_l0_pair_$pair.csv
  rm -f gpurun_out/prof_l0_pair$pair.ncu-rep
done
