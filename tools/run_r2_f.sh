#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 7 -c 7 -o gpurun_out/r2f_lnfold -f python tools/ncu_lnfold.py > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2f_ncu.log
ls -la gpurun_out/r2f_lnfold.ncu-rep
