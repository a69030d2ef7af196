#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_kernel_check.py gemm_ pack > gpurun_out/r2d_check.log 2>&1; echo "kernel cases rc=$?"; grep -E "^(FAIL|EXC)" gpurun_out/r2d_check.log; grep -c PASS gpurun_out/r2d_check.log
timeout 600 python tools/gpu_determinism.py > gpurun_out/r2d_determinism.log 2>&1; echo "determinism rc=$?"; grep -E "run-to-run|emulation|Error" gpurun_out/r2d_determinism.log
timeout 600 python tools/shape_table.py --fusions > gpurun_out/r2d_fusions.log 2>&1; echo "fusion table rc=$?"; grep ROW gpurun_out/r2d_fusions.log
