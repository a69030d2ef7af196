#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gpu_kernel_check.py attn_temporal 2>&1 | grep -E "PASS|FAIL|EXC"
timeout 900 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/r2v_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2v_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python tools/gpu_kernel_check.py --time 2>&1 | grep "TIME attn_temporal"
