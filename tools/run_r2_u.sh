#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gpu_kernel_check.py attn_cross attn_temporal 2>&1 | grep -E "PASS|FAIL|EXC"
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 300 -k "head_dim_128 or reference_own_forward" 2>&1 | tail -15
