#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_ab.py 3 13 3 13 > gpurun_out/r2x_attn_ab.log 2>&1; grep -o '"poly_of_8": "[0-9]*"\|"L0": {"own_ms": [0-9.]*, "torch_ms": [0-9.]*' gpurun_out/r2x_attn_ab.log | paste - -
bash tools/run_r2_p.sh
