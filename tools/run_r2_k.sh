#!/bin/bash
# hang-safe probe of a new attention variant, then the sweep
mkdir -p gpurun_out
TTVDM_ATTN_PINGPONG=1 timeout 40 python tools/attn_one.py 1024 > gpurun_out/r2k_probe.log 2>&1; rc=$?; echo "probe rc=$rc"; tail -2 gpurun_out/r2k_probe.log
if [ $rc -eq 0 ]; then timeout 500 python tools/attn_ab.py 3:1:1 2:1:1 4:1:1 0:1:1 3:1:0; fi
