#!/bin/bash
# launch list of one VGL Euler step on the final tree (the full captures of tools/run_profile.sh are unchanged kernels)
mkdir -p gpurun_out
export TTVDM_STEP_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-only > gpurun_out/profile_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
