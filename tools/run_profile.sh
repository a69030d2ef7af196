#!/bin/bash
# ncu evidence for one VGL Euler step at the bench workload (576x1024): launch list + full captures of the top kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-only > gpurun_out/profile_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel -s 200 -c 3 \
    -o gpurun_out/prof_gemm -f python bench.py --profile-only > gpurun_out/profile_gemm.log 2>&1
echo "gemm full rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_flash -c 2 \
    -o gpurun_out/prof_attn -f python bench.py --profile-only > gpurun_out/profile_attn.log 2>&1
echo "attn full rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:gn_|layernorm|attn_temporal" -c 28 \
    -o gpurun_out/prof_gn -f python bench.py --profile-only > gpurun_out/profile_gn.log 2>&1
echo "gn / layernorm / attn_temporal full rc=$?"
ls -la gpurun_out/
