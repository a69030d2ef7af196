#!/bin/bash
# ncu evidence for one VGL Euler step at the bench workload (576x1024): launch list + full captures of the top kernels.
# The .ncu-rep files are summarised ON THE BOX (tools/ncu_summarize.py) and deleted: gpurun only copies 64 MiB back.
mkdir -p gpurun_out
export TTVDM_STEP_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --profile-only > gpurun_out/profile_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel -s 200 -c 6 \
    -o gpurun_out/prof_gemm -f python bench.py --profile-only > gpurun_out/profile_gemm.log 2>&1
echo "gemm full rc=$?"
python tools/ncu_summarize.py gpurun_out/prof_gemm.ncu-rep gpurun_out/ncu_full_gemm.csv && rm -f gpurun_out/prof_gemm.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_flash -c 2 \
    -o gpurun_out/prof_attn -f python bench.py --profile-only > gpurun_out/profile_attn.log 2>&1
echo "attn full rc=$?"
python tools/ncu_summarize.py gpurun_out/prof_attn.ncu-rep gpurun_out/ncu_full_attn.csv && rm -f gpurun_out/prof_attn.ncu-rep
timeout 600 ncu --set full --clock-control none --profile-from-start off -k "regex:gn_|layernorm|attn_temporal|attn_cross" -c 24 \
    -o gpurun_out/prof_gn -f python bench.py --profile-only > gpurun_out/profile_gn.log 2>&1
echo "gn / layernorm / attn_temporal / attn_cross full rc=$?"
python tools/ncu_summarize.py gpurun_out/prof_gn.ncu-rep gpurun_out/ncu_full_norms_small_attn.csv && rm -f gpurun_out/prof_gn.ncu-rep
ls -la gpurun_out/
