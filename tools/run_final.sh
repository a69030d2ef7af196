#!/bin/bash
# Round-end evidence run on one B200: GPU tests, both bench arms, ncu launch list of one Euler step, full captures of
# the attention kernels (the gemm / norm / temporal captures come from tools/run_profile.sh).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/final_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest_gpu.log
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.log; echo "reference rc=$?"
timeout 600 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.log; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/final_launches.csv python bench.py --profile-only > gpurun_out/final_profile_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/final_launches.csv)"
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:attn_flash|attn_cross" -c 3 \
    -o gpurun_out/final_prof_attn -f python bench.py --profile-only > gpurun_out/final_profile_attn.log 2>&1
echo "attn full rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/final_bench_n1.json')); print(d['value'], d['e2e']['value'], d['clocks']); print({k:(v['ms']) for k,v in d['kernel_shares'].items()}); print(d['roofline'])"
