#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:"gn_apply|gn_stats|gemm_kernel|softmax_rows" --launch-skip 196 -c 16 \
   -o /tmp/vae_full -f python tools/vae_time.py --no-cpu --iters 1 --out gpurun_out/vae_time_ncu_full.json > gpurun_out/vae_ncu_full.log 2>&1
echo "ncu full rc=$?"
python tools/ncu_summarize.py /tmp/vae_full.ncu-rep gpurun_out/vae_ncu_full.csv
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_n1.json
