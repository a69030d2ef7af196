#!/bin/bash
# Evidence run for the scope table's "next" rows (VAE path, conditioning builder, whole drop-in pipeline) on one B200:
#   gpurun --timeout 900 -- 'bash tools/run_next_rows.sh'
# Everything lands in gpurun_out/ (the .ncu-rep stays in /tmp: gpurun only brings back 64 MiB); the files that back
# DESIGN.md §7d-§7f are copied to profiles/r01c_* by hand afterwards.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/next_smi.txt 2>&1
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/next_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/next_pytest_gpu.log
timeout 420 python tools/vae_time.py --out gpurun_out/vae_time.json > gpurun_out/vae_time.log 2>&1; echo "vae_time rc=$?"
timeout 180 python tools/cond_time.py --out gpurun_out/cond_time.json > gpurun_out/cond_time.log 2>&1; echo "cond_time rc=$?"
# launch list of one decode + encode at 256x384 (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/vae_launches.csv \
    python tools/vae_time.py --height 256 --width 384 --iters 1 --no-cpu --out gpurun_out/vae_time_ncu.json > gpurun_out/vae_ncu.log 2>&1
echo "ncu launch list rc=$? lines=$(wc -l < gpurun_out/vae_launches.csv)"
# full capture of 16 launches of the decoder's full-resolution block, summarised on the box
timeout 300 ncu --set full --clock-control none -k regex:"gn_apply|gn_stats|gemm_kernel|softmax_rows" --launch-skip 196 -c 16 \
    -o /tmp/vae_full -f python tools/vae_time.py --no-cpu --iters 1 --out gpurun_out/vae_time_ncu_full.json > gpurun_out/vae_ncu_full.log 2>&1
python tools/ncu_summarize.py /tmp/vae_full.ncu-rep gpurun_out/vae_ncu_full.csv
timeout 330 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.log; echo "bench rc=$?"
