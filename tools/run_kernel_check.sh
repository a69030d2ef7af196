#!/bin/bash
# Runs the kernel numerics sweep group by group (each under its own timeout so one hung kernel cannot eat the call).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in gemm conv3x3 tconv attn_spatial attn_cross "attn_temporal groupnorm layernorm im2col upsample axpy sampler"; do
  tag=$(echo $grp | cut -d' ' -f1)
  timeout -k 5 ${CHECK_TIMEOUT:-150} python tools/gpu_kernel_check.py $grp > gpurun_out/check_$tag.log 2>&1
  echo "== $tag exit=$?"; grep -E "^(PASS|FAIL|EXC)" gpurun_out/check_$tag.log
done
if [ "$1" == "--time" ]; then
  timeout -k 5 400 python tools/gpu_kernel_check.py zzz --time > gpurun_out/check_time.log 2>&1
  echo "== time exit=$?"; grep -E "^TIME" gpurun_out/check_time.log
fi
