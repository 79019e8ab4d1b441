#!/bin/bash
# One GPU call that regenerates the round's profile evidence (outputs under gpurun_out/; summaries are copied to profiles/ by hand):
#   ncu --set full of the attention / LayerNorm / encoder GEMM shapes and of the T = 1 and T = 3 head stages,
#   the two ncu launch lists, and the eager per-launch CUDA-event lists of the three model configurations.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/ncu_capture.sh attn:attention_kernel ln:layernorm gemm_qkv:gemm_kernel gemm_proj:gemm_kernel gemm_fc1:gemm_kernel gemm_fc2:gemm_kernel
for cfg in "t1:1 2 145" "t3:3 3 64"; do
  name=${cfg%%:*}; args=${cfg#*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 220 -c 8 -f -o gpurun_out/ncu_head_$name \
    python tools/forward_launches.py prithvi_eo_v1_100 $args > gpurun_out/ncu_head_$name.log 2>&1
  echo "head $name exit $?"
done
bash tools/ncu_lists.sh
{
  for cfg in "prithvi_eo_v1_100 3 13 64" "prithvi_eo_v1_100 1 2 256" "prithvi_eo_v2_300 3 13 128"; do
    echo "##### python tools/forward_launches.py $cfg"
    timeout 300 python tools/forward_launches.py $cfg 2>&1 | tail -22
  done
} > gpurun_out/r02_forward_launches.txt
tail -3 gpurun_out/r02_forward_launches.txt
