#!/bin/bash
# Run every probe family in its own process (a device trap must not take the others down).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for fam in ${@:-pre stitch ln gemm attn model bench}; do
  echo "=== $fam ===" 
  case $fam in pre|stitch|model) script=tests/probe_parity.py;; *) script=tools/gpu_probe.py;; esac
  timeout 300 python $script $fam > gpurun_out/probe_$fam.log 2>&1
  echo "exit $?" >> gpurun_out/probe_$fam.log
  tail -n 60 gpurun_out/probe_$fam.log
done
