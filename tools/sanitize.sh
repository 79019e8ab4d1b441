#!/bin/bash
# compute-sanitizer over every kernel family on tiny shapes: memcheck, then racecheck on the kernels
# with hand-rolled mbarrier / TMEM / cluster protocols.  Run on a GPU box: gpurun -- bash tools/sanitize.sh
# Logs: gpurun_out/sanitize_{memcheck,racecheck}.log ; summary lines are copied to profiles/ by hand.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export IG_NO_GRAPH=1
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  log=gpurun_out/sanitize_${tool}.log
  : > "$log"
  for part in pre ops model stitch aux; do
    echo "===== $tool / $part" >> "$log"
    timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitize_target.py $part >> "$log" 2>&1
    echo "exit code $?" >> "$log"
  done
  grep -E "=====|ERROR SUMMARY|RACECHECK SUMMARY|exit code|SANITIZE TARGET DONE" "$log"
done
