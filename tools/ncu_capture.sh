#!/bin/bash
# ncu --set full capture of single kernels: tools/ncu_capture.sh <target>:<kernel-regex> ...
mkdir -p gpurun_out
for spec in "$@"; do
  tgt=${spec%%:*}; rx=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f \
     -o gpurun_out/ncu_$tgt python tools/ncu_target.py $tgt > gpurun_out/ncu_$tgt.log 2>&1
  echo "$tgt exit $?"
done
