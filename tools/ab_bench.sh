#!/bin/bash
# A/B of side libraries (tools/variants.sh build ...) on the default bench step, alternating order, N rounds:
#   tools/ab_bench.sh 2 arr1 arr2      (prints ms/step per variant per round; "shipped" = the in-tree library)
cd "$(dirname "$0")/.."
PKG=instageo-e2e-geospatial-ml_b200
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for name in shipped "$@"; do
    lib=""; [ $name != shipped ] && lib=$PWD/$PKG/libig_$name.so
    INSTAGEO_B200_LIB=$lib python bench.py --steps 30 --warmup 5 --no-tile --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=d['kernel_families']
print('round $r %-8s ms/step %.3f  clk %s  lin %.3f conv %.3f attn %.3f ln %.3f' % ('$name', d['ms_per_step'], d['clocks']['sm_mhz'], f['gemm_linear']['ms_per_step'], f['gemm_conv']['ms_per_step'], f['attention']['ms_per_step'], f['layernorm']['ms_per_step']))"
  done
done
