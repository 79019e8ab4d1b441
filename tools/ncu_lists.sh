#!/bin/bash
# ncu launch lists (per-launch device time + DRAM bytes; cold-cache, serialised: compare SHARES) of the default bench
# step and of one tile pass.  Output: gpurun_out/r02_launches_{bench,tile224}.csv
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-tile --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "bench list rc=$?"
timeout 600 ncu --metrics $M --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_tile224.csv \
  python bench.py --workload tile_3660 --stride 224 --steps 1 --warmup 1 > gpurun_out/ncu_tile.log 2>&1
echo "tile list rc=$?"
tail -n 2 gpurun_out/ncu_bench.log; tail -n 2 gpurun_out/ncu_tile.log
