#!/bin/bash
# Multi-GPU bench lines (N = $1): config #5 chip set (12 500 chips/GPU), the default chips workload, the tile.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/smi_${N}gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --workload chipset_100k > gpurun_out/bench_chipset_${N}gpu.json 2> gpurun_out/bench_chipset_${N}gpu.err
tail -c 300 gpurun_out/bench_chipset_${N}gpu.err; cut -c1-400 gpurun_out/bench_chipset_${N}gpu.json
timeout 600 $TR bench.py --gpus $N > gpurun_out/bench_chips_v1_${N}gpu.json 2> gpurun_out/bench_chips_v1_${N}gpu.err
tail -c 300 gpurun_out/bench_chips_v1_${N}gpu.err; cut -c1-300 gpurun_out/bench_chips_v1_${N}gpu.json
if [ -z "$SKIP_TILE" ]; then
timeout 600 $TR bench.py --gpus $N --workload tile_3660 --stride 112 --steps 3 > gpurun_out/bench_tile_s112_${N}gpu.json 2> gpurun_out/bench_tile_s112_${N}gpu.err
tail -c 300 gpurun_out/bench_tile_s112_${N}gpu.err; cut -c1-300 gpurun_out/bench_tile_s112_${N}gpu.json
fi
