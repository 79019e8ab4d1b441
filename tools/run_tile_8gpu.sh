N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29515 tools/check_tile_sharded.py 2>&1 | grep "CHECK\|False"
for s in 112 224; do
timeout 600 $TR --master-port 2952$((s/112)) bench.py --gpus $N --workload tile_3660 --stride $s --steps 5 > gpurun_out/bench_tile_s${s}_${N}gpu.json 2> gpurun_out/bench_tile_s${s}_${N}gpu.err
cut -c1-220 gpurun_out/bench_tile_s${s}_${N}gpu.json
done
