N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR bench.py --gpus $N > gpurun_out/bench_chips_v1_${N}gpu.json 2> gpurun_out/bench_chips_v1_${N}gpu.err
cut -c1-200 gpurun_out/bench_chips_v1_${N}gpu.json
