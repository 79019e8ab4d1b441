#!/bin/bash
# One GPU-box pass: parity tests, the §8(f) kernel probe, smoke, the default bench line.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_probe.py aux > gpurun_out/probe_aux.log 2>&1; tail -24 gpurun_out/probe_aux.log
if [ -z "$SKIP_BENCH" ]; then
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_chips_v1.json 2> gpurun_out/bench_chips_v1.err; tail -c 400 gpurun_out/bench_chips_v1.err; cut -c1-300 gpurun_out/bench_chips_v1.json
fi
