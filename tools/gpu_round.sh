#!/bin/bash
# One GPU call: the -m gpu suite file by file (a CUDA fault in one file does not poison the others), smoke, the
# default bench line.  Output under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/pytest_gpu.log
for f in tests/test_gpu_*.py; do
  echo "##### $f" >> gpurun_out/pytest_gpu.log
  timeout 900 python -m pytest "$f" -x -q -m gpu -s 2>&1 | tail -40 >> gpurun_out/pytest_gpu.log
done
grep -E "#####|passed|failed|error|Error|assert" gpurun_out/pytest_gpu.log | head -80
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
