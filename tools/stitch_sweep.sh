#!/bin/bash
# tuning sweep of the stitch kernel's rows-per-item / rows-per-batch / classes-per-pass
mkdir -p gpurun_out
python -m pytest tests/test_gpu_stitch.py tests/test_gpu_tile.py -x -q -m gpu 2>&1 | tail -3
for rpi in 4 8 16; do for r in 1 2 4; do for ch in 2 4 7; do
  echo "--- RPI=$rpi R=$r CH=$ch"
  IG_STITCH_RPI=$rpi IG_STITCH_R=$r IG_STITCH_CH=$ch python tools/gpu_probe.py stitch 2>&1 | grep "perf"
done; done; done
