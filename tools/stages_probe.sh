#!/bin/bash
cd "$(dirname "$0")/.."
for s in 2 3 4 5; do echo "== stages $s"; IG_GEMM_STAGES=$s python tools/gpu_probe.py perf 2>&1 | grep "v1_b64"; done
