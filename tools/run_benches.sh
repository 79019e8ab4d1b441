#!/bin/bash
# All single-GPU bench lines of a round -> gpurun_out/bench_*.json (copy the ones to keep into profiles/)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_chips_v1.json 2> gpurun_out/bench_chips_v1.err; tail -c 600 gpurun_out/bench_chips_v1.err
timeout 900 python bench.py --workload chips_v2_300m_t3 --steps 10 --no-cpu-baseline > gpurun_out/bench_chips_v2_300m.json 2> gpurun_out/bench_v2.err; tail -c 600 gpurun_out/bench_v2.err
timeout 600 python bench.py --workload tile_3660 --stride 224 --steps 5 > gpurun_out/bench_tile_s224.json 2> gpurun_out/bench_tile.err; tail -c 600 gpurun_out/bench_tile.err
timeout 600 python bench.py --workload tile_3660 --stride 112 --steps 3 > gpurun_out/bench_tile_s112.json 2>> gpurun_out/bench_tile.err
for f in gpurun_out/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "clocks", d["clocks"])
    print({k:round(v["ms_per_step"],3) for k,v in d["kernel_families"].items()})
except Exception as e: print("ERR",e)
PY
done
