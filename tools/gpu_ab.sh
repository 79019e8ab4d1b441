#!/bin/bash
# A/B of launch mechanics on the default bench (short): graph+PDL (default), graph without PDL, eager with PDL, eager plain.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "default:" "nopdl:IG_NO_PDL=1" "nograph:IG_NO_GRAPH=1" "plain:IG_NO_GRAPH=1 IG_NO_PDL=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-tile --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$name.json').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'graph', d['forward_path'], 'parity', d['parity']['ok'], d['parity']['max_abs'], d['parity']['class_hist'])
except Exception as e:
    print('ERR', e)
PY
)"
  tail -2 gpurun_out/ab_$name.err
done
