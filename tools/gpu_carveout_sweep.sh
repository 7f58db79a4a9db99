#!/bin/bash
# Sweep of per-kernel shared-memory carve-out preferences (GSR_CARVEOUT_<NAME> env overrides), C2, pipelined fps.
# CONFIGS: ';'-separated lists of NAME=pct pairs, e.g. "BLEND=25;BLEND=25 PRE=72"
mkdir -p gpurun_out
IFS=';' read -ra CFGS <<< "${CONFIGS:-none}"
for round in $(seq 1 ${ROUNDS:-1}); do
for cfg in "${CFGS[@]}"; do
  envs=""
  for kv in $cfg; do [ "$kv" != "none" ] && envs="$envs GSR_CARVEOUT_$kv"; done
  tag=$(echo "$cfg" | tr ' =' '__')
  env $envs timeout 600 python bench.py --steps ${STEPS:-300} --warmup 10 --no-cpu-baseline --workload ${WL:-C2} > gpurun_out/bench_co_${tag}_$round.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_co_${tag}_$round.log').read().strip().splitlines()[-1])
    s=d['stages']
    print('%-40s fps %.1f e2e %.1f pre %.3f sort %.3f blend %.3f serial %.3f' % ('$cfg', d['value'], d['e2e']['value'], s['preprocess']['ms'], s['sort']['ms'], s['blend']['ms'], s['frame_serial_ms']))
except Exception as e:
    print('$cfg failed', e); print(open('gpurun_out/bench_co_${tag}_$round.log').read()[-600:])
PY
done
done
