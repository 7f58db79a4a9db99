#!/bin/bash
# Generic environment-variable sweep of the C2 bench.  CONFIGS: ';'-separated lists of VAR=value pairs ("none" = defaults).
mkdir -p gpurun_out
IFS=';' read -ra CFGS <<< "${CONFIGS:-none}"
for round in $(seq 1 ${ROUNDS:-1}); do
for cfg in "${CFGS[@]}"; do
  envs="GSR_SWEEP=1"
  for kv in $cfg; do [ "$kv" != "none" ] && envs="$envs $kv"; done
  tag=$(echo "$cfg" | tr ' =/' '___')
  env $envs timeout 600 python bench.py --steps ${STEPS:-300} --warmup 10 --no-cpu-baseline --workload ${WL:-C2} > gpurun_out/bench_env_${tag}_$round.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_env_${tag}_$round.log').read().strip().splitlines()[-1])
    s=d['stages']
    print('%-44s fps %.1f e2e %.1f pre %.3f sort %.3f blend %.3f serial %.3f' % ('$cfg', d['value'], d['e2e']['value'], s['preprocess']['ms'], s['sort']['ms'], s['blend']['ms'], s['frame_serial_ms']))
except Exception as e:
    print('$cfg failed', e); print(open('gpurun_out/bench_env_${tag}_$round.log').read()[-600:])
PY
done
done
