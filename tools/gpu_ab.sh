#!/bin/bash
# A/B of library variants (gsrast_b200/variants/lib_*.so) on one box: GPU tests on the default library first,
# then ROUNDS alternating bench runs per variant (pipelined throughput is what counts; +-1 % run-to-run).
mkdir -p gpurun_out
if [ -z "$NOTEST" ]; then
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
fi
# ENVS="NAME=VALUE NAME2=VALUE2 ...": additionally run the default library once per entry with that variable set
for e in $ENVS; do ln -sf $PWD/gsrast_b200/libgsrast_b200.so gsrast_b200/variants/lib_env_${e//[^A-Za-z0-9_]/_}.so; done
for round in $(seq 1 ${ROUNDS:-2}); do
for lib in gsrast_b200/variants/lib_*.so; do
  n=$(basename $lib .so)
  envset=""
  for e in $ENVS; do [ "lib_env_${e//[^A-Za-z0-9_]/_}" = "$n" ] && envset="$e"; done
  env $envset GSRAST_B200_LIB=$PWD/$lib timeout ${BENCH_TIMEOUT:-150} python bench.py --steps ${STEPS:-300} --warmup 10 --no-cpu-baseline --workload ${WL:-C2} > gpurun_out/bench_${n}_$round.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${n}_$round.log').read().strip().splitlines()[-1])
    s=d['stages']
    print('$n', 'fps %.1f e2e %.1f'%(d['value'],d['e2e']['value']), {k:(round(v['ms'],3) if isinstance(v,dict) else round(v,3)) for k,v in s.items() if k not in ('ranges','preprocess_plus_sort')}, 'hist',round(s['sort']['hist_ms'],3),'passes',[round(x,3) for x in s['sort']['pass_ms']])
except Exception as e:
    print('$n failed', e); print(open('gpurun_out/bench_${n}_$round.log').read()[-800:])
PY
done
done
