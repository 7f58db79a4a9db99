#!/bin/bash
# A/B: bench C2 stage times for each library variant in gsrast_b200/variants (+ parity spot check)
mkdir -p gpurun_out
for lib in gsrast_b200/variants/lib_*.so; do
  n=$(basename $lib .so)
  GSRAST_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --workload ${WL:-C2} > gpurun_out/bench_$n.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$n.log').read().strip().splitlines()[-1])
    s=d['stages']
    print('$n', 'fps %.1f e2e %.1f'%(d['value'],d['e2e']['value']), {k:(round(v['ms'],3) if isinstance(v,dict) else round(v,3)) for k,v in s.items()}, 'hist',round(s['sort']['hist_ms'],3),'passes',[round(x,3) for x in s['sort']['pass_ms']])
except Exception as e:
    print('$n failed', e); print(open('gpurun_out/bench_$n.log').read()[-800:])
PY
done
if [ -n "$TESTLIB" ]; then
GSRAST_B200_LIB=$PWD/gsrast_b200/variants/lib_$TESTLIB.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "sort_pairs or c1_contract" 2>&1 | tail -4
fi
