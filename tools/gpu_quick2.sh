#!/bin/bash
# quick loop: sort/parity spot tests + bench of chosen workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x 2>&1 | tail -6
for w in ${WLS:-C2}; do
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --workload $w > gpurun_out/bench_$w.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$w.log').read().strip().splitlines()[-1])
    s=d['stages']
    print('$w', 'fps %.1f e2e %.1f R=%d'%(d['value'],d['e2e']['value'],d['config']['num_rendered']), {k:(round(v['ms'],3) if isinstance(v,dict) else round(v,3)) for k,v in s.items()}, 'hist',round(s['sort']['hist_ms'],3),'passes',[round(x,3) for x in s['sort']['pass_ms']])
except Exception as e:
    print('$w failed', e); print(open('gpurun_out/bench_$w.log').read()[-1500:])
PY
done
