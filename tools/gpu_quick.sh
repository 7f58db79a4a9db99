#!/bin/bash
# quick loop: small GPU tests + bench (no CPU baseline)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -25 > gpurun_out/pytest_small.log
tail -8 gpurun_out/pytest_small.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 900 -x 2>&1 | tail -25 > gpurun_out/pytest_full.log
tail -5 gpurun_out/pytest_full.log
for w in C2 C5 C3; do
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
