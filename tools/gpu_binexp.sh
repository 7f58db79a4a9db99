#!/bin/bash
# first contact of the bin expansion: parity tests, then C2 bench in both binning modes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bin_expand.py tests/test_gpu_split_sort.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -x 2>&1 | tail -15
for mode in "" "--radix-binning"; do
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --workload C2 $mode > gpurun_out/bench_C2$mode.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_C2$mode.log').read().strip().splitlines()[-1])
    s=d['stages']
    print('C2 $mode', 'fps %.1f e2e %.1f R=%d Rc=%d'%(d['value'],d['e2e']['value'],d['config']['num_rendered'],d['config']['num_coarse']), {k:(round(v['ms'],3) if isinstance(v,dict) else round(v,3)) for k,v in s.items()}, 'hist',round(s['sort']['hist_ms'],3),'passes',[round(x,3) for x in s['sort']['pass_ms']], 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3))
except Exception as e:
    print('C2 $mode failed', e); print(open('gpurun_out/bench_C2$mode.log').read()[-1500:])
PY
done
