#!/bin/bash
# parity spot check + bench + launch list + full ncu captures of the hot kernels (one frame's worth of sort passes)
mkdir -p gpurun_out
WL=${WL:-C2}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --workload $WL > gpurun_out/bench_$WL.log 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$WL.log').read().strip().splitlines()[-1])
s=d['stages']
print('$WL', 'fps %.1f e2e %.1f R=%d'%(d['value'],d['e2e']['value'],d['config']['num_rendered']), {k:(round(v['ms'],3) if isinstance(v,dict) else round(v,3)) for k,v in s.items()}, 'hist',round(s['sort']['hist_ms'],3),'passes',[round(x,3) for x in s['sort']['pass_ms']])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 160 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/ncu_list.log 2>&1
for k in ${KERNELS:-blend_culled_kernel duplicate_sorted_kernel preprocess_kernel}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/ncu_$k.log 2>&1
done
# six consecutive onesweep launches = 4 depth passes + 2 tile passes of one frame
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 48 -c 6 -f -o gpurun_out/prof_onesweep python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/ncu_onesweep.log 2>&1
ls gpurun_out | head -40
