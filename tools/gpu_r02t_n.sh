#!/bin/bash
# Round 2, call T (N = $1 GPUs): the C4 bench line exactly as the driver launches it.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/r02t_bench_C4_n$N.json 2> gpurun_out/r02t_bench_C4_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02t_bench_C4_n$N.json').read().strip().splitlines()[-1])
    print({k:(round(d[k]['value'],1) if isinstance(d.get(k),dict) else d.get(k)) for k in ('value','e2e','e2e_u8','gather','gather_u8')}, d['e2e'].get('d2h_GB/s'), d['gather'].get('GB/s_into_rank0'), d['roofline']['frac'], d['clocks'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02t_bench_C4_n$N.err').read()[-1500:])
PY
