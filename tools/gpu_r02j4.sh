#!/bin/bash
# Round 2, call J4 (4 GPUs): the C4 bench line at N=8 (value, e2e, e2e_u8, overlapped NCCL gather fp32 / u8), then the same
# with write-combined pinned landing buffers for the e2e legs.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02j_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 40 --warmup 5 > gpurun_out/r02j_bench_C4_n4.json 2> gpurun_out/r02j_bench_C4_n4.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02j_bench_C4_n4.json').read().strip().splitlines()[-1])
    print({k:(round(d[k]['value'],1) if isinstance(d.get(k),dict) else d.get(k)) for k in ('value','e2e','e2e_u8','gather','gather_u8')}, d['e2e'].get('d2h_GB/s'), d['gather'].get('GB/s_into_rank0'), d['roofline']['frac'], d['clocks'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02j_bench_C4_n4.err').read()[-1500:])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 40 --warmup 5 --write-combined --no-gather > gpurun_out/r02j_bench_C4_n4_wc.json 2> gpurun_out/r02j_bench_C4_n4_wc.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02j_bench_C4_n4_wc.json').read().strip().splitlines()[-1])
    print('WC', {k:(round(d[k]['value'],1) if isinstance(d.get(k),dict) else d.get(k)) for k in ('value','e2e','e2e_u8')}, d['e2e'].get('d2h_GB/s'))
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02j_bench_C4_n4_wc.err').read()[-1500:])
PY
