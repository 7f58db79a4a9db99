#!/bin/bash
# Round 2, call T (8 GPUs): the 2-GPU bit-equality test, then the C4 bench line at N = 8 exactly as the driver launches it
# (value, e2e, e2e_u8, overlapped NCCL gather fp32 / u8).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_views.py -m gpu -q --timeout 120 -k "two_gpus" 2>&1 | tail -3 | tee gpurun_out/r02t_pytest_2gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r02t_bench_C4_n8.json 2> gpurun_out/r02t_bench_C4_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02t_bench_C4_n8.json').read().strip().splitlines()[-1])
    print({k:(round(d[k]['value'],1) if isinstance(d.get(k),dict) else d.get(k)) for k in ('value','e2e','e2e_u8','gather','gather_u8')}, d['e2e'].get('d2h_GB/s'), d['gather'].get('GB/s_into_rank0'), d['roofline']['frac'], d['clocks'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r02t_bench_C4_n8.err').read()[-1500:])
PY
