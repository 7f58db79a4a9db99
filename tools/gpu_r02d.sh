#!/bin/bash
# Round 2, call D (2 GPUs): the 2-GPU bit-equality test, then both bench arms at N=2 (C4: orbit views sharded,
# overlapped NCCL gather on by default).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_views.py -m gpu -q --timeout 600 2>&1 | tail -5 | tee gpurun_out/r02d_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r02d_bench_C4_n2.json 2> gpurun_out/r02d_bench_C4_n2.err
tail -c 1500 gpurun_out/r02d_bench_C4_n2.json; tail -5 gpurun_out/r02d_bench_C4_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/r02d_bench_C4_n2_reference.json 2> gpurun_out/r02d_ref.err
tail -c 900 gpurun_out/r02d_bench_C4_n2_reference.json; tail -3 gpurun_out/r02d_ref.err
