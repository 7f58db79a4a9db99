#!/bin/bash
# Round 2, call A: the whole GPU test suite (new full-size parity tests), packed-FP32 microbench, the default bench
# line, then an A/B of the library variants in gsrast_b200/variants (2 alternating rounds).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_gpu.txt 2>&1
timeout 120 ./tools/microbench_f32x2 > gpurun_out/r02a_microbench_f32x2.txt 2>&1; cat gpurun_out/r02a_microbench_f32x2.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 --durations=12 -s 2>&1 | tail -60 > gpurun_out/r02a_pytest_gpu.txt
tail -25 gpurun_out/r02a_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r02a_bench_C2.json 2> gpurun_out/r02a_bench_C2.err; tail -c 3000 gpurun_out/r02a_bench_C2.json; tail -3 gpurun_out/r02a_bench_C2.err
NOTEST=1 ROUNDS=2 STEPS=200 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02a_ab.txt
