#!/bin/bash
# Round 2, call B: GPU tests with the two-pixel blend kernel as the default, then A/B of the blend variants
# (C2, and C5 = the blend-bound configuration).
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -x 2>&1 | tail -25 > gpurun_out/r02b_pytest_gpu.txt
tail -8 gpurun_out/r02b_pytest_gpu.txt
NOTEST=1 ROUNDS=2 STEPS=200 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02b_ab_C2.txt
NOTEST=1 ROUNDS=1 STEPS=100 WL=C5 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02b_ab_C5.txt
NOTEST=1 ROUNDS=1 STEPS=60 WL=C3 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02b_ab_C3.txt
