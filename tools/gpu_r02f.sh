#!/bin/bash
# Round 2, call F: GPU suite with the fused sort tail (rect gather in the last depth pass + look-back duplication),
# then A/B against the split form (-DGSR_FUSED_SORT=0) on C2 / C1 / C3.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -30 > gpurun_out/r02f_pytest_gpu.txt
tail -6 gpurun_out/r02f_pytest_gpu.txt
NOTEST=1 ROUNDS=2 STEPS=200 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02f_ab_C2.txt
NOTEST=1 ROUNDS=1 STEPS=200 WL=C1 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02f_ab_C1.txt
NOTEST=1 ROUNDS=1 STEPS=60 WL=C3 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02f_ab_C3.txt
