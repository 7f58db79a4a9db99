#!/bin/bash
# Round 2, call L: one-lane SH evaluation in preprocess (b_new) vs the 4-lane compacting form (a_old), occupancy 5 / 7
# CTAs per SM, prefetch mode 3 (64-byte pull of the shared line); parallel chunk table, 1024-thread expand scan.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -x 2>&1 | tail -6 | tee gpurun_out/r02l_pytest.txt
NOTEST=1 ROUNDS=2 STEPS=200 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02l_ab_C2.txt
NOTEST=1 ROUNDS=1 STEPS=100 WL=C5 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02l_ab_C5.txt
NOTEST=1 ROUNDS=1 STEPS=60 WL=C3 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02l_ab_C3.txt
NOTEST=1 ROUNDS=1 STEPS=200 WL=C1 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02l_ab_C1.txt
