#!/bin/bash
# Round 2, call I: balanced expand_fill (base) vs the thread-per-(tile, quarter) walk (fb0): expansion / full-size
# parity tests, then A/B on C2, C3, C5, C1.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_bin_expand.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q --timeout 1200 2>&1 | tail -6 | tee gpurun_out/r02i_pytest.txt
NOTEST=1 ROUNDS=2 STEPS=200 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02i_ab_C2.txt
NOTEST=1 ROUNDS=1 STEPS=60 WL=C3 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02i_ab_C3.txt
NOTEST=1 ROUNDS=1 STEPS=100 WL=C5 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02i_ab_C5.txt
NOTEST=1 ROUNDS=1 STEPS=200 WL=C1 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02i_ab_C1.txt
