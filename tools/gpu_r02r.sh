#!/bin/bash
# Round 2, call R: expand scan holding its chunk counts in registers (b_regs, default) vs two passes over memory
# (a_base); uniform-digit fast path of the most significant depth pass + vote-aggregated top histogram digit (c_uni).
# Every test run is bounded (a wrong sort can turn a blend loop into minutes): 300 s per pytest call, 60 s per test.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 90 -x 2>&1 | tail -5 | tee gpurun_out/r02r_pytest.txt
GSRAST_B200_LIB=$PWD/gsrast_b200/variants/lib_c_uni.so timeout 300 python -m pytest tests/test_gpu_split_sort.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --timeout 60 -x 2>&1 | tail -5 | tee gpurun_out/r02r_pytest_uni.txt
NOTEST=1 ROUNDS=2 STEPS=200 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02r_ab_C2.txt
NOTEST=1 ROUNDS=1 STEPS=60 WL=C3 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02r_ab_C3.txt
NOTEST=1 ROUNDS=1 STEPS=200 WL=C1 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02r_ab_C1.txt
