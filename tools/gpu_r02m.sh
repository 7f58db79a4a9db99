#!/bin/bash
# Round 2, call M: lean callers skip the sorted 64-bit keys (b_lk) vs writing them (a_base); double-buffered blend
# staging with one barrier per round (c_dbuf at 8 CTAs/SM, d_dbuf7 at 7).  Parity tests on the default library and on
# the double-buffered variant first.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -x 2>&1 | tail -6 | tee gpurun_out/r02m_pytest.txt
GSRAST_B200_LIB=$PWD/gsrast_b200/variants/lib_c_dbuf.so timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_views.py -m gpu -q --timeout 1200 2>&1 | tail -6 | tee gpurun_out/r02m_pytest_dbuf.txt
NOTEST=1 ROUNDS=2 STEPS=200 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02m_ab_C2.txt
NOTEST=1 ROUNDS=1 STEPS=100 WL=C5 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02m_ab_C5.txt
NOTEST=1 ROUNDS=1 STEPS=60 WL=C3 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02m_ab_C3.txt
