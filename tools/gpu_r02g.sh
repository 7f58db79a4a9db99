#!/bin/bash
# Round 2, call G: blend trip variants (3-compare tail = base; unroll 4; predicated scalar colour FFMAs) on C2 / C5,
# after the blend-kernel parity tests.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_views.py -m gpu -q --timeout 900 2>&1 | tail -4 | tee gpurun_out/r02g_pytest.txt
NOTEST=1 ROUNDS=2 STEPS=200 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02g_ab_C2.txt
NOTEST=1 ROUNDS=1 STEPS=100 WL=C5 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02g_ab_C5.txt
