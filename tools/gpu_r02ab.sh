#!/bin/bash
# Round 2, call AB: duplication blocks add up the pair counts before their own (b_self, default) vs the single-CTA
# scan between gather_rects and the duplication (a_base).  Whole GPU suite on the default library first (bounded).
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 120 -x 2>&1 | tail -4 | tee gpurun_out/r02ab_pytest.txt
NOTEST=1 ROUNDS=2 STEPS=200 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02ab_ab_C2.txt
NOTEST=1 ROUNDS=1 STEPS=60 WL=C3 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02ab_ab_C3.txt
NOTEST=1 ROUNDS=1 STEPS=200 WL=C1 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02ab_ab_C1.txt
