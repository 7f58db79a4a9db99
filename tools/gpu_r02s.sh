#!/bin/bash
# Round 2, call S: lean calls without a radii buffer skip internal_radii + uniform-digit top depth pass (b_new = HEAD)
# vs neither (a_base), then the evidence set of HEAD (tools/gpu_evidence.sh, TAG=r02s).
mkdir -p gpurun_out
NOTEST=1 ROUNDS=2 STEPS=200 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02s_ab_C2.txt
TAG=r02s bash tools/gpu_evidence.sh
