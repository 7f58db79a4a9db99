#!/bin/bash
# Round 2, call N: blend occupancy with the lean expansion in place — 8 (a_base), 7, 6 CTAs/SM, the double-buffered
# staging at 7 and 6; 4 chunks per count CTA.
mkdir -p gpurun_out
NOTEST=1 ROUNDS=2 STEPS=200 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02n_ab_C2.txt
NOTEST=1 ROUNDS=1 STEPS=100 WL=C5 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02n_ab_C5.txt
NOTEST=1 ROUNDS=1 STEPS=60 WL=C3 bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02n_ab_C3.txt
