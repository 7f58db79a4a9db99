#!/bin/bash
# Round 2, call U: environment sweeps on the HEAD library — L2 fetch granularity hint (GSR_L2_FETCH=32/64/128) and the
# blend's shared-memory carve-out with the double-buffered staging (50 default, 56, 64, 75).
mkdir -p gpurun_out
NOTEST=1 ROUNDS=2 STEPS=200 ENVS="GSR_L2_FETCH=32 GSR_L2_FETCH=64 GSR_L2_FETCH=128 GSR_CARVEOUT_BLEND=56 GSR_CARVEOUT_BLEND=64 GSR_CARVEOUT_BLEND=75" bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02u_ab_C2.txt
grep -h "L2 fetch" gpurun_out/bench_lib_env_GSR_L2_FETCH*_1.log | sort | uniq -c | tee -a gpurun_out/r02u_ab_C2.txt
NOTEST=1 ROUNDS=1 STEPS=100 WL=C5 ENVS="GSR_L2_FETCH=32 GSR_L2_FETCH=64 GSR_CARVEOUT_BLEND=64" bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02u_ab_C5.txt
