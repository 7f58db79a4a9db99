#!/bin/bash
# Round 2, call C: full GPU suite, A/B of preprocess / blend-occupancy variants (+ carve-out env variants), then a
# launch list and --set full captures of the new blend kernel and of preprocess.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -30 > gpurun_out/r02c_pytest_gpu.txt
tail -6 gpurun_out/r02c_pytest_gpu.txt
NOTEST=1 ROUNDS=2 STEPS=200 ENVS="GSR_CARVEOUT_BLEND=30 GSR_CARVEOUT_BLEND=60" bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/r02c_ab_C2.txt
rm -f gsrast_b200/variants/lib_env_*.so
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 200 --csv --log-file gpurun_out/r02c_launches_C2.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
for k in blend_pair_kernel preprocess_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/r02c_prof_$k python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log | cut -c1-200
done
ls -la gpurun_out | tail -8
