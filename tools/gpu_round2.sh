#!/bin/bash
mkdir -p gpurun_out
./tools/microbench > gpurun_out/microbench.log 2>&1; cat gpurun_out/microbench.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_fullsize.py 2>&1 | tail -40 > gpurun_out/pytest_small.log
tail -15 gpurun_out/pytest_small.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-1500
# launch list of the same command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
# full captures of the top kernels (one launch each)
for k in preprocess_kernel onesweep_kernel histogram_kernel blend_culled_kernel duplicate_kernel identify_ranges_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
