#!/bin/bash
# Round-1 evidence (run as r01j, r01l): GPU tests, smoke, contract bench (CPU + reference-GPU legs), reference arm, C1/C3/C5 lines,
# launch list of the bench command, full captures of the hot kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-300
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-300
for WL in C1 C3 C5; do
  timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --workload $WL > gpurun_out/bench_$WL.log 2>&1
  tail -1 gpurun_out/bench_$WL.log | cut -c1-160
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 200 --csv --log-file gpurun_out/launches_C2.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --workload C2 > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/launches_C2.csv
for k in ${KERNELS:-preprocess_kernel blend_culled_kernel expand_fill_kernel expand_count_kernel duplicate_sorted_kernel gather_rects_kernel}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload C2 > gpurun_out/ncu_$k.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 20 -c 5 -f -o gpurun_out/prof_onesweep python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload C2 > gpurun_out/ncu_onesweep.log 2>&1
ls gpurun_out | wc -l
