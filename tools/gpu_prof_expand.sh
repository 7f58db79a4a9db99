#!/bin/bash
# launch list of the C2 bench + full captures of the bin-expansion kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 200 --csv --log-file gpurun_out/launches_C2.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --workload C2 > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/launches_C2.csv
for k in ${KERNELS:-expand_fill_kernel expand_count_kernel}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-4} -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload C2 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log | cut -c1-100
done
