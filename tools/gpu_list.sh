#!/bin/bash
# launch list only (ncu gpu__time_duration per launch) for one workload
mkdir -p gpurun_out
WL=${WL:-C2}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-80} -c ${COUNT:-200} --csv --log-file gpurun_out/launches_$WL.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-300
