#!/bin/bash
# Round evidence for one workload (default C2): the bench line, the ncu launch list of the same command, and ONE
# `--set full` capture that holds every kernel of one frame (22 launches after warm-up; the report must stay well below the 64 MiB gpurun_out limit).
#   TAG=r02d WL=C2 bash tools/gpu_prof_all.sh   ->   gpurun_out/${TAG}_*
mkdir -p gpurun_out
WL=${WL:-C2}; TAG=${TAG:-prof}
timeout 900 python bench.py --steps 200 --warmup 5 --workload $WL > gpurun_out/${TAG}_bench_$WL.json 2> gpurun_out/${TAG}_bench_$WL.err
tail -c 600 gpurun_out/${TAG}_bench_$WL.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 200 --csv --log-file gpurun_out/${TAG}_launches_$WL.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/ncu_list.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -s 88 -c 22 -f -o gpurun_out/${TAG}_frame_$WL python bench.py --steps 6 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/ncu_frame.log 2>&1
tail -2 gpurun_out/ncu_frame.log | cut -c1-200
ls -la gpurun_out | grep ${TAG}
