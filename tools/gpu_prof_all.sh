#!/bin/bash
# Round evidence: bench line, ncu launch list of the same command, one full capture per hot kernel.
mkdir -p gpurun_out
WL=${WL:-C2}
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --workload $WL > gpurun_out/bench_$WL.log 2>&1
tail -1 gpurun_out/bench_$WL.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 200 --csv --log-file gpurun_out/launches_$WL.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/ncu_list.log 2>&1
for k in preprocess_kernel duplicate_sorted_kernel gather_rects_kernel identify_ranges_kernel blend_culled_kernel histogram_kernel scan_block_sums_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/ncu_$k.log 2>&1
done
# six consecutive onesweep launches of one frame: 4 depth-digit passes over P + 2 tile-digit passes over R
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 48 -c 6 -f -o gpurun_out/prof_onesweep python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/ncu_onesweep.log 2>&1
ls gpurun_out | wc -l
