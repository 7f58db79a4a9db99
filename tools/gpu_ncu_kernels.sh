#!/bin/bash
# full ncu capture of each kernel regex in $KERNELS (one launch each, after warm-up)
mkdir -p gpurun_out
for k in $KERNELS; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-8} -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload ${WL:-C2} > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log | cut -c1-200
done
