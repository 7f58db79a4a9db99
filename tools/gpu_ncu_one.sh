#!/bin/bash
# ncu full capture of one kernel:  KERNEL=regex [LIB=variant] [SKIP=n] bash tools/gpu_ncu_one.sh
mkdir -p gpurun_out
[ -n "$LIB" ] && export GSRAST_B200_LIB=$PWD/gsrast_b200/variants/lib_$LIB.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s ${SKIP:-8} -c 1 -f -o gpurun_out/prof_${KERNEL}_${LIB:-main} python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload ${WL:-C2} > gpurun_out/ncu_$KERNEL.log 2>&1
tail -3 gpurun_out/ncu_$KERNEL.log
