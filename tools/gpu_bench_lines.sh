#!/bin/bash
# Bench lines of HEAD for every single-GPU workload (TAG=... prefix), the driver's own command (--steps 20 --warmup 5)
# and the reference arm.  The library-level evidence (ncu, sanitizers, tests) is tools/gpu_evidence.sh.
mkdir -p gpurun_out
TAG=${TAG:-r02z}
timeout 600 python bench.py --steps 200 --warmup 5 --workload C2 > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench_C2.err; tail -c 300 gpurun_out/${TAG}_bench_C2.json; echo
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_C2_driver_cmd.json 2> gpurun_out/${TAG}_bench_C2_driver_cmd.err
for wl in C1 C3 C5; do
  timeout 600 python bench.py --steps 100 --warmup 5 --workload $wl > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_C2_reference_arm.json 2> gpurun_out/${TAG}_ref.err
python - <<PY
import json
for n in ("C2","C2_driver_cmd","C1","C3","C5"):
    d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%n).read().strip().splitlines()[-1])
    print(n, "value %.1f e2e %.1f u8 %.1f latency %.1f roof %.3f"%(d["value"],d["e2e"]["value"],d["e2e_u8"]["value"],d["latency_fps"],d["roofline"]["frac"]), d["steps"], d["clocks"]["reasons"])
d=json.loads(open("gpurun_out/${TAG}_bench_C2_reference_arm.json").read().strip().splitlines()[-1]); print("reference arm", d["value"], d["cpu_baseline"]["cores"])
PY
