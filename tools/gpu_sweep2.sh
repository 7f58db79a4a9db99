mkdir -p gpurun_out
run() { # name, env...
  n=$1; shift
  env "$@" timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --workload C2 > gpurun_out/bench_s2_$n.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_s2_$n.log').read().strip().splitlines()[-1]); s=d['stages']
    print('%-24s fps %.1f e2e %.1f pre %.3f sort %.3f blend %.3f serial %.3f' % ('$n', d['value'], d['e2e']['value'], s['preprocess']['ms'], s['sort']['ms'], s['blend']['ms'], s['frame_serial_ms']))
except Exception as e:
    print('$n failed', e); print(open('gpurun_out/bench_s2_$n.log').read()[-500:])
PY
}
run default A=1
run lanes3 GSRAST_B200_LIB=$PWD/gsrast_b200/variants/lib_lanes3.so
run lanes1 GSRAST_B200_LIB=$PWD/gsrast_b200/variants/lib_lanes1.so
run sortpad20 GSR_SORT_PAD_KB=20
run sortpad50 GSR_SORT_PAD_KB=50
run default2 A=1
