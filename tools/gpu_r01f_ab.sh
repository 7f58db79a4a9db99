#!/bin/bash
# r01f: parity of the "all" variant (fill fast path + double-buffered blend + fast cull bounds), compute-sanitizer
# memcheck / racecheck of one small frame, then alternating C2 bench rounds of every variant.
mkdir -p gpurun_out
export GSRAST_B200_LIB=$PWD/gsrast_b200/variants/lib_${TESTLIB:-all}.so
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu_${TESTLIB:-all}.log
tail -6 gpurun_out/pytest_gpu_${TESTLIB:-all}.log
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke\]|Error|hazard" gpurun_out/sanitizer_$tool.log | head -8
done
unset GSRAST_B200_LIB
NOTEST=1 ROUNDS=${ROUNDS:-2} STEPS=${STEPS:-300} bash tools/gpu_ab.sh
