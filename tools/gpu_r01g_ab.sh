#!/bin/bash
# r01g: full GPU suite + racecheck on the default library (live-warp mask in the blend cull, 2 chunks per CTA in
# expand_count), then alternating C2 bench rounds of the variants.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|smoke\]|Error" gpurun_out/sanitizer_racecheck.log | cut -c1-300 | head -6
NOTEST=1 ROUNDS=${ROUNDS:-2} STEPS=${STEPS:-300} bash tools/gpu_ab.sh
