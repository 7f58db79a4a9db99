#!/bin/bash
# What the driver runs at round end + sanitizers on the smoke frame: GPU tests, smoke, both bench arms.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke\]|Error|hazard" gpurun_out/sanitizer_$tool.log | cut -c1-200 | head -6
done
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-300
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-200
