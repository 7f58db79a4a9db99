#!/bin/bash
# First GPU pass: parity tests, smoke, sanitizer on the smoke, short bench, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -40 > gpurun_out/pytest_small.log
cat gpurun_out/pytest_small.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; tail -15 gpurun_out/sanitizer.log
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 900 2>&1 | tail -30 > gpurun_out/pytest_full.log
cat gpurun_out/pytest_full.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; tail -3 gpurun_out/bench.log
