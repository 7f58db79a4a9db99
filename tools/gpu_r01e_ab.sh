#!/bin/bash
# r01e A/B: GPU tests on the default library, then C2 bench of every variant in gsrast_b200/variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
bash tools/gpu_variants.sh
