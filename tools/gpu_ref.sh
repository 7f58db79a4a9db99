#!/bin/bash
mkdir -p gpurun_out/golden
timeout 600 python tests/golden/make_gsrast_fixtures.py gpurun_out/golden 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_reference_live.py -m gpu -q --timeout 600 2>&1 | tail -25
