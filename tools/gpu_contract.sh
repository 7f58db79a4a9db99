#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, both bench arms (N=1), then N=2.. if more GPUs are visible
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-600
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-600
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  [ $NG -gt 8 ] && NG=8
  for n in 2 $NG; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 64 --warmup 4 > gpurun_out/bench_n$n.log 2>&1; tail -1 gpurun_out/bench_n$n.log | cut -c1-400
  done
fi
