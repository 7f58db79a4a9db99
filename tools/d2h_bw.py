#!/usr/bin/env python
"""d2h_bw.py: what the box gives a 24.9 MB (one fp32 1080p frame) device -> pinned-host copy: one stream, two streams,
torch-pinned vs cudaHostAlloc'd landing memory.  Context for bench.py's e2e leg (profiles/r02w_d2h_bw.txt)."""
import time

import torch

n = 3 * 1080 * 1920
dev = torch.device("cuda:0")
src = [torch.randn(n, device=dev) for _ in range(4)]
host = torch.empty((40, n), dtype=torch.float32).pin_memory()
s = [torch.cuda.Stream() for _ in range(2)]
torch.cuda.synchronize()


def run(nstreams, reps=40):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        with torch.cuda.stream(s[i % nstreams]):
            host[i % 40].copy_(src[i % 4], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n * 4 / dt / 1e9, dt / reps * 1e3


for k in (1, 2):
    run(k, 8)
    gbs, ms = run(k)
    print("D2H fp32 frame, %d stream(s): %.1f GB/s, %.3f ms per frame" % (k, gbs, ms))
# one copy alone (latency of the last frame of a call)
torch.cuda.synchronize()
t0 = time.perf_counter()
host[0].copy_(src[0], non_blocking=True)
torch.cuda.synchronize()
print("single frame copy: %.3f ms" % ((time.perf_counter() - t0) * 1e3))
h8 = torch.empty((40, n), dtype=torch.uint8).pin_memory()
s8 = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(4)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(40):
    h8[i].copy_(s8[i % 4], non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("D2H u8 frame: %.1f GB/s, %.3f ms per frame" % (40 * n / dt / 1e9, dt / 40 * 1e3))
