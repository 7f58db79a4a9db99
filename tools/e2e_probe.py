#!/usr/bin/env python
"""e2e_probe.py: wall-clock time of one public call for n views — resident (device frames), fp32 to pinned host, 8-bit
to pinned host — to separate the per-call costs of the e2e leg (pipeline fill, drain of the last copy, wall clock vs
CUDA events) from its per-frame cost.  profiles/r02w_e2e_probe.txt"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gsrast_b200 import camera, scene  # noqa: E402
from gsrast_b200.views import ViewRenderer, pack_cameras  # noqa: E402

sc = scene.make_config_scene("C2")[0]
W, H = 1920, 1080
cam = camera.default_camera(W, H)
vr = ViewRenderer.from_scene(sc, W, H)
packed = np.repeat(pack_cameras([cam]), 100, axis=0)
out_dev = torch.empty((80, 3, H, W), device="cuda")
host = torch.empty((80, 3, H, W), dtype=torch.float32).pin_memory()
host8 = torch.empty((80, 3, H, W), dtype=torch.uint8).pin_memory()
tx, ty = cam.tan_fovx, cam.tan_fovy


def wall(fn, reps=6):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[len(ts) // 2], min(ts)


for n in (1, 2, 5, 10, 20, 40, 80):
    blk = packed[:n]
    wall(lambda: vr.render(blk, tx, ty, out=out_dev[:n]), 2)
    r = wall(lambda: vr.render(blk, tx, ty, out=out_dev[:n]))
    wall(lambda: vr.render_host(blk, tx, ty, out_host=host[:n]), 2)
    h = wall(lambda: vr.render_host(blk, tx, ty, out_host=host[:n]))
    wall(lambda: vr.render_host_u8(blk, tx, ty, out_host=host8[:n]), 2)
    u = wall(lambda: vr.render_host_u8(blk, tx, ty, out_host=host8[:n]))
    print("n=%3d  resident %7.3f ms (min %7.3f)  host fp32 %7.3f (min %7.3f)  host u8 %7.3f (min %7.3f)   per view: %.3f / %.3f / %.3f"
          % (n, r[0], r[1], h[0], h[1], u[0], u[1], r[0] / n, h[0] / n, u[0] / n))

# --- first-call effects: a fresh pinned landing buffer, a 5-view warm-up call, then 20-view calls back to back
import gc

for trial in range(2):
    fresh = torch.empty((20, 3, H, W), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    blk5, blk20 = packed[:5], packed[:20]
    vr.render_host(blk5, tx, ty, out_host=fresh[:5])
    torch.cuda.synchronize()
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        vr.render_host(blk20, tx, ty, out_host=fresh)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print("fresh pinned buffer, warm-up of 5 views, then 20-view calls: " + " ".join("%.3f" % t for t in ts) + " ms")
    del fresh
    gc.collect()
