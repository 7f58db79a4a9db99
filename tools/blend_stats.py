#!/usr/bin/env python
"""blend_stats.py [workload] — the blend's work counters for one frame (GSR_FLAG_BLEND_COUNT: the counting instantiation
of the default kernel; include/gsrast_b200.h gsr_stage_times.blend_counters), for the one-pixel and the two-pixel kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gsrast_b200 import _lib, camera, scene
from gsrast_b200 import rasterizer as Rz

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
sc, cfg = scene.make_config_scene(wl)
W, H = cfg["W"], cfg["H"]
cam = camera.default_camera(W, H)
tiles = ((W + 15) // 16) * ((H + 15) // 16)
for name, fl in (("two pixels / thread (default)", 0), ("one pixel / thread", _lib.FLAG_BLEND_ONE_PIXEL)):
    g = Rz.GSGaussians(W, H, device="cuda", use_rects=False, flags=fl | _lib.FLAG_BLEND_COUNT)
    g.configure_from_splat_data(sc)
    g.draw(cam, timings=True)
    R, tm = g.draw(cam, timings=True)
    c = tm["blend_counters"]
    print("%s  %s: R=%d  blend %.3f ms (counting build)" % (wl, name, R, tm["blend_ms"]))
    print("   rounds staged %.2f/tile of %.2f available; splats staged %d (%.1f %% of R)" %
          (c["tile_rounds"] / tiles, R / 128 / tiles, c["splats_staged"], 100.0 * c["splats_staged"] / R))
    print("   warp trips (32-pixel units) %d = %.2f per staged splat; evaluated pairs %d; live %.1f %%, passed %.1f %%, blended %.1f %%" %
          (c["warp_trips"], c["warp_trips"] / max(c["splats_staged"], 1), 32 * c["warp_trips"],
           100.0 * c["pairs_live"] / max(32 * c["warp_trips"], 1), 100.0 * c["pairs_passed"] / max(32 * c["warp_trips"], 1),
           100.0 * c["pairs_blended"] / max(32 * c["warp_trips"], 1)))
    del g
    torch.cuda.empty_cache()
