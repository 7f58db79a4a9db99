#!/usr/bin/env python
"""blend_stats.py [workload] — diagnostics: how many 256-splat batches the blend stages per tile versus how many the
deepest contributing splat of the tile needed.  Needs a library built with -DGSR_BLEND_STATS
(tools/build_variant.sh stats -DGSR_BLEND_STATS; GSRAST_B200_LIB=gsrast_b200/variants/lib_stats.so)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gsrast_b200 import _lib, camera, scene
from gsrast_b200.views import ViewRenderer

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
sc, cfg = scene.make_config_scene(wl)
W, H = cfg["W"], cfg["H"]
vr = ViewRenderer.from_scene(sc, W, H, device=torch.device("cuda", 0))
cam = camera.default_camera(W, H)
packed = np.stack([cam.packed()]).astype(np.float32)
out = torch.empty((1, 3, H, W), dtype=torch.float32, device="cuda")
L = _lib.lib()
st = (C.c_ulonglong * 4)()
vr.render(packed, cam.tan_fovx, cam.tan_fovy, out=out)
L.gsr_debug_blend_stats(st, 1)
_, nr = vr.render(packed, cam.tan_fovx, cam.tan_fovy, out=out)
L.gsr_debug_blend_stats(st, 1)
tiles = ((W + 15) // 16) * ((H + 15) // 16)
R = nr[0]
print("%s: R=%d tiles=%d  rounds available %.2f/tile" % (wl, R, tiles, R / 256 / tiles))
print("tile-rounds staged %d (%.2f/tile)  needed by deepest n_contrib %d (%.2f/tile)" % (st[0], st[0] / tiles, st[3], st[3] / tiles))
print("warp-rounds walking a list %d (%.2f of 8 per staged round)  candidate trips %d (%.3f per pair)" %
      (st[1], st[1] / max(st[0], 1), st[2], st[2] / R))
