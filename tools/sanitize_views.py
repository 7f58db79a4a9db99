#!/usr/bin/env python
"""Small frames through the resident-scene renderer (gsr_renderer_render / _render_host: the lean state, i.e. the
id-only bin expansion and the double-buffered blend that bench.py times) for compute-sanitizer runs; checked against
the single-call path (full state).  tools/gpu_evidence.sh runs it under memcheck / racecheck / synccheck / initcheck."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from gsrast_b200 import camera, scene  # noqa: E402
from gsrast_b200.views import ViewRenderer  # noqa: E402
from helpers import run_cuda  # noqa: E402

sc = scene.make_config_scene("C2", P=30_000)[0]
W, H = 640, 360
cams = camera.orbit_cameras(3, W, H)
vr = ViewRenderer.from_scene(sc, W, H)
out, nr = vr.render(cams, cams[0].tan_fovx, cams[0].tan_fovy)
torch.cuda.synchronize()
host, nr_h = vr.render_host(cams, cams[0].tan_fovx, cams[0].tan_fovy)
assert nr == nr_h and np.array_equal(out.cpu().numpy(), host.numpy())
for v, cam in enumerate(cams):
    single = run_cuda(sc, cam)
    assert single["num_rendered"] == nr[v] and np.array_equal(out[v].cpu().numpy(), single["out_color"]), v
vr.close()
print("[views] ok: %d views, num_rendered %s" % (len(cams), list(nr)))
