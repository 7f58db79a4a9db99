"""Generates tests/golden/gsrast_ref_*.npz by running the REFERENCE ITSELF — the in-tree
apps/gsrast/gscuda/GSCuda.cu compiled unmodified into oracle/_ref (make -C oracle ref) — on a
GPU.  Run on the GPU box:   python tests/golden/make_gsrast_fixtures.py gpurun_out/golden
then copy the .npz files into tests/golden/.  The CPU test tests/test_golden.py checks the
GSRast-mode oracle against them, which is what pins that oracle to the reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gsrast_b200 import camera, scene  # noqa: E402
from oracle import gscuda_ref  # noqa: E402

CASES = {
    # name: (config, P, W, H, camera index (None = default pose), use_rects)
    "a": ("C1", 4000, 320, 240, None, True),
    "b": ("C2", 6000, 400, 225, 3, False),
    "c": ("C5", 2500, 256, 192, None, True),
}


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, (cfgname, P, W, H, cam_idx, use_rects) in CASES.items():
        sc, _ = scene.make_config_scene(cfgname, P=P)
        cam = camera.default_camera(W, H) if cam_idx is None else camera.orbit_cameras(7, W, H)[cam_idx]
        r = gscuda_ref.RefRenderer(sc, W, H, use_rects=use_rects)
        r.draw(cam, background=(0.1, 0.2, 0.3))
        st = r.state()
        vis = st["radii"] > 0
        np.savez_compressed(
            os.path.join(outdir, "gsrast_ref_%s.npz" % name),
            radii=st["radii"], tiles_touched=st["tiles_touched"], point_offsets=st["point_offsets"],
            vis_idx=np.nonzero(vis)[0].astype(np.int32), depths=st["depths"][vis], means2D=st["means2D"][vis],
            conic_opacity=st["conic_opacity"][vis], cov3D=st["cov3D"][vis], rgb=st["rgb"][vis],
            rects=(st["rects"][vis] if use_rects else np.zeros((0, 2), np.int32)),
            keys=st["keys"], values=st["values"], ranges=st["ranges"], n_contrib=st["n_contrib"],
            final_T=st["final_T"].astype(np.float16), out_color=st["out_color"].astype(np.float16),
            meta=np.array([P, W, H, -1 if cam_idx is None else cam_idx, int(use_rects), st["num_rendered"]], np.int64))
        print(name, "P", P, "R", st["num_rendered"], "visible", int(vis.sum()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
