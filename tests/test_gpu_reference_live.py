"""A/B against the reference's own code on the same GPU: oracle/_ref is the in-tree
apps/gsrast/gscuda/GSCuda.cu compiled UNMODIFIED (against a GLM stand-in, -fmad=false) by
`make -C oracle ref`.  Our GSRast-compat path must reproduce its radii, tile counts, sorted
keys/values and tile ranges bit for bit, and its image within tolerance."""
import numpy as np
import pytest

from gsrast_b200 import camera as Cm
from gsrast_b200 import scene as S

from helpers import psnr, run_cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,P,W,H,use_rects", [("C1", 100_000, 1280, 720, True), ("C2", 200_000, 1000, 555, False),
                                                 ("C5", 60_000, 960, 540, True)])
def test_compat_path_matches_reference_binary(cfg, P, W, H, use_rects):
    from oracle import gscuda_ref

    if not gscuda_ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    sc = S.make_config_scene(cfg, P=P)[0]
    cam = Cm.default_camera(W, H)
    ref = gscuda_ref.RefRenderer(sc, W, H, use_rects=use_rects)
    ref.draw(cam)
    st = ref.state()
    cu = run_cuda(sc, cam, compat=True, use_rects=use_rects)
    assert cu["num_rendered"] == st["num_rendered"]
    assert np.array_equal(cu["radii"], st["radii"])
    assert np.array_equal(cu["tiles_touched"], st["tiles_touched"])
    assert np.array_equal(cu["point_offsets"], st["point_offsets"])
    vis = st["radii"] > 0
    for k in ("depths", "means2D", "conic_opacity", "cov3D"):
        assert np.array_equal(cu[k][vis].view(np.uint32), st[k][vis].view(np.uint32)), k
    if sc.colors_precomp is None:
        assert np.array_equal(cu["rgb"][vis].view(np.uint32), st["rgb"][vis].view(np.uint32))
    if use_rects:
        assert np.array_equal(cu["rects"][vis], st["rects"][vis])
    assert np.array_equal(cu["keys"], st["keys"]) and np.array_equal(cu["values"], st["values"])
    assert np.array_equal(cu["ranges"], st["ranges"])
    cmax = float(np.abs(st["rgb"][vis]).max()) if sc.colors_precomp is None else 1.0
    err = np.abs(cu["out_color"] - st["out_color"])
    assert err.max() <= max(1.0, cmax) / 255.0 and float(np.mean(err > 1 / 255.0)) <= 1e-5
    assert psnr(cu["out_color"], st["out_color"]) >= 50.0
    assert float(np.mean(cu["n_contrib"] != st["n_contrib"])) <= 2e-4
