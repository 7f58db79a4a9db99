"""The GSRast-mode oracle against fixtures produced by the REFERENCE'S OWN CODE (see
tests/golden/README.md): bit-exact integers and float scratch, image within tolerance."""
import os

import numpy as np
import pytest

from gsrast_b200 import camera as Cm
from gsrast_b200 import scene as S

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = {"a": "C1", "b": "C2", "c": "C5"}


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_gsrast_mode_oracle_matches_reference_fixture(oracle, name):
    fx = np.load(os.path.join(HERE, "golden", "gsrast_ref_%s.npz" % name))
    P, W, H, cam_idx, use_rects, R = [int(v) for v in fx["meta"]]
    sc, _ = S.make_config_scene(CFG[name], P=P)
    cam = Cm.default_camera(W, H) if cam_idx < 0 else Cm.orbit_cameras(7, W, H)[cam_idx]
    r = oracle.forward_scene(sc, cam, background=(0.1, 0.2, 0.3), use_rects=bool(use_rects), mode=oracle.MODE_GSRAST)
    assert r.num_rendered == R
    assert np.array_equal(r.radii, fx["radii"])
    assert np.array_equal(r.tiles_touched, fx["tiles_touched"])
    assert np.array_equal(r.point_offsets, fx["point_offsets"])
    vis = fx["vis_idx"]
    assert np.array_equal(np.nonzero(r.radii > 0)[0], vis)
    for k in ("depths", "means2D", "conic_opacity", "cov3D"):
        assert np.array_equal(r[k][vis].view(np.uint32), fx[k].view(np.uint32)), k
    if sc.colors_precomp is None:
        assert np.array_equal(r.rgb[vis].view(np.uint32), fx["rgb"].view(np.uint32))
    if use_rects:
        assert np.array_equal(r.rects[vis], fx["rects"])
    assert np.array_equal(r["keys"], fx["keys"]) and np.array_equal(r["values"], fx["values"])
    assert np.array_equal(r.ranges, fx["ranges"])
    # image: the reference blends with CUDA's expf, the oracle with libm's; fixture stored as float16
    cmax = float(np.abs(fx["rgb"]).max()) if sc.colors_precomp is None else 1.0
    ref_img = fx["out_color"].astype(np.float32)
    err = np.abs(r.out_color - ref_img)
    assert err.max() <= max(1.0, cmax) / 255.0 + np.abs(ref_img).max() * 2.0 ** -10
    assert float(np.mean(r.n_contrib != fx["n_contrib"])) <= 2e-4
    assert np.abs(r.final_T - fx["final_T"].astype(np.float32)).max() <= 1 / 255.0 + 2.0 ** -10


def test_contract_mode_regression_hashes(oracle):
    """KAT-11: seeded C1-scale frame through the contract-mode oracle; integer outputs hashed.
    (Self-generated regression pins — they guard the oracle against accidental edits; they are not
    evidence of agreement with upstream, which is absent from the reference tree.)"""
    import hashlib
    import json

    sc, _ = S.make_config_scene("C1", P=20_000)
    cam = Cm.default_camera(640, 360)
    r = oracle.forward_scene(sc, cam, use_rects=True)
    got = {k: hashlib.sha256(np.ascontiguousarray(r[k]).tobytes()).hexdigest()[:16]
           for k in ("radii", "tiles_touched", "keys", "values", "ranges", "rects")}
    got["num_rendered"] = r.num_rendered
    path = os.path.join(HERE, "golden", "contract_c1_20k_hashes.json")
    if not os.path.exists(path):
        pytest.fail("missing %s; expected content: %s" % (path, json.dumps(got)))
    assert got == json.load(open(path))
