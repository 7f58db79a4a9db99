"""The C++ boundary: a GL-free copy of GSRast's host side (tests/cpp/dropin_host.cpp — the
resizeFunctional allocators and the FORWARD call of apps/gsrast/GSGaussians.cpp) built against
the header-only shims include/rasterizer.h and include/gscuda_dropin.h must produce the same
frame as the ctypes path."""
import os
import subprocess

import numpy as np
import pytest

from gsrast_b200 import camera as Cm
from gsrast_b200 import scene as S

from helpers import run_cuda

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "dropin_host")


@pytest.mark.parametrize("mode", ["contract", "gscuda"])
def test_cpp_host_matches_ctypes(tmp_path, mode):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.dirname(EXE)])
    sc = S.make_config_scene("C1", P=40_000)[0]
    W, H = 800, 450
    cam = Cm.default_camera(W, H)
    compat = mode == "gscuda"
    if compat:
        means, scales, rot, opac, shs = sc.gsrast_layout()
    else:
        means, scales, rot, opac, shs = sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs
    path_in, path_out = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    with open(path_in, "wb") as f:
        np.array([sc.P, W, H, 3, 16], np.int32).tofile(f)
        np.array([cam.tan_fovx, cam.tan_fovy], np.float32).tofile(f)
        np.zeros(3, np.float32).tofile(f)
        cam.viewmatrix.astype(np.float32).tofile(f)
        cam.projmatrix.astype(np.float32).tofile(f)
        cam.cam_pos.astype(np.float32).tofile(f)
        for a in (means, scales, rot, opac, shs):
            np.ascontiguousarray(a, np.float32).tofile(f)
    out = subprocess.check_output([EXE, path_in, path_out, mode], timeout=300).decode()
    assert "dropin_host" in out
    head = np.fromfile(path_out, np.int32, 5)
    img = np.fromfile(path_out, np.float32, offset=20).reshape(3, H, W)
    cu = run_cuda(sc, cam, compat=compat, use_rects=True)
    assert head[0] == cu["num_rendered"]
    assert list(head[1:4]) == [2, 2, 2]          # one call per allocator per frame, two frames
    assert head[4] == cu["radii"][0]             # Inspector-style field access through fromChunk
    assert np.array_equal(img, cu["out_color"])
