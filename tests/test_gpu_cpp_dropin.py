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


def test_cpp_views_host_ply_to_u8_frames(tmp_path):
    """tests/cpp/views_host.cpp: .ply -> gsr_ply_load -> gsr_renderer_create -> gsr_renderer_render_host_u8, all from
    C++ through the C ABI; frames and num_rendered must equal the ctypes path's on the same cameras."""
    import torch

    from gsrast_b200.views import ViewRenderer, pack_cameras

    exe = os.path.join(ROOT, "tests", "cpp", "views_host")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "views_host"])
    sc = S.make_config_scene("C1", P=30_000)[0]
    W, H = 640, 360
    cams = Cm.orbit_cameras(3, W, H)
    ply, cam_bin, out_bin = str(tmp_path / "data.ply"), str(tmp_path / "cams.bin"), str(tmp_path / "out.bin")
    S.write_ply(ply, sc)
    with open(cam_bin, "wb") as f:
        np.array([len(cams), W, H], np.int32).tofile(f)
        np.array([cams[0].tan_fovx, cams[0].tan_fovy], np.float32).tofile(f)
        pack_cameras(cams).astype(np.float32).tofile(f)
    out = subprocess.check_output([exe, ply, cam_bin, out_bin], timeout=300).decode()
    assert "views_host: P=30000 views=3" in out
    head = np.fromfile(out_bin, np.int32, 2 + len(cams))
    frames = np.fromfile(out_bin, np.uint8, offset=4 * (2 + len(cams))).reshape(len(cams), 3, H, W)
    # the same scene as the loader stages it (activation round trip through the file), same cameras, ctypes path
    staged = S.load_ply_native(ply)
    vr = ViewRenderer.from_scene(staged, W, H)
    want, nr = vr.render_host_u8(cams, cams[0].tan_fovx, cams[0].tan_fovy)
    torch.cuda.synchronize()
    assert list(head[2:]) == nr and head[0] == sc.P
    assert np.array_equal(frames, want.numpy())
    assert frames.max() > 0
    vr.close()


def test_cpp_interop_stream_ordering_and_inspector_state(tmp_path):
    """tests/cpp/interop_host.cpp: a caller-owned NON-default stream and output buffer, a racing memset queued before
    the render and a D2H copy queued right after it with no synchronisation in between (INTEGRATION.md option C); and
    the Inspector's fields read through gsr_renderer_map_geometry_state on a GSR_FLAG_KEEP_STATE renderer."""
    from gsrast_b200.views import pack_cameras

    exe = os.path.join(ROOT, "tests", "cpp", "interop_host")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "interop_host"])
    sc = S.make_config_scene("C1", P=60_000)[0]
    W, H = 960, 540
    cams = Cm.orbit_cameras(5, W, H)
    ply, cam_bin = str(tmp_path / "data.ply"), str(tmp_path / "cams.bin")
    S.write_ply(ply, sc)
    with open(cam_bin, "wb") as f:
        np.array([len(cams), W, H], np.int32).tofile(f)
        np.array([cams[0].tan_fovx, cams[0].tan_fovy], np.float32).tofile(f)
        pack_cameras(cams).astype(np.float32).tofile(f)
    res = subprocess.run([exe, ply, cam_bin], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "bad_rounds=0 state_ok=1" in res.stdout
