"""View-batch renderer (gsr_renderer_*, the C form of GSGaussians) and the viewer-buffer repack."""
import numpy as np
import pytest

from gsrast_b200 import camera as Cm
from gsrast_b200 import scene as S

from helpers import run_cuda

pytestmark = pytest.mark.gpu


def test_view_batch_matches_single_forward():
    import torch

    from gsrast_b200.views import ViewRenderer

    sc = S.make_config_scene("C2", P=60_000)[0]
    W, H = 800, 448
    cams = Cm.orbit_cameras(5, W, H)
    vr = ViewRenderer.from_scene(sc, W, H)
    out, nr = vr.render(cams, cams[0].tan_fovx, cams[0].tan_fovy)
    torch.cuda.synchronize()
    host, nr_h = vr.render_host(cams, cams[0].tan_fovx, cams[0].tan_fovy)
    assert nr == nr_h
    for v, cam in enumerate(cams):
        single = run_cuda(sc, cam)
        assert single["num_rendered"] == nr[v]
        assert np.array_equal(out[v].cpu().numpy(), single["out_color"])
        assert np.array_equal(host[v].numpy(), single["out_color"])
    out2, nr2, times = vr.render(cams[:2], cams[0].tan_fovx, cams[0].tan_fovy, timings=True)
    assert times["num_rendered"] == nr2[1] and times["sort_passes"] >= 5 and times["depth_passes"] == 4 and times["kernel_launches"] >= 10
    assert times["total_ms"] > 0
    vr.close()


def test_view_batch_compat_mode():
    import torch

    from gsrast_b200.views import ViewRenderer

    sc = S.make_config_scene("C1", P=30_000)[0]
    W, H = 640, 360
    cam = Cm.default_camera(W, H)
    means4, scales4, rot, opac, shs_raw = sc.gsrast_layout()
    t = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    vr = ViewRenderer(P=sc.P, D=3, M=16, means3D=t(means4), shs=t(shs_raw), colors_precomp=None, opacities=t(opac),
                      scales=t(scales4), rotations=t(rot), background=torch.zeros(3, device="cuda"), width=W, height=H,
                      compat=True)
    out, nr = vr.render([cam], cam.tan_fovx, cam.tan_fovy)
    torch.cuda.synchronize()
    # the renderer mirrors GSGaussians, which always passes its _rects buffer (GSGaussians.cpp:137,204)
    single = run_cuda(sc, cam, compat=True, use_rects=True)
    assert nr[0] == single["num_rendered"] and np.array_equal(out[0].cpu().numpy(), single["out_color"])


def test_repack_gsrast_scene():
    """vec4 / raw-PLY viewer buffers -> contract layout on the device."""
    import torch

    from gsrast_b200 import _lib

    sc = S.make_config_scene("C1", P=10_001)[0]
    means4, scales4, rot, opac, shs_raw = sc.gsrast_layout()
    t = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    m4, s4, raw = t(means4), t(scales4), t(shs_raw)
    m3 = torch.empty((sc.P, 3), device="cuda")
    s3 = torch.empty((sc.P, 3), device="cuda")
    sh = torch.empty((sc.P, 16, 3), device="cuda")
    _lib.check(_lib.lib().gsr_repack_gsrast_scene(sc.P, m4.data_ptr(), s4.data_ptr(), raw.data_ptr(), m3.data_ptr(),
                                                  s3.data_ptr(), sh.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(m3.cpu().numpy(), sc.means3D)
    assert np.array_equal(s3.cpu().numpy(), sc.scales)
    assert np.array_equal(sh.cpu().numpy(), sc.shs)


def test_u8_frame_delivery():
    """gsr_renderer_render_host_u8 / gsr_frames_to_u8: round(clamp(x, 0, 1) * 255) of exactly the float frames, planar
    layout kept, odd element counts (tail kernel) included."""
    import torch

    from gsrast_b200.views import ViewRenderer, frames_to_u8

    sc = S.make_config_scene("C2", P=60_000)[0]
    W, H = 800, 448
    cams = Cm.orbit_cameras(5, W, H)
    vr = ViewRenderer.from_scene(sc, W, H, background=(0.25, 1.5, -0.5))  # background outside [0, 1] exercises the clamp
    host_f, nr_f = vr.render_host(cams, cams[0].tan_fovx, cams[0].tan_fovy)
    host_u, nr_u = vr.render_host_u8(cams, cams[0].tan_fovx, cams[0].tan_fovy)
    assert nr_f == nr_u and host_u.dtype == torch.uint8 and tuple(host_u.shape) == (5, 3, H, W)
    f = host_f.numpy()
    want = np.rint(np.clip(f, 0.0, 1.0) * np.float32(255.0)).astype(np.uint8)  # rint = round-half-even, like cvt.rni
    assert np.array_equal(host_u.numpy(), want)
    assert want.min() == 0 and want.max() == 255
    # the bare conversion, 4k + 3 elements
    x = torch.linspace(-0.5, 1.5, 4 * 1000 + 3, device="cuda")
    got = frames_to_u8(x).cpu().numpy()
    assert np.array_equal(got, np.rint(np.clip(x.cpu().numpy(), 0.0, 1.0) * np.float32(255.0)).astype(np.uint8))
    vr.close()


def test_renderer_keep_state_serves_the_inspector_fields():
    """GSR_FLAG_KEEP_STATE + gsr_renderer_map_geometry_state: the nine fields the reference's Inspector reads
    (Inspector.cpp:174-188 through GSGaussians::mapGeometryState, GSGaussians.cpp:214-219) out of the renderer's
    private scratch, bit-identical with a plain gsr_forward of the same view."""
    import torch

    from gsrast_b200.views import ViewRenderer

    sc = S.make_config_scene("C1", P=50_000)[0]
    W, H = 800, 448
    cams = Cm.orbit_cameras(3, W, H)
    vr = ViewRenderer.from_scene(sc, W, H, keep_state=True)
    out, nr = vr.render(cams, cams[0].tan_fovx, cams[0].tan_fovy)
    torch.cuda.synchronize()
    lanes = vr.num_lanes()
    for lane in range(lanes):
        v = max(i for i in range(len(cams)) if i % lanes == lane)  # the last view that ran on this lane
        g = vr.map_geometry_state(lane)
        single = run_cuda(sc, cams[v])
        assert np.array_equal(out[v].cpu().numpy(), single["out_color"])
        vis = single["radii"] > 0
        assert np.array_equal(g["internal_radii"].cpu().numpy(), single["radii"])
        assert np.array_equal(g["tiles_touched"].cpu().numpy().view(np.uint32), single["tiles_touched"])
        assert np.array_equal(g["point_offsets"].cpu().numpy().view(np.uint32), single["point_offsets"])
        for k in ("depths", "means2D", "conic_opacity", "cov3D", "rgb"):
            assert np.array_equal(g[k].cpu().numpy()[vis].view(np.uint32), single[k][vis].view(np.uint32)), k
        assert np.array_equal(g["clamped"].cpu().numpy()[vis], single["clamped"][vis])
    vr.close()
    # the default (lean) renderer renders the same frames
    vr2 = ViewRenderer.from_scene(sc, W, H)
    out2, nr2 = vr2.render(cams, cams[0].tan_fovx, cams[0].tan_fovy)
    torch.cuda.synchronize()
    assert nr2 == nr and torch.equal(out2, out)
    vr2.close()


def test_same_views_on_two_gpus_are_bit_equal():
    """Multi-GPU correctness of the view sharding: the same four views rendered on cuda:0 and on cuda:1
    (scene replicated, no torch.cuda.set_device by the caller) must be the same bits.  Needs >= 2 visible devices."""
    import torch

    from gsrast_b200.views import ViewRenderer

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    sc = S.make_config_scene("C2", P=400_000)[0]
    W, H = 1280, 720
    cams = Cm.orbit_cameras(4, W, H)
    frames, counts = [], []
    for d in (0, 1):
        vr = ViewRenderer.from_scene(sc, W, H, device="cuda:%d" % d)
        out, nr = vr.render(cams, cams[0].tan_fovx, cams[0].tan_fovy)
        torch.cuda.synchronize(d)
        host, nr_h = vr.render_host(cams, cams[0].tan_fovx, cams[0].tan_fovy)
        assert out.device.index == d and nr == nr_h
        assert np.array_equal(out.cpu().numpy(), host.numpy())
        frames.append(out.cpu().numpy())
        counts.append(nr)
        vr.close()
    assert counts[0] == counts[1]
    assert np.array_equal(frames[0], frames[1])
    # and the single-call path with explicit device tensors, current device left at 0
    a = run_cuda(sc, cams[0], device="cuda:0")
    b = run_cuda(sc, cams[0], device="cuda:1")
    assert a["num_rendered"] == b["num_rendered"] == counts[0][0]
    assert np.array_equal(a["keys"], b["keys"]) and np.array_equal(a["values"], b["values"])
    assert np.array_equal(a["out_color"], b["out_color"]) and np.array_equal(a["out_color"], frames[0][0])
