"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes ->
libgsrast_b200.so), against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): radii, per-Gaussian tile counts, sorted key/value arrays and
tile ranges BIT-EXACT; images max-abs <= 1/255 per channel and PSNR >= 50 dB."""
import numpy as np
import pytest

from gsrast_b200 import camera as Cm
from gsrast_b200 import scene as S

from helpers import assert_parity, psnr, run_cuda, run_oracle
from test_oracle_kat import tiny_scene

pytestmark = pytest.mark.gpu


def _scene(name, P):
    return S.make_config_scene(name, P=P)[0]


@pytest.mark.parametrize("use_rects", [False, True])
def test_c1_contract_parity(oracle, use_rects):
    """BASELINE config C1 at its full size: 100k Gaussians SH3, 1280x720."""
    sc, cfg = S.make_config_scene("C1")
    cam = Cm.default_camera(cfg["W"], cfg["H"])
    cu = run_cuda(sc, cam, use_rects=use_rects)
    ref = run_oracle(oracle, sc, cam, use_rects=use_rects)
    assert_parity(cu, ref)
    vis = ref.radii > 0
    assert np.array_equal(cu["cov3D"][vis].view(np.uint32), ref.cov3D[vis].view(np.uint32))
    assert np.array_equal(cu["clamped"][vis], ref.clamped[vis])
    if use_rects:
        assert np.array_equal(cu["rects"][vis], ref.rects[vis])


@pytest.mark.parametrize("use_rects", [False, True])
def test_c1_gsrast_compat_parity(oracle, use_rects):
    """Same scene through gscuda::forward semantics (vec4 strides, NDC cull/depth, DC colour,
    T<0.001, y extent without sqrt) — the viewer as it is wired today passes rects."""
    sc, cfg = S.make_config_scene("C1")
    cam = Cm.default_camera(cfg["W"], cfg["H"])
    cu = run_cuda(sc, cam, compat=True, use_rects=use_rects)
    ref = run_oracle(oracle, sc, cam, compat=True, use_rects=use_rects)
    vis = ref.radii > 0
    # in-tree colours are 0.5 + 0.4*dc, un-clamped (GSCuda.cu:364-365): with dc ~ N(0,1) they reach ~2.3, so
    # a single 1/255 alpha-threshold flip moves a pixel by up to |colour|/255
    assert_parity(cu, ref, colour_max=float(np.abs(ref.rgb[vis]).max()))
    assert np.array_equal(cu["cov3D"][vis].view(np.uint32), ref.cov3D[vis].view(np.uint32))


@pytest.mark.parametrize("deg", [0, 1, 2])
def test_sh_degrees(oracle, deg):
    sc = _scene("C1", 30_000)
    cam = Cm.default_camera(640, 360)
    cu = run_cuda(sc, cam, D=deg)
    ref = run_oracle(oracle, sc, cam, D=deg)
    assert_parity(cu, ref)
    vis = ref.radii > 0
    assert np.array_equal(cu["clamped"][vis], ref.clamped[vis])


def test_c5_precomputed_colours_low_opacity(oracle):
    """C5 scaled down: precomputed colours (SH off), low opacity, heavy tile overlap."""
    sc = _scene("C5", 60_000)
    cam = Cm.default_camera(960, 540)
    cu = run_cuda(sc, cam, background=(0.1, 0.3, 0.5))
    ref = run_oracle(oracle, sc, cam, background=(0.1, 0.3, 0.5))
    assert_parity(cu, ref, colours_from_sh=False)


def test_blend_kernels_agree(oracle):
    """The culled default blend and the plain kernel implement the same semantics."""
    from gsrast_b200.rasterizer import FLAG_BLEND_SIMPLE

    sc = _scene("C5", 60_000)
    cam = Cm.default_camera(960, 540)
    a = run_cuda(sc, cam)
    b = run_cuda(sc, cam, flags=FLAG_BLEND_SIMPLE)
    ref = run_oracle(oracle, sc, cam)
    assert_parity(b, ref, colours_from_sh=False)
    assert np.abs(a["out_color"] - b["out_color"]).max() <= 1.0 / 255.0
    assert psnr(a["out_color"], b["out_color"]) >= 50.0


@pytest.mark.parametrize("cfg,P,W,H,compat", [("C2", 150_000, 1000, 555, False), ("C5", 60_000, 960, 540, False),
                                               ("C1", 60_000, 650, 366, True)])
def test_one_pixel_and_pair_blend_kernels_agree(oracle, cfg, P, W, H, compat):
    """Default = two pixels per thread on the packed FP32 pipe (blend_pair_kernel); GSR_FLAG_BLEND_ONE_PIXEL = the
    one-pixel culled kernel.  Per pixel the arithmetic is the same operation for operation, so the frames, final_T
    and n_contrib must be IDENTICAL; both are held to the oracle (ragged sizes: half tile rows / columns)."""
    from gsrast_b200.rasterizer import FLAG_BLEND_ONE_PIXEL

    sc = _scene(cfg, P)
    cam = Cm.orbit_cameras(5, W, H)[2]
    a = run_cuda(sc, cam, compat=compat, background=(0.2, 0.1, 0.4))
    b = run_cuda(sc, cam, compat=compat, background=(0.2, 0.1, 0.4), flags=FLAG_BLEND_ONE_PIXEL)
    assert a["num_rendered"] == b["num_rendered"] > 0
    assert np.array_equal(a["out_color"], b["out_color"])
    assert np.array_equal(a["final_T"], b["final_T"]) and np.array_equal(a["n_contrib"], b["n_contrib"])
    ref = run_oracle(oracle, sc, cam, compat=compat, background=(0.2, 0.1, 0.4))
    cmax = float(np.abs(ref.rgb[ref.radii > 0]).max()) if compat else 1.0
    assert_parity(a, ref, colours_from_sh=sc.colors_precomp is None, colour_max=cmax)


def test_ragged_resolution_and_orbit_camera(oracle):
    """Width/height not multiples of 16 (1080p-like half tile row) and an off-axis camera."""
    sc = _scene("C2", 80_000)
    cam = Cm.orbit_cameras(7, 1000, 555)[3]
    cu = run_cuda(sc, cam, use_rects=True, radii_external=True)
    ref = run_oracle(oracle, sc, cam, use_rects=True)
    assert_parity(cu, ref)


def test_cov3d_precomp_and_bbox(oracle):
    sc = _scene("C1", 20_000)
    cam = Cm.default_camera(640, 360)
    base = run_oracle(oracle, sc, cam)
    cov = base.cov3D.copy()
    cov[base.radii == 0] = np.array([1e-4, 0, 0, 1e-4, 0, 1e-4], np.float32)
    bmin, bmax = [-2.0, -1.0, -3.0], [1.5, 2.0, 2.5]
    cu = run_cuda(sc, cam, cov3D_precomp=cov, boxmin=bmin, boxmax=bmax)
    ref = run_oracle(oracle, sc, cam, cov3D_precomp=cov, boxmin=bmin, boxmax=bmax)
    assert_parity(cu, ref)
    inside = np.all((sc.means3D >= np.array(bmin)) & (sc.means3D <= np.array(bmax)), axis=1)
    assert np.all(cu["radii"][~inside] == 0)


def test_empty_frame_and_empty_scene(oracle):
    bg = (0.2, 0.4, 0.6)
    sc = tiny_scene([[0, 0, -10.0], [0, 0, -9.0]])  # behind the camera
    cam = Cm.default_camera(320, 240)
    cu = run_cuda(sc, cam, background=bg)
    assert cu["num_rendered"] == 0
    for c in range(3):
        assert np.all(cu["out_color"][c] == np.float32(bg[c]))
    assert np.all(cu["final_T"] == 1.0) and np.all(cu["n_contrib"] == 0)
    # in-tree semantics: early return leaves the image untouched (GSCuda.cu:775-778)
    cc = run_cuda(sc, cam, background=bg, compat=True)
    assert cc["num_rendered"] == 0 and np.all(cc["out_color"] == -7.0)
    # P == 0
    empty = tiny_scene(np.zeros((0, 3)))
    ce = run_cuda(empty, cam, background=bg)
    assert ce["num_rendered"] == 0 and np.all(ce["out_color"][2] == np.float32(0.6))


def test_single_pair_range_quirk(oracle):
    """R == 1: the contract closes the tile range, the in-tree kernel does not (GSCuda.cu:533-536)."""
    sc = tiny_scene([[0.14, 0, 0]], scales=[[0.001] * 3])  # radius-3 splat strictly inside one tile
    cam = Cm.default_camera(320, 240)
    cu = run_cuda(sc, cam)
    ref = run_oracle(oracle, sc, cam)
    assert ref.num_rendered == 1
    assert_parity(cu, ref)
    cc = run_cuda(sc, cam, compat=True)
    rc = run_oracle(oracle, sc, cam, compat=True, out_color_init=np.full((3, 240, 320), -7.0, np.float32))
    assert rc.num_rendered == 1 and np.array_equal(cc["ranges"], rc.ranges) and cc["ranges"].max() == 0


def test_kat_blend_on_gpu(oracle):
    """KAT-8 cases on the CUDA path: alpha cap, 1/255 skip, termination, background."""
    cam = Cm.default_camera(320, 240)
    n = 4
    sc = tiny_scene([[0, 0, 0.001 * i] for i in range(n)], scales=[[0.3] * 3] * n, opac=[1.0] * n,
                    colors=[[1, 0.5, 0]] * n)
    for compat in (False, True):
        cu = run_cuda(sc, cam, background=(0.25, 0.5, 0.75), compat=compat)
        ref = run_oracle(oracle, sc, cam, background=(0.25, 0.5, 0.75), compat=compat)
        assert_parity(cu, ref, colours_from_sh=False, n_contrib_budget=0.0)
    low = tiny_scene([[0, 0, 0]], scales=[[0.3] * 3], opac=[0.003], colors=[[1, 1, 1]])
    cu = run_cuda(low, cam, background=(0.1, 0.2, 0.3))
    assert cu["num_rendered"] > 0 and cu["n_contrib"].max() == 0 and np.allclose(cu["out_color"][1], 0.2)


def test_allocator_protocol(oracle):
    """KAT-10: each allocator is called exactly once per forward, geometry -> image -> binning,
    with the sizes required<T>() reports; the buffers are grow-only across calls."""
    from gsrast_b200 import rasterizer as R

    sc = _scene("C1", 20_000)
    cam = Cm.default_camera(640, 360)
    cu = run_cuda(sc, cam)
    geom, binning, img = cu["allocs"]
    assert (geom.calls, binning.calls, img.calls) == (1, 1, 1)
    assert geom.requests[0] == R.GeometryState.required(sc.P)
    assert img.requests[0] == R.ImageState.required(640, 360)
    assert binning.requests[0] == R.BinningState.required(cu["num_rendered"])
    g = R.GSGaussians(640, 360, use_rects=False)
    g.configure_from_splat_data(sc)
    r1 = g.draw(cam)
    p1 = (g._geom.ptr, g._binning.ptr, g._img.ptr)
    r2 = g.draw(cam)
    assert r1 == r2 == cu["num_rendered"]
    assert p1 == (g._geom.ptr, g._binning.ptr, g._img.ptr)  # no re-allocation on a steady camera
    assert np.abs(g.out_color.cpu().numpy() - cu["out_color"]).max() == 0.0  # deterministic


def test_sort_pairs_standalone(oracle):
    """The in-house radix sort against the oracle's stable sort: ragged sizes around the tile
    size, duplicate-heavy keys, every pass count the rasterizer can ask for."""
    import torch

    from gsrast_b200 import rasterizer as R

    rng = np.random.default_rng(11)
    for n, bits in [(1, 44), (31, 45), (4095, 45), (4096, 45), (4097, 45), (100_003, 47), (1_000_000, 45),
                    (300_000, 40), (50_000, 33), (70_000, 64), (2_000_000, 44)]:
        keys = rng.integers(0, 1 << min(bits, 63), n, dtype=np.uint64)
        if bits == 64:
            keys |= rng.integers(0, 2, n, dtype=np.uint64) << np.uint64(63)
        keys[rng.integers(0, n, n // 2)] = keys[0]               # long runs of equal keys
        keys[:: 7] &= np.uint64(0xFFFFFFFF00000000)                # skewed low digits
        vals = rng.permutation(n).astype(np.uint32)
        ko, vo = R.sort_pairs(torch.from_numpy(keys.view(np.int64)).cuda(), torch.from_numpy(vals.view(np.int32)).cuda(),
                              bits)
        rk, rv = oracle.sort_pairs(keys, vals, bits)
        assert np.array_equal(ko.cpu().numpy().view(np.uint64), rk), (n, bits)
        assert np.array_equal(vo.cpu().numpy().view(np.uint32), rv), (n, bits)


def test_identify_ranges_standalone(oracle):
    import torch

    from gsrast_b200 import rasterizer as R

    rng = np.random.default_rng(12)
    tiles = np.sort(rng.integers(0, 500, 20_000).astype(np.uint64))
    keys = (tiles << np.uint64(32)) | rng.integers(0, 1 << 32, tiles.size, dtype=np.uint64)
    for compat in (False, True):
        got = R.identify_tile_ranges(torch.from_numpy(keys.view(np.int64)).cuda(), 512, compat=compat)
        assert np.array_equal(got.cpu().numpy().view(np.uint32), oracle.identify_tile_ranges(keys, 512, compat=compat))


def test_higher_msb_matches(oracle):
    from gsrast_b200 import rasterizer as R

    for n in list(range(1, 70)) + [3072, 3600, 8160, 32400, 65535, 65536, 2 ** 20 + 1]:
        assert R.get_higher_msb(n) == oracle.get_higher_msb(n)


def test_lean_state_flag_changes_no_output_of_the_pass(oracle):
    """GSR_FLAG_LEAN_STATE (what gsr_renderer_* uses for its private scratch): every output of the forward pass is
    identical; only state nothing in the pass reads back is not materialised — the geometry fields cov3D, clamped,
    tiles_touched, point_offsets and, with the bin expansion (which yields the tile ranges from its scan), the sorted
    64-bit keys.  150 k Gaussians = 147 duplication blocks, so a fused duplication's look-back crosses several 32-block
    windows; both binning modes."""
    from gsrast_b200 import _lib

    sc = S.make_config_scene("C2", P=150_000)[0]
    cam = Cm.default_camera(800, 448)
    for extra in (0, _lib.FLAG_RADIX_BINNING):
        full = run_cuda(sc, cam, flags=extra)
        # (lean calls only fill `radii` when the caller passes a buffer for it: internal_radii has no reader)
        lean = run_cuda(sc, cam, flags=extra | _lib.FLAG_LEAN_STATE, radii_external=True)
        assert full["num_rendered"] == lean["num_rendered"] > 0
        for k in ("radii", "depths", "means2D", "conic_opacity", "rgb", "keys", "values", "ranges", "n_contrib",
                  "final_T", "out_color"):
            if k == "keys" and extra == 0:
                continue  # bin expansion, lean: point_list_keys is not written (values and ranges are, compared here)
            a, b = full[k], lean[k]
            if k == "depths":  # entries of Gaussians that emit nothing hold the all-ones (NaN) pattern: compare bits
                a, b = a.view(np.uint32), b.view(np.uint32)
            assert np.array_equal(a, b), (extra, k)
        vis = full["radii"] > 0
        assert np.abs(full["cov3D"].reshape(-1, 6)[vis]).max() > 0  # the default call does fill it


def test_pair_count_overflow_is_detected_on_the_device():
    """140 000 Gaussians that each cover the whole 4K grid: 4.5e9 pairs.  The 32-bit sum the reference would use
    (GSCuda.cu:771) wraps to ~2.4e8 — a plausible num_rendered for which the binning chunk would be sized and then
    overrun.  The scan kernel's 64-bit total must turn this into GSR_ERR_TOO_MANY_PAIRS before the binning allocator
    is ever called."""
    import torch

    from gsrast_b200 import _lib
    from gsrast_b200 import rasterizer as R

    P, W, H = 140_000, 3840, 2160
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(7)
    means = (torch.rand((P, 3), generator=g) - 0.5) * 0.2
    t = lambda a: a.to(dev).contiguous()  # noqa: E731
    scales = torch.full((P, 3), 50.0)
    rot = torch.tensor([1.0, 0.0, 0.0, 0.0]).repeat(P, 1)
    cam = Cm.default_camera(W, H)
    geom, binning, img = R.resize_functional(dev), R.resize_functional(dev), R.resize_functional(dev)
    out = torch.zeros((3, H, W), device=dev)
    keep = [t(means), t(torch.zeros((P, 16, 3))), t(torch.full((P,), 0.5)), t(scales), t(rot),
            t(torch.from_numpy(cam.viewmatrix)), t(torch.from_numpy(cam.projmatrix)), t(torch.from_numpy(cam.cam_pos)),
            torch.zeros(3, device=dev)]
    with pytest.raises(RuntimeError) as ei:
        R.Rasterizer.forward(geom, binning, img, P, 3, 16, keep[8], W, H, keep[0], keep[1], None, keep[2], keep[3], 1.0,
                             keep[4], None, keep[5], keep[6], keep[7], cam.tan_fovx, cam.tan_fovy, False, out)
    torch.cuda.synchronize()
    assert str(_lib.ERR_TOO_MANY_PAIRS) in str(ei.value)
    assert binning.calls == 0 and geom.calls == 1 and img.calls == 1
    # the library is usable afterwards
    sc = _scene("C1", 20_000)
    cam2 = Cm.default_camera(640, 360)
    assert run_cuda(sc, cam2)["num_rendered"] > 0


def test_oversized_grid_is_rejected_before_any_allocation():
    """More than 2^23 tiles (or more than 65535 tile columns / rows) would overflow the packed rects and the 32-bit
    block sums: GSR_ERR_INVALID_ARG before a single allocator call."""
    import torch

    from gsrast_b200 import _lib
    from gsrast_b200 import rasterizer as R

    dev = torch.device("cuda")
    dummy = torch.zeros(64, device=dev)
    geom, binning, img = R.resize_functional(dev), R.resize_functional(dev), R.resize_functional(dev)
    for W, H in ((50_000, 50_000), (16 * 70_000, 16)):
        with pytest.raises(RuntimeError) as ei:
            R.Rasterizer.forward(geom, binning, img, 1, 3, 16, dummy, W, H, dummy, dummy, None, dummy, dummy, 1.0, dummy,
                                 None, dummy, dummy, dummy, 1.0, 1.0, False, dummy)
        assert str(_lib.ERR_INVALID_ARG) in str(ei.value)
    assert geom.calls == 0 and img.calls == 0 and binning.calls == 0
