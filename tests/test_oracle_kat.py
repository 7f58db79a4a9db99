"""Known-answer tests of the CPU oracle.  The reference ships no
tests or golden vectors, so every expected value here is derived by hand from
the reference source lines cited, or from closed-form math in float64."""
import math

import numpy as np
import pytest

from gsrast_b200 import camera as Cm
from gsrast_b200.scene import SplatScene


def tiny_scene(means, scales=None, rots=None, opac=None, shs=None, colors=None):
    means = np.asarray(means, np.float32).reshape(-1, 3)
    P = means.shape[0]
    scales = np.full((P, 3), 0.05, np.float32) if scales is None else np.asarray(scales, np.float32).reshape(P, 3)
    rots = np.tile(np.array([1, 0, 0, 0], np.float32), (P, 1)) if rots is None else np.asarray(rots, np.float32)
    opac = np.full(P, 0.9, np.float32) if opac is None else np.asarray(opac, np.float32)
    if colors is not None:
        return SplatScene(means, scales, rots, opac, None, np.asarray(colors, np.float32).reshape(P, 3), 0, 0)
    if shs is None:
        shs = np.zeros((P, 16, 3), np.float32)
    return SplatScene(means, scales, rots, opac, np.asarray(shs, np.float32), None, 3, 16)


CAM = Cm.default_camera(320, 240)


# ---- KAT-1 getHigherMsb (GSCuda.cu:481-502) ---------------------------------------------------
@pytest.mark.parametrize("n,expect", [(3072, 12), (3600, 12), (8160, 13), (32400, 15), (1, 1), (2, 2), (255, 8),
                                      (256, 9), (65535, 16), (65536, 17)])
def test_kat1_higher_msb(oracle, n, expect):
    assert oracle.get_higher_msb(n) == expect
    # it is "bits needed to represent n"
    assert oracle.get_higher_msb(n) == max(1, int(n).bit_length())


# ---- KAT-2 getRect (GSCuda.cu:237-259) --------------------------------------------------------
def test_kat2_get_rect_edges(oracle):
    gx, gy = 20, 15
    # splat well inside: p=(100,100), r=10 -> x tiles [int(90/16)=5, int((100+10+15)/16)=7)
    assert oracle.get_rect(100.0, 100.0, 10, 10, gx, gy) == (5, 5, 7, 7)
    # negative numerator truncates toward zero: (3-10)/16 = -0.4375 -> 0 ; clamp keeps 0
    assert oracle.get_rect(3.0, 3.0, 10, 10, gx, gy)[:2] == (0, 0)
    # centre left of the screen but radius reaches in: p.x=-5, r=10 -> max = int((−5+10+15)/16)=1
    assert oracle.get_rect(-5.0, 50.0, 10, 10, gx, gy) == (0, 2, 1, 4)
    # completely off-screen to the left: max clamps to 0 -> zero area
    r = oracle.get_rect(-100.0, 50.0, 10, 10, gx, gy)
    assert (r[2] - r[0]) * (r[3] - r[1]) == 0
    # beyond the right/bottom edge: both clamp to the grid
    assert oracle.get_rect(1000.0, 1000.0, 10, 10, gx, gy) == (20, 15, 20, 15)
    # radius spanning the grid edge
    assert oracle.get_rect(315.0, 235.0, 10, 10, gx, gy) == (19, 14, 20, 15)
    # rect variant with different extents
    assert oracle.get_rect(100.0, 100.0, 20, 4, gx, gy) == (5, 6, 8, 7)
    # NaN centre behaves like CUDA's saturating conversion (NaN -> 0)
    r = oracle.get_rect(float("nan"), 10.0, 3, 3, gx, gy)
    assert r[0] == 0 and r[2] == 0


# ---- KAT-3 isotropic Gaussian on the optical axis ---------------------------------------------
def test_kat3_axis_gaussian_closed_form(oracle):
    s, W, H = 0.05, 320, 240
    sc = tiny_scene([[0, 0, 0]], scales=[[s, s, s]])
    r = oracle.forward_scene(sc, CAM)
    z = 5.0
    fy = H / (2.0 * CAM.tan_fovy)
    fx = W / (2.0 * CAM.tan_fovx)
    assert abs(fx - fy) < 1e-3
    c = (fy * s / z) ** 2 + 0.3
    # conic = (1/c, 0, 1/c)
    co = r.conic_opacity[0]
    assert co[0] == pytest.approx(1.0 / c, rel=1e-5) and co[2] == pytest.approx(1.0 / c, rel=1e-5)
    assert abs(co[1]) < 1e-6 and co[3] == np.float32(0.9)
    # radius: mid=c, det=c^2 -> lambda = c + sqrt(0.1)   (the max(0.1, .) floor is active)
    assert r.radii[0] == math.ceil(3.0 * math.sqrt(c + math.sqrt(0.1)))
    assert r.depths[0] == pytest.approx(z, rel=1e-6)
    # ndc2Pix(0, S) = (S-1)/2
    assert r.means2D[0, 0] == pytest.approx((W - 1) / 2.0, abs=1e-3)
    assert r.means2D[0, 1] == pytest.approx((H - 1) / 2.0, abs=1e-3)
    # cov3D = s^2 I
    assert np.allclose(r.cov3D[0], [s * s, 0, 0, s * s, 0, s * s], atol=1e-9)


def test_kat3_gsrast_mode_pixel_centre_and_depth(oracle):
    sc = tiny_scene([[0, 0, 0]])
    r = oracle.forward_scene(sc, CAM, mode=oracle.MODE_GSRAST)
    # in-tree: pixel = (ndc*0.5+0.5)*S -> exactly S/2 on the axis (GSCuda.cu:342); depth = NDC z in (0,1)
    assert r.means2D[0, 0] == pytest.approx(160.0, abs=1e-3) and r.means2D[0, 1] == pytest.approx(120.0, abs=1e-3)
    assert 0.0 < r.depths[0] < 1.0
    # DC-only colour 0.5 + 0.4*sh (GSCuda.cu:364-365)
    assert np.allclose(r.rgb[0], 0.5)


# ---- KAT-4 cull boundaries ---------------------------------------------------------------------
def test_kat4_near_plane_and_det(oracle):
    # camera at z=-5 looking +z: view z = world z + 5.  Keep iff view z > 0.2.
    zs = np.float32(-5.0) + np.array([0.2, np.nextafter(np.float32(0.2), np.float32(1)), 0.19, 0.5], np.float32)
    sc = tiny_scene([[0, 0, z] for z in zs], scales=[[0.001] * 3] * 4)
    r = oracle.forward_scene(sc, CAM)
    vz = r.depths
    # recompute exactly what the oracle sees
    expected_alive = []
    for z in zs:
        v = CAM.viewmatrix
        pvz = np.float32(np.float32(np.float32(v[2] * np.float32(0)) + np.float32(v[6] * np.float32(0)))
                         + np.float32(v[10] * z)) + v[14]
        expected_alive.append(bool(pvz > np.float32(0.2)))
    assert [bool(x > 0) for x in r.radii] == expected_alive
    assert expected_alive[2] is False and expected_alive[3] is True
    assert vz[3] == pytest.approx(0.5, abs=1e-5)
    # bounding-box cull (SIBR variant): everything outside the box is dropped
    sc2 = tiny_scene([[0, 0, 0], [1.0, 0, 0]])
    r2 = oracle.forward(P=2, D=3, M=16, background=np.zeros(3, np.float32), W=320, H=240, means3D=sc2.means3D,
                        shs=sc2.shs, colors_precomp=None, opacities=sc2.opacities, scales=sc2.scales,
                        scale_modifier=1.0, rotations=sc2.rotations, cov3D_precomp=None, viewmatrix=CAM.viewmatrix,
                        projmatrix=CAM.projmatrix, cam_pos=CAM.cam_pos, tan_fovx=CAM.tan_fovx, tan_fovy=CAM.tan_fovy,
                        boxmin=[-0.5, -0.5, -0.5], boxmax=[0.5, 0.5, 0.5])
    assert r2.radii[0] > 0 and r2.radii[1] == 0


# ---- KAT-5 spherical harmonics -----------------------------------------------------------------
def _sh_basis_f64(d):
    x, y, z = d
    C0 = 0.28209479177387814
    C1 = 0.4886025119029199
    C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
    C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    return np.array([C0, -C1 * y, C1 * z, -C1 * x, C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz,
                     C2[4] * (xx - yy), C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
                     C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
                     C3[6] * x * (xx - 3 * yy)])


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_kat5_sh_against_float64(oracle, deg):
    rng = np.random.default_rng(5)
    P = 64
    means = np.stack([rng.uniform(-1.5, 1.5, P), rng.uniform(-1.0, 1.0, P), rng.uniform(-1, 1, P)], 1)
    shs = rng.normal(0, 0.5, (P, 16, 3)).astype(np.float32)
    sc = tiny_scene(means, shs=shs)
    r = oracle.forward_scene(sc, CAM, D=deg)
    ncoef = (deg + 1) ** 2
    vis = r.radii > 0
    assert vis.sum() >= 48
    for i in np.nonzero(vis)[0]:
        d = sc.means3D[i].astype(np.float64) - CAM.cam_pos.astype(np.float64)
        d /= np.linalg.norm(d)
        b = _sh_basis_f64(d)[:ncoef]
        want = b @ shs[i, :ncoef, :].astype(np.float64) + 0.5
        assert np.array_equal(r.clamped[i].astype(bool), want < 0) or np.abs(want).min() < 1e-5
        assert np.allclose(r.rgb[i], np.maximum(want, 0), atol=2e-6)


# ---- KAT-6 key packing, emission order, stability ------------------------------------------------
def test_kat6_keys_and_stable_order(oracle):
    # two identical Gaussians (same tile set, same depth bits) + a nearer one
    sc = tiny_scene([[0, 0, 0], [0, 0, 0], [0, 0, -1.0]], scales=[[0.2] * 3] * 3)
    r = oracle.forward_scene(sc, CAM)
    gx, gy = r.grid
    assert r.tiles_touched[0] == r.tiles_touched[1] > 1
    # emission: Gaussian 0's tiles row-major, then Gaussian 1's, ... (GSCuda.cu:461-474)
    n0 = int(r.tiles_touched[0])
    tiles0 = (r.keys_unsorted[:n0] >> np.uint64(32)).astype(np.int64)
    assert np.all(np.diff(tiles0) > 0)
    assert np.all(r.values_unsorted[:n0] == 0) and np.all(r.values_unsorted[n0:2 * n0] == 1)
    assert np.all((r.keys_unsorted[:n0] & np.uint64(0xffffffff)) == r.depths[0:1].view(np.uint32)[0])
    assert np.array_equal(r.point_offsets, np.cumsum(r.tiles_touched, dtype=np.uint32))
    # sorted: tile-major, nearer first, ties by ascending Gaussian index
    k, v = r["keys"], r["values"]
    assert np.all(k[1:] >= k[:-1])
    t = int(tiles0[0])
    sel = (k >> np.uint64(32)) == np.uint64(t)
    assert list(v[sel]) == [2, 0, 1]
    # and it is exactly a stable argsort of the unsorted list
    order = np.argsort(r.keys_unsorted, kind="stable")
    assert np.array_equal(k, r.keys_unsorted[order]) and np.array_equal(v, r.values_unsorted[order])


def test_kat6_sort_pairs_matches_numpy_stable(oracle):
    rng = np.random.default_rng(6)
    for n, bits in [(1, 40), (17, 45), (5000, 44), (100_000, 47), (4096, 33)]:
        keys = rng.integers(0, 1 << bits, n, dtype=np.uint64)
        keys[rng.integers(0, n, n // 3)] = keys[0]  # many duplicates
        vals = np.arange(n, dtype=np.uint32)
        ko, vo = oracle.sort_pairs(keys, vals, bits, threads=3)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(ko, keys[order]) and np.array_equal(vo, vals[order])


# ---- KAT-7 identifyTileRanges (GSCuda.cu:504-538) -------------------------------------------------
def test_kat7_ranges(oracle):
    tiles = np.array([2, 2, 2, 5, 7, 7], np.uint64)
    keys = (tiles << np.uint64(32)) | np.arange(6, dtype=np.uint64)
    rg = oracle.identify_tile_ranges(keys, 9)
    want = np.zeros((9, 2), np.uint32)
    want[2], want[5], want[7] = (0, 3), (3, 4), (4, 6)
    assert np.array_equal(rg, want)  # empty tiles stay (0,0); first starts at 0; last closes at R
    # R == 1: contract closes the range; the in-tree kernel leaves (0,0) (close is inside the else)
    one = np.array([np.uint64(4) << np.uint64(32)], np.uint64)
    assert list(oracle.identify_tile_ranges(one, 9)[4]) == [0, 1]
    assert list(oracle.identify_tile_ranges(one, 9, compat=True)[4]) == [0, 0]
    # for R > 1 both modes agree
    assert np.array_equal(oracle.identify_tile_ranges(keys, 9, compat=True), want)


# ---- KAT-8 blend ------------------------------------------------------------------------------------
def _blend_scene(opac, colors, scale=0.3, n=1):
    return tiny_scene([[0, 0, 0.001 * i] for i in range(n)], scales=[[scale] * 3] * n, opac=opac, colors=colors)


def test_kat8_single_splat_alpha_cap_and_background(oracle):
    bg = (0.25, 0.5, 0.75)
    sc = _blend_scene([1.0], [[1.0, 0.5, 0.0]])
    r = oracle.forward_scene(sc, CAM, background=bg)
    # centre pixel (159.5,119.5) is half a pixel from the mean: alpha is capped at 0.99
    px = r.out_color[:, 120, 160]
    d = np.array([r.means2D[0, 0] - 160.0, r.means2D[0, 1] - 120.0], np.float64)
    co = r.conic_opacity[0].astype(np.float64)
    power = -0.5 * (co[0] * d[0] ** 2 + co[2] * d[1] ** 2) - co[1] * d[0] * d[1]
    alpha = min(0.99, 1.0 * math.exp(power))
    assert alpha == 0.99
    want = np.array([1.0, 0.5, 0.0]) * alpha + (1 - alpha) * np.array(bg)
    assert np.allclose(px, want, atol=1e-6)
    assert r.final_T[120 * 320 + 160] == pytest.approx(0.01, abs=1e-7)
    assert r.n_contrib[120 * 320 + 160] == 1
    # a far-away pixel of the same tile sees alpha < 1/255 -> pure background, n_contrib 0
    far = r.out_color[:, 0, 0]
    assert np.allclose(far, bg) and r.n_contrib[0] == 0 and r.final_T[0] == 1.0


def test_kat8_termination_threshold(oracle):
    # stack of opaque splats: T after k blends = 0.01^k.  T*(1-alpha) < 1e-4 stops BEFORE the blend:
    # k=1: T=1e-2; k=2: test_T = 1e-4*(1-eps) ... float: 0.01f*0.01f' -> compare with the oracle's own floats
    n = 4
    sc = _blend_scene([1.0] * n, [[1, 1, 1]] * n, n=n)
    r = oracle.forward_scene(sc, CAM)
    pix = 120 * 320 + 160
    T = np.float32(1.0)
    last = 0
    a = np.float32(0.99)
    for k in range(1, n + 1):
        t = np.float32(T * np.float32(np.float32(1.0) - a))
        if t < np.float32(0.0001):
            break
        T, last = t, k
    assert r.n_contrib[pix] == last and r.final_T[pix] == T
    # 0.99f is 0.99000001, so (1-0.99f)^2 = 9.99998e-5 < 1e-4: the SECOND splat already terminates the pixel
    assert last == 1 and T == np.float32(np.float32(1.0) - a)
    # GSRast mode stops one earlier (T < 0.001, GSCuda.cu:653)
    rc = oracle.forward_scene(sc, CAM, mode=oracle.MODE_GSRAST)
    # (colours come from SH in that mode; only the termination is checked)
    assert rc.n_contrib[pix] == 1


def test_kat8_low_alpha_skip(oracle):
    # opacity below 1/255 never contributes
    sc = _blend_scene([0.003], [[1, 1, 1]])
    r = oracle.forward_scene(sc, CAM, background=(0.1, 0.2, 0.3))
    assert r.num_rendered > 0 and r.n_contrib.max() == 0
    assert np.allclose(r.out_color[0], 0.1) and np.allclose(r.out_color[2], 0.3)


# ---- KAT-9 nothing rendered --------------------------------------------------------------------------
def test_kat9_empty_frame(oracle):
    sc = tiny_scene([[0, 0, -10.0]])  # behind the camera
    bg = (0.2, 0.4, 0.6)
    r = oracle.forward_scene(sc, CAM, background=bg)
    assert r.num_rendered == 0
    assert np.allclose(r.out_color[0], 0.2) and np.allclose(r.out_color[1], 0.4) and np.allclose(r.out_color[2], 0.6)
    # in-tree gscuda returns early and leaves the image stale (GSCuda.cu:775-778)
    stale = np.full((3, 240, 320), 9.0, np.float32)
    rc = oracle.forward_scene(sc, CAM, background=bg, mode=oracle.MODE_GSRAST, out_color_init=stale)
    assert rc.num_rendered == 0 and np.all(rc.out_color == 9.0)


def test_rects_variant(oracle):
    """SIBR fast-culling rects: extents (ceil(3 sqrt(cov.x)), ceil(3 sqrt(cov.z))); GSRast mode
    drops the sqrt on y (GSCuda.cu:352)."""
    sc = tiny_scene([[0.3, -0.2, 0]], scales=[[0.3, 0.05, 0.1]])
    r = oracle.forward_scene(sc, CAM, use_rects=True)
    co = r.conic_opacity[0].astype(np.float64)
    det = co[0] * co[2] - co[1] ** 2
    cov = np.array([co[2] / det, -co[1] / det, co[0] / det])  # invert the conic
    assert r.rects[0, 0] == math.ceil(3 * math.sqrt(cov[0]) - 1e-4) or r.rects[0, 0] == math.ceil(3 * math.sqrt(cov[0]))
    assert r.rects[0, 1] == math.ceil(3 * math.sqrt(cov[2]) - 1e-4) or r.rects[0, 1] == math.ceil(3 * math.sqrt(cov[2]))
    r0 = oracle.forward_scene(sc, CAM, use_rects=False)
    assert r.tiles_touched[0] <= r0.tiles_touched[0]  # the rect is never larger than the radius square
