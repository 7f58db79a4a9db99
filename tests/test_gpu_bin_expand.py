"""GPU parity of the bin expansion (gsrast_b200/csrc/bin_expand.cu) — the default way the tile half
of the (tile << 32 | depth) sort is produced — against the CPU oracle and against the radix path
(GSR_FLAG_RADIX_BINNING), bit for bit: sorted keys, sorted values, tile ranges.

The reference gets these from cub::DeviceRadixSort::SortPairs + identifyTileRanges
(/root/reference/apps/gsrast/gscuda/GSCuda.cu:794-801, 504-538).  The cases move the number of bins
(1, a few, > 256 = two bin-digit passes, > 4096 = automatic fallback to the radix path), make bins and
chunks ragged, and make single chunks emit more pairs than one staging round holds."""
import numpy as np
import pytest

from gsrast_b200 import _lib
from gsrast_b200 import camera as Cm
from gsrast_b200 import scene as S

from helpers import assert_parity, run_cuda, run_oracle
from test_oracle_kat import tiny_scene

pytestmark = pytest.mark.gpu


def _same_lists(a, b):
    assert a["num_rendered"] == b["num_rendered"]
    assert np.array_equal(a["keys"], b["keys"]), "sorted keys"
    assert np.array_equal(a["values"], b["values"]), "sorted values"
    assert np.array_equal(a["ranges"], b["ranges"]), "tile ranges"


def _lean_same(full, sc, cam, **kw):
    """The renderer's lean state (GSR_FLAG_LEAN_STATE: the id-only instantiations of the count / fill passes, no sorted
    64-bit keys) must leave the same sorted Gaussian list, tile ranges and frame."""
    lean = run_cuda(sc, cam, flags=_lib.FLAG_LEAN_STATE, **kw)
    assert lean["num_rendered"] == full["num_rendered"]
    assert np.array_equal(lean["values"], full["values"]), "sorted values (lean)"
    assert np.array_equal(lean["ranges"], full["ranges"]), "tile ranges (lean)"
    assert np.array_equal(lean["n_contrib"], full["n_contrib"]) and np.array_equal(lean["out_color"], full["out_color"])


@pytest.mark.parametrize("W,H,bin_passes", [(100, 60, 1), (128, 128, 1), (1000, 555, 1), (1920, 1080, 1),
                                            (3840, 2160, 2), (4096, 4200, 2)])
def test_bin_counts(oracle, W, H, bin_passes):
    from gsrast_b200.rasterizer import get_higher_msb

    gx, gy = (W + 15) // 16, (H + 15) // 16
    nbins = ((gx + 7) // 8) * ((gy + 7) // 8)
    assert (get_higher_msb(nbins) + 7) // 8 == bin_passes
    sc = S.make_config_scene("C1", P=23_457)[0]
    cam = Cm.default_camera(W, H)
    cu = run_cuda(sc, cam, timings=True)
    ref = run_oracle(oracle, sc, cam)
    assert_parity(cu, ref)
    t = cu["times"]
    assert t["binning_mode"] == 0 and t["sort_passes"] == 4 + bin_passes
    assert 0 < t["num_coarse"] <= t["num_rendered"]
    _same_lists(cu, run_cuda(sc, cam, flags=_lib.FLAG_RADIX_BINNING))
    _lean_same(cu, sc, cam)


def test_more_than_4096_bins_falls_back(oracle):
    W, H = 16385, 4100  # 1025 x 257 tiles -> 129 x 33 = 4257 bins
    sc = S.make_config_scene("C1", P=3_000)[0]
    cam = Cm.default_camera(W, H)
    cu = run_cuda(sc, cam, timings=True)
    assert cu["times"]["binning_mode"] == 1
    ref = run_oracle(oracle, sc, cam)
    assert_parity(cu, ref, check_image=False)


def test_big_splats_multi_round_staging(oracle):
    """Splats covering whole bins: one chunk of 512 records emits up to 32k pairs (> the 4096 staged per
    round), every record touches many bins, bins see every Gaussian."""
    rng = np.random.default_rng(5)
    n = 3000
    xyz = rng.uniform(-1.5, 1.5, size=(n, 3)).astype(np.float32)
    sc = tiny_scene(xyz, scales=rng.uniform(0.3, 1.2, size=(n, 3)).astype(np.float32))
    cam = Cm.default_camera(640, 400)
    cu = run_cuda(sc, cam, timings=True)
    ref = run_oracle(oracle, sc, cam)
    assert ref.num_rendered > 100 * n
    assert cu["times"]["binning_mode"] == 0
    _same_lists(cu, dict(num_rendered=ref.num_rendered, keys=ref["keys"], values=ref["values"], ranges=ref.ranges))
    _same_lists(cu, run_cuda(sc, cam, flags=_lib.FLAG_RADIX_BINNING))
    _lean_same(cu, sc, cam)  # the multi-round path of the id-only fill pass


def test_depth_ties_and_compat(oracle):
    """Exact depth ties (stability decides) and the GSRast-compat mode through the bin expansion."""
    rng = np.random.default_rng(12)
    n = 5000
    xy = rng.uniform(-1.2, 1.2, size=(n, 2)).astype(np.float32)
    sc = tiny_scene(np.concatenate([xy, np.zeros((n, 1), np.float32)], axis=1), scales=np.full((n, 3), 0.08, np.float32))
    cam = Cm.default_camera(640, 400)
    cu = run_cuda(sc, cam)
    ref = run_oracle(oracle, sc, cam)
    _same_lists(cu, dict(num_rendered=ref.num_rendered, keys=ref["keys"], values=ref["values"], ranges=ref.ranges))
    _lean_same(cu, sc, cam)
    sc2 = S.make_config_scene("C2", P=60_001)[0]
    cam2 = Cm.orbit_cameras(5, 1280, 720)[3]
    a = run_cuda(sc2, cam2, compat=True, use_rects=True)
    b = run_cuda(sc2, cam2, compat=True, use_rects=True, flags=_lib.FLAG_RADIX_BINNING)
    _same_lists(a, b)
    assert np.abs(a["out_color"] - b["out_color"]).max() == 0.0
    _lean_same(a, sc2, cam2, compat=True, use_rects=True)


@pytest.mark.parametrize("P", [1, 2, 31, 33, 511, 513, 2049])
def test_ragged_record_counts(oracle, P):
    sc = S.make_config_scene("C1", P=P)[0]
    cam = Cm.default_camera(200, 120)
    cu = run_cuda(sc, cam)
    ref = run_oracle(oracle, sc, cam)
    assert_parity(cu, ref, check_image=ref.num_rendered > 0)
    if ref.num_rendered > 0:
        _lean_same(cu, sc, cam)
