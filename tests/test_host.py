"""Host-side logic on CPU: camera conventions, synthetic scenes, the 62-float PLY, view
sharding and the frame gather (gloo, world_size 2)."""
import math
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from gsrast_b200 import camera as Cm
from gsrast_b200 import scene as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_camera_matches_gsrast_conventions():
    cam = Cm.default_camera(1920, 1080)
    v = cam.viewmatrix.reshape(4, 4)  # [col,row]
    # eye (0,0,-5) looking at the origin: view-space z of the origin is +5 after the row-2 flip
    p = np.array([0, 0, 0, 1], np.float32)
    pv = np.einsum("cr,c->r", v, p)
    assert pv[2] == np.float32(5.0)
    # proj = perspective * (unflipped) view: clip w = distance along the view axis
    ph = np.einsum("cr,c->r", cam.projmatrix.reshape(4, 4), p)
    assert abs(ph[3] - 5.0) < 1e-6 and abs(ph[0]) < 1e-6 and abs(ph[1]) < 1e-6
    assert cam.tan_fovy == np.float32(math.tan(math.radians(45) / 2))
    assert abs(cam.tan_fovx - cam.tan_fovy * 1920 / 1080) < 1e-6
    # invertUp: world +y maps to +y_ndc * (-1) ... i.e. image rows grow along world +y? check sign consistency
    up = np.einsum("cr,c->r", cam.projmatrix.reshape(4, 4), np.array([0, 1, 0, 1], np.float32))
    assert up[1] < 0  # up vector (0,-1,0): world +y points down the image


def test_scene_generator_is_seeded_and_shaped():
    a, cfg = S.make_config_scene("C1", P=5000)
    b, _ = S.make_config_scene("C1", P=5000)
    assert np.array_equal(a.means3D, b.means3D) and np.array_equal(a.shs, b.shs)
    assert a.means3D.shape == (5000, 3) and a.shs.shape == (5000, 16, 3) and a.rotations.shape == (5000, 4)
    assert np.allclose(np.linalg.norm(a.rotations, axis=1), 1, atol=1e-5)
    assert 0 < a.opacities.min() and a.opacities.max() < 1
    assert np.linalg.norm(a.means3D, axis=1).max() < 3.6
    c5, _ = S.make_config_scene("C5", P=1000)
    assert c5.shs is None and c5.colors_precomp.shape == (1000, 3) and c5.opacities.mean() < 0.15
    for name in ("C1", "C2", "C3", "C4", "C5"):
        assert name in S.CONFIGS


def test_ply_round_trip(tmp_path):
    sc, _ = S.make_config_scene("C1", P=777)
    path = str(tmp_path / "data.ply")
    S.write_ply(path, sc)
    assert os.path.getsize(path) > 777 * 62 * 4
    back = S.read_ply(path)
    assert back.P == 777
    assert np.array_equal(back.means3D, sc.means3D) and np.array_equal(back.shs, sc.shs)
    assert np.allclose(back.scales, sc.scales, rtol=1e-6) and np.allclose(back.opacities, sc.opacities, atol=1e-6)
    assert np.allclose(back.rotations, sc.rotations, atol=1e-6)
    # PLY order <-> contract order (the SH layout trap)
    raw = S.contract_sh_to_ply_order(sc.shs)
    assert raw[5, 3 + 1 * 15 + (4 - 1)] == sc.shs[5, 4, 1]
    assert np.array_equal(S.ply_order_to_contract_sh(raw), sc.shs)


def test_native_ply_loader_matches_splatdata_semantics(tmp_path):
    """gsr_ply_count / gsr_ply_load (csrc/ply.cu, host code of the C ABI) against the numpy restatement of
    SplatData::loadFromSplatsPly + activation (SplatData.cpp:114-156, 48-58), and its three error paths."""
    from gsrast_b200 import _lib

    sc, _ = S.make_config_scene("C1", P=20_001)  # > one 16384-record read chunk
    path = str(tmp_path / "data.ply")
    S.write_ply(path, sc)
    ref = S.read_ply(path)
    nat = S.load_ply_native(path)
    assert nat.P == ref.P == 20_001
    assert np.array_equal(nat.means3D, ref.means3D) and np.array_equal(nat.shs, ref.shs)
    assert np.allclose(nat.scales, ref.scales, rtol=2e-7) and np.allclose(nat.opacities, ref.opacities, atol=2e-7)
    assert np.allclose(nat.rotations, ref.rotations, atol=2e-7)
    assert np.allclose(nat.bbox[0], ref.means3D.min(0)) and np.allclose(nat.bbox[1], ref.means3D.max(0))
    assert np.allclose(nat.center, ref.means3D.astype(np.float64).mean(0), atol=1e-5)
    L = _lib.lib()
    # missing file, header without a count on line 3, truncated body (SplatData.cpp:120-124, 146-152)
    with pytest.raises(ValueError, match="cannot open"):
        S.load_ply_native(str(tmp_path / "nope.ply"))
    bad = tmp_path / "bad.ply"
    bad.write_bytes(b"ply\nformat binary_little_endian 1.0\ncomment nothing here\nend_header\n")
    with pytest.raises(ValueError, match="vertex count"):
        S.load_ply_native(str(bad))
    raw = open(path, "rb").read()
    cut = tmp_path / "cut.ply"
    cut.write_bytes(raw[:-4000])
    with pytest.raises(ValueError, match="shorter"):
        S.load_ply_native(str(cut))
    import ctypes as C
    n = C.c_int(-1)
    assert L.gsr_ply_count(path.encode(), C.byref(n)) == 0 and n.value == 20_001
    assert L.gsr_ply_load(path.encode(), 5, None, None, None, None, None, None, None) == _lib.ERR_INVALID_ARG  # capacity
    assert L.gsr_ply_load(path.encode(), 20_001, None, None, None, None, None, None, None) == 20_001  # all outputs optional


def test_shard_views_partition():
    from gsrast_b200.views import shard_views, views_per_rank

    for n in (0, 1, 7, 256):
        for w in (1, 2, 4, 8):
            parts = [shard_views(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            assert max(len(p) for p in parts) <= views_per_rank(n, w) if n else True


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from gsrast_b200.views import shard_views, gather_frames
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, n = dist.get_rank(), 5
mine = shard_views(n, rank, 2)
local = torch.stack([torch.full((3, 4, 6), float(v)) for v in mine])
out = gather_frames(local, n, rank, 2, dst=0)
if rank == 0:
    assert out.shape == (5, 3, 4, 6)
    assert [float(out[v, 0, 0, 0]) for v in range(5)] == [0.0, 1.0, 2.0, 3.0, 4.0]
    print("GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
'''


def test_gather_frames_gloo_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference times the CPU implementation and prints the contract's JSON line."""
    import json

    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0", "--workload", "C1"], timeout=600).decode()
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_numa_placement_is_best_effort():
    """bind_to_gpu_numa_node never raises and leaves the affinity mask alone when it cannot place the process
    (no GPU / no sysfs entry / single node)."""
    from gsrast_b200.views import bind_to_gpu_numa_node

    before = os.sched_getaffinity(0)
    res = bind_to_gpu_numa_node(0)
    assert isinstance(res, dict) and ("skipped" in res or "node" in res)
    if "skipped" in res:
        assert os.sched_getaffinity(0) == before
    else:
        assert os.sched_getaffinity(0) <= before
        os.sched_setaffinity(0, before)


def _load_bench():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_byte_model_follows_the_survey():
    """bench.py rates preprocess on SURVEY 8(d)'s bytes exactly (284 B x visible + 20 B x culled), the reference's sort
    formulation at 152 B/pair, and — physically — the lean expansion at less than the full-state one."""
    b = _load_bench()
    P, vis, R, Rc, tiles = 3_300_000, 2_732_659, 15_515_568, 3_650_085, 8160
    lean = b.algorithmic_bytes(P, vis, P - vis, R, tiles, 4, 1, False, Rc=Rc, lean=True)
    full = b.algorithmic_bytes(P, vis, P - vis, R, tiles, 4, 1, False, Rc=Rc, lean=False)
    assert lean["preprocess"] == 284 * vis + 20 * (P - vis) == 787_421_976
    assert lean["sort_survey"] == 152 * R
    assert lean["expand_fill"] < full["expand_fill"] and lean["expand_count"] < full["expand_count"]
    assert full["expand_fill"] - lean["expand_fill"] == 8 * R + 4 * Rc  # the 64-bit keys and the depth half of the records
    assert lean["preprocess_moved"] > lean["preprocess"]
    radix = b.algorithmic_bytes(P, vis, P - vis, R, tiles, 4, 2, False, Rc=0)
    assert radix["sort_moved"] > lean["sort_moved"]


def test_bench_job_cameras_are_the_same_for_both_arms_and_any_step_count():
    """The views a rank renders depend on (workload, world, rank) only: the reference arm renders what the GPU arm's
    rank 0 renders, and a different --steps only makes the list longer."""
    b = _load_bench()
    short = b.job_cameras("C4", 4, 2, 2, 1, 640, 360)
    long_ = b.job_cameras("C4", 16, 2, 2, 1, 640, 360)
    assert len(short) == 6 and len(long_) == 18
    for a, c in zip(short, long_):
        assert np.array_equal(a.packed(), c.packed())
    other = b.job_cameras("C4", 4, 2, 2, 0, 640, 360)
    assert not np.array_equal(other[0].packed(), short[0].packed())
    single = b.job_cameras("C2", 3, 1, 1, 0, 640, 360)
    assert all(np.array_equal(single[0].packed(), c.packed()) for c in single)


def test_python_flag_and_error_constants_match_the_header():
    """gsrast_b200/_lib.py mirrors include/gsrast_b200.h by hand: every GSR_FLAG_* / GSR_ERR_* value must agree."""
    import re

    from gsrast_b200 import _lib

    text = open(os.path.join(ROOT, "include", "gsrast_b200.h")).read()
    flags = {m.group(1): int(m.group(2), 16) for m in re.finditer(r"#define\s+GSR_FLAG_(\w+)\s+0x([0-9a-fA-F]+)u", text)}
    errs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+GSR_ERR_(\w+)\s+\((-\d+)\)", text)}
    assert len(flags) >= 7 and len(errs) >= 4
    for name, val in flags.items():
        assert getattr(_lib, "FLAG_" + name) == val, name
    for name, val in errs.items():
        assert getattr(_lib, "ERR_" + name) == val, name
    assert len(set(flags.values())) == len(flags)  # no two flags share a bit
