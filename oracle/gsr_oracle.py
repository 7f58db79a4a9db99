"""ctypes wrapper around oracle/libgsr_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module; nothing under gsrast_b200/ does.  See gsr_oracle.cpp for what
it restates and for the parity-pinning status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgsr_oracle.so")
_lib = None

MODE_CONTRACT = 0
MODE_GSRAST = 1


class _In(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("P", "D", "M", "W", "H", "means_stride", "scales_stride", "flags", "prefiltered", "use_rects",
                 "threads")] + \
               [(n, C.c_float) for n in ("scale_modifier", "tan_fovx", "tan_fovy")] + \
               [(n, C.c_void_p) for n in
                ("background", "means3D", "shs", "colors_precomp", "opacities", "scales", "rotations",
                 "cov3D_precomp", "viewmatrix", "projmatrix", "cam_pos", "boxmin", "boxmax")]


class _Out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("depths", "clamped", "radii", "means2D", "cov3D", "conic_opacity", "rgb", "tiles_touched",
                 "point_offsets", "rects", "ranges", "n_contrib", "final_T", "out_color",
                 "keys_unsorted", "values_unsorted", "keys", "values")] + \
               [("num_rendered", C.c_int64), ("pairs_evaluated", C.c_int64)] + \
               [(n, C.c_double) for n in
                ("t_preprocess", "t_scan", "t_duplicate", "t_sort", "t_ranges", "t_blend", "t_total")]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "gsr_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libgsr_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.gsr_oracle_forward.restype = C.c_int64
        _lib.gsr_oracle_forward.argtypes = [C.POINTER(_In), C.POINTER(_Out)]
        _lib.gsr_oracle_free.argtypes = [C.POINTER(_Out)]
        _lib.gsr_oracle_get_higher_msb.restype = C.c_uint32
        _lib.gsr_oracle_get_higher_msb.argtypes = [C.c_uint32]
        _lib.gsr_oracle_get_rect.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.gsr_oracle_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                               C.c_int]
        _lib.gsr_oracle_identify_tile_ranges.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        _lib.gsr_oracle_hardware_threads.restype = C.c_int
    return _lib


def hardware_threads() -> int:
    return int(lib().gsr_oracle_hardware_threads())


def get_higher_msb(n: int) -> int:
    return int(lib().gsr_oracle_get_higher_msb(n))


def get_rect(px, py, ex, ey, gx, gy):
    out = np.zeros(4, dtype=np.uint32)
    lib().gsr_oracle_get_rect(px, py, ex, ey, gx, gy, out.ctypes.data)
    return tuple(int(v) for v in out)  # minx, miny, maxx, maxy


def sort_pairs(keys: np.ndarray, vals: np.ndarray, end_bit: int, threads: int = 1):
    k = np.ascontiguousarray(keys, dtype=np.uint64).copy()
    v = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    ko = np.empty_like(k)
    vo = np.empty_like(v)
    lib().gsr_oracle_sort_pairs(k.ctypes.data, v.ctypes.data, ko.ctypes.data, vo.ctypes.data, k.size, end_bit,
                                threads)
    return ko, vo


def identify_tile_ranges(keys: np.ndarray, num_tiles: int, compat: bool = False) -> np.ndarray:
    k = np.ascontiguousarray(keys, dtype=np.uint64)
    ranges = np.zeros((num_tiles, 2), dtype=np.uint32)
    lib().gsr_oracle_identify_tile_ranges(k.size, k.ctypes.data, ranges.ctypes.data, 1 if compat else 0)
    return ranges


def _p(a):
    return None if a is None else a.ctypes.data


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class OracleResult(dict):
    __getattr__ = dict.__getitem__


def forward(*, P, D, M, background, W, H, means3D, shs, colors_precomp, opacities, scales, scale_modifier,
            rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered=False,
            use_rects=False, boxmin=None, boxmax=None, mode=MODE_CONTRACT, threads=None,
            out_color_init=None) -> OracleResult:
    """Run the CPU restatement of forward() (argument names = the reference's, GSCuda.cuh:103-126)."""
    means3D = _f32(means3D)
    scales = _f32(scales)
    means_stride = int(means3D.shape[1]) if means3D.ndim == 2 else 3
    scales_stride = (int(scales.shape[1]) if scales.ndim == 2 else 3) if scales is not None else 3
    keep = dict(background=_f32(background), means3D=means3D, shs=_f32(shs), colors_precomp=_f32(colors_precomp),
                opacities=_f32(opacities), scales=scales, rotations=_f32(rotations),
                cov3D_precomp=_f32(cov3D_precomp), viewmatrix=_f32(viewmatrix), projmatrix=_f32(projmatrix),
                cam_pos=_f32(cam_pos), boxmin=_f32(boxmin), boxmax=_f32(boxmax))
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    N = W * H
    o = dict(
        depths=np.zeros(P, np.float32), clamped=np.zeros((P, 3), np.uint8), radii=np.zeros(P, np.int32),
        means2D=np.zeros((P, 2), np.float32), cov3D=np.zeros((P, 6), np.float32),
        conic_opacity=np.zeros((P, 4), np.float32), rgb=np.zeros((P, 3), np.float32),
        tiles_touched=np.zeros(P, np.uint32), point_offsets=np.zeros(P, np.uint32),
        rects=np.zeros((P, 2), np.int32), ranges=np.zeros((T, 2), np.uint32), n_contrib=np.zeros(N, np.uint32),
        final_T=np.zeros(N, np.float32),
        out_color=(np.zeros((3, H, W), np.float32) if out_color_init is None
                   else np.ascontiguousarray(out_color_init, dtype=np.float32).copy()))
    if threads is None:
        threads = hardware_threads()
    cin = _In(P=P, D=D, M=M, W=W, H=H, means_stride=means_stride, scales_stride=scales_stride, flags=int(mode),
              prefiltered=int(prefiltered), use_rects=int(use_rects), threads=int(threads),
              scale_modifier=float(scale_modifier), tan_fovx=float(tan_fovx), tan_fovy=float(tan_fovy),
              **{k: _p(v) for k, v in keep.items()})
    cout = _Out(**{k: _p(v) for k, v in o.items()})
    R = int(lib().gsr_oracle_forward(C.byref(cin), C.byref(cout)))
    if R < 0:
        raise RuntimeError("oracle forward failed: %d" % R)
    res = OracleResult(o)
    res["num_rendered"] = R
    res["pairs_evaluated"] = int(cout.pairs_evaluated)
    res["grid"] = (gx, gy)
    for name, dt in (("keys_unsorted", np.uint64), ("values_unsorted", np.uint32), ("keys", np.uint64),
                     ("values", np.uint32)):
        ptr = getattr(cout, name)
        if R > 0 and ptr:
            buf = (C.c_char * (R * np.dtype(dt).itemsize)).from_address(ptr)
            res[name] = np.frombuffer(buf, dtype=dt).copy()
        else:
            res[name] = np.zeros(0, dtype=dt)
    res["timings"] = {k: getattr(cout, "t_" + k) for k in
                      ("preprocess", "scan", "duplicate", "sort", "ranges", "blend", "total")}
    res["threads"] = int(threads)
    lib().gsr_oracle_free(C.byref(cout))
    return res


def forward_scene(scene, cam, background=(0.0, 0.0, 0.0), D=None, use_rects=False, mode=MODE_CONTRACT,
                  threads=None, **kw) -> OracleResult:
    """Convenience: run on a gsrast_b200.scene.SplatScene + gsrast_b200.camera.Camera."""
    if mode == MODE_GSRAST:
        means4, scales4, rot, opac, shs_raw = scene.gsrast_layout()
        return forward(P=scene.P, D=3, M=16, background=np.asarray(background, np.float32), W=cam.width,
                       H=cam.height, means3D=means4, shs=shs_raw, colors_precomp=scene.colors_precomp,
                       opacities=opac, scales=scales4, scale_modifier=1.0, rotations=rot, cov3D_precomp=None,
                       viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, cam_pos=cam.cam_pos,
                       tan_fovx=cam.tan_fovx, tan_fovy=cam.tan_fovy, use_rects=use_rects, mode=mode,
                       threads=threads, **kw)
    return forward(P=scene.P, D=scene.sh_degree if D is None else D, M=scene.max_coeffs,
                   background=np.asarray(background, np.float32), W=cam.width, H=cam.height,
                   means3D=scene.means3D, shs=scene.shs, colors_precomp=scene.colors_precomp,
                   opacities=scene.opacities, scales=scene.scales, scale_modifier=1.0, rotations=scene.rotations,
                   cov3D_precomp=None, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, cam_pos=cam.cam_pos,
                   tan_fovx=cam.tan_fovx, tan_fovy=cam.tan_fovy, use_rects=use_rects, mode=mode, threads=threads,
                   **kw)
