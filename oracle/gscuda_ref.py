"""ctypes wrapper around oracle/_ref/libgscuda_ref.so — the reference's OWN in-tree rasterizer
(/root/reference/apps/gsrast/gscuda/GSCuda.cu, compiled unmodified by `make -C oracle ref`).
TEST INFRASTRUCTURE ONLY (GPU needed).  Used to pin the GSRast-mode oracle and as the
"reference built for sm_100" GPU baseline."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgscuda_ref.so")
# the same sources built with the reference's own flags (nvcc's default FMA contraction): report-only variant
LIB_PATH_FMAD = os.path.join(_HERE, "_ref", "libgscuda_ref_fmad.so")
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)
_libs = {}


class _Geom(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("tilesTouched", "depths", "clamped", "internalRadii", "means2D", "cov3D",
                                          "conicOpacity", "rgb", "pointOffsets", "total")]


class _Bin(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("keysUnsorted", "keys", "valuesUnsorted", "values", "total")]


class _Img(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("ranges", "nContrib", "accumAlpha", "total")]


def available(fmad: bool = False) -> bool:
    return os.path.exists(LIB_PATH_FMAD if fmad else LIB_PATH)


def lib(fmad: bool = False):
    _lib = _libs.get(fmad)
    if _lib is None:
        _lib = _libs[fmad] = C.CDLL(LIB_PATH_FMAD if fmad else LIB_PATH)
        _lib.gscuda_ref_forward.restype = None
        _lib.gscuda_ref_forward.argtypes = [ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p, C.c_int,
                                            C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.gscuda_ref_geometry_layout.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Geom)]
        _lib.gscuda_ref_binning_layout.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Bin)]
        _lib.gscuda_ref_image_layout.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Img)]
    return _lib


class RefRenderer:
    """Holds the viewer-layout device buffers and the grow-only allocators like GSGaussians does
    (GSGaussians.cpp:27-42,109-153) and calls the reference's gscuda::forward."""

    def __init__(self, scene, width, height, device="cuda", use_rects=True, fmad=False):
        import torch

        self.torch = torch
        self.lib = lib(fmad)
        self.dev = torch.device(device)
        self.W, self.H, self.P = width, height, scene.P
        means4, scales4, rot, opac, shs_raw = scene.gsrast_layout()
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)  # noqa: E731
        self.means, self.scales, self.rot, self.opac, self.shs = t(means4), t(scales4), t(rot), t(opac), t(shs_raw)
        self.colors = t(scene.colors_precomp) if scene.colors_precomp is not None else None
        self.bg = torch.zeros(3, device=self.dev)
        self.out = torch.zeros((3, height, width), device=self.dev)
        self.rects = torch.zeros((scene.P, 2), dtype=torch.int32, device=self.dev) if use_rects else None
        self.bufs = [None, None, None]
        self.sizes = [0, 0, 0]
        self.last_req = [0, 0, 0]

        def mk(i):
            def cb(n, _u):
                self.last_req[i] = int(n)
                if n > self.sizes[i]:
                    self.bufs[i] = None
                    self.bufs[i] = torch.zeros(2 * n, dtype=torch.uint8, device=self.dev)
                    self.sizes[i] = 2 * n
                return self.bufs[i].data_ptr()
            return ALLOC_FN(cb)

        self.cbs = [mk(0), mk(1), mk(2)]  # geometry, binning, image

    def draw(self, cam, background=(0.0, 0.0, 0.0)):
        torch = self.torch
        self.bg.copy_(torch.tensor(background, dtype=torch.float32))
        view = torch.from_numpy(cam.viewmatrix).to(self.dev)
        proj = torch.from_numpy(cam.projmatrix).to(self.dev)
        cpos = torch.from_numpy(cam.cam_pos).to(self.dev)
        torch.cuda.synchronize()
        p = lambda x: None if x is None else x.data_ptr()  # noqa: E731
        self.lib.gscuda_ref_forward(self.cbs[0], None, self.cbs[1], None, self.cbs[2], None, self.P, 3, 16, p(self.bg),
                                 self.W, self.H, p(self.means), p(self.shs), p(self.colors), p(self.opac),
                                 p(self.scales), 1.0, p(self.rot), None, p(view), p(proj), p(cpos), cam.tan_fovx,
                                 cam.tan_fovy, 0, p(self.out), None, p(self.rects), None, None)
        torch.cuda.synchronize()
        self._keep = (view, proj, cpos)

    def state(self):
        """Scratch fields of the last draw() as numpy arrays (reference chunk layouts)."""
        torch = self.torch
        P, N = self.P, self.W * self.H
        g = _Geom()
        gb = self.bufs[0]
        self.lib.gscuda_ref_geometry_layout(gb.data_ptr(), P, C.byref(g))

        def view(buf, off, nbytes, dt):
            rel = off
            return buf[rel:rel + nbytes].cpu().numpy().view(dt)

        out = dict(
            tiles_touched=view(gb, g.tilesTouched, 4 * P, np.uint32), depths=view(gb, g.depths, 4 * P, np.float32),
            radii=view(gb, g.internalRadii, 4 * P, np.int32), means2D=view(gb, g.means2D, 8 * P, np.float32).reshape(P, 2),
            cov3D=view(gb, g.cov3D, 24 * P, np.float32).reshape(P, 6),
            conic_opacity=view(gb, g.conicOpacity, 16 * P, np.float32).reshape(P, 4),
            rgb=view(gb, g.rgb, 12 * P, np.float32).reshape(P, 3),
            point_offsets=view(gb, g.pointOffsets, 4 * P, np.uint32))
        R = int(out["point_offsets"][-1]) if P else 0
        out["num_rendered"] = R
        if R > 0:
            b = _Bin()
            bb = self.bufs[1]
            self.lib.gscuda_ref_binning_layout(bb.data_ptr(), R, C.byref(b))
            out["keys_unsorted"] = view(bb, b.keysUnsorted, 8 * R, np.uint64)
            out["keys"] = view(bb, b.keys, 8 * R, np.uint64)
            out["values_unsorted"] = view(bb, b.valuesUnsorted, 4 * R, np.uint32)
            out["values"] = view(bb, b.values, 4 * R, np.uint32)
        im = _Img()
        ib = self.bufs[2]
        self.lib.gscuda_ref_image_layout(ib.data_ptr(), N, C.byref(im))
        T = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        out["ranges"] = view(ib, im.ranges, 8 * T, np.uint32).reshape(T, 2)  # the reference sizes it per pixel
        out["n_contrib"] = view(ib, im.nContrib, 4 * N, np.uint32)
        out["final_T"] = view(ib, im.accumAlpha, 4 * N, np.float32)
        out["out_color"] = self.out.cpu().numpy()
        out["rects"] = self.rects.cpu().numpy() if self.rects is not None else None
        return out
