"""Host-side mirror of the reference's operator interface for the splat draw path.

Python here is plumbing only (device memory via torch, ctypes into libgsrast_b200.so); all
compute is in the CUDA library and there is no CPU fallback.

Mirrors (file:line relative to /root/reference):
  Rasterizer.forward        CudaRasterizer::Rasterizer::forward as invoked at
                            apps/gsrast/GSGaussians.cpp:179-206 (same argument names and order)
  gscuda_forward            gscuda::forward, apps/gsrast/gscuda/GSCuda.cuh:103-126
  resize_functional         resizeFunctional, apps/gsrast/GSGaussians.cpp:27-42
  GeometryState.from_chunk  gscuda::gs::GeometryState::fromChunk, gscuda/AuxBuffer.cu:44-63
  ImageState / BinningState AuxBuffer.cu:65-89
  required                  required<T>, gscuda/AuxBuffer.cuh:8-14
  GSGaussians               apps/gsrast/GSGaussians.{hpp,cpp} (configureFromSplatData, draw,
                            mapGeometryState)
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import FLAG_BLEND_COUNT, FLAG_BLEND_ONE_PIXEL, FLAG_BLEND_SIMPLE, FLAG_GSRAST_COMPAT  # noqa: F401

NUM_CHANNELS = 3  # Config.hpp:46
BLOCK_X = 16      # Config.hpp:47
BLOCK_Y = 16      # Config.hpp:48


def _ptr(x):
    """Device (or host) address of a tensor / ndarray / int / None."""
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return int(x)


class ResizeFunctional:
    """resizeFunctional (GSGaussians.cpp:27-42): grow-only device buffer; a request larger
    than the current size frees it and allocates 2*N; otherwise the same base is returned."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.buf = None
        self.size = 0
        self.calls = 0
        self.requests = []

    def __call__(self, n: int) -> int:
        self.calls += 1
        self.requests.append(int(n))
        if n > self.size:
            self.buf = None  # cudaFree
            self.buf = torch.empty(2 * n, dtype=torch.uint8, device=self.device)
            self.size = 2 * n
        return self.buf.data_ptr() if self.buf is not None else 0

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() if self.buf is not None else 0


def resize_functional(device="cuda") -> ResizeFunctional:
    return ResizeFunctional(device)


def _wrap_alloc(fn):
    def cb(nbytes, _user):
        try:
            p = fn(int(nbytes))
            return int(p) if p else 0
        except Exception:  # never let an exception cross the C boundary
            return 0
    return _lib.ALLOC_FN(cb)


def get_higher_msb(n: int) -> int:
    """getHigherMsb (GSCuda.cu:481-502)."""
    return int(_lib.lib().gsr_get_higher_msb(n))


def _view(chunk: torch.Tensor, base: int, addr: int, nbytes: int, dtype, shape):
    off = addr - base
    return chunk[off:off + nbytes].view(dtype).view(*shape)


class GeometryState:
    """Field views into the geometry chunk (AuxBuffer.cuh:38-54; our field order is the
    CudaRasterizer one — use this accessor, not offsets)."""

    @staticmethod
    def required(P: int) -> int:
        return int(_lib.lib().gsr_geometry_state_required(P))

    @staticmethod
    def from_chunk(chunk: torch.Tensor, P: int) -> dict:
        st = _lib.GeometryState()
        base = chunk.data_ptr()
        _lib.lib().gsr_geometry_state_map(base, P, C.byref(st))
        f32, i32, u8 = torch.float32, torch.int32, torch.uint8
        return dict(
            depths=_view(chunk, base, st.depths, 4 * P, f32, (P,)),
            clamped=_view(chunk, base, st.clamped, 3 * P, u8, (P, 3)),
            internal_radii=_view(chunk, base, st.internal_radii, 4 * P, i32, (P,)),
            means2D=_view(chunk, base, st.means2D, 8 * P, f32, (P, 2)),
            cov3D=_view(chunk, base, st.cov3D, 24 * P, f32, (P, 6)),
            conic_opacity=_view(chunk, base, st.conic_opacity, 16 * P, f32, (P, 4)),
            rgb=_view(chunk, base, st.rgb, 12 * P, f32, (P, 3)),
            tiles_touched=_view(chunk, base, st.tiles_touched, 4 * P, i32, (P,)),
            point_offsets=_view(chunk, base, st.point_offsets, 4 * P, i32, (P,)),
        )


class ImageState:
    @staticmethod
    def required(W: int, H: int) -> int:
        return int(_lib.lib().gsr_image_state_required(W, H))

    @staticmethod
    def from_chunk(chunk: torch.Tensor, W: int, H: int) -> dict:
        st = _lib.ImageState()
        base = chunk.data_ptr()
        _lib.lib().gsr_image_state_map(base, W, H, C.byref(st))
        T = ((W + BLOCK_X - 1) // BLOCK_X) * ((H + BLOCK_Y - 1) // BLOCK_Y)
        N = W * H
        return dict(
            ranges=_view(chunk, base, st.ranges, 8 * T, torch.int32, (T, 2)),
            n_contrib=_view(chunk, base, st.n_contrib, 4 * N, torch.int32, (N,)),
            accum_alpha=_view(chunk, base, st.accum_alpha, 4 * N, torch.float32, (N,)),
        )


class BinningState:
    @staticmethod
    def required(R: int) -> int:
        return int(_lib.lib().gsr_binning_state_required(R))

    @staticmethod
    def from_chunk(chunk: torch.Tensor, R: int) -> dict:
        st = _lib.BinningState()
        base = chunk.data_ptr()
        _lib.lib().gsr_binning_state_map(base, R, C.byref(st))
        return dict(
            point_list_keys_unsorted=_view(chunk, base, st.point_list_keys_unsorted, 8 * R, torch.int64, (R,)),
            point_list_keys=_view(chunk, base, st.point_list_keys, 8 * R, torch.int64, (R,)),
            point_list_unsorted=_view(chunk, base, st.point_list_unsorted, 4 * R, torch.int32, (R,)),
            point_list=_view(chunk, base, st.point_list, 4 * R, torch.int32, (R,)),
        )


def _forward(geometryBuffer, binningBuffer, imageBuffer, P, D, M, background, width, height, means3D, shs,
             colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
             cam_pos, tan_fovx, tan_fovy, prefiltered, out_color, radii, rects, boxmin, boxmax, stream, flags,
             means_stride, scales_stride, timings):
    L = _lib.lib()
    cbs = [_wrap_alloc(geometryBuffer), _wrap_alloc(binningBuffer), _wrap_alloc(imageBuffer)]
    keep = []

    def host3(v):
        if v is None:
            return None
        a = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(3))
        keep.append(a)
        return a.ctypes.data

    times = _lib.StageTimes() if timings else None
    dev = out_color.device if isinstance(out_color, torch.Tensor) else None
    if stream is None:
        stream_ptr = torch.cuda.current_stream(dev).cuda_stream
    else:
        stream_ptr = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
    a = _lib.ForwardArgs(
        geometry_alloc=cbs[0], geometry_user=None, binning_alloc=cbs[1], binning_user=None,
        image_alloc=cbs[2], image_user=None, P=int(P), D=int(D), M=int(M), background=_ptr(background),
        width=int(width), height=int(height), means3D=_ptr(means3D), means_stride=int(means_stride),
        shs=_ptr(shs), colors_precomp=_ptr(colors_precomp), opacities=_ptr(opacities), scales=_ptr(scales),
        scales_stride=int(scales_stride), scale_modifier=float(scale_modifier), rotations=_ptr(rotations),
        cov3D_precomp=_ptr(cov3D_precomp), viewmatrix=_ptr(viewmatrix), projmatrix=_ptr(projmatrix),
        cam_pos=_ptr(cam_pos), tan_fovx=float(tan_fovx), tan_fovy=float(tan_fovy), prefiltered=int(bool(prefiltered)),
        out_color=_ptr(out_color), radii=_ptr(radii), rects=_ptr(rects), boxmin=host3(boxmin), boxmax=host3(boxmax),
        stream=stream_ptr, flags=int(flags),
        timings=C.pointer(times) if times is not None else None)
    # the library works on the CURRENT device (streams, events, the pinned read-back slot): make it the tensors'
    if dev is not None and dev.type == "cuda":
        with torch.cuda.device(dev):
            rc = L.gsr_forward_ex(C.byref(a))
    else:
        rc = L.gsr_forward_ex(C.byref(a))
    _lib.check(rc)
    if timings:
        return rc, times.as_dict()
    return rc


class Rasterizer:
    """CudaRasterizer::Rasterizer — forward only."""

    @staticmethod
    def forward(geometryBuffer, binningBuffer, imageBuffer, P, D, M, background, width, height, means3D, shs,
                colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
                cam_pos, tan_fovx, tan_fovy, prefiltered, out_color, radii=None, rects=None, boxmin=None,
                boxmax=None, *, stream=None, flags=0, timings=False):
        """Same 29 arguments, same order, same meaning as the reference call site
        (GSGaussians.cpp:179-206).  The three buffer arguments are callables
        `f(nbytes) -> device address` (std::function<char*(size_t)>).  Returns num_rendered
        (and the per-stage times dict when `timings=True`)."""
        return _forward(geometryBuffer, binningBuffer, imageBuffer, P, D, M, background, width, height, means3D,
                        shs, colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
                        projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered, out_color, radii, rects, boxmin,
                        boxmax, stream, flags & ~FLAG_GSRAST_COMPAT, 3, 3, timings)


def gscuda_forward(geometryBuffer, binningBuffer, imageBuffer, numGaussians, shDims, M, background, width, height,
                   means3D, shs, colorsPrecomp, opacities, scales, scaleModifier, rotations, cov3DPrecomp,
                   viewMatrix, projMatrix, camPos, tanFOVx, tanFOVy, prefiltered, outColor, radii=None, rects=None,
                   boxMin=None, boxMax=None, *, stream=None, flags=0, timings=False):
    """gscuda::forward (GSCuda.cuh:103-126) with its in-tree semantics and vec4-strided
    means3D / scales, as GSGaussians::draw calls it today."""
    return _forward(geometryBuffer, binningBuffer, imageBuffer, numGaussians, shDims, M, background, width, height,
                    means3D, shs, colorsPrecomp, opacities, scales, scaleModifier, rotations, cov3DPrecomp,
                    viewMatrix, projMatrix, camPos, tanFOVx, tanFOVy, prefiltered, outColor, radii, rects, boxMin,
                    boxMax, stream, flags | FLAG_GSRAST_COMPAT, 4, 4, timings)


def sort_pairs(keys: torch.Tensor, values: torch.Tensor, end_bit: int):
    """Stable (u64 key, u32 value) radix sort over bits [0,end_bit) — the in-house replacement
    of cub::DeviceRadixSort::SortPairs (GSCuda.cu:794-797).  int64 / int32 tensors carry the
    unsigned bit patterns."""
    L = _lib.lib()
    n = keys.numel()
    k_in = keys.contiguous().clone()
    v_in = values.contiguous().clone()
    k_out = torch.empty_like(k_in)
    v_out = torch.empty_like(v_in)
    temp = torch.empty(int(L.gsr_sort_pairs_temp_bytes(n)), dtype=torch.uint8, device=keys.device)
    _lib.check(L.gsr_sort_pairs(k_in.data_ptr(), v_in.data_ptr(), k_out.data_ptr(), v_out.data_ptr(), n, int(end_bit),
                                temp.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return k_out, v_out


def identify_tile_ranges(sorted_keys: torch.Tensor, num_tiles: int, compat: bool = False) -> torch.Tensor:
    """identifyTileRanges (GSCuda.cu:504-538) incl. the memset of GSCuda.cu:800."""
    ranges = torch.empty((num_tiles, 2), dtype=torch.int32, device=sorted_keys.device)
    _lib.check(_lib.lib().gsr_identify_tile_ranges(sorted_keys.data_ptr(), sorted_keys.numel(), ranges.data_ptr(),
                                                   num_tiles, FLAG_GSRAST_COMPAT if compat else 0,
                                                   torch.cuda.current_stream().cuda_stream))
    return ranges


class GSGaussians:
    """The splat draw strategy (apps/gsrast/GSGaussians.{hpp,cpp}) minus GL: holds the
    device-resident scene, the three grow-only scratch buffers and the output buffer, and
    renders one camera per draw().  `compat=True` reproduces the viewer as it is today
    (gscuda::forward on vec4 / raw-PLY buffers); the default is the CudaRasterizer contract."""

    def __init__(self, width: int, height: int, device="cuda", compat: bool = False, use_rects: bool = True,
                 flags: int = 0):
        self.width, self.height = int(width), int(height)
        self.device = torch.device(device)
        self.compat = bool(compat)
        self.flags = int(flags)
        self.num_gaussians = 0
        self._geom = ResizeFunctional(self.device)
        self._binning = ResizeFunctional(self.device)
        self._img = ResizeFunctional(self.device)
        # _interopTex (GSGaussians.cpp:49-50): planar float[3][H][W]
        self.out_color = torch.zeros((NUM_CHANNELS, self.height, self.width), dtype=torch.float32, device=self.device)
        self.background = torch.zeros(3, dtype=torch.float32, device=self.device)  # GSGaussians.cpp:148
        self._cam = torch.zeros(36, dtype=torch.float32, device=self.device)
        self._cam_host = torch.zeros(36, dtype=torch.float32).pin_memory() if self.device.type == "cuda" else None
        self.use_rects = use_rects
        self.rects = None
        self.sh_degree, self.max_coeffs = 3, 16
        self.last_num_rendered = 0

    def configure_from_splat_data(self, scene) -> bool:
        """configureFromSplatData (GSGaussians.cpp:109-153): upload the SoA attributes."""
        if scene is None or scene.P == 0:
            return False
        dev = self.device
        self.num_gaussians = scene.P
        if self.compat:
            means4, scales4, rot, opac, shs_raw = scene.gsrast_layout()
            self.positions = torch.from_numpy(means4).to(dev)
            self.scales = torch.from_numpy(scales4).to(dev)
            self.shs = torch.from_numpy(shs_raw).to(dev)
            self.rotations = torch.from_numpy(rot).to(dev)
            self.opacities = torch.from_numpy(opac).to(dev)
            self.colors_precomp = None
        else:
            self.positions = torch.from_numpy(scene.means3D).to(dev)
            self.scales = torch.from_numpy(scene.scales).to(dev)
            self.shs = torch.from_numpy(scene.shs).to(dev) if scene.shs is not None else None
            self.rotations = torch.from_numpy(scene.rotations).to(dev)
            self.opacities = torch.from_numpy(scene.opacities).to(dev)
            self.colors_precomp = (torch.from_numpy(scene.colors_precomp).to(dev)
                                   if scene.colors_precomp is not None else None)
            self.sh_degree, self.max_coeffs = scene.sh_degree, scene.max_coeffs
        self.rects = (torch.zeros((scene.P, 2), dtype=torch.int32, device=dev)  # GSGaussians.cpp:137
                      if self.use_rects else None)
        return True

    def draw(self, camera, timings: bool = False):
        """draw() (GSGaussians.cpp:155-212): upload view / proj / camPos, run forward."""
        self._cam_host.copy_(torch.from_numpy(camera.packed()))
        self._cam.copy_(self._cam_host, non_blocking=True)
        view, proj, cam_pos = self._cam[0:16], self._cam[16:32], self._cam[32:35]
        if self.compat:
            out = gscuda_forward(self._geom, self._binning, self._img, self.num_gaussians, 3, 16, self.background,
                                 self.width, self.height, self.positions, self.shs, None, self.opacities, self.scales,
                                 1.0, self.rotations, None, view, proj, cam_pos, camera.tan_fovx, camera.tan_fovy,
                                 False, self.out_color, None, self.rects, None, None, flags=self.flags,
                                 timings=timings)
        else:
            out = Rasterizer.forward(self._geom, self._binning, self._img, self.num_gaussians, self.sh_degree,
                                     self.max_coeffs, self.background, self.width, self.height, self.positions,
                                     self.shs, self.colors_precomp, self.opacities, self.scales, 1.0, self.rotations,
                                     None, view, proj, cam_pos, camera.tan_fovx, camera.tan_fovy, False,
                                     self.out_color, None, self.rects, None, None, flags=self.flags, timings=timings)
        self.last_num_rendered = out[0] if timings else out
        return out

    def map_geometry_state(self) -> dict:
        """mapGeometryState (GSGaussians.cpp:214-219)."""
        return GeometryState.from_chunk(self._geom.buf, self.num_gaussians)

    def map_image_state(self) -> dict:
        return ImageState.from_chunk(self._img.buf, self.width, self.height)

    def map_binning_state(self) -> dict:
        return BinningState.from_chunk(self._binning.buf, self.last_num_rendered)
