"""ctypes binding of libgsrast_b200.so (include/gsrast_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises at first use.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GSRAST_B200_LIB selects an alternative build of the same library (A/B runs of kernel variants)
LIB_PATH = os.environ.get("GSRAST_B200_LIB") or os.path.join(_HERE, "libgsrast_b200.so")

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)

FLAG_GSRAST_COMPAT = 0x1
FLAG_BLEND_SIMPLE = 0x2
FLAG_LEAN_STATE = 0x8
FLAG_RADIX_BINNING = 0x4
FLAG_BLEND_COUNT = 0x10
FLAG_KEEP_STATE = 0x20
FLAG_BLEND_ONE_PIXEL = 0x40
BLEND_COUNTERS = ("tile_rounds", "warp_rounds", "candidates_listed", "warp_trips", "pairs_live", "pairs_passed",
                  "pairs_blended", "splats_staged")

ERR_INVALID_ARG = -1000
ERR_ALLOC_FAILED = -1001
ERR_TOO_MANY_PAIRS = -1002
ERR_SORT_STALLED = -1003
ERR_PLY_OPEN = -1004
ERR_PLY_FORMAT = -1005
ERR_PLY_TRUNCATED = -1006


class StageTimes(C.Structure):
    _fields_ = [(n, C.c_float) for n in
                ("preprocess_ms", "scan_ms", "duplicate_ms", "sort_ms", "ranges_ms", "blend_ms", "total_ms")] + \
               [("num_rendered", C.c_int), ("sort_passes", C.c_int), ("kernel_launches", C.c_int),
                ("sort_hist_ms", C.c_float), ("sort_pass_ms", C.c_float * 8),
                ("depth_sort_ms", C.c_float), ("depth_passes", C.c_int),
                ("expand_ms", C.c_float), ("num_coarse", C.c_int), ("binning_mode", C.c_int),
                ("expand_count_ms", C.c_float), ("expand_fill_ms", C.c_float),
                ("blend_counters", C.c_ulonglong * 8)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["sort_pass_ms"] = [float(x) for x in self.sort_pass_ms][: max(self.sort_passes, 0)]
        d["blend_counters"] = {n: int(self.blend_counters[i]) for i, n in enumerate(BLEND_COUNTERS)}
        return d


class ForwardArgs(C.Structure):
    _fields_ = [
        ("geometry_alloc", ALLOC_FN), ("geometry_user", C.c_void_p),
        ("binning_alloc", ALLOC_FN), ("binning_user", C.c_void_p),
        ("image_alloc", ALLOC_FN), ("image_user", C.c_void_p),
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int),
        ("background", C.c_void_p),
        ("width", C.c_int), ("height", C.c_int),
        ("means3D", C.c_void_p), ("means_stride", C.c_int),
        ("shs", C.c_void_p),
        ("colors_precomp", C.c_void_p),
        ("opacities", C.c_void_p),
        ("scales", C.c_void_p), ("scales_stride", C.c_int),
        ("scale_modifier", C.c_float),
        ("rotations", C.c_void_p),
        ("cov3D_precomp", C.c_void_p),
        ("viewmatrix", C.c_void_p),
        ("projmatrix", C.c_void_p),
        ("cam_pos", C.c_void_p),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("prefiltered", C.c_int),
        ("out_color", C.c_void_p),
        ("radii", C.c_void_p),
        ("rects", C.c_void_p),
        ("boxmin", C.c_void_p),
        ("boxmax", C.c_void_p),
        ("stream", C.c_void_p),
        ("flags", C.c_uint),
        ("timings", C.POINTER(StageTimes)),
    ]


class GeometryState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("depths", "clamped", "internal_radii", "means2D", "cov3D", "conic_opacity", "rgb", "tiles_touched",
                 "point_offsets", "block_sums")] + [("scan_size", C.c_size_t)] + \
               [("depth_keys", C.c_void_p), ("tile_rects", C.c_void_p), ("depth_sort_keys", C.c_void_p * 2), ("depth_sort_ids", C.c_void_p * 2),
                ("depth_sort_space", C.c_void_p), ("depth_sort_size", C.c_size_t),
                ("sorted_rects", C.c_void_p), ("sorted_block_sums", C.c_void_p), ("coarse_block_sums", C.c_void_p)]


class ImageState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ranges", "n_contrib", "accum_alpha", "tile_order", "blend_counters")]


class BinningState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("point_list_keys_unsorted", "point_list_keys", "point_list_unsorted", "point_list",
                 "list_sorting_space")] + [("sorting_size", C.c_size_t)]


# every symbol include/gsrast_b200.h declares
EXPORTS = (
    "gsr_forward", "gsr_forward_gscuda", "gsr_forward_ex",
    "gsr_geometry_state_required", "gsr_image_state_required", "gsr_binning_state_required",
    "gsr_geometry_state_map", "gsr_image_state_map", "gsr_binning_state_map",
    "gsr_get_higher_msb", "gsr_sort_pairs_temp_bytes", "gsr_sort_pairs", "gsr_identify_tile_ranges",
    "gsr_error_string", "gsr_version",
    "gsr_renderer_create", "gsr_renderer_destroy", "gsr_renderer_render", "gsr_renderer_render_host",
    "gsr_renderer_last_times", "gsr_repack_gsrast_scene", "gsr_ply_count", "gsr_ply_load",
    "gsr_renderer_render_host_u8", "gsr_frames_to_u8", "gsr_renderer_map_geometry_state", "gsr_renderer_num_lanes",
    "gsr_pinned_alloc", "gsr_pinned_free",
)

_lib = None


def lib():
    """Load the CUDA library.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "gsrast_b200: %s not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C gsrast_b200/csrc). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    _fwd29 = [ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p, C.c_int, C.c_int, C.c_int,
              C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p,
              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsr_forward.restype = C.c_int
    L.gsr_forward.argtypes = _fwd29
    L.gsr_forward_gscuda.restype = C.c_int
    L.gsr_forward_gscuda.argtypes = _fwd29
    L.gsr_forward_ex.restype = C.c_int
    L.gsr_forward_ex.argtypes = [C.POINTER(ForwardArgs)]
    L.gsr_geometry_state_required.restype = C.c_size_t
    L.gsr_geometry_state_required.argtypes = [C.c_int]
    L.gsr_image_state_required.restype = C.c_size_t
    L.gsr_image_state_required.argtypes = [C.c_int, C.c_int]
    L.gsr_binning_state_required.restype = C.c_size_t
    L.gsr_binning_state_required.argtypes = [C.c_size_t]
    L.gsr_geometry_state_map.restype = C.c_size_t
    L.gsr_geometry_state_map.argtypes = [C.c_void_p, C.c_int, C.POINTER(GeometryState)]
    L.gsr_image_state_map.restype = C.c_size_t
    L.gsr_image_state_map.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(ImageState)]
    L.gsr_binning_state_map.restype = C.c_size_t
    L.gsr_binning_state_map.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(BinningState)]
    L.gsr_get_higher_msb.restype = C.c_uint32
    L.gsr_get_higher_msb.argtypes = [C.c_uint32]
    L.gsr_sort_pairs_temp_bytes.restype = C.c_size_t
    L.gsr_sort_pairs_temp_bytes.argtypes = [C.c_size_t]
    L.gsr_sort_pairs.restype = C.c_int
    L.gsr_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                 C.c_void_p]
    L.gsr_identify_tile_ranges.restype = C.c_int
    L.gsr_identify_tile_ranges.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_uint, C.c_void_p]
    L.gsr_error_string.restype = C.c_char_p
    L.gsr_error_string.argtypes = [C.c_int]
    L.gsr_version.restype = C.c_int
    # view-batch renderer (views.cu)
    L.gsr_renderer_create.restype = C.c_void_p
    L.gsr_renderer_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p,
                                      C.c_uint]
    L.gsr_renderer_destroy.restype = None
    L.gsr_renderer_destroy.argtypes = [C.c_void_p]
    L.gsr_renderer_render.restype = C.c_int
    L.gsr_renderer_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
    L.gsr_renderer_render_host.restype = C.c_int
    L.gsr_renderer_render_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                           C.c_void_p]
    L.gsr_renderer_last_times.restype = C.c_int
    L.gsr_renderer_last_times.argtypes = [C.c_void_p, C.POINTER(StageTimes)]
    L.gsr_renderer_render_host_u8.restype = C.c_int
    L.gsr_renderer_render_host_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                              C.c_void_p]
    L.gsr_frames_to_u8.restype = C.c_int
    L.gsr_frames_to_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.gsr_renderer_map_geometry_state.restype = C.c_int
    L.gsr_renderer_map_geometry_state.argtypes = [C.c_void_p, C.c_int, C.POINTER(GeometryState)]
    L.gsr_renderer_num_lanes.restype = C.c_int
    L.gsr_renderer_num_lanes.argtypes = []
    L.gsr_pinned_alloc.restype = C.c_void_p
    L.gsr_pinned_alloc.argtypes = [C.c_size_t, C.c_uint]
    L.gsr_pinned_free.restype = None
    L.gsr_pinned_free.argtypes = [C.c_void_p]
    L.gsr_ply_count.restype = C.c_int
    L.gsr_ply_count.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    L.gsr_ply_load.restype = C.c_int
    L.gsr_ply_load.argtypes = [C.c_char_p, C.c_int] + [C.c_void_p] * 7
    L.gsr_repack_gsrast_scene.restype = C.c_int
    L.gsr_repack_gsrast_scene.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
    _lib = L
    return L


def error_string(code: int) -> str:
    return lib().gsr_error_string(code).decode()


def check(code: int) -> int:
    if code < 0:
        raise RuntimeError("gsrast_b200 error %d: %s" % (code, error_string(code)))
    return code
