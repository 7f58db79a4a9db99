/* gsrast_b200.h — C ABI of libgsrast_b200.so: the B200 (sm_100a) splat forward rasterizer
 * that replaces GSRast's splat draw path.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the
 * 42yeah/GSRast tree, commit bee842bd).  Plain pointers and sizes only: no C++ types, no
 * torch types.  Unless stated otherwise every data pointer is a DEVICE pointer, exactly as
 * in the reference (apps/gsrast/GSGaussians.cpp:179-206).
 *
 * Error convention: functions returning int return >= 0 on success and -(int)cudaError_t or
 * one of the GSR_ERR_* codes (<= -1000) on failure; nothing throws across this boundary and
 * the sticky CUDA error stays visible to cudaPeekAtLastError() so GSRast's CHECK_CUDA_ERROR
 * macro (apps/gsrast/CudaBuffer.hpp:8-12) keeps working.
 */
#ifndef GSRAST_B200_H
#define GSRAST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_VERSION 200 /* round 2 */

/* Scratch allocator callback: the C form of the reference's std::function<char*(size_t)>
 * (apps/gsrast/gscuda/GSCuda.cuh:103-105; resizeFunctional, GSGaussians.cpp:27-42).
 * Must return a device pointer valid for `bytes`; called synchronously on the caller's
 * thread, at most once per forward call, in the order geometry -> image -> binning. */
typedef char* (*gsr_alloc_fn)(size_t bytes, void* user);

/* flags for gsr_forward_ex */
#define GSR_FLAG_GSRAST_COMPAT 0x1u /* reproduce in-tree gscuda::forward (GSCuda.cu:695-811) instead of the
                                       CudaRasterizer contract: NDC cull, NDC-z keys, DC-only colour,
                                       T<0.001 termination, y-extent w/o sqrt, R==1 range quirk, stale image
                                       when nothing is rendered */
#define GSR_FLAG_BLEND_SIMPLE 0x2u  /* use the plain per-tile blend kernel (no sub-tile culling); for A/B tests */
#define GSR_FLAG_RADIX_BINNING 0x4u /* sort the tile half of the keys with radix passes over all num_rendered pairs
                                       (the path CUB takes in the reference, GSCuda.cu:794-797) instead of the default
                                       bin expansion (sort per 8x8-tile bin, expand each bin into its tiles); same
                                       output bit for bit; also taken automatically for grids of more than 4096 bins */
#define GSR_FLAG_LEAN_STATE 0x8u    /* do not materialise state nothing in this forward pass reads back: the geometry fields
                                       cov3D[6P], clamped[3P], tiles_touched[P], point_offsets[P] and — when the call
                                       passes no `radii` buffer — internal_radii[P] (39 B/Gaussian of
                                       stores + the 4 B/Gaussian re-read of the index-order scan; the reference keeps them
                                       for its Inspector and for duplicateWithKeys, apps/gsrast/Inspector.cpp:174-188,
                                       GSCuda.cu:445) and, on the default bin-expansion path, the sorted 64-bit keys
                                       point_list_keys[R] (8 of the 12 B/pair the tile sort writes; the reference re-reads
                                       them once, in identifyTileRanges GSCuda.cu:504-538 — here the ranges come out of the
                                       expansion's scan).  point_list, ranges and every output are the same bits.
                                       gsr_renderer_* sets it for its private per-lane scratch; gsr_forward* never does
                                       on its own */

#define GSR_FLAG_BLEND_COUNT 0x10u  /* with `timings`: run the counting instantiation of the blend kernel (same arithmetic)
                                       and return its work counters in gsr_stage_times.blend_counters — the unit
                                       SURVEY 8(d) rates the blend in.  Slower; for reporting, never on a timed path */
#define GSR_FLAG_BLEND_ONE_PIXEL 0x40u /* use the one-pixel-per-thread culled blend kernel instead of the default
                                          two-pixels-per-thread kernel on the packed FP32 pipe; for A/B tests */
#define GSR_FLAG_KEEP_STATE 0x20u   /* gsr_renderer_create only: keep every geometry-state field materialised (no
                                       GSR_FLAG_LEAN_STATE) so gsr_renderer_map_geometry_state serves the reference's
                                       Inspector panel (apps/gsrast/Inspector.cpp:174-188) */

#define GSR_ERR_INVALID_ARG (-1000)
#define GSR_ERR_ALLOC_FAILED (-1001)   /* an allocator callback returned NULL */
#define GSR_ERR_TOO_MANY_PAIRS (-1002) /* num_rendered >= 2^30: beyond the sort's look-back counters (detected on the
                                          device with a 64-bit sum, so a wrapped 32-bit total cannot slip through) */
#define GSR_ERR_SORT_STALLED (-1003)   /* a look-back watchdog tripped (never expected).  Asynchronous like a CUDA
                                          error: reported by the call that notices it — the same call when it
                                          synchronises (timings, *_render_host*), otherwise the next call of the same
                                          host thread / renderer.  The frame that stalled is invalid. */
#define GSR_ERR_PLY_OPEN (-1004)       /* the file cannot be opened ("Bad PLY reader", SplatData.cpp:120-124) */
#define GSR_ERR_PLY_FORMAT (-1005)     /* no vertex count on the third header line / no end_header */
#define GSR_ERR_PLY_TRUNCATED (-1006)  /* fewer than P records in the body ("Reader is EOF?", SplatData.cpp:146-152) */

/* Replaces CudaRasterizer::Rasterizer::forward (deps/diff-gaussian-rasterization, absent
 * submodule; call shape at apps/gsrast/GSGaussians.cpp:179-206) — identical argument order
 * and meaning, std::function allocators flattened to (fn, user) pairs, plus a trailing
 * cudaStream_t (void*; NULL = legacy default stream, which keeps the reference's
 * post-call cudaDeviceSynchronize semantics intact).
 *
 *   means3D float[P][3], shs float[P][M][3], colors_precomp float[P][3] or NULL,
 *   opacities float[P], scales float[P][3], rotations float[P][4] (r,x,y,z),
 *   cov3D_precomp float[P][6] or NULL, viewmatrix/projmatrix float[16] column-major (device),
 *   cam_pos float[3] (device), background float[3] (device),
 *   out_color float[3][H][W] planar (device), radii int[P] or NULL, rects int[P][2] or NULL,
 *   boxmin/boxmax HOST float[3] or NULL (SIBR bounding-box cull).
 * Returns num_rendered (tile-Gaussian pairs) or a negative error. */
int gsr_forward(gsr_alloc_fn geometry_alloc, void* geometry_user, gsr_alloc_fn binning_alloc, void* binning_user,
                gsr_alloc_fn image_alloc, void* image_user, int P, int D, int M, const float* background, int width,
                int height, const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, float scale_modifier, const float* rotations,
                const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                float tan_fovx, float tan_fovy, int prefiltered, float* out_color, int* radii, int* rects,
                const float* boxmin, const float* boxmax, void* stream);

/* Replaces gscuda::forward (apps/gsrast/gscuda/GSCuda.cuh:103-126, GSCuda.cu:695-811) as
 * GSRast calls it today: same argument list, but means3D / scales are the viewer's
 * vec4-strided buffers (GSGaussians.cpp:121-125) and `shs` is the raw 48-float PLY block;
 * semantics are the in-tree ones (GSR_FLAG_GSRAST_COMPAT).  Returns num_rendered. */
int gsr_forward_gscuda(gsr_alloc_fn geometry_alloc, void* geometry_user, gsr_alloc_fn binning_alloc,
                       void* binning_user, gsr_alloc_fn image_alloc, void* image_user, int P, int D, int M,
                       const float* background, int width, int height, const float* means3D, const float* shs,
                       const float* colors_precomp, const float* opacities, const float* scales,
                       float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                       float tan_fovy, int prefiltered, float* out_color, int* radii, int* rects,
                       const float* boxmin, const float* boxmax, void* stream);

/* Per-stage device times of one forward call, filled when `timings` is non-NULL in
 * gsr_forward_ex (CUDA events on the call's stream; forces a stream sync at the end). */
typedef struct gsr_stage_times {
    float preprocess_ms, scan_ms, duplicate_ms, sort_ms, ranges_ms, blend_ms, total_ms;
    int num_rendered;
    int sort_passes; /* depth passes + tile passes; sort_ms covers the tile passes only */
    int kernel_launches; /* kernels this library launched inside the call */
    float sort_hist_ms;     /* digit-histogram scans of both sort halves (+ the depth-key histogram) */
    float sort_pass_ms[8];  /* each onesweep digit pass: depth_passes passes over P Gaussians first,
                               then the tile-digit passes over the num_rendered pairs */
    float depth_sort_ms;    /* depth half of the LSD sort (P records), overlaps the num_rendered round trip */
    int depth_passes;
    /* bin expansion (default binning mode): sort_ms then covers the bin-digit passes over the
     * num_coarse (Gaussian, 8x8-tile bin) records, expand_ms the count / scan / fill kernels that turn
     * each bin's depth-ordered list into the sorted per-tile lists and the tile ranges (ranges_ms = 0). */
    float expand_ms;
    int num_coarse;
    int binning_mode;       /* 0 = bin expansion, 1 = radix passes over the pairs */
    float expand_count_ms;  /* expand_count_kernel alone (one launch) */
    float expand_fill_ms;   /* expand_fill_kernel alone (one launch): the pass that writes the 12 B/pair result */
    /* GSR_FLAG_BLEND_COUNT: [0] tile-rounds staged (128 splats each), [1] warp-rounds that walked a candidate list,
     * [2] candidates listed (warp x splat), [3] candidate trips executed (warp x splat; x32 = evaluated pixel-splat
     * pairs, the unit of SURVEY 8(d)), [4] of those pairs, the ones whose pixel was still live, [5] pairs that passed
     * the power / alpha tests, [6] pairs blended, [7] splats staged */
    unsigned long long blend_counters[8];
} gsr_stage_times;
#define GSR_BLEND_COUNTERS 8

/* Same call with explicit strides / flags (superset of the two above). */
typedef struct gsr_forward_args {
    gsr_alloc_fn geometry_alloc; void* geometry_user;
    gsr_alloc_fn binning_alloc;  void* binning_user;
    gsr_alloc_fn image_alloc;    void* image_user;
    int P, D, M;
    const float* background;
    int width, height;
    const float* means3D;  int means_stride;  /* floats per record: 3 (contract) or 4 (GSRast vec4) */
    const float* shs;
    const float* colors_precomp;
    const float* opacities;
    const float* scales;   int scales_stride; /* 3 or 4 */
    float scale_modifier;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* cam_pos;
    float tan_fovx, tan_fovy;
    int prefiltered;
    float* out_color;
    int* radii;
    int* rects;
    const float* boxmin; /* host */
    const float* boxmax; /* host */
    void* stream;
    unsigned flags;
    gsr_stage_times* timings; /* optional, host */
} gsr_forward_args;

int gsr_forward_ex(const gsr_forward_args* args);

/* ---- scratch-buffer layout as API ------------------------------------------------------
 * The reference re-derives field pointers from the raw chunks with
 * gscuda::gs::{Geometry,Image,Binning}State::fromChunk (apps/gsrast/gscuda/AuxBuffer.cu:44-89,
 * AuxBuffer.cuh:36-78); the Inspector reads nine geometry fields that way
 * (GSGaussians.cpp:214-219, Inspector.cpp:174-188).  These are the equivalents. */
typedef struct gsr_geometry_state {
    float* depths;            /* [P]; entries of Gaussians that emit no pair hold the bit pattern 0xffffffff (the
                                 reference leaves them stale, GSCuda.cu:356-359 returns before :369) */
    unsigned char* clamped;   /* [3P] */
    int* internal_radii;      /* [P] */
    float* means2D;           /* [P][2] */
    float* cov3D;             /* [P][6] */
    float* conic_opacity;     /* [P][4] */
    float* rgb;               /* [P][3] */
    uint32_t* tiles_touched;  /* [P] */
    uint32_t* point_offsets;  /* [P] inclusive scan of tiles_touched */
    uint32_t* block_sums;     /* scan scratch (replaces the CUB temp storage) */
    size_t scan_size;
    /* depth half of the radix sort, run per Gaussian before duplication */
    uint32_t* depth_keys;        /* = (uint32_t*)depths: the low half of the sort key is the depth's bit pattern, so the
                                    depth sort reads that array directly (no second 4 B/Gaussian store) */
    uint32_t* tile_rects;        /* [P][2] miny<<16|minx, height<<16|width of the tile rect (0 = emits nothing) */
    uint32_t* depth_sort_keys[2];/* [P] ping-pong; [1] ends up holding the sorted depth keys */
    uint32_t* depth_sort_ids[2]; /* [P] ping-pong; [1] ends up holding the Gaussian ids in depth order */
    char* depth_sort_space;      /* histograms + look-back state of that sort */
    size_t depth_sort_size;
    uint32_t* sorted_rects;      /* [P][2] tile_rects gathered into depth order */
    uint32_t* sorted_block_sums; /* scan scratch of the pair counts in depth order (scan_size bytes) */
    uint32_t* coarse_block_sums; /* per preprocess block: (Gaussian, 8x8-tile bin) records it will emit */
} gsr_geometry_state;

typedef struct gsr_image_state {
    uint32_t* ranges;     /* [tiles][2]  (start, end) into the sorted lists */
    uint32_t* n_contrib;  /* [W*H] */
    float* accum_alpha;   /* [W*H] final transmittance */
    uint32_t* tile_order; /* [tiles] scratch: per-tile pair counts, then their exclusive scan (bin expansion) */
    unsigned long long* blend_counters; /* [GSR_BLEND_COUNTERS] work counters of the blend (GSR_FLAG_BLEND_COUNT) */
} gsr_image_state;

typedef struct gsr_binning_state {
    uint64_t* point_list_keys_unsorted; /* [R] x 8 B: two ping-pong arrays of 32-bit tile keys */
    uint64_t* point_list_keys;          /* [R] sorted (tile << 32 | depth bits) */
    uint32_t* point_list_unsorted;      /* [R] Gaussian ids in depth-ordered emission order */
    uint32_t* point_list;               /* [R] sorted Gaussian ids */
    char* list_sorting_space;           /* second id ping-pong array [R] + histograms + look-back state
                                           + chunk tables of the bin expansion */
    size_t sorting_size;
} gsr_binning_state;

/* required<T>(n) (AuxBuffer.cuh:8-14) */
size_t gsr_geometry_state_required(int P);
size_t gsr_image_state_required(int width, int height);
size_t gsr_binning_state_required(size_t num_rendered);
/* T::fromChunk(chunk, n) (AuxBuffer.cu:44-89); return the bytes consumed */
size_t gsr_geometry_state_map(char* chunk, int P, gsr_geometry_state* out);
size_t gsr_image_state_map(char* chunk, int width, int height, gsr_image_state* out);
size_t gsr_binning_state_map(char* chunk, size_t num_rendered, gsr_binning_state* out);

/* getHigherMsb (GSCuda.cu:481-502) */
uint32_t gsr_get_higher_msb(uint32_t n);

/* ---- building blocks, exported for parity tests and reuse ------------------------------ */
/* Stable LSD radix sort of (u64 key, u32 value) pairs over bits [0, end_bit): the in-house
 * replacement of cub::DeviceRadixSort::SortPairs (GSCuda.cu:794-797).  `temp` must hold
 * gsr_sort_pairs_temp_bytes(n) bytes.  Sorted output lands in keys_out / vals_out;
 * keys_in / vals_in are clobbered. */
size_t gsr_sort_pairs_temp_bytes(size_t n);
int gsr_sort_pairs(uint64_t* keys_in, uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out, size_t n,
                   int end_bit, char* temp, void* stream);
/* identifyTileRanges (GSCuda.cu:504-538, 800-801): zeroes ranges[num_tiles][2] then fills. */
int gsr_identify_tile_ranges(const uint64_t* sorted_keys, size_t n, uint32_t* ranges, int num_tiles, unsigned flags,
                             void* stream);

/* ---- resident-scene renderer for streams / batches of views -----------------------------
 * The C form of the reference's GSGaussians object (apps/gsrast/GSGaussians.{hpp,cpp}):
 * create = constructor + configureFromSplatData (GSGaussians.cpp:44-153; the scene arrays stay
 * caller-owned DEVICE buffers and must outlive the renderer), render = draw() (:155-212) for
 * n_views cameras.  It owns the grow-only scratch chunks (resizeFunctional, :27-42) and two
 * internal streams that alternate views so the num_rendered round trip of one view hides
 * behind the sort/blend of the previous one.  With GSR_FLAG_GSRAST_COMPAT the scene arrays are
 * the viewer's vec4 / raw-PLY buffers.  `cameras` is a HOST array of 36 floats per view:
 * viewmatrix[16], projmatrix[16], cam_pos[3], pad — what draw() uploads per frame. */
void* gsr_renderer_create(int P, int D, int M, const float* means3D, const float* shs, const float* colors_precomp,
                          const float* opacities, const float* scales, const float* rotations,
                          const float* background, float scale_modifier, int width, int height, void* stream,
                          unsigned flags);
void gsr_renderer_destroy(void* renderer);
/* out_color: DEVICE float[n_views][3][H][W]; num_rendered: HOST int[n_views] or NULL;
 * timings: gsr_stage_times* of the last view (serialises the views) or NULL.
 * Returns the summed num_rendered or a negative error. */
int gsr_renderer_render(void* renderer, const float* cameras, int n_views, float tan_fovx, float tan_fovy,
                        float* out_color, int* num_rendered, void* timings);
/* Frames are delivered to HOST memory out_color[n_views][3][H][W] (pinned => the copy of view
 * k overlaps the render of view k+1); returns once every frame has landed. */
int gsr_renderer_render_host(void* renderer, const float* cameras, int n_views, float tan_fovx, float tan_fovy,
                             float* out_color_host, int* num_rendered);
/* 8-bit delivery: frames are quantised on the device — round(clamp(x, 0, 1) * 255), same planar [3][H][W] layout, what
 * the viewer's RGBA8 framebuffer / screenshot path holds (apps/gsrast/Inspector.cpp:222-257) — and a quarter of the
 * bytes cross PCIe.  out_color_host: HOST uint8[n_views][3][H][W].  gsr_frames_to_u8 is the bare conversion of any
 * device float array (n_values elements; `frames` 16-byte aligned). */
int gsr_renderer_render_host_u8(void* renderer, const float* cameras, int n_views, float tan_fovx, float tan_fovy,
                                unsigned char* out_color_host, int* num_rendered);
int gsr_frames_to_u8(const float* frames, unsigned char* out, size_t n_values, void* stream);
/* Page-locked HOST landing buffers for the *_render_host* calls (the reference keeps its frame in a GL buffer,
 * apps/gsrast/CudaBuffer.cpp:21-32; a host-side consumer needs pinned memory for the copies to overlap rendering).
 * GSR_PINNED_WRITE_COMBINED: not snooped during the device's writes and slow to READ from the CPU — for frames that
 * are forwarded (network, encoder DMA), not inspected.  Returns NULL on failure. */
#define GSR_PINNED_WRITE_COMBINED 0x1u
#define GSR_PINNED_PORTABLE 0x2u
void* gsr_pinned_alloc(size_t bytes, unsigned flags);
void gsr_pinned_free(void* ptr);
int gsr_renderer_last_times(void* renderer, gsr_stage_times* out);
/* The renderer's form of GSGaussians::mapGeometryState (apps/gsrast/GSGaussians.cpp:214-219): field pointers into the
 * private geometry chunk of `lane` (view v of a render call ran on lane v % gsr_renderer_num_lanes(); with `timings`
 * every view runs on lane 0).  Valid until the next render call on that renderer; synchronise the renderer's stream
 * before reading.  The renderer must have been created with GSR_FLAG_KEEP_STATE for cov3D / clamped / tiles_touched /
 * point_offsets to hold data (otherwise those four are stale scratch).  Returns 0, or GSR_ERR_INVALID_ARG when the
 * lane has not rendered yet. */
int gsr_renderer_map_geometry_state(void* renderer, int lane, gsr_geometry_state* out);
int gsr_renderer_num_lanes(void);

/* One-shot device repack of the buffers GSRast's viewer uploads (vec4 means / scales,
 * apps/gsrast/GSGaussians.cpp:121-125; SH in raw PLY order, SplatData.hpp:17-25 — f_dc[3]
 * then f_rest[c*15+k-1]) into the contract layout (float3, sh[k][c]).  Any pair may be NULL. */
int gsr_repack_gsrast_scene(int P, const float* means4, const float* scales4, const float* shs_raw, float* means3,
                            float* scales3, float* shs, void* stream);

/* Scene staging from the .ply the viewer loads (HOST code, no GPU needed): replaces SplatData::loadFromSplatsPly +
 * the activation loop (apps/gsrast/SplatData.cpp:114-156, :48-58; record layout SplatData.hpp:17-25) and writes the
 * rasterizer's contract layout directly — means3D[P][3], scales[P][3] = exp, rotations[P][4] = normalised (w,x,y,z),
 * opacities[P] = sigmoid, shs[P][16][3] re-interleaved from the file's f_dc[3], f_rest[c*15+k-1] order.
 * gsr_ply_count reads the header only.  gsr_ply_load fills caller-owned HOST arrays sized for `capacity` Gaussians
 * (any output may be NULL), optionally the bounding box (min[3], max[3]) and the mean position the viewer uses to
 * place its camera (GSRastWindow.cpp:26-36), and returns the number of Gaussians read or a GSR_ERR_PLY_* code. */
int gsr_ply_count(const char* path, int* num_gaussians);
int gsr_ply_load(const char* path, int capacity, float* means3D, float* scales, float* rotations, float* opacities,
                 float* shs, float* bbox_min_max, float* center);

const char* gsr_error_string(int code);
int gsr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GSRAST_B200_H */
