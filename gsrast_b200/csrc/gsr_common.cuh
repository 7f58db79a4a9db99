// gsr_common.cuh — shared declarations of the sm_100a splat forward rasterizer.
// Internal; the public surface is include/gsrast_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gsrast_b200.h"

namespace gsr {

constexpr int TILE_X = 16;  // BLOCK_X / BLOCK_Y of the reference (GSCuda.cu:20-21, Config.hpp:47-48)
constexpr int TILE_Y = 16;
constexpr int PRE_THREADS = 256;  // Gaussians per preprocess / duplicate block

// ---- individually rounded binary32 arithmetic ------------------------------------------
// The integer outputs of preprocess (radii, tile rects, depth key bits) are defined by
// IEEE operations in the source's association order; these intrinsics are never contracted
// into FMAs by nvcc, so the kernel matches the CPU oracle bit for bit.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float frcp(float a) { return __frcp_rn(a); }
// a*b + c*d + e*f, left to right (glm mat3 products, transformPoint)
__device__ __forceinline__ float dot3(float a, float b, float c, float d, float e, float f) {
    return fadd(fadd(fmul(a, b), fmul(c, d)), fmul(e, f));
}

struct PreprocessParams {
    int P, D, M;
    int W, H;
    int grid_x, grid_y;
    const float* means3D; int means_stride;
    const float* scales;  int scales_stride;
    const float* rotations;
    const float* opacities;
    const float* shs;
    const float* colors_precomp;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* cam_pos;
    float scale_modifier;
    float tan_fovx, tan_fovy;
    float focal_x, focal_y;
    float boxmin[3], boxmax[3];
    int prefiltered;
    // outputs
    int* radii;
    int* rects;  // int2[P] or null
    float* depths;  // [P]; doubles as the low half of the sort key: 0xffffffff where nothing is emitted
    unsigned char* clamped;
    float* means2D;
    float* cov3D;
    float* conic_opacity;
    float* rgb;
    uint32_t* tiles_touched;
    uint32_t* block_sums;  // [ceil(P/256)]
    uint32_t* coarse_block_sums;  // [ceil(P/256)] (Gaussian, bin) records of the block, or null
    uint32_t* tile_rects;  // uint2[P]: (miny<<16|minx, height<<16|width) of the tile rect, 0 when nothing is emitted
};

// stage launchers (each returns the number of kernels it launched, or <0 on error)
int launch_preprocess(const PreprocessParams& p, bool compat, cudaStream_t s);
// sums2 (optional): a second per-block array that is only reduced; its total goes to total_host_mapped[1].
int launch_scan_block_sums(uint32_t* block_sums, int num_blocks, uint32_t* total_dev, uint32_t* total_host_mapped,
                           cudaStream_t s, const uint32_t* sums2 = nullptr);
// Duplication blocks take num_dup_blocks(P) groups of consecutive depth ranks.
int num_dup_blocks(int P);
// sorted_rects[i] = tile_rects[sorted_ids[i]] and block_sums[b] = pairs emitted by the depth ranks of block b.
// Also materialises point_offsets[i] = inclusive scan of tiles_touched in index order (GSCuda.cu:771) from
// block_offsets = the exclusive offsets of preprocess' 256-Gaussian blocks.
// coarse: sorted_rects receives the rect in units of 8x8-tile bins (same packing) and block_sums the number of
// (Gaussian, bin) records — the input of the bin-expansion path.
// n_sorted (device, optional): only the first *n_sorted entries of sorted_ids are defined (sort32 with drop_pad).
int launch_gather_rects(int P, const uint32_t* sorted_ids, const uint32_t* tile_rects, uint32_t* sorted_rects,
                        uint32_t* block_sums, const uint32_t* tiles_touched, const uint32_t* block_offsets,
                        uint32_t* point_offsets, bool coarse, cudaStream_t s, const uint32_t* n_sorted = nullptr);
// Emits the (tile, Gaussian) pairs of the Gaussians taken in depth order: 32-bit tile keys + Gaussian
// ids, and accumulates the tile-digit histograms of the following radix passes into `hist`
// ([passes][256], zeroed).  `block_offsets` = exclusive scan of launch_gather_rects' block_sums, or — self_offsets —
// those block sums themselves, unscanned (every block then adds up the counts before its own).
int launch_duplicate_sorted(int P, int grid_x, const uint32_t* sorted_ids, const uint32_t* sorted_rects,
                            const uint32_t* block_offsets, uint32_t* keys32_out, uint32_t* vals_out, uint32_t* hist,
                            int tile_bits, cudaStream_t s, const uint32_t* n_sorted = nullptr, bool self_offsets = false);
// Fused form for callers that do not need point_offsets: gathers the rects by id itself (tile_rects; coarse = emit
// (bin, Gaussian) records) — or, rects_presorted, reads them already in depth order and final units from `tile_rects`
// (the last depth pass of the sort left them there, Sort32Plan::rect_dst) — and finds its offsets by decoupled look-back; `fuse_state` = duplicate_fused_state_bytes(P)
// zeroed bytes.  Replaces launch_gather_rects + the scan of its block sums + launch_duplicate_sorted.
size_t duplicate_fused_state_bytes(int P);
int launch_duplicate_fused(int P, int grid_x, const uint32_t* sorted_ids, const uint32_t* tile_rects, bool coarse,
                           void* fuse_state, uint32_t* keys32_out, uint32_t* vals_out, uint32_t* hist, int tile_bits,
                           cudaStream_t s, const uint32_t* n_sorted = nullptr, uint32_t* error_flag = nullptr,
                           bool rects_presorted = false);
// zero_first: clear ranges[num_tiles] here (otherwise the caller has already done it)
int launch_identify_ranges(const uint64_t* keys, size_t n, uint32_t* ranges, int num_tiles, bool compat,
                           cudaStream_t s, bool zero_first = true);

// radix sort
size_t sort_temp_bytes(size_t n);
int sort_num_passes(int end_bit);
// Sorts over bits [0,end_bit). The result lands in (keys_b, vals_b) when the pass count is odd and in
// (keys_a, vals_a) when it is even; `*result_in_a` says which.  The forward path picks its buffers so
// that the sorted lists land in point_list_keys / point_list without a final copy.
// `events` (optional, passes+2 entries) are recorded before the histogram, after it, and after every pass.
int launch_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, size_t n, int end_bit,
                      char* temp, bool* result_in_a, cudaStream_t s, cudaEvent_t* events = nullptr,
                      uint32_t* error_flag = nullptr);

// 32-bit-key LSD sort over bits [0,end_bit) (<= 4 passes).  Pass p reads the input (p == 0) or
// buffer (p-1)&1 and writes buffer p&1, the last pass writes the outputs (which may alias a buffer the
// last pass does not read).  vals_in == nullptr: values are the input positions.  expand_low != nullptr:
// the last pass writes 64-bit keys (key32 << 32 | expand_low[value]) to keys_out64 instead of keys_out.
struct Sort32Plan {
    size_t n;
    int end_bit;
    const uint32_t* keys_in;
    const uint32_t* vals_in;
    uint32_t* kbuf[2];
    uint32_t* vbuf[2];
    uint32_t* keys_out;
    uint32_t* vals_out;
    const uint32_t* expand_low;
    uint64_t* keys_out64;
    char* temp;       // sort_temp_bytes(n), prepared by sort32_prepare
    bool hist_ready;  // the producer of keys_in already accumulated the digit histograms
    // Keys equal to 0xffffffff are dropped: the histogram does not count them, the first pass compacts them away,
    // the later passes and every consumer work on the first sort32_kept_count() entries of the output only
    // (the rest of the output arrays is undefined).  Needs >= 2 passes and !hist_ready.
    bool drop_pad;
    // last pass only, optional: also write rect_dst[g] = rect_src[value] (uint2; coarse: converted to bin units) for the
    // item landing at g — the tile rects in sorted order; keys_out may then be nullptr.  Needs >= 2 passes.
    const uint32_t* rect_src;
    uint32_t* rect_dst;
    bool rect_coarse;
    // device word the look-back watchdog sets when it trips (nullptr: a word inside `temp`, which nobody reads back)
    uint32_t* error_flag;
};
// Device word that holds the number of kept keys after launch_sort32 with drop_pad (same temp, n, end_bit).
const uint32_t* sort32_kept_count(char* temp, size_t n, int end_bit);
// Zeroes the histograms / tickets / look-back state in `temp`; returns the histogram array
// ([pass][256]) for producers that count digits themselves, or nullptr on error.
uint32_t* sort32_prepare(char* temp, size_t n, int end_bit, cudaStream_t s);
// `events` (optional, passes+2 entries): before the histogram, after its scan, after every pass.
int launch_sort32(const Sort32Plan& plan, cudaStream_t s, cudaEvent_t* events = nullptr);

// ---- bin expansion (bin_expand.cu) -------------------------------------------------------------
// Bins are 8x8 tiles.  The (Gaussian, bin) records, stably sorted by bin (so each bin lists its Gaussians in
// depth order), are expanded into the per-tile sorted lists: point_list_keys / point_list / ranges.
constexpr int BIN_SHIFT = 3;
constexpr int BIN_SIDE = 1 << BIN_SHIFT;
constexpr int BIN_TILES = BIN_SIDE * BIN_SIDE;  // 64: one bit per tile in a 64-bit mask
constexpr int MAX_BINS = 4096;                  // larger grids fall back to radix passes over the pairs
// the rect of tile_rects (miny<<16|minx, height<<16|width) in units of bins; empty stays empty
__host__ __device__ __forceinline__ uint2 coarse_rect(const uint2 r) {
    if (r.y == 0u) return make_uint2(0u, 0u);
    const uint32_t x0 = r.x & 0xffffu, y0 = r.x >> 16, w = r.y & 0xffffu, h = r.y >> 16;
    const uint32_t bx0 = x0 >> BIN_SHIFT, bx1 = (x0 + w - 1u) >> BIN_SHIFT;
    const uint32_t by0 = y0 >> BIN_SHIFT, by1 = (y0 + h - 1u) >> BIN_SHIFT;
    return make_uint2((by0 << 16) | bx0, ((by1 - by0 + 1u) << 16) | (bx1 - bx0 + 1u));
}
size_t expand_temp_bytes(size_t R);
struct ExpandPlan {
    size_t n_records;             // (Gaussian, bin) records
    size_t num_rendered;          // R
    const uint32_t* rec_bins;     // [n_records] sorted bin ids
    const uint32_t* rec_ids;      // [n_records] Gaussian ids, depth order inside each bin
    const uint32_t* bin_counts;   // [nbins] records per bin (the digit histogram of a single-pass sort), or null
    const uint32_t* tile_rects;   // uint2[P]
    const uint32_t* depths;       // [P] depth bits
    int grid_x, grid_y, bins_x, bins_y;
    char* temp;                   // expand_temp_bytes(R)
    uint32_t* tile_counts;        // [tiles] scratch
    uint32_t* ranges;             // uint2[tiles]
    uint64_t* keys_out;           // [R]
    uint32_t* vals_out;           // [R]
    bool r1_quirk;                // GSRast-compat: a frame of exactly one pair never closes its range (GSCuda.cu:533-536)
};
// ev (optional, 4 events): recorded before / after the count kernel and before / after the fill kernel
int launch_bin_expand(const ExpandPlan& plan, cudaStream_t s, cudaEvent_t* ev = nullptr);

struct BlendParams {
    int W, H, grid_x, grid_y;
    const uint32_t* ranges;      // uint2[tiles]
    const uint32_t* point_list;  // sorted Gaussian ids
    const float* means2D;
    const float* colors;         // float3[P]
    const float* conic_opacity;
    const float* background;     // device float[3]
    float* final_T;
    uint32_t* n_contrib;
    float* out_color;
    float t_min;                 // 0.0001f contract, 0.001f GSRast
    const uint32_t* tile_order;  // optional
    unsigned long long* counters;  // optional [GSR_BLEND_COUNTERS]: run the counting instantiation (GSR_FLAG_BLEND_COUNT)
};
// simple: the reference-structured kernel; one_pixel: blend_culled_kernel (one pixel per thread) instead of the default
// blend_pair_kernel (two pixels per thread on the packed FP32 pipe)
int launch_blend(const BlendParams& p, bool simple, cudaStream_t s, bool one_pixel = false);
int launch_fill_background(int W, int H, const float* background, float* out_color, float* final_T,
                           uint32_t* n_contrib, cudaStream_t s);

// Mapped pinned words (64 bytes) the device reports into; SLOT_R is the pipeline's only host round trip.
constexpr int SLOT_R = 0;      // num_rendered; 0xffffffff = the pair count does not fit (scan_block_sums_kernel)
constexpr int SLOT_RC = 1;     // (Gaussian, bin) records of the bin-expansion path
constexpr int SLOT_ERROR = 4;  // sticky: set by the look-back watchdogs (radix_sort.cu, binning.cu), taken by the host
struct HostSlot {
    uint32_t* host = nullptr;
    uint32_t* dev = nullptr;
    cudaEvent_t landed = nullptr;  // recorded right after the kernel that writes the word
    int device = -1;
};
int ensure_slot(HostSlot& s);
void release_slot(HostSlot& s);
// GSR_ERR_SORT_STALLED (and the word cleared) when a watchdog reported into the slot since the last look, else 0
int take_async_error(HostSlot& s);
// gsr_forward_ex with an explicit readback slot (nullptr = the calling thread's own).
int forward_impl(const gsr_forward_args* args, HostSlot* slot);

}  // namespace gsr

// ---- programmatic dependent launch (sm_90+) ---------------------------------------------------
// The stages of one frame are a chain of short kernels on one stream.  A kernel launched through
// launch_pdl may be scheduled while the previous kernel of the stream drains (its CTAs take the SM
// slots the predecessor frees, set up shared memory, fetch their ticket...) and blocks in
// gsr_pdl_wait() until the predecessor has completed and its writes are visible.  EVERY kernel calls
// gsr_pdl_wait() before it touches global data, so the chain stays transitively ordered; in a kernel
// launched the ordinary way the instruction returns at once.
#ifdef __CUDACC__
__device__ __forceinline__ void gsr_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gsr_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
#ifndef GSR_PDL
    // Default: ordinary launches.  Measured on B200 (C2): PDL shortens a lone frame by 1.6 % (1.160 -> 1.142 ms)
    // but costs 5 % of the pipelined throughput (838 -> 798 frames/s), because parked dependent CTAs take SM
    // slots from the other view's kernels, which already fill the drain gaps.  Build with -DGSR_PDL for
    // latency-critical single-view use.
    kernel<<<grid, block, smem, s>>>(static_cast<KArgs>(args)...);
    return cudaPeekAtLastError();
#else
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
#endif
}
#endif

// Raise the sticky error of a call from the device.  The word usually lives in mapped pinned HOST memory (HostSlot),
// so this is a plain system-scope store + fence, not an atomic (idempotent: every reporter stores 1).
#ifdef __CUDACC__
__device__ __forceinline__ void gsr_raise_error(uint32_t* flag) {
    if (flag) {
        *reinterpret_cast<volatile uint32_t*>(flag) = 1u;
        __threadfence_system();
    }
}
#endif

// ---- shared-memory carve-out per kernel -----------------------------------------------------------
// The SM's unified L1 / shared memory is split per kernel; CTAs of two kernels only share an SM when the SM's
// current split fits both, so the preference decides how well the stages of the two view lanes overlap
// (measured: +8 % pipelined frames/s from the blend's preference alone, profiles/r01i_carveout.txt).
// GSR_CARVEOUT(kernel, "NAME", default): percentage of the unified memory preferred as shared memory, -1 = leave
// the driver's choice; the environment variable GSR_CARVEOUT_<NAME> overrides the default (tuning runs).
inline int gsr_set_carveout(const void* fn, const char* env_name, int dflt) {
    int v = dflt;
    if (const char* e = getenv(env_name)) v = atoi(e);
    if (v >= 0) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, v > 100 ? 100 : v);
    return v;
}
#define GSR_CARVEOUT(kernel, NAME, DFLT)                                                              \
    do {                                                                                              \
        static bool done_[32] = {};                                                                   \
        int dev_ = 0;                                                                                 \
        cudaGetDevice(&dev_);                                                                         \
        if (!done_[dev_ & 31]) {                                                                      \
            gsr_set_carveout(reinterpret_cast<const void*>(kernel), "GSR_CARVEOUT_" NAME, DFLT);      \
            done_[dev_ & 31] = true;                                                                  \
        }                                                                                             \
    } while (0)

#define GSR_CUDA_TRY(expr)                          \
    do {                                            \
        cudaError_t _e = (expr);                    \
        if (_e != cudaSuccess) return -(int)_e;     \
    } while (0)
