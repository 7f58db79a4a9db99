// forward.cu — host orchestration + C ABI of the forward rasterizer.
//
// Replaces gscuda::forward (/root/reference/apps/gsrast/gscuda/GSCuda.cu:695-811) and the
// scratch-buffer carving of AuxBuffer.cu:13-89 / AuxBuffer.cuh:8-14.  Stage order and the
// allocator protocol are the reference's: geometry chunk -> image chunk -> preprocess -> scan
// -> (num_rendered to the host) -> binning chunk -> duplicate -> sort -> ranges -> blend.
// Differences that are deliberate:
//   * the only host<->device round trip is the 4-byte num_rendered, written by the scan
//     kernel straight into mapped pinned memory (the reference does a blocking cudaMemcpy,
//     GSCuda.cu:772, and its caller a cudaDeviceSynchronize per call, CudaBuffer.hpp:8-12);
//     the host waits on an event recorded right behind that kernel, and the GPU spends the
//     wait on the depth half of the radix sort, which needs neither num_rendered nor the
//     binning chunk;
//   * the LSD radix sort is split: depth digits are sorted per Gaussian (P records) before
//     duplication, only the tile digits per pair (radix_sort.cu explains why this is the
//     same sort);
//   * `ranges` is one entry per tile, not per pixel (GSCuda.cu:800 clears W*H entries);
//   * everything runs on the caller's stream.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <limits>

#include "gsr_common.cuh"

namespace gsr {

namespace {

// obtain() of AuxBuffer.cu:13-21: bump allocation at 128-byte alignment.
template <typename T>
void obtain(char*& chunk, T*& out, size_t bytes, size_t align = 128) {
    size_t off = reinterpret_cast<size_t>(chunk);
    size_t aligned = (off + align - 1) / align * align;
    out = reinterpret_cast<T*>(aligned);
    chunk = reinterpret_cast<char*>(aligned + bytes);
}

int num_pre_blocks(int P) { return (P + PRE_THREADS - 1) / PRE_THREADS; }

// Mapped pinned slot for the num_rendered readback, one per host thread and device; released when the thread ends
// (at process exit the runtime may already be gone: the calls then fail harmlessly).
struct ThreadSlot : HostSlot {
    ~ThreadSlot() { release_slot(*this); }
};
thread_local ThreadSlot g_slot;

}  // namespace

int ensure_slot(HostSlot& s) {
    int dev = 0;
    GSR_CUDA_TRY(cudaGetDevice(&dev));
    if (s.host && s.device == dev) return 0;
    if (s.host) cudaFreeHost(s.host);
    s.host = nullptr;
    void* h = nullptr;
    GSR_CUDA_TRY(cudaHostAlloc(&h, 64, cudaHostAllocMapped));
    memset(h, 0, 64);
    s.host = static_cast<uint32_t*>(h);
    void* d = nullptr;
    GSR_CUDA_TRY(cudaHostGetDevicePointer(&d, h, 0));
    s.dev = static_cast<uint32_t*>(d);
    if (s.landed) cudaEventDestroy(s.landed);
    s.landed = nullptr;
    GSR_CUDA_TRY(cudaEventCreateWithFlags(&s.landed, cudaEventDisableTiming));
    s.device = dev;
    return 0;
}

void release_slot(HostSlot& s) {
    if (s.host) cudaFreeHost(s.host);
    if (s.landed) cudaEventDestroy(s.landed);
    s.host = s.dev = nullptr;
    s.landed = nullptr;
    s.device = -1;
}

// The sort's look-back watchdog reports into word SLOT_ERROR of the slot (sticky, like a CUDA asynchronous error): the
// host looks at it wherever it already holds the slot in its hands — on entry of the next call, behind the
// num_rendered wait, and after every stream synchronisation the library does itself.
int take_async_error(HostSlot& s) {
    if (!s.host) return 0;
    volatile uint32_t* w = s.host;
    if (w[SLOT_ERROR] == 0u) return 0;
    w[SLOT_ERROR] = 0u;
    return GSR_ERR_SORT_STALLED;
}

namespace {

// Stage events of one call (only when the caller asked for timings); destroyed on every way out of forward_impl.
struct StageTimer {
    bool on = false;
    cudaStream_t s = nullptr;
    cudaEvent_t ev[10] = {};
    cudaEvent_t sort_ev[16] = {};  // [0..11] the two sort halves, [12..15] count / fill kernels of the bin expansion
    int n = 0;
    int init(bool enable, cudaStream_t stream) {
        s = stream;
        if (!enable) return 0;
        on = true;
        for (auto& e : ev) GSR_CUDA_TRY(cudaEventCreate(&e));
        for (auto& e : sort_ev) GSR_CUDA_TRY(cudaEventCreate(&e));
        return 0;
    }
    void mark() {
        if (on && n < 10) cudaEventRecord(ev[n++], s);
    }
    float ms(int a, int b) {
        float t = 0.f;
        if (on && a < n && b < n) cudaEventElapsedTime(&t, ev[a], ev[b]);
        return t;
    }
    float sort_ms(int a, int b) {
        float t = 0.f;
        if (on) cudaEventElapsedTime(&t, sort_ev[a], sort_ev[b]);
        return t;
    }
    ~StageTimer() {
        if (!on) return;
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : sort_ev)
            if (e) cudaEventDestroy(e);
    }
};

}  // namespace

}  // namespace gsr

using namespace gsr;

extern "C" {

uint32_t gsr_get_higher_msb(uint32_t n) {
    // GSCuda.cu:481-502: binary search for the position above the highest set bit
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step;
        else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

size_t gsr_geometry_state_map(char* chunk, int P, gsr_geometry_state* g) {
    char* c = chunk;
    const size_t n = (size_t)(P > 0 ? P : 0);
    gsr_geometry_state st;
    obtain(c, st.depths, n * sizeof(float));
    obtain(c, st.clamped, n * 3);
    obtain(c, st.internal_radii, n * sizeof(int));
    obtain(c, st.means2D, n * 2 * sizeof(float));
    obtain(c, st.cov3D, n * 6 * sizeof(float));
    obtain(c, st.conic_opacity, n * 4 * sizeof(float));
    obtain(c, st.rgb, n * 3 * sizeof(float));
    obtain(c, st.tiles_touched, n * sizeof(uint32_t));
    st.scan_size = ((size_t)num_pre_blocks(P) + 4) * sizeof(uint32_t);
    obtain(c, st.block_sums, st.scan_size);
    obtain(c, st.point_offsets, n * sizeof(uint32_t));
    st.depth_keys = reinterpret_cast<uint32_t*>(st.depths);  // the same array: see gsrast_b200.h
    obtain(c, st.tile_rects, n * 2 * sizeof(uint32_t));
    for (int i = 0; i < 2; ++i) obtain(c, st.depth_sort_keys[i], n * sizeof(uint32_t));
    for (int i = 0; i < 2; ++i) obtain(c, st.depth_sort_ids[i], n * sizeof(uint32_t));
    st.depth_sort_size = sort_temp_bytes(n);
    obtain(c, st.depth_sort_space, st.depth_sort_size);
    obtain(c, st.sorted_rects, n * 2 * sizeof(uint32_t));
    obtain(c, st.sorted_block_sums, st.scan_size);
    obtain(c, st.coarse_block_sums, st.scan_size);
    if (g) *g = st;
    return (size_t)(c - chunk);
}

size_t gsr_image_state_map(char* chunk, int width, int height, gsr_image_state* im) {
    char* c = chunk;
    const size_t N = (size_t)width * (size_t)height;
    const size_t T = (size_t)((width + TILE_X - 1) / TILE_X) * (size_t)((height + TILE_Y - 1) / TILE_Y);
    gsr_image_state st;
    obtain(c, st.ranges, T * 2 * sizeof(uint32_t));
    obtain(c, st.n_contrib, N * sizeof(uint32_t));
    obtain(c, st.accum_alpha, N * sizeof(float));
    obtain(c, st.tile_order, T * sizeof(uint32_t));
    obtain(c, st.blend_counters, GSR_BLEND_COUNTERS * sizeof(unsigned long long));
    if (im) *im = st;
    return (size_t)(c - chunk);
}

size_t gsr_binning_state_map(char* chunk, size_t R, gsr_binning_state* b) {
    char* c = chunk;
    gsr_binning_state st;
    obtain(c, st.point_list_keys_unsorted, R * sizeof(uint64_t));
    obtain(c, st.point_list_keys, R * sizeof(uint64_t));
    obtain(c, st.point_list_unsorted, R * sizeof(uint32_t));
    obtain(c, st.point_list, R * sizeof(uint32_t));
    st.sorting_size = (R * sizeof(uint32_t) + 127) / 128 * 128 + sort_temp_bytes(R) + expand_temp_bytes(R);
    obtain(c, st.list_sorting_space, st.sorting_size);
    if (b) *b = st;
    return (size_t)(c - chunk);
}

// required<T>() probes with a null chunk; +128 covers a chunk base that is not 128-aligned
// (the reference over-asks the same way, GSCuda.cu:735,783).
size_t gsr_geometry_state_required(int P) { return gsr_geometry_state_map(nullptr, P, nullptr) + 128; }
size_t gsr_image_state_required(int width, int height) { return gsr_image_state_map(nullptr, width, height, nullptr) + 128; }
size_t gsr_binning_state_required(size_t R) { return gsr_binning_state_map(nullptr, R, nullptr) + 128; }

int gsr_forward_ex(const gsr_forward_args* a) { return gsr::forward_impl(a, nullptr); }

}  // extern "C"

// GSR_LEAN_KEYS=1: lean calls do not materialise the sorted 64-bit keys (bin_expand.cu, KEYS=false)
#ifndef GSR_LEAN_KEYS
#define GSR_LEAN_KEYS 1
#endif
// GSR_LEAN_RADII=1: lean calls that pass no `radii` buffer do not fill internal_radii (4 B/Gaussian of preprocess stores)
#ifndef GSR_LEAN_RADII
#define GSR_LEAN_RADII 1
#endif
// GSR_DUP_SELF_OFFSETS=1: the duplication blocks add up the pair counts before their own themselves (no scan launch)
#ifndef GSR_DUP_SELF_OFFSETS
#define GSR_DUP_SELF_OFFSETS 1
#endif
#ifndef GSR_DUP_SELF_MAX
#define GSR_DUP_SELF_MAX 4096
#endif
// GSR_FUSED_DUP=1: lean calls run the fused duplication (gather + look-back + emit in one kernel)
#ifndef GSR_FUSED_DUP
#define GSR_FUSED_DUP 0
#endif
// GSR_FUSED_SORT=1: lean calls let the LAST depth pass of the sort bring the tile rects into depth order
// (Sort32Plan::rect_dst) and the duplication find its offsets by look-back over those presorted rects, so the
// gather_rects launch and the single-CTA scan behind it disappear from the frame.  Built, parity-green (76 GPU tests),
// and measured SLOWER (profiles/r02f_ab_C2.txt: the pass with 16 random 8-byte gathers per thread takes 93-106 us
// instead of 32, the look-back duplication 48 instead of 33; 1315 vs 1383 frames/s): gather_rects at 88 % occupancy
// hides the DRAM latency of that gather far better than 32 warps per SM in the tail of a sort pass can.  Off.
#ifndef GSR_FUSED_SORT
#define GSR_FUSED_SORT 0
#endif
// every stage launcher returns the number of kernels it launched, or a negative error
#define GSR_STAGE(call)       \
    do {                      \
        rc = (call);          \
        if (rc < 0) return rc;\
        launches += rc;       \
    } while (0)
int gsr::forward_impl(const gsr_forward_args* a, gsr::HostSlot* slot_in) {
    HostSlot& slot = slot_in ? *slot_in : g_slot;
    if (!a || !a->geometry_alloc || !a->binning_alloc || !a->image_alloc) return GSR_ERR_INVALID_ARG;
    if (a->P < 0 || a->width <= 0 || a->height <= 0 || !a->out_color || !a->background) return GSR_ERR_INVALID_ARG;
    if (a->P > 0 && (!a->means3D || !a->opacities || !a->viewmatrix || !a->projmatrix)) return GSR_ERR_INVALID_ARG;
    if (a->P > 0 && !a->cov3D_precomp && (!a->scales || !a->rotations)) return GSR_ERR_INVALID_ARG;
    if (a->P > 0 && !a->colors_precomp && !a->shs) return GSR_ERR_INVALID_ARG;
    const bool compat = (a->flags & GSR_FLAG_GSRAST_COMPAT) != 0;
    if (a->P > 0 && !compat && !a->colors_precomp && !a->cam_pos) return GSR_ERR_INVALID_ARG;
    const int ms = a->means_stride ? a->means_stride : 3, ss = a->scales_stride ? a->scales_stride : 3;
    if ((ms != 3 && ms != 4) || (ss != 3 && ss != 4)) return GSR_ERR_INVALID_ARG;
    if (!compat && !a->colors_precomp && (a->D < 0 || a->D > 3 || a->M < 1)) return GSR_ERR_INVALID_ARG;

    cudaStream_t s = static_cast<cudaStream_t>(a->stream);
    const int P = a->P, W = a->width, H = a->height;
    const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
    // tile rects are packed 16 + 16 bits, and 256 Gaussians x (tiles per Gaussian <= gx*gy) must fit the 32-bit block
    // sums of the scan (GSCuda.cu:771 scans in 32 bits as well); 2^23 tiles = a 46 000 x 46 000 pixel image
    if (gx > 0xffff || gy > 0xffff || (long long)gx * gy > (1ll << 23)) return GSR_ERR_INVALID_ARG;
    const int tiles = gx * gy;
    int launches = 0, rc = 0;

    StageTimer tm;
    if ((rc = tm.init(a->timings != nullptr, s)) < 0) return rc;
    cudaEvent_t* sort_ev = tm.on ? tm.sort_ev : nullptr;

    // geometry chunk (GSCuda.cu:723-729)
    char* gchunk = a->geometry_alloc(gsr_geometry_state_required(P), a->geometry_user);
    if (!gchunk) return GSR_ERR_ALLOC_FAILED;
    gsr_geometry_state geom;
    gsr_geometry_state_map(gchunk, P, &geom);
    int* radii = a->radii ? a->radii : geom.internal_radii;

    // image chunk (GSCuda.cu:734-736)
    char* ichunk = a->image_alloc(gsr_image_state_required(W, H), a->image_user);
    if (!ichunk) return GSR_ERR_ALLOC_FAILED;
    gsr_image_state img;
    gsr_image_state_map(ichunk, W, H, &img);

    const int tile_bits = (int)gsr_get_higher_msb((uint32_t)tiles);  // GSCuda.cu:791-797: end_bit = 32 + this
    // tile half of the sort: bin expansion (bin_expand.cu) unless asked otherwise or the grid has too many bins
    const int bins_x = (gx + BIN_SIDE - 1) >> BIN_SHIFT, bins_y = (gy + BIN_SIDE - 1) >> BIN_SHIFT;
    const int nbins = bins_x * bins_y;
    const bool bin_mode = !(a->flags & GSR_FLAG_RADIX_BINNING) && nbins <= MAX_BINS;
    const int bin_bits = (int)gsr_get_higher_msb((uint32_t)nbins);
    const int tile_passes = sort_num_passes(bin_mode ? bin_bits : tile_bits);
    const int depth_passes = sort_num_passes(32);

    uint32_t R = 0, Rc = 0;
    bool fused_dup = false, fused_sort = false, dup_self = false;
    const uint32_t* n_depth = nullptr;  // device: Gaussians the depth sort kept
    uint32_t* sort_error = nullptr;     // device address of the slot's sticky error word
    tm.mark();  // 0
    if (P > 0) {
        if ((rc = ensure_slot(slot)) < 0) return rc;
        if ((rc = take_async_error(slot)) < 0) return rc;  // a watchdog of an earlier call on this slot tripped
        sort_error = slot.dev + SLOT_ERROR;
        PreprocessParams pp;
        memset(&pp, 0, sizeof(pp));
        pp.P = P; pp.D = a->D; pp.M = a->M; pp.W = W; pp.H = H; pp.grid_x = gx; pp.grid_y = gy;
        pp.means3D = a->means3D; pp.means_stride = ms;
        pp.scales = a->scales; pp.scales_stride = ss;
        pp.rotations = a->rotations; pp.opacities = a->opacities; pp.shs = a->shs;
        pp.colors_precomp = a->colors_precomp; pp.cov3D_precomp = a->cov3D_precomp;
        pp.viewmatrix = a->viewmatrix; pp.projmatrix = a->projmatrix; pp.cam_pos = a->cam_pos;
        pp.scale_modifier = a->scale_modifier;
        pp.tan_fovx = a->tan_fovx; pp.tan_fovy = a->tan_fovy;
        pp.focal_y = (float)H / (2.0f * a->tan_fovy);  // GSCuda.cu:721
        pp.focal_x = (float)W / (2.0f * a->tan_fovx);
        const float fm = std::numeric_limits<float>::max();  // GSCuda.cu:738-741
        for (int i = 0; i < 3; ++i) {
            pp.boxmin[i] = a->boxmin ? a->boxmin[i] : -fm;
            pp.boxmax[i] = a->boxmax ? a->boxmax[i] : fm;
        }
        pp.prefiltered = a->prefiltered;
        pp.radii = radii; pp.rects = a->rects; pp.depths = geom.depths; pp.clamped = geom.clamped;
        pp.means2D = geom.means2D; pp.cov3D = geom.cov3D; pp.conic_opacity = geom.conic_opacity;
        pp.rgb = geom.rgb; pp.tiles_touched = geom.tiles_touched; pp.block_sums = geom.block_sums;
        pp.tile_rects = geom.tile_rects;
        pp.coarse_block_sums = bin_mode ? geom.coarse_block_sums : nullptr;
        const bool lean = (a->flags & GSR_FLAG_LEAN_STATE) != 0;
        if (lean) {
            pp.cov3D = nullptr; pp.clamped = nullptr; pp.tiles_touched = nullptr;
            if (GSR_LEAN_RADII && !a->radii) pp.radii = nullptr;  // no caller buffer: internal_radii has no reader either
        }
        // lean callers need no point_offsets, so the duplication can gather its rects and find its offsets itself
        // (binning.cu, duplicate_sorted_kernel<true>) instead of running behind gather_rects + a single-CTA scan
        fused_sort = lean && GSR_FUSED_SORT != 0;
        fused_dup = fused_sort || (lean && GSR_FUSED_DUP != 0);
        if (fused_dup) GSR_CUDA_TRY(cudaMemsetAsync(geom.sorted_block_sums, 0, duplicate_fused_state_bytes(P), s));
        // all clears of the frame up front, so the kernels behind them form uninterrupted dependent-launch chains
        if (!sort32_prepare(geom.depth_sort_space, (size_t)P, 32, s)) return -(int)cudaGetLastError();
        GSR_CUDA_TRY(cudaMemsetAsync(img.ranges, 0, sizeof(uint32_t) * 2 * (size_t)tiles, s));
        GSR_STAGE(launch_preprocess(pp, compat, s));
        tm.mark();  // 1
        const int nb = num_pre_blocks(P);
        GSR_STAGE(launch_scan_block_sums(geom.block_sums, nb, geom.block_sums + nb, slot.dev, s,
                                         bin_mode ? geom.coarse_block_sums : nullptr));
        GSR_CUDA_TRY(cudaEventRecord(slot.landed, s));
        tm.mark();  // 2
        // ---- depth half of the sort: P records, queued before the host waits --------------------
        Sort32Plan dp;
        memset(&dp, 0, sizeof(dp));
        dp.n = (size_t)P; dp.end_bit = 32;
        dp.keys_in = geom.depth_keys; dp.vals_in = nullptr;
        dp.kbuf[0] = geom.depth_sort_keys[0]; dp.kbuf[1] = geom.depth_sort_keys[1];
        dp.vbuf[0] = geom.depth_sort_ids[0]; dp.vbuf[1] = geom.depth_sort_ids[1];
        dp.keys_out = geom.depth_sort_keys[1]; dp.vals_out = geom.depth_sort_ids[1];
        dp.temp = geom.depth_sort_space; dp.hist_ready = false;
        dp.error_flag = sort_error;
        // Gaussians that emit nothing carry the key 0xffffffff (preprocess): the sort drops them, so its later
        // passes, the rect gather and the duplication only see the n_depth <= P Gaussians that are on screen
        dp.drop_pad = true;
        if (fused_sort) {
            dp.rect_src = geom.tile_rects; dp.rect_dst = geom.sorted_rects; dp.rect_coarse = bin_mode;
            dp.keys_out = nullptr;  // nothing downstream reads the sorted depth keys
        }
        n_depth = sort32_kept_count(geom.depth_sort_space, (size_t)P, 32);
        GSR_STAGE(launch_sort32(dp, s, sort_ev));
        // tile rects into depth order + scan of the per-block pair counts (same total, other order)
        if (!fused_dup) GSR_STAGE(launch_gather_rects(P, geom.depth_sort_ids[1], geom.tile_rects, geom.sorted_rects,
                                      geom.sorted_block_sums, geom.tiles_touched, geom.block_sums,
                                      lean ? nullptr : geom.point_offsets, bin_mode, s, n_depth));
        const int ndb = num_dup_blocks(P);
        // up to GSR_DUP_SELF_MAX duplication blocks (4 M Gaussians; beyond, the quadratic adds cost what the scan does) every block adds up the counts before its own
        // instead of waiting for a single-CTA scan of them (binning.cu, self_offsets)
        dup_self = !fused_dup && GSR_DUP_SELF_OFFSETS && ndb <= GSR_DUP_SELF_MAX;
        if (!fused_dup && !dup_self)
            GSR_STAGE(launch_scan_block_sums(geom.sorted_block_sums, ndb, geom.sorted_block_sums + ndb, nullptr, s));
        tm.mark();  // 3
        // the one host round trip: num_rendered decides the binning allocation (GSCuda.cu:772,782)
        GSR_CUDA_TRY(cudaEventSynchronize(slot.landed));
        R = static_cast<volatile uint32_t*>(slot.host)[SLOT_R];
        Rc = bin_mode ? static_cast<volatile uint32_t*>(slot.host)[SLOT_RC] : 0u;
        if ((rc = take_async_error(slot)) < 0) return rc;
    } else {
        tm.mark();
        tm.mark();
        tm.mark();
    }

    if (R == 0) {
        // GSCuda.cu:775-778 returns here leaving out_color stale; the contract renders background.
        if (!compat) GSR_STAGE(launch_fill_background(W, H, a->background, a->out_color, img.accum_alpha, img.n_contrib, s));
        if (a->timings) {
            GSR_CUDA_TRY(cudaStreamSynchronize(s));
            memset(a->timings, 0, sizeof(*a->timings));
            a->timings->preprocess_ms = tm.ms(0, 1);
            a->timings->scan_ms = tm.ms(1, 2);
            a->timings->depth_sort_ms = tm.ms(2, 3);
            a->timings->total_ms = tm.ms(0, 3);
            a->timings->kernel_launches = launches;
            if (P > 0 && (rc = take_async_error(slot)) < 0) return rc;
        }
        return 0;
    }
    // the scan kernel adds the per-block counts up in 64 bits and publishes 0xffffffff when the sum does not fit the
    // sort's 30-bit counters (the 32-bit sum the reference would use, GSCuda.cu:771, may have wrapped to anything)
    if (R >= (1u << 30)) return GSR_ERR_TOO_MANY_PAIRS;

    // binning chunk (GSCuda.cu:782-784)
    char* bchunk = a->binning_alloc(gsr_binning_state_required(R), a->binning_user);
    if (!bchunk) return GSR_ERR_ALLOC_FAILED;
    gsr_binning_state bin;
    gsr_binning_state_map(bchunk, R, &bin);

    // ---- duplication in depth order + tile half of the sort ----------------------------------
    uint32_t* k32[2] = {reinterpret_cast<uint32_t*>(bin.point_list_keys_unsorted),
                        reinterpret_cast<uint32_t*>(bin.point_list_keys_unsorted) + R};
    uint32_t* v32[2] = {bin.point_list_unsorted, reinterpret_cast<uint32_t*>(bin.list_sorting_space)};
    char* tile_temp = bin.list_sorting_space + ((size_t)R * sizeof(uint32_t) + 127) / 128 * 128;
    if (bin_mode) {
        // (Gaussian, bin) records in depth order -> stable sort by bin id -> expansion of every bin into the
        // sorted per-tile lists and the tile ranges (bin_expand.cu)
        if (Rc == 0 || Rc > R) return GSR_ERR_INVALID_ARG;  // cannot happen: every pair lies in one of its Gaussian's bins
        uint32_t* bin_hist = sort32_prepare(tile_temp, (size_t)Rc, bin_bits, s);
        if (!bin_hist) return -(int)cudaGetLastError();
        if (fused_dup)
            GSR_STAGE(launch_duplicate_fused(P, bins_x, geom.depth_sort_ids[1], fused_sort ? geom.sorted_rects : geom.tile_rects,
                                             /*coarse=*/true, geom.sorted_block_sums, k32[0], v32[0], bin_hist, bin_bits, s,
                                             n_depth, sort_error, /*rects_presorted=*/fused_sort));
        else
            GSR_STAGE(launch_duplicate_sorted(P, bins_x, geom.depth_sort_ids[1], geom.sorted_rects, geom.sorted_block_sums,
                                              k32[0], v32[0], bin_hist, bin_bits, s, n_depth, dup_self));
        tm.mark();  // 4
        const int out = tile_passes & 1;  // one pass: [0] -> [1]; two: [0] -> [1] -> [0]
        {
            Sort32Plan tp;
            memset(&tp, 0, sizeof(tp));
            tp.n = (size_t)Rc; tp.end_bit = bin_bits;
            tp.keys_in = k32[0]; tp.vals_in = v32[0];
            tp.kbuf[0] = k32[1]; tp.kbuf[1] = k32[0];
            tp.vbuf[0] = v32[1]; tp.vbuf[1] = v32[0];
            tp.keys_out = k32[out]; tp.vals_out = v32[out];
            tp.temp = tile_temp; tp.hist_ready = true;
            tp.error_flag = sort_error;
            GSR_STAGE(launch_sort32(tp, s, sort_ev ? sort_ev + 6 : nullptr));
        }
        tm.mark();  // 5
        ExpandPlan ep;
        memset(&ep, 0, sizeof(ep));
        ep.n_records = (size_t)Rc; ep.num_rendered = (size_t)R;
        ep.rec_bins = k32[out]; ep.rec_ids = v32[out];
        ep.bin_counts = tile_passes == 1 ? bin_hist : nullptr;  // single pass: the digit histogram is the bin histogram
        ep.tile_rects = geom.tile_rects; ep.depths = reinterpret_cast<const uint32_t*>(geom.depths);
        ep.grid_x = gx; ep.grid_y = gy; ep.bins_x = bins_x; ep.bins_y = bins_y;
        ep.temp = tile_temp + sort_temp_bytes((size_t)R);
        ep.tile_counts = img.tile_order; ep.ranges = img.ranges;
        // lean callers (the renderer's private scratch): nothing reads the sorted 64-bit keys back — the ranges come out of
        // the expansion's scan — so they are not materialised: 8 of the 12 bytes per pair the tile sort writes (GSR_LEAN_KEYS)
        ep.keys_out = ((a->flags & GSR_FLAG_LEAN_STATE) && GSR_LEAN_KEYS) ? nullptr : bin.point_list_keys;
        ep.vals_out = bin.point_list;
        ep.r1_quirk = compat && R == 1;
        GSR_STAGE(launch_bin_expand(ep, s, sort_ev ? sort_ev + 12 : nullptr));
        tm.mark();  // 6
    } else {
        uint32_t* tile_hist = sort32_prepare(tile_temp, (size_t)R, tile_bits, s);
        if (!tile_hist) return -(int)cudaGetLastError();
        if (fused_dup)
            GSR_STAGE(launch_duplicate_fused(P, gx, geom.depth_sort_ids[1], fused_sort ? geom.sorted_rects : geom.tile_rects,
                                             /*coarse=*/false, geom.sorted_block_sums, k32[0], v32[0], tile_hist, tile_bits, s,
                                             n_depth, sort_error, /*rects_presorted=*/fused_sort));
        else
            GSR_STAGE(launch_duplicate_sorted(P, gx, geom.depth_sort_ids[1], geom.sorted_rects, geom.sorted_block_sums, k32[0],
                                              v32[0], tile_hist, tile_bits, s, n_depth, dup_self));
        tm.mark();  // 4
        {
            Sort32Plan tp;
            memset(&tp, 0, sizeof(tp));
            tp.n = (size_t)R; tp.end_bit = tile_bits;
            tp.keys_in = k32[0]; tp.vals_in = v32[0];
            tp.kbuf[0] = k32[1]; tp.kbuf[1] = k32[0];
            tp.vbuf[0] = v32[1]; tp.vbuf[1] = v32[0];
            tp.keys_out = nullptr; tp.vals_out = bin.point_list;
            tp.expand_low = reinterpret_cast<const uint32_t*>(geom.depths);  // key = tile << 32 | depth bits
            tp.keys_out64 = bin.point_list_keys;
            tp.temp = tile_temp; tp.hist_ready = true;
            tp.error_flag = sort_error;
            GSR_STAGE(launch_sort32(tp, s, sort_ev ? sort_ev + 6 : nullptr));
        }
        tm.mark();  // 5
        GSR_STAGE(launch_identify_ranges(bin.point_list_keys, R, img.ranges, tiles, compat, s, /*zero_first=*/false));
        tm.mark();  // 6
    }

    BlendParams bp;
    memset(&bp, 0, sizeof(bp));
    bp.W = W; bp.H = H; bp.grid_x = gx; bp.grid_y = gy;
    bp.ranges = img.ranges; bp.point_list = bin.point_list; bp.means2D = geom.means2D;
    bp.colors = a->colors_precomp ? a->colors_precomp : geom.rgb;  // GSCuda.cu:803
    bp.conic_opacity = geom.conic_opacity; bp.background = a->background;
    bp.final_T = img.accum_alpha; bp.n_contrib = img.n_contrib; bp.out_color = a->out_color;
    bp.t_min = compat ? 0.001f : 0.0001f;  // GSCuda.cu:653 vs contract
    bp.tile_order = nullptr;
    // GSR_FLAG_BLEND_COUNT: the counting instantiation of the blend (same arithmetic + work counters), for reporting
    const bool count = (a->flags & GSR_FLAG_BLEND_COUNT) != 0 && a->timings != nullptr;
    if (count) {
        GSR_CUDA_TRY(cudaMemsetAsync(img.blend_counters, 0, GSR_BLEND_COUNTERS * sizeof(unsigned long long), s));
        bp.counters = img.blend_counters;
    }
    GSR_STAGE(launch_blend(bp, (a->flags & GSR_FLAG_BLEND_SIMPLE) != 0, s, (a->flags & GSR_FLAG_BLEND_ONE_PIXEL) != 0));
    tm.mark();  // 7

    if (a->timings) {
        GSR_CUDA_TRY(cudaStreamSynchronize(s));
        gsr_stage_times* t = a->timings;
        memset(t, 0, sizeof(*t));
        t->preprocess_ms = tm.ms(0, 1);
        t->scan_ms = tm.ms(1, 2);
        t->depth_sort_ms = tm.ms(2, 3);
        t->duplicate_ms = tm.ms(3, 4);
        t->sort_ms = tm.ms(4, 5);
        t->ranges_ms = bin_mode ? 0.f : tm.ms(5, 6);
        t->expand_ms = bin_mode ? tm.ms(5, 6) : 0.f;
        if (bin_mode && Rc) {
            t->expand_count_ms = tm.sort_ms(12, 13);
            t->expand_fill_ms = tm.sort_ms(14, 15);
        }
        t->num_coarse = (int)Rc;
        t->binning_mode = bin_mode ? 0 : 1;
        t->blend_ms = tm.ms(6, 7);
        t->total_ms = tm.ms(0, 7);
        t->num_rendered = (int)R;
        t->depth_passes = depth_passes;
        t->sort_passes = depth_passes + tile_passes;
        t->kernel_launches = launches;
        t->sort_hist_ms = tm.sort_ms(0, 1) + tm.sort_ms(6, 7);
        for (int i = 0; i < depth_passes && i < 4; ++i) t->sort_pass_ms[i] = tm.sort_ms(1 + i, 2 + i);
        for (int i = 0; i < tile_passes && depth_passes + i < 8; ++i)
            t->sort_pass_ms[depth_passes + i] = tm.sort_ms(7 + i, 8 + i);
        if (count)
            GSR_CUDA_TRY(cudaMemcpy(t->blend_counters, img.blend_counters, GSR_BLEND_COUNTERS * sizeof(unsigned long long),
                                    cudaMemcpyDeviceToHost));
        if ((rc = take_async_error(slot)) < 0) return rc;  // the stream is idle: this call's own passes have reported
    }
    return (int)R;
}
#undef GSR_STAGE

extern "C" {

static void fill_args(gsr_forward_args& x, gsr_alloc_fn ga, void* gu, gsr_alloc_fn ba, void* bu, gsr_alloc_fn ia,
                      void* iu, int P, int D, int M, const float* background, int width, int height,
                      const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
                      const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
                      const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                      float tan_fovy, int prefiltered, float* out_color, int* radii, int* rects, const float* boxmin,
                      const float* boxmax, void* stream) {
    memset(&x, 0, sizeof(x));
    x.geometry_alloc = ga; x.geometry_user = gu; x.binning_alloc = ba; x.binning_user = bu;
    x.image_alloc = ia; x.image_user = iu;
    x.P = P; x.D = D; x.M = M; x.background = background; x.width = width; x.height = height;
    x.means3D = means3D; x.shs = shs; x.colors_precomp = colors_precomp; x.opacities = opacities;
    x.scales = scales; x.scale_modifier = scale_modifier; x.rotations = rotations;
    x.cov3D_precomp = cov3D_precomp; x.viewmatrix = viewmatrix; x.projmatrix = projmatrix; x.cam_pos = cam_pos;
    x.tan_fovx = tan_fovx; x.tan_fovy = tan_fovy; x.prefiltered = prefiltered; x.out_color = out_color;
    x.radii = radii; x.rects = rects; x.boxmin = boxmin; x.boxmax = boxmax; x.stream = stream;
}

int gsr_forward(gsr_alloc_fn ga, void* gu, gsr_alloc_fn ba, void* bu, gsr_alloc_fn ia, void* iu, int P, int D, int M,
                const float* background, int width, int height, const float* means3D, const float* shs,
                const float* colors_precomp, const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered, float* out_color, int* radii,
                int* rects, const float* boxmin, const float* boxmax, void* stream) {
    gsr_forward_args x;
    fill_args(x, ga, gu, ba, bu, ia, iu, P, D, M, background, width, height, means3D, shs, colors_precomp, opacities,
              scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy,
              prefiltered, out_color, radii, rects, boxmin, boxmax, stream);
    x.means_stride = 3;
    x.scales_stride = 3;
    return gsr_forward_ex(&x);
}

int gsr_forward_gscuda(gsr_alloc_fn ga, void* gu, gsr_alloc_fn ba, void* bu, gsr_alloc_fn ia, void* iu, int P, int D,
                       int M, const float* background, int width, int height, const float* means3D, const float* shs,
                       const float* colors_precomp, const float* opacities, const float* scales,
                       float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                       float tan_fovy, int prefiltered, float* out_color, int* radii, int* rects, const float* boxmin,
                       const float* boxmax, void* stream) {
    gsr_forward_args x;
    fill_args(x, ga, gu, ba, bu, ia, iu, P, D, M, background, width, height, means3D, shs, colors_precomp, opacities,
              scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy,
              prefiltered, out_color, radii, rects, boxmin, boxmax, stream);
    x.means_stride = 4;  // GSGaussians.cpp:121-125: vec4 positions and scales
    x.scales_stride = 4;
    x.flags = GSR_FLAG_GSRAST_COMPAT;
    return gsr_forward_ex(&x);
}

size_t gsr_sort_pairs_temp_bytes(size_t n) { return sort_temp_bytes(n); }

int gsr_sort_pairs(uint64_t* keys_in, uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out, size_t n, int end_bit,
                   char* temp, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    bool in_a = false;
    // the watchdog reports into the calling thread's slot: asynchronous, surfaced by this thread's next call
    int rc = ensure_slot(g_slot);
    if (rc < 0) return rc;
    if ((rc = take_async_error(g_slot)) < 0) return rc;
    rc = launch_sort_pairs(keys_in, vals_in, keys_out, vals_out, n, end_bit, temp, &in_a, s, nullptr,
                           g_slot.dev + SLOT_ERROR);
    if (rc < 0) return rc;
    if (in_a && n > 0) {
        GSR_CUDA_TRY(cudaMemcpyAsync(keys_out, keys_in, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
        GSR_CUDA_TRY(cudaMemcpyAsync(vals_out, vals_in, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    }
    return rc;
}

int gsr_identify_tile_ranges(const uint64_t* sorted_keys, size_t n, uint32_t* ranges, int num_tiles, unsigned flags,
                             void* stream) {
    return launch_identify_ranges(sorted_keys, n, ranges, num_tiles, (flags & GSR_FLAG_GSRAST_COMPAT) != 0,
                                  static_cast<cudaStream_t>(stream));
}

const char* gsr_error_string(int code) {
    if (code >= 0) return "success";
    switch (code) {
        case GSR_ERR_INVALID_ARG: return "gsrast_b200: invalid argument";
        case GSR_ERR_ALLOC_FAILED: return "gsrast_b200: scratch allocator returned NULL";
        case GSR_ERR_TOO_MANY_PAIRS: return "gsrast_b200: num_rendered >= 2^30 (or the image has more than 2^23 tiles)";
        case GSR_ERR_SORT_STALLED: return "gsrast_b200: radix sort look-back watchdog tripped";
        case GSR_ERR_PLY_OPEN: return "gsrast_b200: cannot open the .ply file";
        case GSR_ERR_PLY_FORMAT: return "gsrast_b200: .ply header has no vertex count on line 3 or no end_header";
        case GSR_ERR_PLY_TRUNCATED: return "gsrast_b200: .ply body is shorter than the header's vertex count";
        default: return cudaGetErrorString(static_cast<cudaError_t>(-code));
    }
}

int gsr_version(void) { return GSR_VERSION; }

}  // extern "C"
