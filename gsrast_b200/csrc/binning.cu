// binning.cu — scan of tiles_touched, tile-key duplication and tile-range identification.
//
// Replaces (file:line in /root/reference/apps/gsrast/gscuda):
//   cub::DeviceScan::InclusiveSum + the 4-byte D2H readback   GSCuda.cu:771-772
//   duplicateWithKeys                                          GSCuda.cu:422-475 (launch :787)
//   cudaMemset(ranges) + identifyTileRanges                    GSCuda.cu:800-801, 504-538
//
// num_rendered comes from per-block sums (written by preprocess) and a single-CTA scan of those
// sums (scan_block_sums_kernel, which publishes the total straight into mapped pinned host memory).
// Duplication runs on the Gaussians in DEPTH order (they are radix-sorted by their depth key while
// the host waits for num_rendered, see radix_sort.cu): their tile rects are gathered into that order
// and the per-block pair counts scanned in the same window; the duplication kernel itself is
// load-balanced (the 256 Gaussians of a block pool their tile counts) and emits 32-bit tile keys.
#include "gsr_common.cuh"

namespace gsr {

namespace {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 16;
constexpr int DUP_ITEMS = 4;  // consecutive outputs per thread and window in the duplication kernel
constexpr int DUP_GPT = 4;    // Gaussians (consecutive depth ranks) per thread of a duplication block
constexpr int DUP_GAUSS = DUP_GPT * PRE_THREADS;

// Exclusive scan of block_sums[n] in place; total -> *total_dev and *total_host (mapped).
// One CTA; every thread owns SCAN_ITEMS consecutive entries (vector loads), so up to 16K entries
// (4.2 M Gaussians) take a single round of one warp scan + one cross-warp scan.
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(uint32_t* __restrict__ block_sums, int n,
                                                                       uint32_t* __restrict__ total_dev,
                                                                       volatile uint32_t* total_host,
                                                                       const uint32_t* __restrict__ sums2) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s_total2;
    __shared__ unsigned long long s_total64;  // the same total in 64 bits: tells a wrapped 32-bit sum from a small one
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long my64 = 0;
    if (tid == 0) { s_carry = 0; s_total2 = 0; s_total64 = 0; }
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    __syncthreads();
    if (sums2) {  // plain reduction of the second array (its loads overlap the scan below)
        uint32_t t2 = 0;
        for (int i = tid; i < n; i += SCAN_THREADS) t2 += __ldg(sums2 + i);
        t2 = __reduce_add_sync(0xffffffffu, t2);
        if (lane == 0 && t2) atomicAdd(&s_total2, t2);
    }
    constexpr int CHUNK = SCAN_THREADS * SCAN_ITEMS;
    for (int base = 0; base < n; base += CHUNK) {
        const int i0 = base + tid * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS];
        if (i0 + SCAN_ITEMS <= n) {
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
                const uint4 x = reinterpret_cast<const uint4*>(block_sums + i0)[q];
                v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j) v[j] = (i0 + j < n) ? block_sums[i0 + j] : 0u;
        }
        uint32_t tsum = 0;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) { tsum += v[j]; my64 += v[j]; }
        uint32_t incl = tsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += t;
            }
            s_warp[lane] = wi - w;  // exclusive warp offsets
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        uint32_t run = carry + s_warp[warp] + incl - tsum;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) {
            const uint32_t x = v[j];
            v[j] = run;
            run += x;
        }
        if (i0 + SCAN_ITEMS <= n) {
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q)
                reinterpret_cast<uint4*>(block_sums + i0)[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j)
                if (i0 + j < n) block_sums[i0 + j] = v[j];
        }
        __syncthreads();
        if (tid == SCAN_THREADS - 1) s_carry = run;
        __syncthreads();
    }
    if (total_host) {  // uniform
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) my64 += __shfl_xor_sync(0xffffffffu, my64, d);
        if (lane == 0 && my64) atomicAdd(&s_total64, my64);
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t total = s_carry;
        *total_dev = total;
        if (total_host) {
            if (sums2) total_host[SLOT_RC] = s_total2;
            // a sum beyond the sort's 30-bit counters is published as 0xffffffff (GSR_ERR_TOO_MANY_PAIRS on the host)
            total_host[SLOT_R] = s_total64 >= (1ull << 30) ? 0xffffffffu : total;
            __threadfence_system();
        }
    }
}

// Two independent jobs of the window in which the host waits for num_rendered, in one launch:
//  (a) gather the tile rects into depth order (one 8-byte gather per Gaussian, then everything
//      downstream is coalesced) and leave the pair count of every duplication block (DUP_GAUSS depth
//      ranks) for the scan;
//  (b) materialise point_offsets = inclusive scan of tiles_touched in INDEX order — the array
//      cub::DeviceScan::InclusiveSum leaves in pointOffsets (GSCuda.cu:771), which the Inspector reads
//      (Inspector.cpp:174-188).  The pipeline itself consumes the scan in depth order, so this is on
//      the side.  `block_offsets` are the exclusive offsets of preprocess' 256-Gaussian blocks.
template <bool COARSE>
__global__ void __launch_bounds__(PRE_THREADS) gather_rects_kernel(const int P, const uint32_t* __restrict__ sorted_ids,
                                                                   const uint2* __restrict__ tile_rects,
                                                                   uint2* __restrict__ sorted_rects,
                                                                   uint32_t* __restrict__ block_sums,
                                                                   const uint32_t* __restrict__ tiles_touched,
                                                                   const uint32_t* __restrict__ block_offsets,
                                                                   uint32_t* __restrict__ point_offsets,
                                                                   const uint32_t* __restrict__ n_sorted) {
    __shared__ uint32_t s_warp[PRE_THREADS / 32];
    __shared__ uint32_t s_pw[PRE_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    // the depth sort drops the Gaussians that emit nothing: only the first Pn depth ranks exist
    const int Pn = n_sorted ? min(P, (int)__ldg(n_sorted)) : P;
    // (a)
    uint32_t cnt = 0;
#pragma unroll
    for (int c = 0; c < DUP_GPT; ++c) {
        const int i = blockIdx.x * DUP_GAUSS + c * PRE_THREADS + tid;
        if (i < Pn) {
            // the tile rect preprocess computed with getRect (GSCuda.cu:237-259; duplicateWithKeys recomputes
            // the same rect, :445-458).  Gaussians that emit nothing (radii <= 0, :440-443) carry an empty rect.
            uint2 rec = __ldg(tile_rects + __ldg(sorted_ids + i));
            if (COARSE) rec = coarse_rect(rec);  // bin expansion: the duplication emits (bin, Gaussian) records
            sorted_rects[i] = rec;
            cnt += (rec.y >> 16) * (rec.y & 0xffffu);
        }
    }
    // (b) thread t owns indices 4t..4t+3 of the block's 1024; 64 threads = one preprocess block.  Skipped (block
    // uniform) when the caller does not want point_offsets (GSR_FLAG_LEAN_STATE): nothing in the pipeline reads it.
    const int i0 = blockIdx.x * DUP_GAUSS + 4 * tid;
    uint32_t tt[4] = {0u, 0u, 0u, 0u};
    if (!point_offsets) {
    } else if (i0 + 4 <= P) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(tiles_touched + i0));
        tt[0] = x.x; tt[1] = x.y; tt[2] = x.z; tt[3] = x.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j < P) tt[j] = __ldg(tiles_touched + i0 + j);
    }
    const uint32_t tsum = tt[0] + tt[1] + tt[2] + tt[3];
    uint32_t incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_pw[warp] = incl;
    const uint32_t wsum = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) s_warp[warp] = wsum;
    __syncthreads();
    if (point_offsets && i0 < P) {
        uint32_t run = __ldg(block_offsets + (i0 / PRE_THREADS)) + ((warp & 1) ? s_pw[warp - 1] : 0u) + incl - tsum;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            run += tt[j];
            o[j] = run;
        }
        if (i0 + 4 <= P) {
            reinterpret_cast<uint4*>(point_offsets + i0)[0] = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i0 + j < P) point_offsets[i0 + j] = o[j];
        }
    }
    if (tid == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < PRE_THREADS / 32; ++w) t += s_warp[w];
        block_sums[blockIdx.x] = t;
    }
}

// Duplication in depth order.  Block b takes the DUP_GAUSS Gaussians at depth ranks DUP_GAUSS*b ..,
// (thread t owns ranks 4t..4t+3 of the block), scans their tile counts and emits every (tile, Gaussian)
// pair behind the block's scanned offset: rows outer, columns inner (GSCuda.cu:461-474).  The work is
// pooled: outputs are produced in windows of 4*256, thread t producing outputs 4t..4t+3 of the window (one
// binary search over the pooled counts, then a walk along the rect rows); the window is transposed
// through shared memory so the stores are fully coalesced and one huge splat cannot serialise a
// thread.  The digit histograms of the tile passes are counted here (shared-memory atomics, flushed once
// per block), so the sort never re-reads the keys.
//
// FUSED (GSR_FLAG_LEAN_STATE callers, which do not need point_offsets): the kernel also does what gather_rects_kernel
// and the single-CTA scan behind it do — it gathers the tile rects of its depth ranks itself (`sorted_rects` is then
// tile_rects indexed by Gaussian id, `coarse` turns them into bin rects) and finds its output offset by decoupled
// look-back over the pair counts of the blocks before it: `fuse_state` = one 64-bit word per block (count << 2 | flag;
// flag 1 = this block's count, 2 = inclusive prefix) followed by the ticket counter that hands out block numbers in
// start order (a block only ever waits for blocks that have started), all zeroed by the caller.
constexpr unsigned long long FUSE_AGG = 1ull, FUSE_INC = 2ull;
template <bool FUSED>
__global__ void __launch_bounds__(PRE_THREADS) duplicate_sorted_kernel(
    const int P_all, const int grid_x, const uint32_t* __restrict__ sorted_ids, const uint2* __restrict__ sorted_rects,
    const uint32_t* __restrict__ block_offsets, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
    uint32_t* __restrict__ hist, const int tile_bits, const uint32_t* __restrict__ n_sorted, const int coarse,
    unsigned long long* __restrict__ fuse_state, const int fuse_blocks, uint32_t* error_flag, const int presorted,
    const int self_offsets) {
    __shared__ uint32_t s_excl[DUP_GAUSS + 1];
    __shared__ __align__(16) uint32_t s_bpart[PRE_THREADS / 32];  // 16-byte aligned: read with LDS.128 like s_warp below
    // 16-byte aligned: the compiler reads the 8 warp sums with LDS.128, which otherwise straddles s_excl[DUP_GAUSS]
    // (unused lane of the vector, but compute-sanitizer racecheck rightly flags the overlap with its later store)
    __shared__ __align__(16) uint32_t s_warp[PRE_THREADS / 32];
    __shared__ uint32_t s_gid[DUP_GAUSS];
    __shared__ uint32_t s_origin[DUP_GAUSS];  // miny << 16 | minx
    __shared__ uint32_t s_width[DUP_GAUSS];
    __shared__ uint32_t s_hist[4 * 256];
    __shared__ __align__(16) uint32_t s_okey[DUP_ITEMS * PRE_THREADS];
    __shared__ __align__(16) uint32_t s_oval[DUP_ITEMS * PRE_THREADS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int passes = (tile_bits + 7) >> 3;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    __shared__ uint32_t s_block;
    int block = blockIdx.x;
    if (FUSED) {
        if (tid == 0) s_block = atomicAdd(reinterpret_cast<uint32_t*>(fuse_state + fuse_blocks), 1u);
        __syncthreads();
        block = (int)s_block;
    }
    const int i0 = block * DUP_GAUSS + DUP_GPT * tid;
    const int P = n_sorted ? min(P_all, (int)__ldg(n_sorted)) : P_all;  // depth ranks that exist
    if (block * DUP_GAUSS >= P) return;  // (fused: nothing before this block ever looks at it)
    // every global load of the block is issued here, before anything waits
    // self_offsets: `block_offsets` holds the UNSCANNED pair counts of the duplication blocks and every block adds up the
    // counts before its own (a few coalesced loads per thread, in flight with the rect loads below) — the single-CTA scan
    // between gather_rects and this kernel leaves the frame's dependency chain (7 us + a launch at C2)
    uint32_t boff = (FUSED || self_offsets) ? 0u : __ldg(block_offsets + block);
    uint32_t bpart = 0;
    if (!FUSED && self_offsets)
        for (int j = tid; j < block; j += PRE_THREADS) bpart += __ldg(block_offsets + j);
    uint2 rec[DUP_GPT];
    uint32_t gid[DUP_GPT];
    if (FUSED && presorted) {
        // the rects arrive in depth order and final units (the last depth pass of the sort wrote them): coalesced loads,
        // offsets by look-back below — no gather, no separate scan
        if (i0 + DUP_GPT <= P) {
            const uint4 r01 = __ldg(reinterpret_cast<const uint4*>(sorted_rects + i0));
            const uint4 r23 = __ldg(reinterpret_cast<const uint4*>(sorted_rects + i0) + 1);
            const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(sorted_ids + i0));
            rec[0] = make_uint2(r01.x, r01.y); rec[1] = make_uint2(r01.z, r01.w);
            rec[2] = make_uint2(r23.x, r23.y); rec[3] = make_uint2(r23.z, r23.w);
            gid[0] = g4.x; gid[1] = g4.y; gid[2] = g4.z; gid[3] = g4.w;
        } else {
#pragma unroll
            for (int c = 0; c < DUP_GPT; ++c) {
                const bool ok = i0 + c < P;
                rec[c] = ok ? __ldg(sorted_rects + i0 + c) : make_uint2(0u, 0u);
                gid[c] = ok ? __ldg(sorted_ids + i0 + c) : 0u;
            }
        }
    } else if (FUSED) {
        if (i0 + DUP_GPT <= P) {
            const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(sorted_ids + i0));
            gid[0] = g4.x; gid[1] = g4.y; gid[2] = g4.z; gid[3] = g4.w;
#pragma unroll
            for (int c = 0; c < DUP_GPT; ++c) rec[c] = __ldg(sorted_rects + gid[c]);
        } else {
#pragma unroll
            for (int c = 0; c < DUP_GPT; ++c) {
                const bool ok = i0 + c < P;
                gid[c] = ok ? __ldg(sorted_ids + i0 + c) : 0u;
                rec[c] = ok ? __ldg(sorted_rects + gid[c]) : make_uint2(0u, 0u);
            }
        }
        if (coarse) {
#pragma unroll
            for (int c = 0; c < DUP_GPT; ++c) rec[c] = coarse_rect(rec[c]);
        }
    } else if (i0 + DUP_GPT <= P) {
        const uint4 r01 = __ldg(reinterpret_cast<const uint4*>(sorted_rects + i0));
        const uint4 r23 = __ldg(reinterpret_cast<const uint4*>(sorted_rects + i0) + 1);
        const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(sorted_ids + i0));
        rec[0] = make_uint2(r01.x, r01.y); rec[1] = make_uint2(r01.z, r01.w);
        rec[2] = make_uint2(r23.x, r23.y); rec[3] = make_uint2(r23.z, r23.w);
        gid[0] = g4.x; gid[1] = g4.y; gid[2] = g4.z; gid[3] = g4.w;
    } else {
#pragma unroll
        for (int c = 0; c < DUP_GPT; ++c) {
            const bool ok = i0 + c < P;
            rec[c] = ok ? __ldg(sorted_rects + i0 + c) : make_uint2(0u, 0u);
            gid[c] = ok ? __ldg(sorted_ids + i0 + c) : 0u;
        }
    }
    for (int j = tid; j < passes * 256; j += PRE_THREADS) s_hist[j] = 0;

    uint32_t cnt[DUP_GPT], tsum = 0;
#pragma unroll
    for (int c = 0; c < DUP_GPT; ++c) {
        cnt[c] = (rec[c].y >> 16) * (rec[c].y & 0xffffu);
        tsum += cnt[c];
    }
    uint32_t incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    if (!FUSED && self_offsets) {
        bpart = __reduce_add_sync(0xffffffffu, bpart);
        if (lane == 0) s_bpart[warp] = bpart;
    }
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < PRE_THREADS / 32; ++w) {
        if (w < warp) woff += s_warp[w];
        total += s_warp[w];
    }
    if (!FUSED && self_offsets) {
#pragma unroll
        for (int w = 0; w < PRE_THREADS / 32; ++w) boff += s_bpart[w];
    }
    if (FUSED) {
        // decoupled look-back (warp 0): publish this block's count, add up the counts of the blocks before it until one
        // of them carries an inclusive prefix, publish the own inclusive prefix.  One 64-bit word per block, so count
        // and flag arrive together; volatile loads + relaxed stores, __threadfence orders nothing else that matters
        // (the words carry their whole payload).
        __shared__ uint32_t s_boff;
        if (warp == 0) {
            volatile unsigned long long* st = fuse_state;
            if (lane == 0) st[block] = ((unsigned long long)total << 2) | (block == 0 ? FUSE_INC : FUSE_AGG);
            uint32_t excl = 0;
            int look = block - 1;
            while (look >= 0) {
                const int j = look - lane;
                unsigned long long v = FUSE_INC;  // lanes before block 0: an inclusive prefix of zero
                if (j >= 0) {
                    // blocks are ticketed, so block j has started; bounded like the sort's look-back (radix_sort.cu)
                    uint32_t spins = 0;
                    while (((v = st[j]) & 3ull) == 0ull) {
                        if (++spins > (1u << 22)) {
                            gsr_raise_error(error_flag);
                            v = FUSE_INC;  // give up: the frame is invalid, the call reports GSR_ERR_SORT_STALLED
                            break;
                        }
                        __nanosleep(32);
                    }
                }
                const unsigned inc = __ballot_sync(0xffffffffu, (v & 3ull) == FUSE_INC);
                const int first = inc ? (__ffs((int)inc) - 1) : 32;  // nearest block holding an inclusive prefix
                const uint32_t part = (lane <= first) ? (uint32_t)(v >> 2) : 0u;
                excl += __reduce_add_sync(0xffffffffu, part);
                if (inc) break;
                look -= 32;
            }
            if (lane == 0) {
                if (block != 0) st[block] = ((unsigned long long)(excl + total) << 2) | FUSE_INC;
                s_boff = excl;
            }
        }
        __syncthreads();
        boff = s_boff;
    }
    if (total == 0) return;  // the culled tail of the depth order emits nothing (block-uniform)
    {
        uint32_t run = woff + incl - tsum;
#pragma unroll
        for (int c = 0; c < DUP_GPT; ++c) {
            const int q = DUP_GPT * tid + c;
            s_excl[q] = run;
            run += cnt[c];
            s_gid[q] = gid[c];
            s_origin[q] = rec[c].x;
            s_width[q] = rec[c].y & 0xffffu;
        }
        if (tid == PRE_THREADS - 1) s_excl[DUP_GAUSS] = run;
    }
    __syncthreads();

    const uint32_t lowmask = (1u << min(8, tile_bits)) - 1u;
    for (uint32_t kb = 0; kb < total; kb += DUP_ITEMS * PRE_THREADS) {
        const uint32_t k0 = kb + DUP_ITEMS * tid;
        uint32_t okey[DUP_ITEMS], oval[DUP_ITEMS];
#pragma unroll
        for (int j = 0; j < DUP_ITEMS; ++j) okey[j] = oval[j] = 0;
        if (k0 < total) {
            // largest j with s_excl[j] <= k0
            int lo = 0, hi = DUP_GAUSS - 1;
#pragma unroll
            for (int it = 0; it < 10; ++it) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_excl[mid] <= k0) lo = mid; else hi = mid - 1;
            }
            uint32_t w = s_width[lo], org = s_origin[lo], g = s_gid[lo], next = s_excl[lo + 1];
            const uint32_t t = k0 - s_excl[lo];
            const uint32_t ty = t / w;
            uint32_t tx = t - ty * w;
            uint32_t row_tile = ((org >> 16) + ty) * (uint32_t)grid_x + (org & 0xffffu);
            uint32_t run_d = 0xffffffffu, run_n = 0;  // run of equal second digits inside this thread
#pragma unroll
            for (int j = 0; j < DUP_ITEMS; ++j) {
                const uint32_t k = k0 + j;
                if (k < total) {
                    if (k >= next) {  // next Gaussian that emits anything (empty ones have equal offsets)
                        do { ++lo; next = s_excl[lo + 1]; } while (k >= next);
                        w = s_width[lo]; org = s_origin[lo]; g = s_gid[lo];
                        tx = 0;
                        row_tile = (org >> 16) * (uint32_t)grid_x + (org & 0xffffu);
                    }
                    // key = tile id (GSCuda.cu:466-471: the depth half is re-attached by the last sort pass)
                    const uint32_t tile = row_tile + tx;
                    okey[j] = tile; oval[j] = g;
                    if (++tx == w) { tx = 0; row_tile += (uint32_t)grid_x; }
                    atomicAdd(&s_hist[tile & lowmask], 1u);
                    if (passes > 1) {
                        // a row of tiles shares the higher digits: count runs, not items
                        const uint32_t d = (tile >> 8) & ((1u << min(8, tile_bits - 8)) - 1u);
                        if (d != run_d) {
                            if (run_n) atomicAdd(&s_hist[256 + run_d], run_n);
                            run_d = d; run_n = 0;
                        }
                        ++run_n;
                        for (int ps = 2; ps < passes; ++ps)
                            atomicAdd(&s_hist[ps * 256 + ((tile >> (8 * ps)) & ((1u << min(8, tile_bits - 8 * ps)) - 1u))], 1u);
                    }
                }
            }
            if (run_n) atomicAdd(&s_hist[256 + run_d], run_n);
        }
        reinterpret_cast<uint4*>(s_okey)[tid] = make_uint4(okey[0], okey[1], okey[2], okey[3]);
        reinterpret_cast<uint4*>(s_oval)[tid] = make_uint4(oval[0], oval[1], oval[2], oval[3]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < DUP_ITEMS; ++j) {
            const uint32_t q = tid + j * PRE_THREADS;
            if (kb + q < total) {
                const size_t o = (size_t)boff + kb + q;
                keys_out[o] = s_okey[q];
                vals_out[o] = s_oval[q];
            }
        }
        __syncthreads();
    }
    for (int j = tid; j < passes * 256; j += PRE_THREADS) {
        const uint32_t c = s_hist[j];
        if (c) atomicAdd(hist + j, c);
    }
}

// Four consecutive sorted keys per thread (two 16-byte loads + the left neighbour): boundary
// detection in the tile-id half of the key.  ranges[] must be zeroed beforehand.
template <bool COMPAT>
__global__ void __launch_bounds__(256) identify_ranges_kernel(const size_t n, const uint64_t* __restrict__ keys,
                                                              uint2* __restrict__ ranges) {
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    if (i0 >= n) return;
    uint64_t k[4];
    if (i0 + 3 < n) {
        const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(keys + i0));
        const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2*>(keys + i0 + 2));
        k[0] = a.x; k[1] = a.y; k[2] = b.x; k[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) k[j] = (i0 + j < n) ? __ldg(keys + i0 + j) : 0ull;
    }
    uint32_t prev = (i0 > 0) ? (uint32_t)(__ldg(keys + i0 - 1) >> 32) : 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const size_t idx = i0 + j;
        if (idx >= n) break;
        const uint32_t cur = (uint32_t)(k[j] >> 32);
        if (idx == 0) {
            ranges[cur].x = 0;
        } else {
            if (prev != cur) {
                ranges[prev].y = (uint32_t)idx;
                ranges[cur].x = (uint32_t)idx;
            }
            if (COMPAT && idx == n - 1) ranges[cur].y = (uint32_t)n;  // GSCuda.cu:533-536 (inside the else)
        }
        if (!COMPAT && idx == n - 1) ranges[cur].y = (uint32_t)n;
        prev = cur;
    }
}

}  // namespace

int launch_scan_block_sums(uint32_t* block_sums, int num_blocks, uint32_t* total_dev, uint32_t* total_host_mapped,
                           cudaStream_t s, const uint32_t* sums2) {
    cudaError_t e = launch_pdl(scan_block_sums_kernel, dim3(1), dim3(SCAN_THREADS), 0, s, block_sums, num_blocks, total_dev,
                               (volatile uint32_t*)total_host_mapped, sums2);
    return e == cudaSuccess ? 1 : -(int)e;
}

int num_dup_blocks(int P) { return ((P > 0 ? P : 0) + DUP_GAUSS - 1) / DUP_GAUSS; }

int launch_gather_rects(int P, const uint32_t* sorted_ids, const uint32_t* tile_rects, uint32_t* sorted_rects,
                        uint32_t* block_sums, const uint32_t* tiles_touched, const uint32_t* block_offsets,
                        uint32_t* point_offsets, bool coarse, cudaStream_t s, const uint32_t* n_sorted) {
    if (P <= 0) return 0;
    const int blocks = num_dup_blocks(P);
    cudaError_t e =
        coarse ? launch_pdl(gather_rects_kernel<true>, dim3(blocks), dim3(PRE_THREADS), 0, s, P, sorted_ids,
                            reinterpret_cast<const uint2*>(tile_rects), reinterpret_cast<uint2*>(sorted_rects),
                            block_sums, tiles_touched, block_offsets, point_offsets, n_sorted)
               : launch_pdl(gather_rects_kernel<false>, dim3(blocks), dim3(PRE_THREADS), 0, s, P, sorted_ids,
                            reinterpret_cast<const uint2*>(tile_rects), reinterpret_cast<uint2*>(sorted_rects),
                            block_sums, tiles_touched, block_offsets, point_offsets, n_sorted);
    return e == cudaSuccess ? 1 : -(int)e;
}

int launch_duplicate_sorted(int P, int grid_x, const uint32_t* sorted_ids, const uint32_t* sorted_rects,
                            const uint32_t* block_offsets, uint32_t* keys32_out, uint32_t* vals_out, uint32_t* hist,
                            int tile_bits, cudaStream_t s, const uint32_t* n_sorted, bool self_offsets) {
    if (P <= 0) return 0;
    if (tile_bits < 1 || tile_bits > 32) return GSR_ERR_INVALID_ARG;
    const int blocks = num_dup_blocks(P);
    GSR_CARVEOUT(duplicate_sorted_kernel<false>, "DUP", -1);
    duplicate_sorted_kernel<false><<<blocks, PRE_THREADS, 0, s>>>(
        P, grid_x, sorted_ids, reinterpret_cast<const uint2*>(sorted_rects), block_offsets, keys32_out, vals_out, hist,
        tile_bits, n_sorted, 0, nullptr, 0, nullptr, 0, self_offsets ? 1 : 0);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

size_t duplicate_fused_state_bytes(int P) { return ((size_t)num_dup_blocks(P) + 1) * sizeof(unsigned long long); }

int launch_duplicate_fused(int P, int grid_x, const uint32_t* sorted_ids, const uint32_t* tile_rects, bool coarse,
                           void* fuse_state, uint32_t* keys32_out, uint32_t* vals_out, uint32_t* hist, int tile_bits,
                           cudaStream_t s, const uint32_t* n_sorted, uint32_t* error_flag, bool rects_presorted) {
    if (P <= 0) return 0;
    if (tile_bits < 1 || tile_bits > 32 || !fuse_state) return GSR_ERR_INVALID_ARG;
    const int blocks = num_dup_blocks(P);
    GSR_CARVEOUT(duplicate_sorted_kernel<true>, "DUP", -1);
    duplicate_sorted_kernel<true><<<blocks, PRE_THREADS, 0, s>>>(
        P, grid_x, sorted_ids, reinterpret_cast<const uint2*>(tile_rects), nullptr, keys32_out, vals_out, hist, tile_bits,
        n_sorted, coarse ? 1 : 0, static_cast<unsigned long long*>(fuse_state), blocks, error_flag, rects_presorted ? 1 : 0, 0);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

int launch_identify_ranges(const uint64_t* keys, size_t n, uint32_t* ranges, int num_tiles, bool compat,
                           cudaStream_t s, bool zero_first) {
    cudaError_t e = cudaSuccess;
    if (zero_first) e = cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)num_tiles, s);
    if (e != cudaSuccess) return -(int)e;
    if (n == 0) return 0;
    const unsigned blocks = (unsigned)((n + 1023) / 1024);
    if (compat)
        e = launch_pdl(identify_ranges_kernel<true>, dim3(blocks), dim3(256), 0, s, n, keys, reinterpret_cast<uint2*>(ranges));
    else
        e = launch_pdl(identify_ranges_kernel<false>, dim3(blocks), dim3(256), 0, s, n, keys, reinterpret_cast<uint2*>(ranges));
    return e == cudaSuccess ? 1 : -(int)e;
}

}  // namespace gsr
