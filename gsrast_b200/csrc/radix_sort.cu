// radix_sort.cu — in-house stable LSD radix sort, "onesweep" style: digit histograms up front,
// then one read+write pass per 8-bit digit with the inter-tile prefix resolved by decoupled
// look-back.  No CUB / Thrust.
//
// Replaces cub::DeviceRadixSort::SortPairs as called at
// /root/reference/apps/gsrast/gscuda/GSCuda.cu:794-797 (temp-size probe: AuxBuffer.cu:83-85):
// ascending, stable, over key bits [0, end_bit) of (tile << 32 | depth bits) with
// end_bit = 32 + getHigherMsb(tiles).  Stability is what makes the sorted (key, value) lists
// unique — equal (tile, depth) keys keep their emission order, i.e. ascending Gaussian index —
// so the output is bit-identical to the reference's by construction.
//
// The forward path runs this LSD sort SPLIT in two (forward.cu):
//   * the four depth-digit passes run on P Gaussians (32-bit depth keys, values = Gaussian
//     index generated on the fly), BEFORE duplication, because the low 32 key bits are a
//     per-Gaussian quantity; duplication then emits tile-Gaussian pairs in that order;
//   * only the tile-digit passes (ceil(msb/8) = 2 at 720p..4K) run on the R pairs, on 32-bit
//     tile keys; the last one re-attaches the depth bits and writes the 64-bit sorted keys.
// An LSD sort applies its digit passes in order of significance and every pass is stable, so
// sorting P records by depth, expanding each record into its pairs in place, and continuing
// with the tile digits yields exactly the list the 6-pass sort of R pairs yields.
// The generic 64-bit entry (gsr_sort_pairs) is the same kernel instantiated for uint2 keys.
//
// Per pass and per tile of SORT_TILE items (one CTA of 256 threads, 16 items per thread):
//   1. coalesced warp-striped load of keys and values into registers; per-warp digit counts
//      with shared-memory atomics;
//   2. cross-warp exclusive prefix per digit -> tile histogram, published at once;
//   3. stable in-warp ranking: peers with the same digit find each other through an
//      atomicOr'd lane mask in shared memory; items go straight to their staged slot;
//   4. decoupled look-back over the per-tile status words (2 flag bits + 30-bit count); tiles
//      are handed out by an atomic ticket so a tile only ever waits for tiles that started;
//   5. keys/values are written out from shared memory as contiguous per-digit runs.
#include <algorithm>
#include <atomic>

#include "gsr_common.cuh"

namespace gsr {

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;  // == RADIX: thread d owns digit d in the scan / look-back steps
constexpr int SORT_WARPS = SORT_THREADS / 32;
// Items per thread: 16 (4096-pair tiles) for the big pair-level passes, 8 for the Gaussian-level
// depth passes, whose few hundred tiles would otherwise not fill the machine twice.
#ifndef GSR_ITEMS_LARGE
#define GSR_ITEMS_LARGE 16
#endif
#ifndef GSR_MINB_LARGE
#define GSR_MINB_LARGE 4
#endif
constexpr int ITEMS_LARGE = GSR_ITEMS_LARGE;
constexpr int ITEMS_SMALL = 8;
constexpr int MAX_PASSES = 8;
#ifndef GSR_LOOKBACK_W
#define GSR_LOOKBACK_W 8
#endif
constexpr int LOOKBACK_W = GSR_LOOKBACK_W;
// GSR_SORT_UNIFORM: the most significant depth pass counts and ranks whole-warp-uniform steps by vote (UNI below).
// Measured (profiles/r02r_ab_*.txt): that pass 0.033 -> 0.030 ms at C2, 0.044 -> 0.039 at C3.  The same vote in the
// histogram kernel (one add per warp for the top digit) made it SLOWER, 0.017 -> 0.020 ms, and was dropped.
#ifndef GSR_SORT_UNIFORM
#define GSR_SORT_UNIFORM 1
#endif

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_INC = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VAL_MASK = ~FLAG_MASK;
constexpr int GROUP_SHIFT = 4;  // look-back groups of 16 tiles (16 * 6144 items < 2^24: sum and arrivals share a word)
constexpr int GROUP_TILES = 1 << GROUP_SHIFT;

constexpr int HIST_THREADS = 256;
// Look-back watchdog: polls (32 ns sleeps) a waiting digit thread tolerates before it gives up and raises the call's
// sticky error (GSR_ERR_SORT_STALLED).  Tiles are ticketed, so a predecessor has always started; 2^22 polls = ~0.2 s.
// -DGSR_FORCE_STALL: test build in which tile 1 of every pass reports a stall unconditionally (tests/test_gpu_stall.py).
constexpr uint32_t WATCHDOG_SPINS = 1u << 22;

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- key types ------------------------------------------------------------------------------
// The digit of a pass never straddles the two 32-bit halves of a 64-bit key (shift is a multiple
// of 8 and the digit has <= 8 bits), so all digit arithmetic is 32-bit.
__device__ __forceinline__ uint32_t digit_of(const uint2 k, const int shift, const uint32_t dmask) {
    return ((shift >= 32 ? k.y : k.x) >> (shift & 31)) & dmask;
}
__device__ __forceinline__ uint32_t digit_of(const uint32_t k, const int shift, const uint32_t dmask) {
    return (k >> shift) & dmask;
}
__device__ __forceinline__ void pad_key(uint2& k) { k = make_uint2(~0u, ~0u); }
__device__ __forceinline__ void pad_key(uint32_t& k) { k = ~0u; }
__device__ __forceinline__ bool is_pad(const uint2 k) { return (k.x & k.y) == ~0u; }
__device__ __forceinline__ bool is_pad(const uint32_t k) { return k == ~0u; }
__device__ __forceinline__ uint32_t low_word(const uint2 k) { return k.x; }
__device__ __forceinline__ uint32_t low_word(const uint32_t k) { return k; }

// ---- up-front histograms of every digit place -------------------------------------------
// Block-shared counters updated with shared-memory atomics: measured on B200 (tools/microbench.cu)
// at ~2.4 SM-cycles per warp-wide ATOMS against ~60 for match.any and ~25 for an 8-step ballot
// match, so plain atomics are the right tool for an order-independent count.
// kept != nullptr: keys equal to the all-ones pad value are NOT counted (the depth sort of the forward path drops
// the Gaussians that emit nothing) and *kept receives the number of keys that were.
template <typename KeyT, int PASSES>
__global__ void __launch_bounds__(HIST_THREADS) histogram_kernel(const KeyT* __restrict__ keys, const size_t n,
                                                                 const int end_bit, uint32_t* __restrict__ hist,
                                                                 uint32_t* __restrict__ kept) {
    __shared__ uint32_t s_hist[PASSES * RADIX];
    __shared__ uint32_t s_kept;
    const int tid = threadIdx.x;
    if (tid == 0) s_kept = 0;
    uint32_t my_kept = 0;
    for (int i = tid; i < PASSES * RADIX; i += HIST_THREADS) s_hist[i] = 0;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    __syncthreads();

    constexpr int PER_THREAD = 8;
    const size_t per_block = (size_t)HIST_THREADS * PER_THREAD;
    for (size_t base = (size_t)blockIdx.x * per_block; base < n; base += (size_t)gridDim.x * per_block) {
        KeyT k[PER_THREAD];
#pragma unroll
        for (int i = 0; i < PER_THREAD; ++i) {
            const size_t pos = base + (size_t)i * HIST_THREADS + tid;
            if (pos < n) k[i] = __ldg(keys + pos);
            else pad_key(k[i]);
        }
#pragma unroll
        for (int i = 0; i < PER_THREAD; ++i) {
            if (base + (size_t)i * HIST_THREADS + tid < n && !(kept && is_pad(k[i]))) {
                ++my_kept;
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    const int shift = ps * RADIX_BITS;
                    const int nb = min(RADIX_BITS, end_bit - shift);
                    atomicAdd(&s_hist[ps * RADIX + digit_of(k[i], shift, (1u << nb) - 1u)], 1u);
                }
            }
        }
    }
    if (kept) {
        my_kept = __reduce_add_sync(0xffffffffu, my_kept);
        if ((tid & 31) == 0 && my_kept) atomicAdd(&s_kept, my_kept);
    }
    __syncthreads();
    for (int i = tid; i < PASSES * RADIX; i += HIST_THREADS) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(hist + i, c);
    }
    if (kept && tid == 0 && s_kept) atomicAdd(kept, s_kept);
}

// Status words carry flag and count in ONE 32-bit word, so relaxed gpu-scope accesses suffice.
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- one digit pass ------------------------------------------------------------------------
struct PassArgs {
    const void* keys_in;
    const uint32_t* vals_in;  // nullptr: the value of item i is i (first pass over an implicit iota)
    void* keys_out;           // KeyT[n]; unused when expand_low != nullptr
    uint32_t* vals_out;
    size_t n;
    const uint32_t* n_dev;    // non-null: the item count lives on the device (<= n); tiles past it exit at once
    int shift, nbits;
    const uint32_t* digit_counts;   // [RADIX] global digit histogram of this pass (every CTA scans it itself)
    uint32_t* status;               // [num_tiles][RADIX], zero-initialised
    uint32_t* gstat;                // [ceil(num_tiles/16)][RADIX], zero-initialised: arrivals << 24 | sum of counts
    uint32_t* ticket;
    uint32_t* error_flag;
    // last tile-digit pass of the forward path: key64 = key32 << 32 | expand_low[value]
    const uint32_t* expand_low;
    uint64_t* keys_out64;
    // RECTS (last depth pass of the forward path, lean callers): the pass also brings the tile rects into the sorted
    // order — rect_dst[g] = rect_src[value] for the item that lands at position g (coarse: in units of bins) — which is
    // what gather_rects_kernel did in a launch of its own; keys_out may then be null (nothing reads the sorted keys)
    const uint2* rect_src;
    uint2* rect_dst;
    int rect_coarse;
};
// UNI (template parameter of the pass, GSR_SORT_UNIFORM): the digit of the pass is expected to be the same for whole
// warps — the most significant byte of float depth keys takes 3-5 values — so a warp whose 32 items of a step agree
// counts and ranks them with one vote instead of 32 shared-memory atomics on one address.

// Shared memory: [SORT_TILE] key staging | [SORT_TILE] u32 value staging | per-warp digit counters
// [SORT_WARPS][RADIX] | per-warp match masks [SORT_WARPS][RADIX] | global bases [RADIX] | misc[16].
template <typename KeyT, int SORT_ITEMS>
constexpr size_t onesweep_smem() {
    return (size_t)SORT_THREADS * SORT_ITEMS * (sizeof(KeyT) + 4) + (size_t)(2 * SORT_WARPS * RADIX + RADIX + 32) * 4;
}

// DROP: items whose key is the all-ones pad value are neither counted nor written (the output is compacted);
// the tile then stages n_stage <= n_tile items.
template <typename KeyT, bool EXPAND, bool FULL, int SORT_ITEMS, bool DROP, bool RECTS, bool UNI = false>
__device__ __forceinline__ void onesweep_tile(const PassArgs& a, const uint32_t n_tile, const uint32_t tile,
                                              unsigned char* s_raw) {
    static_assert(!(DROP && FULL), "a dropping pass tests every item");
    static_assert(!(RECTS && EXPAND), "the rect gather belongs to the last depth pass, the key expansion to the last tile pass");
    constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
    KeyT* s_keys = reinterpret_cast<KeyT*>(s_raw);                                             // [SORT_TILE]
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_raw + (size_t)SORT_TILE * sizeof(KeyT));  // [SORT_TILE]
    uint32_t* s_whist = s_vals + SORT_TILE;                                                    // [2][SORT_WARPS][RADIX]
    uint32_t* s_global = s_whist + 2 * SORT_WARPS * RADIX;                                     // [RADIX]
    uint32_t* s_misc = s_global + RADIX;                                                       // [16]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t warp_base = warp * (32 * SORT_ITEMS);
    uint32_t* my_hist = s_whist + warp * RADIX;  // my_mask = my_hist + SORT_WARPS*RADIX
    const size_t tile_base = (size_t)tile * SORT_TILE;
    const KeyT* kin = reinterpret_cast<const KeyT*>(a.keys_in) + tile_base;
    const uint32_t dmask = (1u << a.nbits) - 1u;
    const int shift = a.shift;

    // 1. warp-striped loads (item i of this thread sits at warp_base + i*32 + lane) + early counts
    KeyT key[SORT_ITEMS];
    uint32_t val[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const uint32_t loc = warp_base + i * 32 + lane;
        if (FULL || loc < n_tile) key[i] = __ldg(kin + loc);
        else pad_key(key[i]);
    }
    if (a.vals_in) {
        const uint32_t* vin = a.vals_in + tile_base;
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            const uint32_t loc = warp_base + i * 32 + lane;
            val[i] = (FULL || loc < n_tile) ? __ldg(vin + loc) : 0u;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) val[i] = (uint32_t)tile_base + warp_base + i * 32 + lane;
    }
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        // out-of-range items carry the pad key, so in a dropping pass one test covers both
        const bool ok = DROP ? !is_pad(key[i]) : (FULL || (warp_base + i * 32 + lane) < n_tile);
        const uint32_t d = digit_of(key[i], shift, dmask);
        if (UNI) {  // every lane takes part in the shuffle and the vote, whatever its `ok`
            const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
            const bool same = ok && d == d0;
            if (__all_sync(0xffffffffu, same)) {
                if (lane == 0) atomicAdd(&my_hist[d0], 32u);
                continue;
            }
        }
        if (ok) atomicAdd(&my_hist[d], 1u);
    }
    __syncthreads();

    // 2. thread d: tile count of digit d, published at once; exclusive scan over digits; per-warp
    //    start offsets of digit d inside the tile's staging order
    uint32_t bins = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) bins += s_whist[w * RADIX + tid];
    uint32_t* my_status = a.status + (size_t)tile * RADIX + tid;
    st_relaxed_u32(my_status, (tile == 0 ? FLAG_INC : FLAG_AGG) | bins);
    // group aggregate: lets a look-back skip 16 tiles with one word once all of them have arrived
    red_add_relaxed_u32(a.gstat + (size_t)(tile >> GROUP_SHIFT) * RADIX + tid, (1u << 24) | bins);
    // first look-back window (own group only): these loads fly while the block ranks
    const int64_t gstart = (int64_t)(tile >> GROUP_SHIFT) << GROUP_SHIFT;
    uint32_t look[LOOKBACK_W];
#pragma unroll
    for (int w = 0; w < LOOKBACK_W; ++w) {
        const int64_t tw = (int64_t)tile - 1 - w;
        look[w] = (tw >= gstart) ? ld_relaxed_u32(a.status + (size_t)tw * RADIX + tid) : FLAG_AGG;
    }
    // exclusive scans over the digits: of the tile's counts, and of the global histogram (start of every
    // digit's output range) — two values through the same shuffles
    uint32_t block_off, digit_off;
    uint32_t n_stage = n_tile;  // items this tile stages and writes
    {
        const uint32_t gcnt = a.digit_counts[tid];
        uint32_t incl = bins, gincl = gcnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            const uint32_t g = __shfl_up_sync(0xffffffffu, gincl, d);
            if (lane >= d) { incl += t; gincl += g; }
        }
        if (lane == 31) { s_misc[1 + warp] = incl; s_misc[1 + SORT_WARPS + warp] = gincl; }
        __syncthreads();
        uint32_t off = 0, goff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            if (w < warp) { off += s_misc[1 + w]; goff += s_misc[1 + SORT_WARPS + w]; }
            tot += s_misc[1 + w];
        }
        if (DROP) n_stage = tot;
        block_off = off + incl - bins;
        digit_off = goff + gincl - gcnt;
    }
    {
        uint32_t run = block_off;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t c = s_whist[w * RADIX + tid];
            s_whist[w * RADIX + tid] = run;
            run += c;
        }
    }
    __syncthreads();

    // 3. stable rank inside the warp, items in order (i major, lane minor).  Peers with the same digit
    //    find each other through an atomicOr'd lane mask in shared memory; the lowest peer advances the
    //    warp's running position of that digit and clears the mask.  Items go straight to their slot.
    const uint32_t lane_lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const bool ok = DROP ? !is_pad(key[i]) : (FULL || (warp_base + i * 32 + lane) < n_tile);
        const uint32_t dg = digit_of(key[i], shift, dmask);
        uint32_t* slot = my_hist + dg;
        if (UNI) {
            const uint32_t d0 = __shfl_sync(0xffffffffu, dg, 0);  // (not inside the && below: every lane must shuffle)
            const bool same = ok && dg == d0;
            if (__all_sync(0xffffffffu, same)) {  // one digit for the whole warp: the ranks are the lane numbers
                const uint32_t base = slot[0];
                __syncwarp();
                if (lane == 0) slot[0] = base + 32u;
                s_keys[base + lane] = key[i];
                s_vals[base + lane] = val[i];
                __syncwarp();
                continue;
            }
        }
        if (ok) atomicOr(slot + SORT_WARPS * RADIX, 1u << lane);
        __syncwarp();
        uint32_t peers = 0, base = 0;
        if (ok) {
            peers = slot[SORT_WARPS * RADIX];
            base = slot[0];
        }
        __syncwarp();
        const uint32_t lower = __popc(peers & lane_lt);
        if (ok) {
            if (lower == 0) {
                slot[0] = base + __popc(peers);
                slot[SORT_WARPS * RADIX] = 0;
            }
            s_keys[base + lower] = key[i];
            s_vals[base + lower] = val[i];
        }
        __syncwarp();
    }

    // 4. decoupled look-back for digit `tid`.  Phase 1 walks the earlier tiles of the own group of 16
    //    (individual status words, LOOKBACK_W per round trip); phase 2 walks whole groups, 4 per round
    //    trip: a group contributes either the inclusive prefix of its last tile (then the walk ends) or,
    //    once all 16 tiles have arrived, its aggregate.
    {
        uint32_t excl = 0;
#ifdef GSR_FORCE_STALL
        if (tile == 1 && tid == 0) gsr_raise_error(a.error_flag);
#endif
        if (tile != 0) {
            bool done = false;
            uint32_t spins = 0;
            int64_t t = (int64_t)tile - 1;
            while (!done && t >= gstart) {
#pragma unroll
                for (int w = 0; w < LOOKBACK_W; ++w) {
                    if (done || t - w < gstart) break;
                    uint32_t v = look[w];
                    while ((v & FLAG_MASK) == 0) {  // predecessor has not published yet
                        if (++spins > WATCHDOG_SPINS) {  // never expected to trip; the frame is then invalid
                            gsr_raise_error(a.error_flag);
                            v = FLAG_INC;
                            break;
                        }
                        __nanosleep(32);
                        v = ld_relaxed_u32(a.status + (size_t)(t - w) * RADIX + tid);
                    }
                    excl += v & VAL_MASK;
                    if ((v & FLAG_MASK) == FLAG_INC) done = true;  // tile 0 always publishes FLAG_INC
                }
                t -= LOOKBACK_W;
                if (!done && t >= gstart) {
#pragma unroll
                    for (int w = 0; w < LOOKBACK_W; ++w)
                        look[w] = (t - w >= gstart) ? ld_relaxed_u32(a.status + (size_t)(t - w) * RADIX + tid) : FLAG_AGG;
                }
            }
            int64_t gi = (int64_t)(tile >> GROUP_SHIFT) - 1;
            while (!done && gi >= 0) {
                uint32_t li[4], gs[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int64_t gq = gi - q;
                    li[q] = gs[q] = 0;
                    if (gq >= 0) {
                        li[q] = ld_relaxed_u32(a.status + (size_t)((gq << GROUP_SHIFT) + GROUP_TILES - 1) * RADIX + tid);
                        gs[q] = ld_relaxed_u32(a.gstat + (size_t)gq * RADIX + tid);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int64_t gq = gi - q;
                    if (done || gq < 0) break;
                    uint32_t l = li[q], g = gs[q];
                    while (true) {
                        if ((l & FLAG_MASK) == FLAG_INC) {
                            excl += l & VAL_MASK;
                            done = true;
                            break;
                        }
                        if ((g >> 24) == GROUP_TILES) {
                            excl += g & 0xffffffu;
                            break;
                        }
                        if (++spins > WATCHDOG_SPINS) {
                            gsr_raise_error(a.error_flag);
                            done = true;
                            break;
                        }
                        __nanosleep(32);
                        l = ld_relaxed_u32(a.status + (size_t)((gq << GROUP_SHIFT) + GROUP_TILES - 1) * RADIX + tid);
                        g = ld_relaxed_u32(a.gstat + (size_t)gq * RADIX + tid);
                    }
                }
                gi -= 4;
            }
            st_relaxed_u32(my_status, FLAG_INC | ((excl + bins) & VAL_MASK));
        }
        // global index of staged item j of digit d:  s_global[d] + j
        s_global[tid] = digit_off + excl - block_off;
    }
    __syncthreads();

    // 5. write contiguous digit runs
    if (EXPAND) {
        // gather the low key halves first so the dependent loads overlap
        uint32_t lo[SORT_ITEMS];
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; ++k) {
            const uint32_t j = tid + k * SORT_THREADS;
            lo[k] = (FULL || j < n_stage) ? __ldg(a.expand_low + s_vals[j]) : 0u;
        }
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; ++k) {
            const uint32_t j = tid + k * SORT_THREADS;
            if (FULL || j < n_stage) {
                const KeyT kk = s_keys[j];
                const uint32_t g = s_global[digit_of(kk, shift, dmask)] + j;
                a.keys_out64[g] = ((uint64_t)low_word(kk) << 32) | (uint64_t)lo[k];
                a.vals_out[g] = s_vals[j];
            }
        }
    } else if (RECTS) {
        // every rect gather of the thread is in flight before the first one is used (the only random access of the
        // pass; 8 bytes per Gaussian out of an array preprocess wrote a few hundred microseconds ago)
        uint2 rc[SORT_ITEMS];
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; ++k) {
            const uint32_t j = tid + k * SORT_THREADS;
            rc[k] = (FULL || j < n_stage) ? __ldg(a.rect_src + s_vals[j]) : make_uint2(0u, 0u);
        }
        KeyT* kout = reinterpret_cast<KeyT*>(a.keys_out);
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; ++k) {
            const uint32_t j = tid + k * SORT_THREADS;
            if (FULL || j < n_stage) {
                const KeyT kk = s_keys[j];
                const uint32_t g = s_global[digit_of(kk, shift, dmask)] + j;
                if (kout) kout[g] = kk;
                a.vals_out[g] = s_vals[j];
                a.rect_dst[g] = a.rect_coarse ? coarse_rect(rc[k]) : rc[k];
            }
        }
    } else {
        KeyT* kout = reinterpret_cast<KeyT*>(a.keys_out);
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; ++k) {
            const uint32_t j = tid + k * SORT_THREADS;
            if (FULL || j < n_stage) {
                const KeyT kk = s_keys[j];
                const uint32_t g = s_global[digit_of(kk, shift, dmask)] + j;
                kout[g] = kk;
                a.vals_out[g] = s_vals[j];
            }
        }
    }
}

template <typename KeyT, bool EXPAND, int MIN_BLOCKS, int SORT_ITEMS, bool DROP = false, bool RECTS = false, bool UNI = false>
__global__ void __launch_bounds__(SORT_THREADS, MIN_BLOCKS) onesweep_kernel(const PassArgs a) {
    constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint32_t* s_whist = reinterpret_cast<uint32_t*>(s_raw + (size_t)SORT_TILE * (sizeof(KeyT) + 4));
    uint32_t* s_misc = s_whist + 2 * SORT_WARPS * RADIX + RADIX;
    const int tid = threadIdx.x;
    // tiles are handed out in order of execution, so a tile only ever waits for tiles that started
    // (the ticket word and the look-back state were zeroed before the previous kernel of the stream started,
    // so the ticket and the shared-memory setup may overlap that kernel's tail)
    if (tid == 0) s_misc[0] = atomicAdd(a.ticket, 1u);
    for (int i = tid; i < 2 * SORT_WARPS * RADIX; i += SORT_THREADS) s_whist[i] = 0;  // counters + masks
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    __syncthreads();
    const uint32_t tile = s_misc[0];
    const size_t tile_base = (size_t)tile * SORT_TILE;
    // tickets are handed out in tile order and a tile only looks back, so the tiles past a device-side count
    // can leave without publishing anything
    const size_t n = a.n_dev ? min(a.n, (size_t)__ldg(a.n_dev)) : a.n;
    if (tile_base >= n) return;
    const uint32_t n_tile = (uint32_t)min((size_t)SORT_TILE, n - tile_base);
    if (DROP)
        onesweep_tile<KeyT, EXPAND, false, SORT_ITEMS, true, RECTS, UNI>(a, n_tile, tile, s_raw);
    else if (n_tile == SORT_TILE)
        onesweep_tile<KeyT, EXPAND, true, SORT_ITEMS, false, RECTS, UNI>(a, n_tile, tile, s_raw);
    else
        onesweep_tile<KeyT, EXPAND, false, SORT_ITEMS, false, RECTS, UNI>(a, n_tile, tile, s_raw);
}

size_t num_sort_tiles(size_t n, int items) { return (n + (size_t)SORT_THREADS * items - 1) / ((size_t)SORT_THREADS * items); }

struct TempLayout {
    uint32_t* hist;     // [MAX_PASSES][RADIX]  global digit histograms (raw counts)
    uint32_t* tickets;  // [MAX_PASSES] tile tickets, + [MAX_PASSES] error flag
    uint32_t* status;   // [passes][num_tiles][RADIX]
    uint32_t* gstat;    // [passes][num_groups][RADIX]
    size_t zero_bytes;
};

size_t num_groups(size_t tiles) { return (tiles + GROUP_TILES - 1) >> GROUP_SHIFT; }

TempLayout carve_temp(char* temp, size_t n, int passes, int items) {
    TempLayout L;
    char* t = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(temp), 128));
    L.hist = reinterpret_cast<uint32_t*>(t);
    t += align_up((size_t)MAX_PASSES * RADIX * 4, 128);
    L.tickets = reinterpret_cast<uint32_t*>(t);
    t += align_up((size_t)MAX_PASSES * 2 * 4, 128);
    L.status = reinterpret_cast<uint32_t*>(t);
    const size_t tiles = num_sort_tiles(n, items);
    L.gstat = L.status + (size_t)passes * tiles * RADIX;
    L.zero_bytes = (size_t)(reinterpret_cast<char*>(L.gstat) - reinterpret_cast<char*>(L.hist)) +
                   (size_t)passes * num_groups(tiles) * RADIX * 4;
    return L;
}

template <typename KeyT>
int launch_histogram(const KeyT* keys, size_t n, int end_bit, int passes, uint32_t* hist, cudaStream_t s,
                     uint32_t* kept = nullptr) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t per_block = (size_t)HIST_THREADS * 8;
    const unsigned hblocks = (unsigned)std::min<size_t>((n + per_block - 1) / per_block, (size_t)sms * 4);
#define GSR_HIST(PS) launch_pdl(histogram_kernel<KeyT, PS>, dim3(hblocks), dim3(HIST_THREADS), 0, s, keys, n, end_bit, hist, kept)
    switch (passes) {
        case 1: GSR_HIST(1); break;
        case 2: GSR_HIST(2); break;
        case 3: GSR_HIST(3); break;
        case 4: GSR_HIST(4); break;
        case 5: GSR_HIST(5); break;
        case 6: GSR_HIST(6); break;
        case 7: GSR_HIST(7); break;
        default: GSR_HIST(8); break;
    }
#undef GSR_HIST
    return 1;
}

template <typename KeyT, bool EXPAND, int MIN_BLOCKS, int ITEMS, bool DROP = false, bool RECTS = false, bool UNI = false>
int launch_pass(const PassArgs& a, cudaStream_t s) {
    constexpr size_t smem = onesweep_smem<KeyT, ITEMS>();
    // per-device attribute (one process may drive several GPUs): set once per device and instantiation
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    GSR_CUDA_TRY(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        GSR_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel<KeyT, EXPAND, MIN_BLOCKS, ITEMS, DROP, RECTS, UNI>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.fetch_or(bit, std::memory_order_release);
    }
    GSR_CARVEOUT((onesweep_kernel<KeyT, EXPAND, MIN_BLOCKS, ITEMS, DROP, RECTS, UNI>), "SORT", -1);
    GSR_CUDA_TRY(launch_pdl(onesweep_kernel<KeyT, EXPAND, MIN_BLOCKS, ITEMS, DROP, RECTS, UNI>,
                            dim3((unsigned)num_sort_tiles(a.n, ITEMS)), dim3(SORT_THREADS), smem, s, a));
    return 1;
}

// Small inputs get small tiles (more CTAs, shorter per-CTA dependency chains).
#ifndef GSR_SMALL_N_LOG2
#define GSR_SMALL_N_LOG2 0
#endif
inline int sort32_items(size_t n) { return n < ((size_t)1 << GSR_SMALL_N_LOG2) ? ITEMS_SMALL : ITEMS_LARGE; }

}  // namespace

int sort_num_passes(int end_bit) { return (end_bit + RADIX_BITS - 1) / RADIX_BITS; }

size_t sort_temp_bytes(size_t n) {
    size_t b = 0;
    b += align_up((size_t)MAX_PASSES * RADIX * 4, 128);
    b += align_up((size_t)MAX_PASSES * 2 * 4, 128);
    // 8 passes of large tiles (64-bit keys) >= 4 passes of small tiles (32-bit keys): same bytes + rounding
    const size_t tiles = num_sort_tiles(n, ITEMS_LARGE) + 1;
    b += align_up((size_t)MAX_PASSES * (tiles + num_groups(tiles) + 1) * RADIX * 4, 128);
    return b + 128;
}

// ---- generic 64-bit-key sort (gsr_sort_pairs) ------------------------------------------------
int launch_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, size_t n, int end_bit,
                      char* temp, bool* result_in_a, cudaStream_t s, cudaEvent_t* events, uint32_t* error_flag) {
    const int passes = sort_num_passes(end_bit);
    if (result_in_a) *result_in_a = (passes % 2) == 0;
    if (n == 0) return 0;
    if (passes < 1 || passes > MAX_PASSES || end_bit > 64) return GSR_ERR_INVALID_ARG;
    if (n >= ((size_t)1 << 30)) return GSR_ERR_TOO_MANY_PAIRS;
    const size_t tiles = num_sort_tiles(n, ITEMS_LARGE);
    TempLayout L = carve_temp(temp, n, passes, ITEMS_LARGE);
    GSR_CUDA_TRY(cudaMemsetAsync(L.hist, 0, L.zero_bytes, s));

    int launches = 0;
    if (events) cudaEventRecord(events[0], s);
    launches += launch_histogram<uint2>(reinterpret_cast<const uint2*>(keys_a), n, end_bit, passes, L.hist, s);
    if (events) cudaEventRecord(events[1], s);
    uint64_t* kin = keys_a; uint32_t* vin = vals_a;
    uint64_t* kout = keys_b; uint32_t* vout = vals_b;
    for (int ps = 0; ps < passes; ++ps) {
        PassArgs a;
        a.keys_in = kin; a.vals_in = vin; a.keys_out = kout; a.vals_out = vout; a.n = n; a.n_dev = nullptr;
        a.shift = ps * RADIX_BITS;
        a.nbits = std::min(RADIX_BITS, end_bit - a.shift);
        a.digit_counts = L.hist + ps * RADIX;
        a.status = L.status + (size_t)ps * tiles * RADIX;
        a.gstat = L.gstat + (size_t)ps * num_groups(tiles) * RADIX;
        a.ticket = L.tickets + ps;
        a.error_flag = error_flag ? error_flag : L.tickets + MAX_PASSES;
        a.expand_low = nullptr; a.keys_out64 = nullptr;
        a.rect_src = nullptr; a.rect_dst = nullptr; a.rect_coarse = 0;
        int rc = launch_pass<uint2, false, 2, ITEMS_LARGE>(a, s);
        if (rc < 0) return rc;
        ++launches;
        if (events) cudaEventRecord(events[2 + ps], s);
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return -(int)e;
    return launches;
}

// ---- 32-bit-key sort used by the forward path --------------------------------------------------
const uint32_t* sort32_kept_count(char* temp, size_t n, int end_bit) {
    return carve_temp(temp, n, sort_num_passes(end_bit), sort32_items(n)).tickets + MAX_PASSES + 1;
}

uint32_t* sort32_prepare(char* temp, size_t n, int end_bit, cudaStream_t s) {
    const int passes = sort_num_passes(end_bit);
    TempLayout L = carve_temp(temp, n, passes, sort32_items(n));
    if (cudaMemsetAsync(L.hist, 0, L.zero_bytes, s) != cudaSuccess) return nullptr;
    return L.hist;
}

int launch_sort32(const Sort32Plan& p, cudaStream_t s, cudaEvent_t* events) {
    const int passes = sort_num_passes(p.end_bit);
    if (p.n == 0) return 0;
    if (passes < 1 || passes > 4 || p.end_bit > 32) return GSR_ERR_INVALID_ARG;
    if (p.n >= ((size_t)1 << 30)) return GSR_ERR_TOO_MANY_PAIRS;
    const int items = sort32_items(p.n);
    const size_t tiles = num_sort_tiles(p.n, items);
    TempLayout L = carve_temp(p.temp, p.n, passes, items);
    int launches = 0;
    if (events) cudaEventRecord(events[0], s);
    if (p.drop_pad && (p.hist_ready || passes < 2)) return GSR_ERR_INVALID_ARG;
    uint32_t* kept = p.drop_pad ? L.tickets + MAX_PASSES + 1 : nullptr;
    if (!p.hist_ready) {
        launches += launch_histogram<uint32_t>(p.keys_in, p.n, p.end_bit, passes, L.hist, s, kept);
    }
    if (events) cudaEventRecord(events[1], s);
    const uint32_t* kin = p.keys_in;
    const uint32_t* vin = p.vals_in;
    for (int ps = 0; ps < passes; ++ps) {
        const bool last = ps == passes - 1;
        PassArgs a;
        a.keys_in = kin; a.vals_in = vin; a.n = p.n;
        a.n_dev = (kept && ps > 0) ? kept : nullptr;  // the first pass compacts, the later ones see only what it kept
        a.shift = ps * RADIX_BITS;
        a.nbits = std::min(RADIX_BITS, p.end_bit - a.shift);
        a.digit_counts = L.hist + ps * RADIX;
        a.status = L.status + (size_t)ps * tiles * RADIX;
        a.gstat = L.gstat + (size_t)ps * num_groups(tiles) * RADIX;
        a.ticket = L.tickets + ps;
        a.error_flag = p.error_flag ? p.error_flag : L.tickets + MAX_PASSES;
        a.expand_low = nullptr; a.keys_out64 = nullptr;
        a.rect_src = nullptr; a.rect_dst = nullptr; a.rect_coarse = 0;
        int rc;
        if (kept && ps == 0) {
            a.keys_out = p.kbuf[0]; a.vals_out = p.vbuf[0];
            rc = items == ITEMS_SMALL ? launch_pass<uint32_t, false, 6, ITEMS_SMALL, true>(a, s)
                                      : launch_pass<uint32_t, false, GSR_MINB_LARGE, ITEMS_LARGE, true>(a, s);
            kin = p.kbuf[0]; vin = p.vbuf[0];
        } else if (last) {
            a.keys_out = p.keys_out; a.vals_out = p.vals_out;
            if (p.expand_low) {
                a.expand_low = p.expand_low; a.keys_out64 = p.keys_out64;
                rc = items == ITEMS_SMALL ? launch_pass<uint32_t, true, 4, ITEMS_SMALL>(a, s)
                                          : launch_pass<uint32_t, true, GSR_MINB_LARGE - 1, ITEMS_LARGE>(a, s);
            } else if (p.rect_src) {
                a.rect_src = reinterpret_cast<const uint2*>(p.rect_src);
                a.rect_dst = reinterpret_cast<uint2*>(p.rect_dst);
                a.rect_coarse = p.rect_coarse ? 1 : 0;
                rc = items == ITEMS_SMALL ? launch_pass<uint32_t, false, 6, ITEMS_SMALL, false, true>(a, s)
                                          : launch_pass<uint32_t, false, GSR_MINB_LARGE, ITEMS_LARGE, false, true>(a, s);
            } else if (GSR_SORT_UNIFORM && p.end_bit == 32 && a.shift == 24) {
                // most significant byte of a full 32-bit key (float depth bits: sign + 7 exponent bits)
                rc = items == ITEMS_SMALL ? launch_pass<uint32_t, false, 6, ITEMS_SMALL, false, false, true>(a, s)
                                          : launch_pass<uint32_t, false, GSR_MINB_LARGE, ITEMS_LARGE, false, false, true>(a, s);
            } else {
                rc = items == ITEMS_SMALL ? launch_pass<uint32_t, false, 6, ITEMS_SMALL>(a, s)
                                          : launch_pass<uint32_t, false, GSR_MINB_LARGE, ITEMS_LARGE>(a, s);
            }
        } else {
            a.keys_out = p.kbuf[ps & 1]; a.vals_out = p.vbuf[ps & 1];
            rc = items == ITEMS_SMALL ? launch_pass<uint32_t, false, 6, ITEMS_SMALL>(a, s)
                                      : launch_pass<uint32_t, false, GSR_MINB_LARGE, ITEMS_LARGE>(a, s);
            kin = p.kbuf[ps & 1]; vin = p.vbuf[ps & 1];
        }
        if (rc < 0) return rc;
        ++launches;
        if (events) cudaEventRecord(events[2 + ps], s);
    }
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return -(int)e;
    return launches;
}

}  // namespace gsr
