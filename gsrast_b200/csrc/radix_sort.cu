// radix_sort.cu — in-house stable LSD radix sort of (u64 key, u32 value) pairs, "onesweep"
// style: one up-front histogram pass over the keys, then one read+write pass per 8-bit
// digit with the inter-tile prefix resolved by decoupled look-back.  No CUB / Thrust.
//
// Replaces cub::DeviceRadixSort::SortPairs as called at
// /root/reference/apps/gsrast/gscuda/GSCuda.cu:794-797 (temp-size probe: AuxBuffer.cu:83-85):
// ascending, stable, over key bits [0, end_bit) with end_bit = 32 + getHigherMsb(tiles).
// Stability is what makes the sorted (key, value) lists unique — equal (tile, depth) keys keep
// their emission order, i.e. ascending Gaussian index — so the output is bit-identical to the
// reference's by construction.
//
// Per pass and per tile of SORT_TILE items (one CTA of 256 threads, 16 items per thread):
//   1. coalesced warp-striped load of keys and values;
//   2. per-warp ranking with match.any (no shared atomics): each warp keeps a private
//      256-bin counter row in shared memory;
//   3. cross-warp exclusive prefix per digit -> tile histogram;
//   4. decoupled look-back over the per-tile status words (2 flag bits + 30-bit count) to
//      obtain the number of same-digit items in all earlier tiles; tiles are handed out by
//      an atomic ticket so a waiting tile only ever waits for tiles that already started;
//   5. keys/values are permuted into digit order through shared memory and written out as
//      contiguous per-digit runs.
// HBM traffic per pair: 8 B (histogram) + passes x (12 B read + 12 B write).
#include <algorithm>

#include "gsr_common.cuh"

namespace gsr {

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;  // == RADIX: thread d owns digit d in the scan / look-back steps
constexpr int SORT_ITEMS = 16;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 pairs
constexpr int MAX_PASSES = 8;
#ifndef GSR_LOOKBACK_W
#define GSR_LOOKBACK_W 8
#endif
#ifndef GSR_SORT_MIN_BLOCKS
#define GSR_SORT_MIN_BLOCKS 3
#endif
constexpr int LOOKBACK_W = GSR_LOOKBACK_W;

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_INC = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VAL_MASK = ~FLAG_MASK;

constexpr int HIST_THREADS = 256;
constexpr int HIST_WARPS = HIST_THREADS / 32;

struct SortTemp {
    uint32_t* hist;     // [MAX_PASSES][RADIX]  global digit histograms -> exclusive offsets
    uint32_t* tickets;  // [MAX_PASSES] tile tickets, + [MAX_PASSES] error flag
    uint32_t* status;   // [passes][num_tiles][RADIX]
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- up-front histograms of every digit place -------------------------------------------
// Block-shared counters updated with shared-memory atomics: measured on B200 (tools/microbench.cu)
// at ~2.4 SM-cycles per warp-wide ATOMS against ~60 for match.any and ~25 for an 8-step ballot
// match, so plain atomics are the right tool for an order-independent count.
template <int PASSES>
__global__ void __launch_bounds__(HIST_THREADS) histogram_kernel(const uint64_t* __restrict__ keys, const size_t n,
                                                                 const int end_bit, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[PASSES * RADIX];
    const int tid = threadIdx.x;
    for (int i = tid; i < PASSES * RADIX; i += HIST_THREADS) s_hist[i] = 0;
    __syncthreads();

    constexpr int PER_THREAD = 8;
    const size_t per_block = (size_t)HIST_THREADS * PER_THREAD;
    for (size_t base = (size_t)blockIdx.x * per_block; base < n; base += (size_t)gridDim.x * per_block) {
        uint64_t k[PER_THREAD];
#pragma unroll
        for (int i = 0; i < PER_THREAD; ++i) {
            const size_t pos = base + (size_t)i * HIST_THREADS + tid;
            k[i] = pos < n ? __ldg(keys + pos) : 0ull;
        }
#pragma unroll
        for (int i = 0; i < PER_THREAD; ++i) {
            if (base + (size_t)i * HIST_THREADS + tid < n) {
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    const int shift = ps * RADIX_BITS;
                    const int nb = min(RADIX_BITS, end_bit - shift);
                    atomicAdd(&s_hist[ps * RADIX + (uint32_t)((k[i] >> shift) & ((1u << nb) - 1u))], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < PASSES * RADIX; i += HIST_THREADS) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(hist + i, c);
    }
}

// hist[pass][d] -> exclusive prefix over d, in place.  One CTA, one pass per iteration.
__global__ void __launch_bounds__(RADIX) scan_histograms_kernel(uint32_t* __restrict__ hist, const int passes) {
    __shared__ uint32_t s_w[RADIX / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int ps = 0; ps < passes; ++ps) {
        const uint32_t v = hist[ps * RADIX + tid];
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        uint32_t off = 0;
#pragma unroll
        for (int w = 0; w < RADIX / 32; ++w)
            if (w < warp) off += s_w[w];
        hist[ps * RADIX + tid] = off + incl - v;
        __syncthreads();
    }
}

// Status words carry flag and count in ONE 32-bit word, so relaxed gpu-scope accesses suffice.
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- one digit pass ------------------------------------------------------------------------
// Shared memory: [SORT_TILE] u64 key staging (re-used as u32 for the values) | [SORT_TILE] u32 value
// prefetch (cp.async) | per-warp digit counters [SORT_WARPS][RADIX] | per-warp match masks
// [SORT_WARPS][RADIX] | global bases [RADIX] | misc.
constexpr size_t ONESWEEP_SMEM = (size_t)SORT_TILE * 12 + (size_t)(2 * SORT_WARPS * RADIX + RADIX + 16) * 4;

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// The digit of a pass never straddles the two 32-bit halves of the key (shift is a multiple of 8 and
// the digit has <= 8 bits), so all digit arithmetic is 32-bit: pick the half, shift, mask.
__device__ __forceinline__ uint32_t digit_of(const uint2 k, const bool hi, const int sh, const uint32_t dmask) {
    return ((hi ? k.y : k.x) >> sh) & dmask;
}

template <bool FULL>
__device__ __forceinline__ void onesweep_tile(const uint2* __restrict__ kin, const uint32_t* __restrict__ vin,
                                              uint2* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                              const uint32_t n_tile, const uint32_t tile, const bool hi, const int sh,
                                              const uint32_t dmask, const uint32_t* __restrict__ digit_offsets,
                                              uint32_t* __restrict__ status, uint32_t* __restrict__ error_flag,
                                              const bool vec_vals, unsigned char* s_raw) {
    uint2* s_keys = reinterpret_cast<uint2*>(s_raw);                                  // [SORT_TILE]
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_raw);                            // aliases s_keys
    uint32_t* s_vpre = reinterpret_cast<uint32_t*>(s_raw + (size_t)SORT_TILE * 8);    // [SORT_TILE]
    uint32_t* s_whist = s_vpre + SORT_TILE;                                           // [SORT_WARPS][RADIX]
    uint32_t* s_global = s_whist + 2 * SORT_WARPS * RADIX;                            // [RADIX]
    uint32_t* s_misc = s_global + RADIX;                                              // [16]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t warp_base = warp * (32 * SORT_ITEMS);
    uint32_t* my_hist = s_whist + warp * RADIX;  // my_mask = my_hist + SORT_WARPS*RADIX

    // 0. values: asynchronous prefetch straight into shared memory (no registers held)
    if (FULL && vec_vals) {
#pragma unroll
        for (int c = 0; c < SORT_TILE / 4 / SORT_THREADS; ++c) {
            const int ch = tid + c * SORT_THREADS;
            cp_async_16(s_vpre + ch * 4, vin + ch * 4);
        }
    } else {
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            const uint32_t loc = tid + i * SORT_THREADS;
            if (loc < n_tile) cp_async_4(s_vpre + loc, vin + loc);
        }
    }

    // 1. warp-striped key load (item i of this thread sits at warp_base + i*32 + lane) + early counts
    uint2 key[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const uint32_t loc = warp_base + i * 32 + lane;
        key[i] = (FULL || loc < n_tile) ? __ldg(kin + loc) : make_uint2(~0u, ~0u);
    }
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        if (FULL || (warp_base + i * 32 + lane) < n_tile) atomicAdd(&my_hist[digit_of(key[i], hi, sh, dmask)], 1u);
    }
    __syncthreads();

    // 2. thread d: tile count of digit d, published at once; exclusive scan over digits; per-warp
    //    start offsets of digit d inside the tile's staging order
    uint32_t bins = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) bins += s_whist[w * RADIX + tid];
    uint32_t* my_status = status + (size_t)tile * RADIX + tid;
    st_volatile_u32(my_status, (tile == 0 ? FLAG_INC : FLAG_AGG) | bins);
    // first look-back window: these loads fly while the block ranks
    uint32_t look[LOOKBACK_W];
#pragma unroll
    for (int w = 0; w < LOOKBACK_W; ++w) {
        const int64_t tw = (int64_t)tile - 1 - w;
        look[w] = (tw >= 0) ? ld_volatile_u32(status + (size_t)tw * RADIX + tid) : FLAG_INC;
    }
    uint32_t block_off;
    {
        uint32_t incl = bins;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_misc[1 + warp] = incl;
        __syncthreads();
        uint32_t off = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w)
            if (w < warp) off += s_misc[1 + w];
        block_off = off + incl - bins;
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
    //    warp's running position of that digit and clears the mask.
    uint32_t pos[SORT_ITEMS];
    const uint32_t lane_lt = (1u << lane) - 1u;
#ifdef GSR_DBG_NO_RANK
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) pos[i] = warp_base + i * 32 + lane;
#else
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const bool ok = FULL || (warp_base + i * 32 + lane) < n_tile;
        uint32_t* slot = my_hist + digit_of(key[i], hi, sh, dmask);
        if (ok) atomicOr(slot + SORT_WARPS * RADIX, 1u << lane);
        __syncwarp();
        uint32_t peers = 0, base = 0;
        if (ok) {
            peers = slot[SORT_WARPS * RADIX];
            base = slot[0];
        }
        __syncwarp();
        const uint32_t lower = __popc(peers & lane_lt);
        if (ok && lower == 0) {
            slot[0] = base + __popc(peers);
            slot[SORT_WARPS * RADIX] = 0;
        }
        pos[i] = base + lower;
        __syncwarp();
    }
#endif

    // 4. decoupled look-back for digit `tid`, LOOKBACK_W predecessor words per round trip
    {
        uint32_t excl = 0;
#ifdef GSR_DBG_NO_LOOKBACK
        if (false) {
#else
        if (tile != 0) {
#endif
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            uint32_t spins = 0;
            while (!done) {
#pragma unroll
                for (int w = 0; w < LOOKBACK_W; ++w) {
                    if (done) break;
                    uint32_t v = look[w];
                    while ((v & FLAG_MASK) == 0) {  // predecessor has not published yet
                        if (++spins > (1u << 22)) {  // watchdog: never expected to trip
                            atomicExch(error_flag, 1u);
                            v = FLAG_INC;
                            break;
                        }
                        __nanosleep(32);
                        v = ld_volatile_u32(status + (size_t)(t - w) * RADIX + tid);
                    }
                    excl += v & VAL_MASK;
                    if ((v & FLAG_MASK) == FLAG_INC) done = true;  // tile 0 always publishes FLAG_INC
                }
                if (!done) {
                    t -= LOOKBACK_W;
#pragma unroll
                    for (int w = 0; w < LOOKBACK_W; ++w)
                        look[w] = (t - w >= 0) ? ld_volatile_u32(status + (size_t)(t - w) * RADIX + tid) : FLAG_INC;
                }
            }
            st_volatile_u32(my_status, FLAG_INC | ((excl + bins) & VAL_MASK));
        }
        // global index of staged item j of digit d:  s_global[d] + j
        s_global[tid] = digit_offsets[tid] + excl - block_off;
    }

    // 5. keys: permute through shared memory into digit order, write contiguous digit runs
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i)
        if (FULL || (warp_base + i * 32 + lane) < n_tile) s_keys[pos[i]] = key[i];
    cp_async_wait_all();
    __syncthreads();
    uint32_t gidx[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const uint32_t j = tid + k * SORT_THREADS;
        if (FULL || j < n_tile) {
            const uint2 kk = s_keys[j];
            const uint32_t g = s_global[digit_of(kk, hi, sh, dmask)] + j;
            gidx[k] = g;
#ifdef GSR_DBG_NO_STORE
            if (g == 0xffffffffu)
#endif
            keys_out[g] = kk;
        }
    }
    __syncthreads();
    // 6. values: same permutation through the same staging buffer
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const uint32_t loc = warp_base + i * 32 + lane;
        if (FULL || loc < n_tile) s_vals[pos[i]] = s_vpre[loc];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const uint32_t j = tid + k * SORT_THREADS;
#ifdef GSR_DBG_NO_STORE
        if (gidx[k] == 0xffffffffu)
#endif
        if (FULL || j < n_tile) vals_out[gidx[k]] = s_vals[j];
    }
}

__global__ void __launch_bounds__(SORT_THREADS, GSR_SORT_MIN_BLOCKS) onesweep_kernel(
    const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, const size_t n, const int shift, const int nbits,
    const uint32_t* __restrict__ digit_offsets,  // [RADIX] exclusive offsets of this pass
    uint32_t* __restrict__ status,               // [num_tiles][RADIX], zero-initialised
    uint32_t* __restrict__ ticket, uint32_t* __restrict__ error_flag, const int vec_vals) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint32_t* s_whist = reinterpret_cast<uint32_t*>(s_raw + (size_t)SORT_TILE * 12);
    uint32_t* s_misc = s_whist + 2 * SORT_WARPS * RADIX + RADIX;
    const int tid = threadIdx.x;
    // tiles are handed out in launch order of execution, so a tile only ever waits for tiles that started
    if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
    for (int i = tid; i < 2 * SORT_WARPS * RADIX; i += SORT_THREADS) s_whist[i] = 0;  // counters + masks
    __syncthreads();
    const uint32_t tile = s_misc[0];
    const size_t tile_base = (size_t)tile * SORT_TILE;
    const uint32_t n_tile = (uint32_t)min((size_t)SORT_TILE, n - tile_base);
    const uint2* kin = reinterpret_cast<const uint2*>(keys_in) + tile_base;
    const uint32_t* vin = vals_in + tile_base;
    const bool hi = shift >= 32;
    const int sh = shift & 31;
    const uint32_t dmask = (1u << nbits) - 1u;
    if (n_tile == SORT_TILE)
        onesweep_tile<true>(kin, vin, reinterpret_cast<uint2*>(keys_out), vals_out, n_tile, tile, hi, sh, dmask,
                            digit_offsets, status, error_flag, vec_vals != 0, s_raw);
    else
        onesweep_tile<false>(kin, vin, reinterpret_cast<uint2*>(keys_out), vals_out, n_tile, tile, hi, sh, dmask,
                             digit_offsets, status, error_flag, vec_vals != 0, s_raw);
}

size_t num_sort_tiles(size_t n) { return (n + SORT_TILE - 1) / SORT_TILE; }

}  // namespace

int sort_num_passes(int end_bit) { return (end_bit + RADIX_BITS - 1) / RADIX_BITS; }

size_t sort_temp_bytes(size_t n) {
    size_t b = 0;
    b += align_up((size_t)MAX_PASSES * RADIX * 4, 128);
    b += align_up((size_t)MAX_PASSES * 2 * 4, 128);
    b += align_up((size_t)MAX_PASSES * num_sort_tiles(n) * RADIX * 4, 128);
    return b + 128;
}

int launch_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, size_t n, int end_bit,
                      char* temp, bool* result_in_a, cudaStream_t s, cudaEvent_t* events) {
    const int passes = sort_num_passes(end_bit);
    if (result_in_a) *result_in_a = (passes % 2) == 0;
    if (n == 0) return 0;
    if (passes < 1 || passes > MAX_PASSES || end_bit > 64) return GSR_ERR_INVALID_ARG;
    if (n >= ((size_t)1 << 30)) return GSR_ERR_TOO_MANY_PAIRS;
    const size_t tiles = num_sort_tiles(n);

    char* t = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(temp), 128));
    uint32_t* hist = reinterpret_cast<uint32_t*>(t);
    t += align_up((size_t)MAX_PASSES * RADIX * 4, 128);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(t);
    t += align_up((size_t)MAX_PASSES * 2 * 4, 128);
    uint32_t* status = reinterpret_cast<uint32_t*>(t);
    const size_t zero_bytes = (size_t)(reinterpret_cast<char*>(status) - reinterpret_cast<char*>(hist)) +
                              (size_t)passes * tiles * RADIX * 4;
    GSR_CUDA_TRY(cudaMemsetAsync(hist, 0, zero_bytes, s));

    int launches = 0;
    if (events) cudaEventRecord(events[0], s);
    {
        // per-device attributes; cheap enough to set on every call (one process may drive several GPUs)
        GSR_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)ONESWEEP_SMEM));
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t per_block = (size_t)HIST_THREADS * 8;
        const unsigned hblocks = (unsigned)std::min<size_t>((n + per_block - 1) / per_block, (size_t)sms * 8);
#define GSR_HIST(PS) \
    histogram_kernel<PS><<<hblocks, HIST_THREADS, 0, s>>>(keys_a, n, end_bit, hist)
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
        scan_histograms_kernel<<<1, RADIX, 0, s>>>(hist, passes);
        launches += 2;
    }
    if (events) cudaEventRecord(events[1], s);
    uint64_t* kin = keys_a; uint32_t* vin = vals_a;
    uint64_t* kout = keys_b; uint32_t* vout = vals_b;
    for (int ps = 0; ps < passes; ++ps) {
        const int shift = ps * RADIX_BITS;
        const int nbits = std::min(RADIX_BITS, end_bit - shift);
        onesweep_kernel<<<(unsigned)tiles, SORT_THREADS, ONESWEEP_SMEM, s>>>(
            kin, vin, kout, vout, n, shift, nbits, hist + ps * RADIX, status + (size_t)ps * tiles * RADIX, tickets + ps,
            tickets + MAX_PASSES, (int)((reinterpret_cast<uintptr_t>(vin) & 15) == 0));
        ++launches;
        if (events) cudaEventRecord(events[2 + ps], s);
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return -(int)e;
    return launches;
}

}  // namespace gsr
