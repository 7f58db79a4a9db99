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
// Warp-private counter rows, updated by the leader of each match.any group with plain
// loads/stores (the rows are private to the warp, so no atomics are needed).
template <int PASSES>
__global__ void __launch_bounds__(HIST_THREADS) histogram_kernel(const uint64_t* __restrict__ keys, const size_t n,
                                                                 const int end_bit, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t s_hist[];  // [HIST_WARPS][PASSES][RADIX]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < HIST_WARPS * PASSES * RADIX; i += HIST_THREADS) s_hist[i] = 0;
    __syncthreads();
    uint32_t* my = s_hist + warp * PASSES * RADIX;

    const size_t per_block = (size_t)HIST_THREADS * 8;
    for (size_t base = (size_t)blockIdx.x * per_block; base < n; base += (size_t)gridDim.x * per_block) {
        uint64_t k[8];
        bool ok[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const size_t pos = base + (size_t)i * HIST_THREADS + tid;
            ok[i] = pos < n;
            k[i] = ok[i] ? __ldg(keys + pos) : 0ull;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int ps = 0; ps < PASSES; ++ps) {
                const int shift = ps * RADIX_BITS;
                const int nb = min(RADIX_BITS, end_bit - shift);
                const uint32_t d = ok[i] ? (uint32_t)((k[i] >> shift) & ((1u << nb) - 1u)) : (0x100u | lane);
                const unsigned peers = __match_any_sync(0xffffffffu, d);
                if (ok[i] && lane == (__ffs(peers) - 1)) my[ps * RADIX + d] += __popc(peers);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int i = tid; i < PASSES * RADIX; i += HIST_THREADS) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < HIST_WARPS; ++w) s += s_hist[w * PASSES * RADIX + i];
        if (s) atomicAdd(hist + i, s);
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

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- one digit pass ------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS) onesweep_kernel(
    const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, const size_t n, const int shift, const int nbits,
    const uint32_t* __restrict__ digit_offsets,  // [RADIX] exclusive offsets of this pass
    uint32_t* __restrict__ status,               // [num_tiles][RADIX], zero-initialised
    uint32_t* __restrict__ ticket, uint32_t* __restrict__ error_flag) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(s_raw);                                // [SORT_TILE]
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_raw + (size_t)SORT_TILE * 8);        // [SORT_TILE]
    uint32_t* s_whist = reinterpret_cast<uint32_t*>(s_raw + (size_t)SORT_TILE * 12);      // [SORT_WARPS][RADIX]
    uint32_t* s_block_off = s_whist + SORT_WARPS * RADIX;                                 // [RADIX]
    uint32_t* s_global = s_block_off + RADIX;                                             // [RADIX]
    uint32_t* s_misc = s_global + RADIX;                                                  // [16]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) s_whist[i] = 0;
    __syncthreads();
    const uint32_t tile = s_misc[0];
    const size_t tile_base = (size_t)tile * SORT_TILE;
    const uint32_t n_tile = (uint32_t)min((size_t)SORT_TILE, n - tile_base);
    const uint32_t dmask = (1u << nbits) - 1u;

    // 1. warp-striped load: item i of this thread sits at warp_base + i*32 + lane
    uint64_t key[SORT_ITEMS];
    uint32_t val[SORT_ITEMS];
    const uint32_t warp_base = warp * (32 * SORT_ITEMS);
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const uint32_t loc = warp_base + i * 32 + lane;
        if (loc < n_tile) {
            key[i] = keys_in[tile_base + loc];
            val[i] = vals_in[tile_base + loc];
        } else {
            key[i] = ~0ull;
            val[i] = 0;
        }
    }

    // 2. rank inside the warp, in item order (i major, lane minor) -> stable
    uint32_t* my_hist = s_whist + warp * RADIX;
    uint32_t rank[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const bool ok = (warp_base + i * 32 + lane) < n_tile;
        const uint32_t d = ok ? ((uint32_t)(key[i] >> shift) & dmask) : (0x100u | lane);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (ok && lane == leader) {
            base = my_hist[d];
            my_hist[d] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        rank[i] = base + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();

    // 3. thread d: exclusive prefix of digit d over the warps, tile count of digit d
    uint32_t bins = 0;
    {
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t c = s_whist[w * RADIX + tid];
            s_whist[w * RADIX + tid] = bins;
            bins += c;
        }
    }

    // 4. decoupled look-back for digit `tid`
    {
        uint32_t* my_status = status + (size_t)tile * RADIX + tid;
        uint32_t excl = 0;
        if (tile == 0) {
            st_volatile_u32(my_status, FLAG_INC | bins);
        } else {
            st_volatile_u32(my_status, FLAG_AGG | bins);
            int64_t t = (int64_t)tile - 1;
            uint32_t spins = 0;
            while (true) {
                const uint32_t v = ld_volatile_u32(status + (size_t)t * RADIX + tid);
                const uint32_t f = v & FLAG_MASK;
                if (f == 0) {
                    if (++spins > (1u << 22)) {  // watchdog: never expected to trip
                        atomicExch(error_flag, 1u);
                        break;
                    }
                    __nanosleep(40);
                    continue;
                }
                excl += v & VAL_MASK;
                if (f == FLAG_INC) break;
                --t;  // f == FLAG_AGG: keep walking back (tile 0 always publishes FLAG_INC)
            }
            st_volatile_u32(my_status, FLAG_INC | ((excl + bins) & VAL_MASK));
        }
        s_global[tid] = digit_offsets[tid] + excl;
    }

    // block-wide exclusive scan of bins over the 256 digits -> position of each digit run in the tile
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
        s_block_off[tid] = off + incl - bins;
    }
    __syncthreads();

    // 5. permute through shared memory into digit order
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const bool ok = (warp_base + i * 32 + lane) < n_tile;
        if (ok) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & dmask;
            const uint32_t pos = s_block_off[d] + my_hist[d] + rank[i];
            s_keys[pos] = key[i];
            s_vals[pos] = val[i];
        }
    }
    __syncthreads();
#pragma unroll 4
    for (uint32_t j = tid; j < n_tile; j += SORT_THREADS) {
        const uint64_t k = s_keys[j];
        const uint32_t d = (uint32_t)(k >> shift) & dmask;
        const size_t g = (size_t)s_global[d] + (j - s_block_off[d]);
        keys_out[g] = k;
        vals_out[g] = s_vals[j];
    }
}

constexpr size_t ONESWEEP_SMEM = (size_t)SORT_TILE * 12 + (size_t)(SORT_WARPS * RADIX + RADIX + RADIX + 16) * 4;

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
        if (passes == 7)
            GSR_CUDA_TRY(cudaFuncSetAttribute(histogram_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              HIST_WARPS * 7 * RADIX * 4));
        if (passes == 8)
            GSR_CUDA_TRY(cudaFuncSetAttribute(histogram_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              HIST_WARPS * 8 * RADIX * 4));
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t per_block = (size_t)HIST_THREADS * 8;
        const unsigned hblocks = (unsigned)std::min<size_t>((n + per_block - 1) / per_block, (size_t)sms * 4);
#define GSR_HIST(PS) \
    histogram_kernel<PS><<<hblocks, HIST_THREADS, HIST_WARPS * PS * RADIX * 4, s>>>(keys_a, n, end_bit, hist)
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
            tickets + MAX_PASSES);
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
