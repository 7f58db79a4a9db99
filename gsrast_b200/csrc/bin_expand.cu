// bin_expand.cu — tile half of the (tile | depth) sort without radix passes over the pairs (sm_100a).
//
// Replaces, together with radix_sort.cu / binning.cu, what the reference does with
//   duplicateWithKeys                      /root/reference/apps/gsrast/gscuda/GSCuda.cu:422-475
//   cub::DeviceRadixSort::SortPairs        GSCuda.cu:794-797
//   cudaMemset(ranges) + identifyTileRanges GSCuda.cu:800-801, 504-538
// and produces the same three outputs bit for bit: point_list_keys (tile << 32 | depth bits, sorted),
// point_list (Gaussian ids in that order) and ranges (per tile: [start, end) into those lists).
//
// Why this is the same sort.  The sorted list is unique: ascending tile, then ascending depth bits,
// ties in ascending Gaussian index (the radix sort is stable and pairs are emitted in index order).
// The depth half of that order is established on the P Gaussians before any pair exists
// (radix_sort.cu).  What remains is a STABLE partition of the depth-ordered pair stream by tile id.
// Doing that with radix passes moves every one of the R pairs twice (R ~ 5 P); here the stream is
// partitioned at Gaussian granularity instead:
//   1. every Gaussian is emitted once per 8x8-tile BIN its rect touches (~1.15 records per Gaussian
//      at 1080p) and those records are radix-sorted by bin id (one 8-bit pass up to 256 bins) — each
//      bin now lists its Gaussians in depth order;
//   2. the record list of a bin is cut into chunks of 512; a chunk turns each record's rect into a
//      64-bit tile mask of the bin, and a 32x32 bit-matrix transpose across the lanes of a warp turns 32
//      masks into 64 ballots (one per tile: which of the 32 records touch it).  popc of a ballot is a
//      count, popc below the own lane is a stable rank;
//   3. expand_count: per chunk and tile, pairs emitted; a scan over the chunks of a bin and over all
//      tiles (row-major tile id) gives every (chunk, tile) its final offset and every tile its range —
//      identifyTileRanges falls out of the scan, the sorted keys are never re-read;
//   4. expand_fill: every chunk recomputes its ballots, ranks its pairs, stages them per tile in shared
//      memory and writes 12 bytes per pair as contiguous runs straight into their final position.
// Pair-level HBM traffic drops from 56 B/pair (8 written by the duplication, 2 x 16 by the first
// pass, 8 + 4 + 12 by the last, 8 re-read for the ranges) to the 12 B/pair of the result itself.
#include <algorithm>

#include "gsr_common.cuh"

namespace gsr {

namespace {

constexpr int EXP_THREADS = 256;
constexpr int EXP_WARPS = EXP_THREADS / 32;
constexpr int EXP_CHUNK = 512;                 // records per chunk
constexpr int EXP_WS = EXP_CHUNK / 32;         // warp-steps (32 records each) per chunk
constexpr int EXP_PER_WARP = EXP_WS / EXP_WARPS;
constexpr int EXP_CAP = 4096;                  // pairs staged per round (>= 32 * 64, the most one warp-step emits)
constexpr int TABLE_THREADS = 1024;

__host__ __device__ inline size_t align128(size_t v) { return (v + 127) / 128 * 128; }

struct ExpandTemp {
    uint32_t* bin_start;        // [MAX_BINS + 1] first record of every bin (+ n_records)
    uint32_t* bin_chunk_first;  // [MAX_BINS + 1] first chunk of every bin (+ number of chunks)
    uint4* chunk_desc;          // [max_chunks] bin, first record, end record, -
    uint32_t* chunk_counts;     // [max_chunks][64] pairs per tile, then exclusive prefix over the bin's chunks
};

size_t max_chunks(size_t R) { return R / EXP_CHUNK + MAX_BINS + 1; }

ExpandTemp carve(char* temp, size_t R) {
    ExpandTemp t;
    char* c = reinterpret_cast<char*>(align128(reinterpret_cast<size_t>(temp)));
    t.bin_start = reinterpret_cast<uint32_t*>(c);
    c += align128((MAX_BINS + 1) * sizeof(uint32_t));
    t.bin_chunk_first = reinterpret_cast<uint32_t*>(c);
    c += align128((MAX_BINS + 1) * sizeof(uint32_t));
    t.chunk_desc = reinterpret_cast<uint4*>(c);
    c += align128(max_chunks(R) * sizeof(uint4));
    t.chunk_counts = reinterpret_cast<uint32_t*>(c);
    return t;
}

// ---- bin boundaries -----------------------------------------------------------------------------
// One warp per bin b in [0, nbins]: bin_start[b] = first record whose bin id is >= b (32-ary search
// in the sorted bin ids: 5 round trips for 2^25 records instead of 25).
__global__ void __launch_bounds__(EXP_THREADS) bin_bounds_kernel(const uint32_t* __restrict__ rec_bins, const uint32_t n,
                                                                  const int nbins, uint32_t* __restrict__ bin_start) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * EXP_WARPS + (threadIdx.x >> 5);
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    if (b > nbins) return;
    uint32_t lo = 0, hi = n;  // answer in [lo, hi]
    if (b == 0) hi = 0;
    if (b == nbins) lo = n;
    while (lo < hi) {
        const uint32_t step = (hi - lo + 31u) / 32u;
        const uint32_t pos = lo + (uint32_t)lane * step;
        const bool probe = pos < hi;
        const bool less = probe && (__ldg(rec_bins + pos) < (uint32_t)b);
        const uint32_t c = __popc(__ballot_sync(0xffffffffu, less));       // sorted: the `less` lanes are a prefix
        const uint32_t nprobe = __popc(__ballot_sync(0xffffffffu, probe));
        if (c == 0) {
            hi = lo;  // the first probe already is >= b
        } else {
            const uint32_t new_lo = lo + (c - 1u) * step + 1u;
            if (c < nprobe) hi = lo + c * step;  // first probe that is >= b
            lo = new_lo;
        }
    }
    if (lane == 0) bin_start[b] = lo;
}

// ---- chunk table --------------------------------------------------------------------------------
// One CTA: chunks per bin, their exclusive scan, and one descriptor per chunk.
__global__ void __launch_bounds__(TABLE_THREADS) chunk_table_kernel(const int nbins, const uint32_t* __restrict__ bin_start,
                                                                    uint32_t* __restrict__ bin_chunk_first,
                                                                    uint4* __restrict__ chunk_desc) {
    constexpr int PER = MAX_BINS / TABLE_THREADS;  // 4 consecutive bins per thread
    __shared__ uint32_t s_warp[TABLE_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    uint32_t st[PER + 1], nch[PER], tsum = 0;
#pragma unroll
    for (int j = 0; j <= PER; ++j) {
        const int b = tid * PER + j;
        st[j] = (b <= nbins) ? bin_start[b] : 0u;
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int b = tid * PER + j;
        nch[j] = (b < nbins) ? (st[j + 1] - st[j] + EXP_CHUNK - 1) / EXP_CHUNK : 0u;
        tsum += nch[j];
    }
    uint32_t incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + incl - tsum;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int b = tid * PER + j;
        if (b < nbins) {
            bin_chunk_first[b] = run;
            for (uint32_t k = 0; k < nch[j]; ++k) {
                const uint32_t s0 = st[j] + k * EXP_CHUNK;
                chunk_desc[run + k] = make_uint4((uint32_t)b, s0, min(s0 + (uint32_t)EXP_CHUNK, st[j + 1]), 0u);
            }
            run += nch[j];
        }
        if (b == nbins - 1) bin_chunk_first[nbins] = run;  // total number of chunks
    }
}

// ---- masks and ballots --------------------------------------------------------------------------
// Tiles of bin (bx8, by8 = its first tile column / row) touched by the rect: bit (8*ly + lx).
__device__ __forceinline__ uint64_t bin_mask(const uint2 rec, const int bx8, const int by8) {
    int x0 = (int)(rec.x & 0xffffu) - bx8, y0 = (int)(rec.x >> 16) - by8;
    int x1 = x0 + (int)(rec.y & 0xffffu) - 1, y1 = y0 + (int)(rec.y >> 16) - 1;
    x0 = max(x0, 0); y0 = max(y0, 0);
    x1 = min(x1, BIN_SIDE - 1); y1 = min(y1, BIN_SIDE - 1);
    if (x1 < x0 || y1 < y0) return 0ull;
    const uint32_t col = ((2u << x1) - 1u) & ~((1u << x0) - 1u);  // bits x0..x1
    const uint64_t ones = 0x0101010101010101ull;
    const uint64_t rows = (ones << (8 * y0)) & (ones >> (8 * (BIN_SIDE - 1 - y1)));  // byte y0..y1 = 1
    return rows * col;
}

// 32x32 bit-matrix transpose across a warp: lane i holds row i, afterwards lane j holds column j
// (bit i of the result = bit j of lane i's input).  Five butterfly stages, one shuffle each.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, const int lane) {
    uint32_t m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? (((y >> j) & m) | (x & ~m)) : ((x & m) | ((y & m) << j));
        m ^= m << (j >> 1);
    }
    return x;
}

struct ExpandArgs {
    const uint32_t* rec_ids;
    const uint4* chunk_desc;
    const uint32_t* num_chunks;  // device: bin_chunk_first[nbins]
    uint32_t* chunk_counts;
    const uint2* tile_rects;
    const uint32_t* depths;
    const uint32_t* tile_start;  // [tiles] exclusive scan of the per-tile totals
    uint64_t* keys_out;
    uint32_t* vals_out;
    int grid_x, grid_y, bins_x;
};

// ---- pass 1: pairs per (chunk, tile) ----------------------------------------------------------------
__global__ void __launch_bounds__(EXP_THREADS) expand_count_kernel(const ExpandArgs a) {
    __shared__ uint32_t s_cnt[BIN_TILES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    const uint32_t c = blockIdx.x;
    if (c >= __ldg(a.num_chunks)) return;
    const uint4 d = __ldg(a.chunk_desc + c);
    if (tid < BIN_TILES) s_cnt[tid] = 0;
    __syncthreads();
    const int bx8 = (int)(d.x % (uint32_t)a.bins_x) << BIN_SHIFT, by8 = (int)(d.x / (uint32_t)a.bins_x) << BIN_SHIFT;
    uint64_t m[EXP_PER_WARP];
#pragma unroll
    for (int k = 0; k < EXP_PER_WARP; ++k) {
        const uint32_t r = d.y + (uint32_t)((warp + k * EXP_WARPS) * 32 + lane);
        m[k] = 0ull;
        if (r < d.z) m[k] = bin_mask(__ldg(a.tile_rects + __ldg(a.rec_ids + r)), bx8, by8);
    }
    uint32_t c_lo = 0, c_hi = 0;
#pragma unroll
    for (int k = 0; k < EXP_PER_WARP; ++k) {
        if (d.y + (uint32_t)((warp + k * EXP_WARPS) * 32) < d.z) {  // warp-uniform
            c_lo += __popc(warp_transpose32((uint32_t)m[k], lane));
            c_hi += __popc(warp_transpose32((uint32_t)(m[k] >> 32), lane));
        }
    }
    if (c_lo) atomicAdd(&s_cnt[lane], c_lo);
    if (c_hi) atomicAdd(&s_cnt[32 + lane], c_hi);
    __syncthreads();
    if (tid < BIN_TILES) a.chunk_counts[(size_t)c * BIN_TILES + tid] = s_cnt[tid];
}

// ---- scan 1: over the chunks of every bin, per tile ---------------------------------------------------
// One CTA of 64 threads per bin; thread t walks the bin's chunks: counts -> exclusive prefix (in place),
// total -> tile_counts[tile id].  Every tile of the grid belongs to exactly one bin, so tile_counts is
// written completely (no clear needed).
__global__ void __launch_bounds__(BIN_TILES) expand_scan_chunks_kernel(const uint32_t* __restrict__ bin_chunk_first,
                                                                       uint32_t* __restrict__ chunk_counts,
                                                                       uint32_t* __restrict__ tile_counts, const int grid_x,
                                                                       const int grid_y, const int bins_x) {
    const int t = threadIdx.x, b = blockIdx.x;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    const uint32_t first = __ldg(bin_chunk_first + b), last = __ldg(bin_chunk_first + b + 1);
    uint32_t run = 0;
    uint32_t* col = chunk_counts + (size_t)first * BIN_TILES + t;
    uint32_t k = first;
    for (; k + 4 <= last; k += 4, col += 4 * BIN_TILES) {
        const uint32_t v0 = col[0], v1 = col[BIN_TILES], v2 = col[2 * BIN_TILES], v3 = col[3 * BIN_TILES];
        col[0] = run; run += v0;
        col[BIN_TILES] = run; run += v1;
        col[2 * BIN_TILES] = run; run += v2;
        col[3 * BIN_TILES] = run; run += v3;
    }
    for (; k < last; ++k, col += BIN_TILES) {
        const uint32_t v = col[0];
        col[0] = run;
        run += v;
    }
    const int tx = ((b % bins_x) << BIN_SHIFT) + (t & (BIN_SIDE - 1)), ty = ((b / bins_x) << BIN_SHIFT) + (t >> BIN_SHIFT);
    if (tx < grid_x && ty < grid_y) tile_counts[ty * grid_x + tx] = run;
}

// ---- scan 2: over the tiles (row-major tile id) -> ranges ---------------------------------------------
// One CTA.  tile_counts becomes its exclusive scan (the start of every tile's list) and ranges[tile] =
// (start, start + count); a tile nothing touches keeps (0, 0) exactly like the reference's cleared and
// never written entry (GSCuda.cu:800, 504-538).
__global__ void __launch_bounds__(TABLE_THREADS) tile_ranges_kernel(uint32_t* __restrict__ tile_counts, const int tiles,
                                                                    uint2* __restrict__ ranges, const int r1_quirk) {
    constexpr int ITEMS = 8;
    __shared__ uint32_t s_warp[TABLE_THREADS / 32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    __syncthreads();
    for (int base = 0; base < tiles; base += TABLE_THREADS * ITEMS) {
        const int i0 = base + tid * ITEMS;
        uint32_t v[ITEMS], tsum = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            v[j] = (i0 + j < tiles) ? tile_counts[i0 + j] : 0u;
            tsum += v[j];
        }
        uint32_t incl = tsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_warp[lane];
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += t;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        uint32_t run = s_carry + s_warp[warp] + incl - tsum;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (i0 + j < tiles) {
                tile_counts[i0 + j] = run;
                // GSRast-compat with a single pair in the frame: .x = 0 is written, .y never is (GSCuda.cu:533-536)
                ranges[i0 + j] = (v[j] && !r1_quirk) ? make_uint2(run, run + v[j]) : make_uint2(0u, 0u);
            }
            run += v[j];
        }
        __syncthreads();
        if (tid == TABLE_THREADS - 1) s_carry = run;
        __syncthreads();
    }
}

// ---- pass 2: rank, stage, write ------------------------------------------------------------------------
__global__ void __launch_bounds__(EXP_THREADS, 4) expand_fill_kernel(const ExpandArgs a) {
    __shared__ uint32_t s_bal[EXP_WS][BIN_TILES];      // ballot of tile t in warp-step ws
    __shared__ uint32_t s_pre[EXP_WS + 1][BIN_TILES];  // pairs of tile t before warp-step ws; [EXP_WS] = chunk total
    __shared__ uint32_t s_wstot[EXP_WS + 1];           // pairs before warp-step ws (all tiles)
    __shared__ uint32_t s_gbase[BIN_TILES];            // final position of the chunk's first pair of tile t
    __shared__ uint32_t s_tileid[BIN_TILES];
    __shared__ uint32_t s_soff[BIN_TILES];             // per round: staged offset of tile t
    __shared__ uint32_t s_gadj[BIN_TILES];             // per round: final position = s_gadj[t] + staged index
    __shared__ uint32_t s_id[EXP_CAP];
    __shared__ uint32_t s_dep[EXP_CAP];
    __shared__ unsigned char s_t[EXP_CAP];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    const uint32_t c = blockIdx.x;
    if (c >= __ldg(a.num_chunks)) return;
    const uint4 d = __ldg(a.chunk_desc + c);
    const int bx8 = (int)(d.x % (uint32_t)a.bins_x) << BIN_SHIFT, by8 = (int)(d.x / (uint32_t)a.bins_x) << BIN_SHIFT;
    const int nws = (int)((d.z - d.y + 31u) >> 5);

    // 1. records of this thread: warp-steps warp and warp + 8
    uint32_t id[EXP_PER_WARP], dep[EXP_PER_WARP];
    uint64_t m[EXP_PER_WARP];
#pragma unroll
    for (int k = 0; k < EXP_PER_WARP; ++k) {
        const uint32_t r = d.y + (uint32_t)((warp + k * EXP_WARPS) * 32 + lane);
        id[k] = dep[k] = 0u;
        m[k] = 0ull;
        if (r < d.z) {
            id[k] = __ldg(a.rec_ids + r);
            m[k] = bin_mask(__ldg(a.tile_rects + id[k]), bx8, by8);
            dep[k] = __ldg(a.depths + id[k]);
        }
    }
    // 2. ballots per tile and pairs per warp-step
#pragma unroll
    for (int k = 0; k < EXP_PER_WARP; ++k) {
        const int ws = warp + k * EXP_WARPS;
        const uint32_t bl = warp_transpose32((uint32_t)m[k], lane);
        const uint32_t bh = warp_transpose32((uint32_t)(m[k] >> 32), lane);
        s_bal[ws][lane] = bl;
        s_bal[ws][32 + lane] = bh;
        const uint32_t tot = __reduce_add_sync(0xffffffffu, (uint32_t)(__popc(bl) + __popc(bh)));
        if (lane == 0) s_wstot[ws + 1] = tot;
    }
    __syncthreads();
    // 3. per tile: prefix over the warp-steps, tile id, final base; prefix of the warp-step totals
    if (tid < BIN_TILES) {
        uint32_t run = 0;
#pragma unroll
        for (int ws = 0; ws < EXP_WS; ++ws) {
            s_pre[ws][tid] = run;
            run += __popc(s_bal[ws][tid]);
        }
        s_pre[EXP_WS][tid] = run;
        const int tx = bx8 + (tid & (BIN_SIDE - 1)), ty = by8 + (tid >> BIN_SHIFT);
        const bool inside = tx < a.grid_x && ty < a.grid_y;
        const uint32_t tile = (uint32_t)(ty * a.grid_x + tx);
        s_tileid[tid] = tile;
        s_gbase[tid] = inside ? __ldg(a.tile_start + tile) + a.chunk_counts[(size_t)c * BIN_TILES + tid] : 0u;
    } else if (tid == BIN_TILES) {
        uint32_t run = 0;
        s_wstot[0] = 0;
#pragma unroll
        for (int ws = 1; ws <= EXP_WS; ++ws) {
            run += s_wstot[ws];
            s_wstot[ws] = run;
        }
    }
    __syncthreads();

    // 4. rounds of at most EXP_CAP pairs: whole warp-steps [ws_a, ws_b)
    const uint32_t lane_lt = (1u << lane) - 1u;
    int ws_a = 0;
    while (ws_a < nws) {
        int ws_b = ws_a + 1;
        while (ws_b < nws && s_wstot[ws_b + 1] - s_wstot[ws_a] <= (uint32_t)EXP_CAP) ++ws_b;
        const uint32_t round_pairs = s_wstot[ws_b] - s_wstot[ws_a];
        if (warp == 0) {
            // staged offsets: exclusive scan over the 64 tiles of the round's per-tile counts
            const uint32_t v0 = s_pre[ws_b][lane] - s_pre[ws_a][lane];
            const uint32_t v1 = s_pre[ws_b][32 + lane] - s_pre[ws_a][32 + lane];
            uint32_t i0 = v0, i1 = v1;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, dd);
                const uint32_t t1 = __shfl_up_sync(0xffffffffu, i1, dd);
                if (lane >= dd) { i0 += t0; i1 += t1; }
            }
            const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31);
            const uint32_t o0 = i0 - v0, o1 = tot0 + i1 - v1;
            s_soff[lane] = o0 - s_pre[ws_a][lane];            // slot = s_soff[t] + s_pre[ws][t] + rank
            s_soff[32 + lane] = o1 - s_pre[ws_a][32 + lane];
            s_gadj[lane] = s_gbase[lane] + s_pre[ws_a][lane] - o0;
            s_gadj[32 + lane] = s_gbase[32 + lane] + s_pre[ws_a][32 + lane] - o1;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < EXP_PER_WARP; ++k) {
            const int ws = warp + k * EXP_WARPS;
            if (ws >= ws_a && ws < ws_b) {
                uint64_t mm = m[k];
                while (mm) {
                    const int t = __ffsll((long long)mm) - 1;
                    mm &= mm - 1ull;
                    const uint32_t slot = s_soff[t] + s_pre[ws][t] + __popc(s_bal[ws][t] & lane_lt);
                    s_id[slot] = id[k];
                    s_dep[slot] = dep[k];
                    s_t[slot] = (unsigned char)t;
                }
            }
        }
        __syncthreads();
        for (uint32_t j = tid; j < round_pairs; j += EXP_THREADS) {
            const int t = s_t[j];
            const uint32_t g = s_gadj[t] + j;
            a.keys_out[g] = ((uint64_t)s_tileid[t] << 32) | (uint64_t)s_dep[j];  // GSCuda.cu:466-471
            a.vals_out[g] = s_id[j];
        }
        __syncthreads();
        ws_a = ws_b;
    }
}

}  // namespace

size_t expand_temp_bytes(size_t R) {
    return 128 + 2 * align128((MAX_BINS + 1) * sizeof(uint32_t)) + align128(max_chunks(R) * sizeof(uint4)) +
           align128(max_chunks(R) * BIN_TILES * sizeof(uint32_t));
}

int launch_bin_expand(const ExpandPlan& p, cudaStream_t s) {
    const int nbins = p.bins_x * p.bins_y;
    const int tiles = p.grid_x * p.grid_y;
    if (nbins < 1 || nbins > MAX_BINS || p.n_records == 0 || p.n_records > p.num_rendered) return GSR_ERR_INVALID_ARG;
    if (p.n_records >= ((size_t)1 << 31)) return GSR_ERR_TOO_MANY_PAIRS;
    ExpandTemp t = carve(p.temp, p.num_rendered);
    const unsigned nchunk_bound =
        (unsigned)(p.n_records / EXP_CHUNK + std::min<size_t>((size_t)nbins, p.n_records) + 1);  // <= max_chunks(R)
    int launches = 0;
    GSR_CUDA_TRY(launch_pdl(bin_bounds_kernel, dim3((nbins + 1 + EXP_WARPS - 1) / EXP_WARPS), dim3(EXP_THREADS), 0, s,
                            p.rec_bins, (uint32_t)p.n_records, nbins, t.bin_start));
    ++launches;
    GSR_CUDA_TRY(launch_pdl(chunk_table_kernel, dim3(1), dim3(TABLE_THREADS), 0, s, nbins, (const uint32_t*)t.bin_start,
                            t.bin_chunk_first, t.chunk_desc));
    ++launches;
    ExpandArgs a;
    a.rec_ids = p.rec_ids;
    a.chunk_desc = t.chunk_desc;
    a.num_chunks = t.bin_chunk_first + nbins;
    a.chunk_counts = t.chunk_counts;
    a.tile_rects = reinterpret_cast<const uint2*>(p.tile_rects);
    a.depths = p.depths;
    a.tile_start = p.tile_counts;
    a.keys_out = p.keys_out;
    a.vals_out = p.vals_out;
    a.grid_x = p.grid_x; a.grid_y = p.grid_y; a.bins_x = p.bins_x;
    GSR_CUDA_TRY(launch_pdl(expand_count_kernel, dim3(nchunk_bound), dim3(EXP_THREADS), 0, s, a));
    ++launches;
    GSR_CUDA_TRY(launch_pdl(expand_scan_chunks_kernel, dim3(nbins), dim3(BIN_TILES), 0, s,
                            (const uint32_t*)t.bin_chunk_first, t.chunk_counts, p.tile_counts, p.grid_x, p.grid_y,
                            p.bins_x));
    ++launches;
    GSR_CUDA_TRY(launch_pdl(tile_ranges_kernel, dim3(1), dim3(TABLE_THREADS), 0, s, p.tile_counts, tiles,
                            reinterpret_cast<uint2*>(p.ranges), p.r1_quirk ? 1 : 0));
    ++launches;
    GSR_CUDA_TRY(launch_pdl(expand_fill_kernel, dim3(nchunk_bound), dim3(EXP_THREADS), 0, s, a));
    ++launches;
    return launches;
}

}  // namespace gsr
