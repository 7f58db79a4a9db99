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
//   2. the record list of a bin is cut into chunks of 256; a chunk turns each record's rect into a
//      64-bit tile mask of the bin, and a 32x32 bit-matrix transpose across the lanes of a warp turns 32
//      masks into 64 ballots (one per tile: which of the 32 records touch it).  popc of a ballot is a
//      count, popc below the own lane is a stable rank;
//   3. expand_count: per chunk and tile, pairs emitted; a scan over the chunks of a bin and over all
//      tiles (row-major tile id) gives every (chunk, tile) its final offset and every tile its range —
//      identifyTileRanges falls out of the scan, the sorted keys are never re-read;
//   4. expand_fill: every chunk walks its ballots tile by tile (that IS the sorted order), stages the pairs
//      in shared memory and writes 12 bytes per pair as contiguous runs straight into their final position.
// Pair-level HBM traffic drops from 56 B/pair (8 written by the duplication, 2 x 16 by the first
// pass, 8 + 4 + 12 by the last, 8 re-read for the ranges) to the 12 B/pair of the result itself.
#include <algorithm>

#include "gsr_common.cuh"

namespace gsr {

namespace {

#ifndef GSR_FILL_FAST
#define GSR_FILL_FAST 1
#endif
// GSR_FILL_BALANCED=1: every thread of the fill pass takes 8 CONSECUTIVE staged slots and finds its position by search
// (tile offset, warp-step prefix, rank in the ballot) instead of walking the ballots of one (tile, quarter).  Built,
// parity-green, measured SLOWER (profiles/r02i_ab_*.txt: fill 0.107 vs 0.088 ms at C2, 0.381 vs 0.353 ms expansion at
// C3): the nine dependent shared-memory loads of the search and the single-warp table build cost more than the idle
// lanes of the walk they replace.  Off.
#ifndef GSR_FILL_BALANCED
#define GSR_FILL_BALANCED 0
#endif
constexpr int EXP_THREADS = 256;
constexpr int EXP_WARPS = EXP_THREADS / 32;
constexpr int EXP_CHUNK = 256;                 // records per chunk: one per thread, one warp-step (32 records) per warp
constexpr int EXP_WS = EXP_CHUNK / 32;         // warp-steps per chunk
constexpr int EXP_CAP = 2048;                  // pairs staged per round (>= 32 * 64, the most one warp-step emits)
constexpr int EXP_QUARTERS = EXP_THREADS / BIN_TILES;   // fill: thread = (tile, quarter of the warp-steps)
constexpr int EXP_WS_PER_Q = EXP_WS / EXP_QUARTERS;
constexpr int TABLE_THREADS = 1024;
#ifndef GSR_ESCAN_THREADS
#define GSR_ESCAN_THREADS 1024   // 256: 25 us at C2 (the longest bin's column of chunk counts walked by 4 threads per tile)
#endif
constexpr int ESCAN_THREADS = GSR_ESCAN_THREADS;
constexpr int ESCAN_WARPS = ESCAN_THREADS / 32;
constexpr int SCAN_PARTS = ESCAN_THREADS / BIN_TILES;   // chunk scan: thread = (tile, part of the bin's chunks)

__host__ __device__ inline size_t align128(size_t v) { return (v + 127) / 128 * 128; }

struct ExpandTemp {
    uint32_t* bin_start;        // [MAX_BINS + 1] first record of every bin (+ n_records)
    uint32_t* bin_chunk_first;  // [MAX_BINS + 1] first chunk of every bin (+ number of chunks)
    uint32_t* done_counter;     // [1] bins whose chunk scan has finished (the last one scans the tiles)
    uint4* chunk_desc;          // [max_chunks] bin, first record, end record, -
    uint32_t* chunk_counts;     // [max_chunks][64] pairs per tile, then exclusive prefix over the bin's chunks
    uint32_t* chunk_ballots;    // [max_chunks][EXP_WS][64] per warp-step and tile: which of the 32 records touch it
    uint2* rec_data;            // [R] per record: Gaussian id, depth bits (gathered once, by the count pass)
};

size_t max_chunks(size_t R) { return R / EXP_CHUNK + std::min<size_t>((size_t)MAX_BINS, R) + 2; }  // + 1 spare: kernels read the descriptor before the bound check

ExpandTemp carve(char* temp, size_t R) {
    ExpandTemp t;
    char* c = reinterpret_cast<char*>(align128(reinterpret_cast<size_t>(temp)));
    t.bin_start = reinterpret_cast<uint32_t*>(c);
    c += align128((MAX_BINS + 1) * sizeof(uint32_t));
    t.bin_chunk_first = reinterpret_cast<uint32_t*>(c);
    c += align128((MAX_BINS + 1) * sizeof(uint32_t));
    t.done_counter = reinterpret_cast<uint32_t*>(c);
    c += 128;
    t.chunk_desc = reinterpret_cast<uint4*>(c);
    c += align128(max_chunks(R) * sizeof(uint4));
    t.chunk_counts = reinterpret_cast<uint32_t*>(c);
    c += align128(max_chunks(R) * BIN_TILES * sizeof(uint32_t));
    t.chunk_ballots = reinterpret_cast<uint32_t*>(c);
    c += align128(max_chunks(R) * EXP_WS * BIN_TILES * sizeof(uint32_t));
    t.rec_data = reinterpret_cast<uint2*>(c);
    return t;
}

// ---- bin boundaries -----------------------------------------------------------------------------
// One warp per bin b in [0, nbins]: bin_start[b] = first record whose bin id is >= b (32-ary search
// in the sorted bin ids: 5 round trips for 2^25 records instead of 25).  Only needed when the bin ids
// took more than one radix pass; with a single pass the digit histogram IS the per-bin count.
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

// exclusive scan of one value per thread over a TABLE_THREADS-wide CTA; returns the exclusive prefix
__device__ __forceinline__ uint32_t table_scan(const uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads();  // s_warp may still be read from a previous call
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
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    if (total) *total = s_warp[32];
    return s_warp[warp] + incl - v;
}

// ---- chunk table --------------------------------------------------------------------------------
// Per-bin record ranges (from the single-pass digit histogram `bin_counts`, or from bin_start when the bin
// ids took two passes), chunks per bin, their exclusive scan, one descriptor per chunk.  Every CTA of the
// grid rebuilds the (tiny) per-bin tables in its shared memory and writes its share of the descriptors —
// one CTA writing all 14 k descriptors of a C2 frame was a 12 us serial step of the frame's dependency chain;
// CTA 0 alone publishes the per-bin tables and resets the scan's arrival counter.
__global__ void __launch_bounds__(TABLE_THREADS) chunk_table_kernel(const int nbins, const uint32_t* __restrict__ bin_counts,
                                                                    uint32_t* __restrict__ bin_start,
                                                                    uint32_t* __restrict__ bin_chunk_first,
                                                                    uint4* __restrict__ chunk_desc,
                                                                    uint32_t* __restrict__ done_counter) {
    constexpr int PER = MAX_BINS / TABLE_THREADS;  // 4 consecutive bins per thread
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_start[MAX_BINS + 1];
    __shared__ uint32_t s_first[MAX_BINS + 1];
    const int tid = threadIdx.x;
    const bool lead = blockIdx.x == 0;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    if (lead && tid == 0) *done_counter = 0;
    uint32_t cnt[PER], st[PER + 1], nch[PER], tsum = 0;
    if (bin_counts) {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int b = tid * PER + j;
            cnt[j] = (b < nbins) ? __ldg(bin_counts + b) : 0u;
            tsum += cnt[j];
        }
        uint32_t run = table_scan(tsum, s_warp, nullptr);
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            st[j] = run;
            run += cnt[j];
        }
        st[PER] = run;
    } else {
#pragma unroll
        for (int j = 0; j <= PER; ++j) {
            const int b = tid * PER + j;
            st[j] = (b <= nbins) ? bin_start[b] : 0u;
        }
    }
    tsum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int b = tid * PER + j;
        nch[j] = (b < nbins) ? (st[j + 1] - st[j] + EXP_CHUNK - 1) / EXP_CHUNK : 0u;
        tsum += nch[j];
    }
    uint32_t total = 0;
    uint32_t run = table_scan(tsum, s_warp, &total);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int b = tid * PER + j;
        if (b < nbins) {
            s_start[b] = st[j];
            s_first[b] = run;
            if (lead) {
                bin_chunk_first[b] = run;
                if (bin_counts) bin_start[b] = st[j];
            }
            run += nch[j];
            if (b == nbins - 1) {
                s_start[nbins] = st[j + 1];
                s_first[nbins] = run;
                if (lead) {
                    bin_chunk_first[nbins] = run;  // number of chunks
                    if (bin_counts) bin_start[nbins] = st[j + 1];
                }
            }
        }
    }
    __syncthreads();
    // descriptors, one chunk per thread and round: bin by binary search over the first-chunk table
    for (uint32_t k = blockIdx.x * TABLE_THREADS + tid; k < total; k += gridDim.x * TABLE_THREADS) {
        int lo = 0, hi = nbins - 1;  // largest b with s_first[b] <= k (empty bins share their successor's value)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_first[mid] <= k) lo = mid; else hi = mid - 1;
        }
        const uint32_t s0 = s_start[lo] + (k - s_first[lo]) * EXP_CHUNK;
        chunk_desc[k] = make_uint4((uint32_t)lo, s0, min(s0 + (uint32_t)EXP_CHUNK, s_start[lo + 1]), 0u);
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
// (bit i of the result = bit j of lane i's input).  Five butterfly stages, one shuffle each: at stage j a
// lane keeps the half of every 2j-bit group that stays (`keep`) and takes the other half, shifted by j,
// from lane ^ j.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, const int lane) {
    uint32_t m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        const bool up = (lane & j) != 0;
        const uint32_t keep = up ? ~m : m;
        const uint32_t t = up ? (y >> j) : (y << j);
        x = (x & keep) | (t & ~keep);
        m ^= m << (j >> 1);
    }
    return x;
}

// KEYS=false (lean callers, ExpandPlan::keys_out == nullptr): the 64-bit sorted keys are not materialised — the forward
// pass never reads them back (the tile ranges fall out of the scan) — so the expansion carries Gaussian ids only: no
// depth gather and no (id, depth) records in the count pass, 4-byte staging and 4 instead of 12 bytes per pair written
// by the fill pass.  point_list and ranges are the same bits either way.
template <bool KEYS> struct StagedRec { using type = uint2; };
template <> struct StagedRec<false> { using type = uint32_t; };

struct ExpandArgs {
    const uint32_t* rec_ids;
    const uint4* chunk_desc;
    const uint32_t* num_chunks;  // device: bin_chunk_first[nbins]
    uint32_t* chunk_counts;
    uint32_t* chunk_ballots;
    uint2* rec_data;
    const uint2* tile_rects;
    const uint32_t* depths;
    const uint32_t* tile_start;  // [tiles] exclusive scan of the per-tile totals
    uint64_t* keys_out;
    uint32_t* vals_out;
    int grid_x, grid_y, bins_x;
};

// ---- pass 1: ballots and pairs per (chunk, tile) --------------------------------------------------------
// Thread i of chunk c owns record i: gathers its tile rect and depth bits (the only random accesses of the
// expansion), builds the tile mask of the bin; every warp transposes its 32 masks into 64 ballots.  Left for
// the fill pass: the ballots, (id, depth) per record, and the chunk's pairs per tile.
#ifndef GSR_COUNT_CHUNKS
#define GSR_COUNT_CHUNKS 2   // chunks per CTA: their dependent load chains (descriptor -> id -> rect, depth) overlap
#endif
constexpr int CNT_CHUNKS = GSR_COUNT_CHUNKS;
template <bool KEYS>
__global__ void __launch_bounds__(EXP_THREADS) expand_count_kernel(const ExpandArgs a) {
    __shared__ uint32_t s_cnt[CNT_CHUNKS][BIN_TILES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    // descriptors fetched alongside the chunk count (the table holds max_chunks(R) entries and the grid is
    // rounded up inside that bound): one dependent round trip less in a kernel that is a chain of them
    const uint32_t nchunks = __ldg(a.num_chunks);
    const uint32_t cbase = blockIdx.x * CNT_CHUNKS;
    uint4 d[CNT_CHUNKS];
#pragma unroll
    for (int k = 0; k < CNT_CHUNKS; ++k) d[k] = __ldg(a.chunk_desc + cbase + k);
    if (cbase >= nchunks) return;
    if (tid < CNT_CHUNKS * BIN_TILES) (&s_cnt[0][0])[tid] = 0;
    __syncthreads();
    // every chunk's id load, then every chunk's gathers: the chains of the CTA's chunks run side by side
    uint32_t id[CNT_CHUNKS];
    bool have[CNT_CHUNKS];
#pragma unroll
    for (int k = 0; k < CNT_CHUNKS; ++k) {
        const uint32_t r = d[k].y + (uint32_t)tid;
        have[k] = (cbase + k < nchunks) && r < d[k].z;
        id[k] = have[k] ? __ldg(a.rec_ids + r) : 0u;
    }
    uint2 rect[CNT_CHUNKS];
    uint32_t dep[CNT_CHUNKS];
#pragma unroll
    for (int k = 0; k < CNT_CHUNKS; ++k) {
        rect[k] = have[k] ? __ldg(a.tile_rects + id[k]) : make_uint2(0u, 0u);
        dep[k] = (KEYS && have[k]) ? __ldg(a.depths + id[k]) : 0u;
    }
#pragma unroll
    for (int k = 0; k < CNT_CHUNKS; ++k) {
        const uint32_t c = cbase + k;
        if (c >= nchunks) break;  // block-uniform
        const int bx8 = (int)(d[k].x % (uint32_t)a.bins_x) << BIN_SHIFT, by8 = (int)(d[k].x / (uint32_t)a.bins_x) << BIN_SHIFT;
        uint64_t m = 0ull;
        if (have[k]) {
            m = bin_mask(rect[k], bx8, by8);
            if (KEYS) a.rec_data[d[k].y + (uint32_t)tid] = make_uint2(id[k], dep[k]);
        }
        const uint32_t bl = warp_transpose32((uint32_t)m, lane);
        const uint32_t bh = warp_transpose32((uint32_t)(m >> 32), lane);
        uint32_t* bal = a.chunk_ballots + ((size_t)c * EXP_WS + warp) * BIN_TILES;
        bal[lane] = bl;
        bal[32 + lane] = bh;
        if (bl) atomicAdd(&s_cnt[k][lane], (uint32_t)__popc(bl));
        if (bh) atomicAdd(&s_cnt[k][32 + lane], (uint32_t)__popc(bh));
    }
    __syncthreads();
    if (tid < CNT_CHUNKS * BIN_TILES) {
        const uint32_t c = cbase + (uint32_t)(tid / BIN_TILES);
        if (c < nchunks) a.chunk_counts[(size_t)c * BIN_TILES + (tid % BIN_TILES)] = (&s_cnt[0][0])[tid];
    }
}

// ---- scans: over the chunks of every bin per tile, then over the tiles -> ranges ---------------------------
// One CTA per bin; thread = (tile t, part): the bin's chunks are split into SCAN_PARTS contiguous parts, each
// summed (independent loads), combined through shared memory and rewritten as exclusive prefixes; the bin's
// per-tile totals go to tile_counts[tile id].  Every tile of the grid belongs to exactly one bin, so
// tile_counts is written completely.  The CTA that finishes last then scans tile_counts in row-major tile
// order: tile_counts becomes the start of every tile's list and ranges[tile] = (start, start + count); a tile
// nothing touches keeps (0, 0) exactly like the reference's cleared and never written entry (GSCuda.cu:800,
// 504-538).
__global__ void __launch_bounds__(ESCAN_THREADS) expand_scan_kernel(const uint32_t* __restrict__ bin_chunk_first,
                                                                   uint32_t* __restrict__ chunk_counts,
                                                                   uint32_t* tile_counts, const int grid_x,
                                                                   const int grid_y, const int bins_x, const int nbins,
                                                                   uint2* __restrict__ ranges, const int r1_quirk,
                                                                   uint32_t* done_counter) {
    __shared__ uint32_t s_part[SCAN_PARTS][BIN_TILES];
    __shared__ uint32_t s_warp[ESCAN_WARPS];
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = tid & (BIN_TILES - 1), part = tid >> 6, b = blockIdx.x;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    const uint32_t first = __ldg(bin_chunk_first + b), last = __ldg(bin_chunk_first + b + 1);
    const uint32_t n = last - first, per = (n + SCAN_PARTS - 1) / SCAN_PARTS;
    const uint32_t k0 = first + min(n, (uint32_t)part * per), k1 = first + min(n, (uint32_t)(part + 1) * per);
    uint32_t sum = 0;
    {
        const uint32_t* col = chunk_counts + (size_t)k0 * BIN_TILES + t;
        uint32_t k = k0;
        for (; k + 8 <= k1; k += 8, col += 8 * BIN_TILES) {
            uint32_t v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = col[q * BIN_TILES];
#pragma unroll
            for (int q = 0; q < 8; ++q) sum += v[q];
        }
        for (; k < k1; ++k, col += BIN_TILES) sum += col[0];
    }
    s_part[part][t] = sum;
    __syncthreads();
    uint32_t run = 0, total = 0;
#pragma unroll
    for (int q = 0; q < SCAN_PARTS; ++q) {
        const uint32_t v = s_part[q][t];
        if (q < part) run += v;
        total += v;
    }
    {
        uint32_t* col = chunk_counts + (size_t)k0 * BIN_TILES + t;
        uint32_t k = k0;
        for (; k + 4 <= k1; k += 4, col += 4 * BIN_TILES) {
            const uint32_t v0 = col[0], v1 = col[BIN_TILES], v2 = col[2 * BIN_TILES], v3 = col[3 * BIN_TILES];
            col[0] = run; run += v0;
            col[BIN_TILES] = run; run += v1;
            col[2 * BIN_TILES] = run; run += v2;
            col[3 * BIN_TILES] = run; run += v3;
        }
        for (; k < k1; ++k, col += BIN_TILES) {
            const uint32_t v = col[0];
            col[0] = run;
            run += v;
        }
    }
    if (part == 0) {
        const int tx = ((b % bins_x) << BIN_SHIFT) + (t & (BIN_SIDE - 1)), ty = ((b / bins_x) << BIN_SHIFT) + (t >> BIN_SHIFT);
        if (tx < grid_x && ty < grid_y) tile_counts[ty * grid_x + tx] = total;
    }
    // ---- the last bin to finish scans the tiles ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(done_counter, 1u) == (uint32_t)(nbins - 1)) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const int tiles = grid_x * grid_y;
    constexpr int ITEMS = 8;
    for (int base = 0; base < tiles; base += ESCAN_THREADS * ITEMS) {
        const int i0 = base + tid * ITEMS;
        uint32_t v[ITEMS], tsum = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            v[j] = (i0 + j < tiles) ? __ldcg(tile_counts + i0 + j) : 0u;
            tsum += v[j];
        }
        uint32_t incl = tsum;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, incl, dd);
            if (lane >= dd) incl += x;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
        {
            // exclusive prefix of the warp totals: one shuffle scan per warp over (up to) 32 totals
            const uint32_t wv = (lane < ESCAN_WARPS) ? s_warp[lane] : 0u;
            uint32_t wi = wv;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const uint32_t x = __shfl_up_sync(0xffffffffu, wi, dd);
                if (lane >= dd) wi += x;
            }
            woff = __shfl_sync(0xffffffffu, wi - wv, warp);
        }
        uint32_t r = s_carry + woff + incl - tsum;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (i0 + j < tiles) {
                tile_counts[i0 + j] = r;
                // GSRast-compat with a single pair in the frame: .x = 0 is written, .y never is (GSCuda.cu:533-536)
                ranges[i0 + j] = (v[j] && !r1_quirk) ? make_uint2(r, r + v[j]) : make_uint2(0u, 0u);
            }
            r += v[j];
        }
        __syncthreads();
        if (tid == ESCAN_THREADS - 1) s_carry = r;
        __syncthreads();
    }
}

// ---- pass 2: rank, stage, write ------------------------------------------------------------------------
// Thread = (tile t of the bin, quarter q of the chunk's warp-steps).  The pairs of tile t are the set bits of
// its ballots, warp-step by warp-step, lane by lane — exactly the order of the sorted list — so a thread walks
// its ballots and appends (id, depth) records to the tile's run in the staging buffer; the buffer is then
// written out as one contiguous run per tile.  Chunks that emit more than EXP_CAP pairs take several rounds of
// whole warp-steps.
template <bool KEYS>
__device__ __forceinline__ void emit_pair(const ExpandArgs& a, const uint32_t g, const uint32_t tile, const uint2 o) {
    a.keys_out[g] = ((uint64_t)tile << 32) | (uint64_t)o.y;  // GSCuda.cu:466-471
    a.vals_out[g] = o.x;
}
template <bool KEYS>
__device__ __forceinline__ void emit_pair(const ExpandArgs& a, const uint32_t g, const uint32_t, const uint32_t o) {
    a.vals_out[g] = o;
}

template <bool KEYS>
__global__ void __launch_bounds__(EXP_THREADS) expand_fill_kernel(const ExpandArgs a) {
    using Rec = typename StagedRec<KEYS>::type;
    __shared__ uint32_t s_bal[EXP_WS][BIN_TILES];      // ballot of tile t in warp-step ws
    __shared__ uint32_t s_pre[EXP_WS + 1][BIN_TILES];  // pairs of tile t before warp-step ws; [EXP_WS] = chunk total
    __shared__ uint32_t s_wspart[2][EXP_WS];           // pairs of warp-step ws, tiles 0-31 / 32-63
    __shared__ uint32_t s_round[EXP_WS + 1];           // round r covers warp-steps [s_round[r], s_round[r+1])
    __shared__ uint32_t s_nrounds;
    __shared__ uint32_t s_round_pairs;
    __shared__ uint32_t s_gbase[BIN_TILES];            // final position of the chunk's first pair of tile t
    __shared__ uint32_t s_tileid[BIN_TILES];
    __shared__ uint32_t s_soff[BIN_TILES];             // per round: staged offset of tile t (minus s_pre[ws_a][t])
    __shared__ uint32_t s_gadj[BIN_TILES];             // per round: final position = s_gadj[t] + staged index
    __shared__ Rec s_rec[EXP_CHUNK];                   // id (, depth bits) of the chunk's records
    __shared__ Rec s_out[EXP_CAP];                     // staged pairs: id (, depth bits)
    __shared__ unsigned char s_t[EXP_CAP];             // staged pairs: tile of the bin

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();
    const uint32_t c = blockIdx.x;
    const uint32_t nchunks = __ldg(a.num_chunks);
    const uint4 d = __ldg(a.chunk_desc + c);  // in bounds for every CTA of the grid (see expand_count_kernel)
    if (c >= nchunks) return;
    const int bx8 = (int)(d.x % (uint32_t)a.bins_x) << BIN_SHIFT, by8 = (int)(d.x / (uint32_t)a.bins_x) << BIN_SHIFT;
    const int nws = (int)((d.z - d.y + 31u) >> 5);

    // 1. everything this chunk needs, all loads independent of each other
    {
        const uint32_t* bal = a.chunk_ballots + (size_t)c * EXP_WS * BIN_TILES;
        const uint32_t b0 = __ldg(bal + tid), b1 = __ldg(bal + EXP_THREADS + tid);
        const uint32_t r = d.y + (uint32_t)tid;
        Rec rec = Rec();
        if (r < d.z) {
            if constexpr (KEYS) rec = a.rec_data[r];
            else rec = __ldg(a.rec_ids + r);
        }
        uint32_t gb = 0, tile = 0;
        if (tid < BIN_TILES) {
            const int tx = bx8 + (tid & (BIN_SIDE - 1)), ty = by8 + (tid >> BIN_SHIFT);
            tile = (uint32_t)(ty * a.grid_x + tx);
            if (tx < a.grid_x && ty < a.grid_y) gb = __ldg(a.tile_start + tile) + a.chunk_counts[(size_t)c * BIN_TILES + tid];
        }
        (&s_bal[0][0])[tid] = b0;
        (&s_bal[0][0])[EXP_THREADS + tid] = b1;
        s_rec[tid] = rec;
        if (tid < BIN_TILES) { s_gbase[tid] = gb; s_tileid[tid] = tile; }
    }
    __syncthreads();
#if GSR_FILL_BALANCED
    // Balanced single-round path (chunks of at most EXP_CAP pairs, i.e. nearly all of them).  The pairs of the chunk
    // form one staged sequence — tile-major, then warp-step, then lane: the sorted order — and every thread takes
    // FILL_ITEMS CONSECUTIVE slots of it: one search (tile by its offset, warp-step by the tile's prefix, rank inside
    // the ballot) positions the thread, then it walks set bits across ballot / tile boundaries.  Every thread does the
    // same amount of work whatever the shape of the rects, where the thread-per-(tile, quarter) walk below leaves most
    // lanes of a warp idle (tiles at the edge of a rect list few records): ~130 thread-slots per pair there.
    {
        constexpr int FILL_ITEMS = EXP_CAP / EXP_THREADS;  // 8
        if (warp == 0) {
            // per tile: prefix over the warp-steps; exclusive scan of the tile totals over the 64 tiles (lane, 32 + lane)
            uint32_t run0 = 0, run1 = 0;
#pragma unroll
            for (int ws = 0; ws < EXP_WS; ++ws) {
                s_pre[ws][lane] = run0;
                s_pre[ws][32 + lane] = run1;
                run0 += (uint32_t)__popc(s_bal[ws][lane]);
                run1 += (uint32_t)__popc(s_bal[ws][32 + lane]);
            }
            s_pre[EXP_WS][lane] = run0;
            s_pre[EXP_WS][32 + lane] = run1;
            uint32_t i0 = run0, i1 = run1;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, dd);
                const uint32_t t1 = __shfl_up_sync(0xffffffffu, i1, dd);
                if (lane >= dd) { i0 += t0; i1 += t1; }
            }
            const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31), tot1 = __shfl_sync(0xffffffffu, i1, 31);
            const uint32_t o0 = i0 - run0, o1 = tot0 + i1 - run1;
            s_soff[lane] = o0;
            s_soff[32 + lane] = o1;
            s_gadj[lane] = s_gbase[lane] - o0;  // final position = s_gadj[t] + staged index
            s_gadj[32 + lane] = s_gbase[32 + lane] - o1;
            if (lane == 0) s_round_pairs = tot0 + tot1;
        }
        __syncthreads();
        const uint32_t total = s_round_pairs;
        if (total <= (uint32_t)EXP_CAP) {
            const uint32_t s0 = (uint32_t)tid * FILL_ITEMS;
            if (s0 < total) {
                // tile: the largest t with s_soff[t] <= s0 (empty tiles share their successor's offset, so this is the owner)
                int t = 0;
#pragma unroll
                for (int step = BIN_TILES / 2; step >= 1; step >>= 1)
                    if (s_soff[t + step] <= s0) t += step;
                uint32_t r = s0 - s_soff[t];
                int ws = 0;
#pragma unroll
                for (int step = EXP_WS / 2; step >= 1; step >>= 1)
                    if (s_pre[ws + step][t] <= r) ws += step;
                r -= s_pre[ws][t];
                uint32_t bits = s_bal[ws][t];
                for (; r > 0; --r) bits &= bits - 1u;  // drop the set bits that belong to the slots before s0
                const uint32_t n_mine = min((uint32_t)FILL_ITEMS, total - s0);
                for (uint32_t i = 0; i < n_mine; ++i) {
                    while (bits == 0u) {  // next warp-step, next tile (s0 + i < total: there is one)
                        if (++ws == EXP_WS) { ws = 0; ++t; }
                        bits = s_bal[ws][t];
                    }
                    const int l = __ffs((int)bits) - 1;
                    bits &= bits - 1u;
                    const uint32_t slot = s0 + i;
                    s_out[slot ^ ((slot >> 4) & 7u)] = s_rec[ws * 32 + l];  // XOR swizzle: 2-way instead of 16-way conflicts
                    s_t[slot] = (unsigned char)t;
                }
            }
            __syncthreads();
            for (uint32_t j = tid; j < total; j += EXP_THREADS) {
                const int tt = s_t[j];
                const uint32_t g = s_gadj[tt] + j;
                emit_pair<KEYS>(a, g, s_tileid[tt], s_out[j ^ ((j >> 4) & 7u)]);
            }
            return;
        }
        __syncthreads();  // a big chunk: the rounds below rebuild their tables
    }
#elif GSR_FILL_FAST
    // Fast path (chunks of at most EXP_CAP pairs, i.e. nearly all of them): a single round, no further table
    // building.  Thread (tile t = tid & 63, quarter q) has t = 32 * (warp & 1) + lane, so every warp derives what
    // its threads need from the ballots on its own — its half's per-tile counts (scanned across the lanes) and the
    // other half's total — instead of steps 2-3 and the round prologue below with their three barriers.
    {
        const int half = warp & 1, q2 = (warp >> 1) * EXP_WS_PER_Q;
        uint32_t mine_before = 0, mine_tot = 0, other_tot = 0;
#pragma unroll
        for (int ws = 0; ws < EXP_WS; ++ws) {
            const uint32_t n_own = (uint32_t)__popc(s_bal[ws][half * 32 + lane]);
            const uint32_t n_oth = (uint32_t)__popc(s_bal[ws][(half ^ 1) * 32 + lane]);
            if (ws < q2) mine_before += n_own;
            mine_tot += n_own;
            other_tot += n_oth;
        }
        uint32_t inc = mine_tot;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, dd);
            if (lane >= dd) inc += up;
        }
        const uint32_t own_T = __shfl_sync(0xffffffffu, inc, 31);
        const uint32_t other_T = __reduce_add_sync(0xffffffffu, other_tot);
        const uint32_t total = own_T + other_T;  // identical in every warp
        if (total <= (uint32_t)EXP_CAP) {
            const int t = half * 32 + lane;
            const uint32_t o = inc - mine_tot + (half ? other_T : 0u);  // staged offset of tile t's run
            if (q2 == 0) s_gadj[t] = s_gbase[t] - o;                    // final position = s_gadj[t] + staged index
            uint32_t slot = o + mine_before;
            const int w1 = min(nws, q2 + EXP_WS_PER_Q);
            for (int ws = q2; ws < w1; ++ws) {
                uint32_t bits = s_bal[ws][t];
                while (bits) {
                    const int l = __ffs((int)bits) - 1;
                    bits &= bits - 1u;
                    s_out[slot] = s_rec[ws * 32 + l];
                    s_t[slot] = (unsigned char)t;
                    ++slot;
                }
            }
            __syncthreads();
            for (uint32_t j = tid; j < total; j += EXP_THREADS) {
                const int tt = s_t[j];
                const uint32_t g = s_gadj[tt] + j;
                emit_pair<KEYS>(a, g, s_tileid[tt], s_out[j]);
            }
            return;
        }
    }
#endif
    // 2. per tile: prefix over the warp-steps; pairs per warp-step
    if (tid < BIN_TILES) {
        uint32_t run = 0;
#pragma unroll
        for (int ws = 0; ws < EXP_WS; ++ws) {
            const uint32_t n = (uint32_t)__popc(s_bal[ws][tid]);
            s_pre[ws][tid] = run;
            run += n;
            const uint32_t tot = __reduce_add_sync(0xffffffffu, n);
            if (lane == 0) s_wspart[warp][ws] = tot;
        }
        s_pre[EXP_WS][tid] = run;
    }
    __syncthreads();
    // 3. rounds of whole warp-steps, at most EXP_CAP pairs each (one thread; usually a single round)
    if (tid == 0) {
        uint32_t nr = 0, acc = 0;
        s_round[0] = 0;
        for (int ws = 0; ws < nws; ++ws) {
            const uint32_t n = s_wspart[0][ws] + s_wspart[1][ws];
            if (acc + n > (uint32_t)EXP_CAP) {
                s_round[++nr] = (uint32_t)ws;
                acc = 0;
            }
            acc += n;
        }
        s_round[++nr] = (uint32_t)nws;
        s_nrounds = nr;
    }
    __syncthreads();
    const int nrounds = (int)s_nrounds;
    const int t = tid & (BIN_TILES - 1), q = tid >> 6;
    for (int rd = 0; rd < nrounds; ++rd) {
        const int ws_a = (int)s_round[rd], ws_b = (int)s_round[rd + 1];
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
            const uint32_t tot1 = __shfl_sync(0xffffffffu, i1, 31);
            const uint32_t o0 = i0 - v0, o1 = tot0 + i1 - v1;
            s_soff[lane] = o0 - s_pre[ws_a][lane];            // slot = s_soff[t] + s_pre[ws][t] + rank in the ballot
            s_soff[32 + lane] = o1 - s_pre[ws_a][32 + lane];
            s_gadj[lane] = s_gbase[lane] + s_pre[ws_a][lane] - o0;
            s_gadj[32 + lane] = s_gbase[32 + lane] + s_pre[ws_a][32 + lane] - o1;
            if (lane == 0) s_round_pairs = tot0 + tot1;
        }
        __syncthreads();
        const uint32_t round_pairs = s_round_pairs;
        {
            const int w0 = max(ws_a, q * EXP_WS_PER_Q), w1 = min(ws_b, (q + 1) * EXP_WS_PER_Q);
            if (w0 < w1) {
                uint32_t slot = s_soff[t] + s_pre[w0][t];
                for (int ws = w0; ws < w1; ++ws) {
                    uint32_t bits = s_bal[ws][t];
                    while (bits) {
                        const int l = __ffs((int)bits) - 1;
                        bits &= bits - 1u;
                        s_out[slot] = s_rec[ws * 32 + l];
                        s_t[slot] = (unsigned char)t;
                        ++slot;
                    }
                }
            }
        }
        __syncthreads();
        for (uint32_t j = tid; j < round_pairs; j += EXP_THREADS) {
            const int tt = s_t[j];
            const uint32_t g = s_gadj[tt] + j;
            emit_pair<KEYS>(a, g, s_tileid[tt], s_out[j]);
        }
        __syncthreads();
    }
}

}  // namespace

size_t expand_temp_bytes(size_t R) {
    return 256 + 2 * align128((MAX_BINS + 1) * sizeof(uint32_t)) + align128(max_chunks(R) * sizeof(uint4)) +
           align128(max_chunks(R) * BIN_TILES * sizeof(uint32_t)) +
           align128(max_chunks(R) * EXP_WS * BIN_TILES * sizeof(uint32_t)) + align128(R * sizeof(uint2));
}

int launch_bin_expand(const ExpandPlan& p, cudaStream_t s, cudaEvent_t* ev) {
    const int nbins = p.bins_x * p.bins_y;
    if (nbins < 1 || nbins > MAX_BINS || p.n_records == 0 || p.n_records > p.num_rendered) return GSR_ERR_INVALID_ARG;
    if (p.n_records >= ((size_t)1 << 31)) return GSR_ERR_TOO_MANY_PAIRS;
    ExpandTemp t = carve(p.temp, p.num_rendered);
    const unsigned nchunk_bound =
        (unsigned)(p.n_records / EXP_CHUNK + std::min<size_t>((size_t)nbins, p.n_records) + 1);  // <= max_chunks(R)
    int launches = 0;
    if (!p.bin_counts) {
        GSR_CUDA_TRY(launch_pdl(bin_bounds_kernel, dim3((nbins + 1 + EXP_WARPS - 1) / EXP_WARPS), dim3(EXP_THREADS), 0, s,
                                p.rec_bins, (uint32_t)p.n_records, nbins, t.bin_start));
        ++launches;
    }
    const unsigned table_ctas = std::min(64u, (nchunk_bound + TABLE_THREADS - 1) / TABLE_THREADS);
    GSR_CUDA_TRY(launch_pdl(chunk_table_kernel, dim3(table_ctas), dim3(TABLE_THREADS), 0, s, nbins, p.bin_counts, t.bin_start,
                            t.bin_chunk_first, t.chunk_desc, t.done_counter));
    ++launches;
    ExpandArgs a;
    a.rec_ids = p.rec_ids;
    a.chunk_desc = t.chunk_desc;
    a.num_chunks = t.bin_chunk_first + nbins;
    a.chunk_counts = t.chunk_counts;
    a.chunk_ballots = t.chunk_ballots;
    a.rec_data = t.rec_data;
    a.tile_rects = reinterpret_cast<const uint2*>(p.tile_rects);
    a.depths = p.depths;
    a.tile_start = p.tile_counts;
    a.keys_out = p.keys_out;
    a.vals_out = p.vals_out;
    a.grid_x = p.grid_x; a.grid_y = p.grid_y; a.bins_x = p.bins_x;
    const bool keys = p.keys_out != nullptr;
    GSR_CARVEOUT(expand_count_kernel<true>, "COUNT", -1);
    GSR_CARVEOUT(expand_fill_kernel<true>, "FILL", -1);
    GSR_CARVEOUT(expand_count_kernel<false>, "COUNT", -1);
    GSR_CARVEOUT(expand_fill_kernel<false>, "FILL", -1);
    GSR_CARVEOUT(expand_scan_kernel, "ESCAN", -1);
    if (ev) cudaEventRecord(ev[0], s);
    const dim3 cgrid((nchunk_bound + CNT_CHUNKS - 1) / CNT_CHUNKS);
    if (keys) GSR_CUDA_TRY(launch_pdl(expand_count_kernel<true>, cgrid, dim3(EXP_THREADS), 0, s, a));
    else GSR_CUDA_TRY(launch_pdl(expand_count_kernel<false>, cgrid, dim3(EXP_THREADS), 0, s, a));
    if (ev) cudaEventRecord(ev[1], s);
    ++launches;
    GSR_CUDA_TRY(launch_pdl(expand_scan_kernel, dim3(nbins), dim3(ESCAN_THREADS), 0, s, (const uint32_t*)t.bin_chunk_first,
                            t.chunk_counts, p.tile_counts, p.grid_x, p.grid_y, p.bins_x, nbins,
                            reinterpret_cast<uint2*>(p.ranges), p.r1_quirk ? 1 : 0, t.done_counter));
    ++launches;
    if (ev) cudaEventRecord(ev[2], s);
    if (keys) GSR_CUDA_TRY(launch_pdl(expand_fill_kernel<true>, dim3(nchunk_bound), dim3(EXP_THREADS), 0, s, a));
    else GSR_CUDA_TRY(launch_pdl(expand_fill_kernel<false>, dim3(nchunk_bound), dim3(EXP_THREADS), 0, s, a));
    if (ev) cudaEventRecord(ev[3], s);
    ++launches;
    return launches;
}

}  // namespace gsr
