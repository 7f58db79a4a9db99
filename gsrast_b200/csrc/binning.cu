// binning.cu — scan of tiles_touched, tile-key duplication and tile-range identification.
//
// Replaces (file:line in /root/reference/apps/gsrast/gscuda):
//   cub::DeviceScan::InclusiveSum + the 4-byte D2H readback   GSCuda.cu:771-772
//   duplicateWithKeys                                          GSCuda.cu:422-475 (launch :787)
//   cudaMemset(ranges) + identifyTileRanges                    GSCuda.cu:800-801, 504-538
//
// The scan is split in three: per-block sums (written by preprocess), a single-CTA scan of
// those sums (scan_block_sums_kernel, which also publishes num_rendered straight into mapped
// pinned host memory), and the intra-block scan, which is fused into the duplication kernel.
// Duplication is load-balanced: the 256 Gaussians of a block pool their tile counts and the
// block's threads walk the pooled output range item by item, so one huge splat does not
// serialise a thread and the key/value stores are fully coalesced.
#include "gsr_common.cuh"

namespace gsr {

namespace {

constexpr int SCAN_THREADS = 1024;

// Exclusive scan of block_sums[n] in place; total -> *total_dev and *total_host (mapped).
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(uint32_t* __restrict__ block_sums, int n,
                                                                       uint32_t* __restrict__ total_dev,
                                                                       volatile uint32_t* total_host) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += SCAN_THREADS) {
        const int i = base + tid;
        const uint32_t v = (i < n) ? block_sums[i] : 0u;
        uint32_t incl = v;
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
        const uint32_t excl = carry + s_warp[warp] + incl - v;
        if (i < n) block_sums[i] = excl;
        __syncthreads();
        if (tid == SCAN_THREADS - 1) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t total = s_carry;
        *total_dev = total;
        if (total_host) {
            *total_host = total;
            __threadfence_system();
        }
    }
}

__device__ __forceinline__ void get_rect_dev(float px, float py, int ex, int ey, int gx, int gy, int& minx, int& miny,
                                             int& maxx, int& maxy) {
    // x / 16.0f == x * 0.0625f bit for bit (exact power-of-two scaling)
    minx = min(gx, max(0, __float2int_rz(fmul(fsub(px, (float)ex), 1.0f / TILE_X))));
    miny = min(gy, max(0, __float2int_rz(fmul(fsub(py, (float)ey), 1.0f / TILE_Y))));
    maxx = min(gx, max(0, __float2int_rz(fmul(fsub(fadd(fadd(px, (float)ex), (float)TILE_X), 1.0f), 1.0f / TILE_X))));
    maxy = min(gy, max(0, __float2int_rz(fmul(fsub(fadd(fadd(py, (float)ey), (float)TILE_Y), 1.0f), 1.0f / TILE_Y))));
}

// One block = the same 256 Gaussians as in preprocess.
__global__ void __launch_bounds__(PRE_THREADS) duplicate_kernel(
    const int P, const int grid_x, const int grid_y, const float2* __restrict__ means2D,
    const float* __restrict__ depths, const uint32_t* __restrict__ tiles_touched,
    const uint32_t* __restrict__ block_offsets, const int* __restrict__ radii, const int2* __restrict__ rects,
    uint32_t* __restrict__ point_offsets, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t s_excl[PRE_THREADS + 1];
    __shared__ uint32_t s_warp[PRE_THREADS / 32];
    __shared__ uint32_t s_depth[PRE_THREADS];
    __shared__ uint32_t s_origin[PRE_THREADS];  // miny << 16 | minx
    __shared__ uint32_t s_width[PRE_THREADS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int base = blockIdx.x * PRE_THREADS;
    const int idx = base + tid;
    const bool valid = idx < P;

    // Culled Gaussians have tiles_touched == 0 (radii <= 0), so they drop out naturally.
    const uint32_t cnt = valid ? tiles_touched[idx] : 0u;
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < PRE_THREADS / 32; ++w)
        if (w < warp) woff += s_warp[w];
    incl += woff;
    const uint32_t boff = block_offsets[blockIdx.x];
    if (valid) point_offsets[idx] = boff + incl;  // inclusive scan, as GSCuda.cu:771 produces
    s_excl[tid] = incl - cnt;
    if (tid == PRE_THREADS - 1) s_excl[PRE_THREADS] = incl;

    if (cnt > 0) {
        const float2 m = means2D[idx];
        int minx, miny, maxx, maxy;
        if (rects == nullptr) {
            const int r = radii[idx];
            get_rect_dev(m.x, m.y, r, r, grid_x, grid_y, minx, miny, maxx, maxy);
        } else {
            const int2 e = rects[idx];
            get_rect_dev(m.x, m.y, e.x, e.y, grid_x, grid_y, minx, miny, maxx, maxy);
        }
        s_depth[tid] = __float_as_uint(depths[idx]);
        s_origin[tid] = ((uint32_t)miny << 16) | (uint32_t)minx;
        // GSCuda.cu:440-443: Gaussians with radii <= 0 emit nothing (their slots stay unwritten)
        s_width[tid] = (radii[idx] > 0) ? (uint32_t)(maxx - minx) : 0u;
    }
    __syncthreads();

    const uint32_t total = s_excl[PRE_THREADS];
    for (uint32_t k = tid; k < total; k += PRE_THREADS) {
        // largest g with s_excl[g] <= k
        int lo = 0, hi = PRE_THREADS - 1;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_excl[mid] <= k) lo = mid; else hi = mid - 1;
        }
        const int g = lo;
        const uint32_t t = k - s_excl[g];
        const uint32_t w = s_width[g];
        if (w == 0) continue;
        const uint32_t ty = t / w, tx = t - ty * w;
        const uint32_t org = s_origin[g];
        const uint32_t y = (org >> 16) + ty, x = (org & 0xffffu) + tx;
        // key = tile id << 32 | depth bits  (GSCuda.cu:466-471); rows outer, columns inner
        const uint64_t key = ((uint64_t)(y * (uint32_t)grid_x + x) << 32) | (uint64_t)s_depth[g];
        const size_t o = (size_t)boff + k;
        keys_out[o] = key;
        vals_out[o] = (uint32_t)(base + g);
    }
}

// Four consecutive sorted keys per thread (two 16-byte loads + the left neighbour): boundary
// detection in the tile-id half of the key.  ranges[] must be zeroed beforehand.
template <bool COMPAT>
__global__ void __launch_bounds__(256) identify_ranges_kernel(const size_t n, const uint64_t* __restrict__ keys,
                                                              uint2* __restrict__ ranges) {
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
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
                           cudaStream_t s) {
    scan_block_sums_kernel<<<1, SCAN_THREADS, 0, s>>>(block_sums, num_blocks, total_dev, total_host_mapped);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

int launch_duplicate(int P, int grid_x, int grid_y, const float* means2D, const float* depths,
                     const uint32_t* tiles_touched, const uint32_t* block_sums, const int* radii, const int* rects,
                     uint32_t* point_offsets, uint64_t* keys_out, uint32_t* vals_out, cudaStream_t s) {
    if (P <= 0) return 0;
    const int blocks = (P + PRE_THREADS - 1) / PRE_THREADS;
    duplicate_kernel<<<blocks, PRE_THREADS, 0, s>>>(P, grid_x, grid_y, reinterpret_cast<const float2*>(means2D), depths,
                                                    tiles_touched, block_sums, radii,
                                                    reinterpret_cast<const int2*>(rects), point_offsets, keys_out,
                                                    vals_out);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

int launch_identify_ranges(const uint64_t* keys, size_t n, uint32_t* ranges, int num_tiles, bool compat,
                           cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)num_tiles, s);
    if (e != cudaSuccess) return -(int)e;
    if (n == 0) return 0;
    const unsigned blocks = (unsigned)((n + 1023) / 1024);
    if (compat)
        identify_ranges_kernel<true><<<blocks, 256, 0, s>>>(n, keys, reinterpret_cast<uint2*>(ranges));
    else
        identify_ranges_kernel<false><<<blocks, 256, 0, s>>>(n, keys, reinterpret_cast<uint2*>(ranges));
    e = cudaPeekAtLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

}  // namespace gsr
