// microbench.cu — throughput of the warp-level primitives the radix sort can rank with
// (B200): shared atomics, match.any, 8-step ballot matching.  Build: nvcc -arch=sm_100a.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int ITERS = 4096;

__global__ void k_atoms_block(uint32_t* out, int mode) {
    __shared__ uint32_t bins[8 * 256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) bins[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
    uint32_t* b = (mode & 1) ? bins + warp * 256 : bins;  // private row per warp vs block-shared
    for (int i = 0; i < ITERS; ++i) {
        x = x * 1664525u + 1013904223u;
        uint32_t d = (mode & 2) ? ((x >> 24) & 3) : (x >> 24);  // skewed (4 bins) vs uniform (256 bins)
        atomicAdd(&b[d], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256) out[blockIdx.x * 256 + threadIdx.x] = bins[threadIdx.x] + lane;
}

__global__ void k_match(uint32_t* out, int mode) {
    __shared__ uint32_t bins[8 * 256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) bins[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
    uint32_t* b = bins + warp * 256;
    uint32_t acc = 0;
    for (int i = 0; i < ITERS; ++i) {
        x = x * 1664525u + 1013904223u;
        uint32_t d = (mode & 2) ? ((x >> 24) & 3) : (x >> 24);
        unsigned peers;
        if (mode & 4) {  // CUB-style 8-step ballot matching
            peers = 0xffffffffu;
#pragma unroll
            for (int bit = 0; bit < 8; ++bit) {
                const bool p = (d >> bit) & 1;
                const unsigned m = __ballot_sync(0xffffffffu, p);
                peers &= p ? m : ~m;
            }
        } else {
            peers = __match_any_sync(0xffffffffu, d);
        }
        if (mode & 8) {  // + the leader's counter update, as the sort does
            if (lane == __ffs(peers) - 1) b[d] += __popc(peers);
            __syncwarp();
        }
        acc += peers;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + b[lane];
}

int main() {
    uint32_t* out;
    cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = 148 * 4, threads = 256;
    const double lane_ops = (double)blocks * threads * ITERS;
    auto run = [&](const char* name, int which, int mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) k_atoms_block<<<blocks, threads>>>(out, mode);
            else k_match<<<blocks, threads>>>(out, mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %8.3f ms  %8.1f G lane-ops/s  (%.2f SM-cycles per warp-op @1.9GHz)\n", name, ms,
               lane_ops / ms / 1e6, ms * 1e-3 * 1.9e9 * 148 / (lane_ops / 32));
    };
    run("atoms block-shared uniform256", 0, 0);
    run("atoms warp-private uniform256", 0, 1);
    run("atoms block-shared skewed4", 0, 2);
    run("atoms warp-private skewed4", 0, 3);
    run("match.any uniform256", 1, 0);
    run("match.any skewed4", 1, 2);
    run("ballot8 uniform256", 1, 4);
    run("ballot8 skewed4", 1, 6);
    run("match.any + leader RMW uniform256", 1, 8);
    run("match.any + leader RMW skewed4", 1, 10);
    run("ballot8 + leader RMW uniform256", 1, 12);
    run("ballot8 + leader RMW skewed4", 1, 14);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
