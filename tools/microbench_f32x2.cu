// microbench_f32x2.cu — issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on B200, and of a
// mixed FFMA2 + MUFU.EX2 stream shaped like the blend inner loop.  Build: nvcc -arch=sm_100a.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int ITERS = 4096;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

__global__ void k_ffma(float* out, float s) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < ITERS; ++i) {
        a0 = fmaf(a0, s, 1.f); a1 = fmaf(a1, s, 1.f); a2 = fmaf(a2, s, 1.f); a3 = fmaf(a3, s, 1.f);
        a4 = fmaf(a4, s, 1.f); a5 = fmaf(a5, s, 1.f); a6 = fmaf(a6, s, 1.f); a7 = fmaf(a7, s, 1.f);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void k_ffma2(float* out, float s) {
    float2 a0 = make_float2(threadIdx.x, 1.f), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
    const float2 ss = make_float2(s, s), one = make_float2(1.f, 1.f);
    for (int i = 0; i < ITERS; ++i) {
        a0 = ffma2(a0, ss, one); a1 = ffma2(a1, ss, one); a2 = ffma2(a2, ss, one); a3 = ffma2(a3, ss, one);
        a4 = ffma2(a4, ss, one); a5 = ffma2(a5, ss, one); a6 = ffma2(a6, ss, one); a7 = ffma2(a7, ss, one);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a1.x + a2.x + a3.x + a4.y + a5.y + a6.y + a7.y;
}

template <typename F>
void run(const char* name, F launch, double flop_per_thread) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    const double threads = 148.0 * 8 * 1024;
    printf("%-28s %8.3f ms  %8.2f TFLOP/s  %8.2f T inst-lanes/s\n", name, ms, threads * flop_per_thread / ms / 1e9,
           threads * (flop_per_thread / (name[6] == '2' ? 4 : 2)) / ms / 1e9);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    run("k_ffma  (scalar)", [&] { k_ffma<<<148 * 8, 1024>>>(out, 0.999f); }, 2.0 * 8 * ITERS);
    run("k_ffma2 (packed f32x2)", [&] { k_ffma2<<<148 * 8, 1024>>>(out, 0.999f); }, 4.0 * 8 * ITERS);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
