// microbench_f32x2.cu — issue cost of the packed FP32 instructions of sm_100 (fma/mul/add.rn.f32x2 -> FFMA2/FMUL2/FADD2)
// against their scalar forms, alone and mixed 1:1 with integer ALU work, on B200.  The blend's candidate trip is
// issue-bound (profiles/*_ncu_full_summary.txt: issue ~80 %, FMA pipe ~39 %), so what matters is warp-instructions
// issued per SM-cycle, not FLOP/s: a packed instruction pays off only if it costs ONE issue slot.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench_f32x2.cu -o microbench_f32x2
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int ITERS = 2048;
typedef unsigned long long u64;

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 pack(float x, float y) { return ((u64)__float_as_uint(y) << 32) | __float_as_uint(x); }

// MODE 0: 8 scalar FFMA   1: 8 FFMA2   2: 8 FMUL2   3: 8 FADD2   4: 8 scalar FFMA + 8 LOP3   5: 8 FFMA2 + 8 LOP3
// 6: 8 scalar FMUL   7: 4 FFMA2 + 4 MUFU.EX2 + 8 LOP3 (trip-like mix)
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float s, uint32_t m) {
    float a[8]; u64 p[8]; uint32_t q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; p[i] = pack(a[i], a[i] + 1.f); q[i] = threadIdx.x * 7 + i; }
    const u64 ss = pack(s, s), one = pack(1e-3f, 1e-3f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 4) a[i] = fmaf(a[i], s, 1e-3f);
            if (MODE == 6) a[i] = a[i] * s;
            if (MODE == 1 || MODE == 5) p[i] = ffma2(p[i], ss, one);
            if (MODE == 2) p[i] = fmul2(p[i], ss);
            if (MODE == 3) p[i] = fadd2(p[i], one);
            if (MODE == 4 || MODE == 5 || MODE == 7) q[i] = (q[i] ^ m) + (q[i] >> 3);
            if (MODE == 7) {
                if (i < 4) p[i] = ffma2(p[i], ss, one);
                else { float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a[i])); a[i] = e; }
            }
        }
    }
    float r = 0; uint32_t rq = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { r += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32)); rq += q[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + (float)rq;
}

template <int MODE>
void run(const char* name, float* out, double instr_per_iter, double flop_per_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k<MODE><<<grid, block>>>(out, 0.999f, 0x5bd1e995u); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, block>>>(out, 0.999f, 0x5bd1e995u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double warps = (double)grid * block / 32;
    const double winstr = warps * ITERS * instr_per_iter;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk_khz * 1e3;  // at the max clock; the device may run lower
    printf("%-44s %7.3f ms  %6.2f warp-instr/clk/SM (at max clock %d MHz)  %7.2f TFLOP/s\n", name, ms,
           winstr / cycles / 148.0, clk_khz / 1000, (double)grid * block * ITERS * flop_per_iter / ms / 1e9);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    run<0>("8 FFMA (scalar)", out, 8, 16);
    run<6>("8 FMUL (scalar)", out, 8, 8);
    run<1>("8 FFMA2 (fma.rn.f32x2)", out, 8, 32);
    run<2>("8 FMUL2 (mul.rn.f32x2)", out, 8, 16);
    run<3>("8 FADD2 (add.rn.f32x2)", out, 8, 16);
    run<4>("8 FFMA + 16 int ALU (xor/shift-add)", out, 8 + 16, 16);
    run<5>("8 FFMA2 + 16 int ALU", out, 8 + 16, 32);
    run<7>("4 FFMA2 + 4 MUFU.EX2 + 16 int ALU", out, 8 + 16, 16);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
