// preprocess.cu — per-Gaussian stage of the forward rasterizer (sm_100a).
//
// Replaces preprocessCUDA / preprocess (apps/gsrast/gscuda/GSCuda.cu:261-415) together with
// quatToMat / computeCov3D (:157-195), computeCov2D (:197-231) and getRect (:237-259), and
// the per-block partial sums of the tiles_touched scan (cub::DeviceScan::InclusiveSum, :771).
//
// Two arithmetic modes, selected at compile time:
//   COMPAT=false  CudaRasterizer contract: near-plane cull on view z,
//                 un-normalised quaternion, focal_x/focal_y, ndc2Pix in double, view-space
//                 depth, SH degree 0..3 with clamping.
//   COMPAT=true   the in-tree gscuda semantics, restated operation by operation.
//
// HBM traffic: one pass over the attribute stream.  means/scales (float3 records) are
// fetched as coalesced float4 and re-sliced through shared memory; rotations are float4
// per thread; the 192-B SH block is only touched for Gaussians that survive culling, by
// 4 lanes per Gaussian x 3 float4 each (each lane owns 4 coefficients x RGB), after a
// warp-level compaction of the survivors.
//
// All arithmetic that feeds an integer output goes through the individually rounded
// intrinsics of gsr_common.cuh so the results are bit-identical to oracle/gsr_oracle.cpp.
#include "gsr_common.cuh"

namespace gsr {

namespace {

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// getRect (GSCuda.cu:237-259): tile-space AABB, truncation toward zero, clamped to the grid.
__device__ __forceinline__ void get_rect(float px, float py, int ex, int ey, int gx, int gy, int& minx, int& miny,
                                         int& maxx, int& maxy) {
    // x / 16.0f == x * 0.0625f bit for bit (exact power-of-two scaling)
    minx = min(gx, max(0, __float2int_rz(fmul(fsub(px, (float)ex), 1.0f / TILE_X))));
    miny = min(gy, max(0, __float2int_rz(fmul(fsub(py, (float)ey), 1.0f / TILE_Y))));
    maxx = min(gx, max(0, __float2int_rz(fmul(fsub(fadd(fadd(px, (float)ex), (float)TILE_X), 1.0f), 1.0f / TILE_X))));
    maxy = min(gy, max(0, __float2int_rz(fmul(fsub(fadd(fadd(py, (float)ey), (float)TILE_Y), 1.0f), 1.0f / TILE_Y))));
}

// upstream ndc2Pix: ((v + 1.0) * S - 1.0) * 0.5 evaluated in double, rounded to float once.
__device__ __forceinline__ float ndc2pix(float v, int S) {
    double t = __dadd_rn((double)v, 1.0);
    t = __dmul_rn(t, (double)S);
    t = __dadd_rn(t, -1.0);
    t = __dmul_rn(t, 0.5);
    return __double2float_rn(t);
}

__device__ __forceinline__ float glm_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float glm_max(float a, float b) { return (a < b) ? b : a; }

// Sigma^T-projected 2D covariance; shared tail of both modes.
// T0[r] = T.c[0][r], T1[r] = T.c[1][r]; c[] = packed symmetric cov3D.
__device__ __forceinline__ void project_cov(const float T0[3], const float T1[3], const float c[6], float& cov00,
                                            float& cov01, float& cov11) {
    float A00 = dot3(T0[0], c[0], T0[1], c[1], T0[2], c[2]);
    float A10 = dot3(T0[0], c[1], T0[1], c[3], T0[2], c[4]);
    float A20 = dot3(T0[0], c[2], T0[1], c[4], T0[2], c[5]);
    float A01 = dot3(T1[0], c[0], T1[1], c[1], T1[2], c[2]);
    float A11 = dot3(T1[0], c[1], T1[1], c[3], T1[2], c[4]);
    float A21 = dot3(T1[0], c[2], T1[1], c[4], T1[2], c[5]);
    cov00 = fadd(dot3(A00, T0[0], A10, T0[1], A20, T0[2]), 0.3f);
    cov01 = dot3(A01, T0[0], A11, T0[1], A21, T0[2]);
    cov11 = fadd(dot3(A01, T1[0], A11, T1[1], A21, T1[2]), 0.3f);
}

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
constexpr float SH_C2_0 = 1.0925484305920792f, SH_C2_1 = -1.0925484305920792f, SH_C2_2 = 0.31539156525252005f,
                SH_C2_3 = -1.0925484305920792f, SH_C2_4 = 0.5462742152960396f;
constexpr float SH_C3_0 = -0.5900435899266435f, SH_C3_1 = 2.890611442640554f, SH_C3_2 = -0.4570457994644658f,
                SH_C3_3 = 0.3731763325901154f, SH_C3_4 = -0.4570457994644658f, SH_C3_5 = 1.445305721320277f,
                SH_C3_6 = -0.5900435899266435f;

#ifndef GSR_PRE_MINB
#define GSR_PRE_MINB 6
#endif
template <bool COMPAT>
__global__ void __launch_bounds__(PRE_THREADS, GSR_PRE_MINB) preprocess_kernel(const PreprocessParams p, const int vec_means,
                                                                 const int vec_scales, const int vec_sh) {
    __shared__ __align__(16) float s_means[PRE_THREADS * 3];
    __shared__ __align__(16) float s_scales[PRE_THREADS * 3];
    __shared__ float s_cam[36];                       // view[16] proj[16] cam_pos[3]
    __shared__ int s_queue[PRE_THREADS / 32][32];     // per-warp survivors (Gaussian index), compacted
    __shared__ float4 s_basis[PRE_THREADS / 32][32 * 4];  // their 16 SH basis factors (contract mode only)
    __shared__ uint32_t s_wsum[PRE_THREADS / 32];
    __shared__ uint32_t s_wsum2[PRE_THREADS / 32];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int base = blockIdx.x * PRE_THREADS;
    const int idx = base + tid;
    const int nvalid = min(PRE_THREADS, p.P - base);
    const bool valid = tid < nvalid;
    gsr_pdl_wait();
    gsr_pdl_launch_dependents();  // the scan behind this kernel may take freed SM slots early (it waits)

    // ---- stage camera + float3 streams -------------------------------------------------
    if (tid < 16) s_cam[tid] = __ldg(p.viewmatrix + tid);
    else if (tid < 32) s_cam[tid] = __ldg(p.projmatrix + (tid - 16));
    else if (tid < 35 && p.cam_pos) s_cam[tid] = __ldg(p.cam_pos + (tid - 32));

    if (p.means_stride == 3) {
        const float* g = p.means3D + (size_t)base * 3;
        const int nf = nvalid * 3;
        if (vec_means) {
            const int nv = nf >> 2;
            if (tid < nv) reinterpret_cast<float4*>(s_means)[tid] = ldg_f4(g + tid * 4);
            for (int i = (nv << 2) + tid; i < nf; i += PRE_THREADS) s_means[i] = __ldg(g + i);
        } else {
            for (int i = tid; i < nf; i += PRE_THREADS) s_means[i] = __ldg(g + i);
        }
    }
    const bool need_scales = (p.cov3D_precomp == nullptr);
    // The rotation and the opacity of every Gaussian are requested HERE, with the mean / scale streams, although they
    // are only used by Gaussians that survive the frustum test (rotation) or emit pairs (opacity): issued behind those
    // tests they were two more dependent DRAM round trips per CTA (ncu: 55 % long-scoreboard stalls); up front the
    // four coalesced streams fly together, at the price of 20 B of reads per culled Gaussian.
#ifndef GSR_PRE_HOIST
#define GSR_PRE_HOIST 1
#endif
    float4 q_early = make_float4(0.f, 0.f, 0.f, 0.f);
    float opac_early = 0.f;
    if (GSR_PRE_HOIST && valid) {
        if (need_scales) q_early = ldg_f4(p.rotations + (size_t)idx * 4);
        opac_early = __ldg(p.opacities + idx);
    }
    if (need_scales && p.scales_stride == 3) {
        const float* g = p.scales + (size_t)base * 3;
        const int nf = nvalid * 3;
        if (vec_scales) {
            const int nv = nf >> 2;
            if (tid < nv) reinterpret_cast<float4*>(s_scales)[tid] = ldg_f4(g + tid * 4);
            for (int i = (nv << 2) + tid; i < nf; i += PRE_THREADS) s_scales[i] = __ldg(g + i);
        } else {
            for (int i = tid; i < nf; i += PRE_THREADS) s_scales[i] = __ldg(g + i);
        }
    }
    __syncthreads();

    const float* v = s_cam;
    const float* pm = s_cam + 16;

    uint32_t tiles = 0;
    uint2 rec = make_uint2(0u, 0u);  // tile rect for the duplication kernel: miny<<16|minx, height<<16|width
    bool need_sh = false;
    float dirx = 0.f, diry = 0.f, dirz = 0.f;
    uint32_t pf_word = 0;  // prefetch mode 3: the word the 64-byte pull returned (one SH coefficient, used below)
    bool pf_low = false, pf_have = false;

    if (valid) {
        int radius_out = 0;
        float px, py, pz, pw = 1.0f;
        if (p.means_stride == 3) {
            px = s_means[3 * tid]; py = s_means[3 * tid + 1]; pz = s_means[3 * tid + 2];
        } else {
            float4 m = ldg_f4(p.means3D + (size_t)idx * 4);
            px = m.x; py = m.y; pz = m.z; pw = m.w;
        }

        bool alive = true;
        float tvx, tvy, tvz;  // view-space point used by the EWA projection
        float prx, pry;       // NDC x, y
        float depth;

        if (!COMPAT) {
            // in_frustum: view-space z only (upstream auxiliary.h); transformPoint4x3 left to right
            tvx = fadd(dot3(v[0], px, v[4], py, v[8], pz), v[12]);
            tvy = fadd(dot3(v[1], px, v[5], py, v[9], pz), v[13]);
            tvz = fadd(dot3(v[2], px, v[6], py, v[10], pz), v[14]);
            if (tvz <= 0.2f) alive = false;
            float hx = fadd(dot3(pm[0], px, pm[4], py, pm[8], pz), pm[12]);
            float hy = fadd(dot3(pm[1], px, pm[5], py, pm[9], pz), pm[13]);
            float hw = fadd(dot3(pm[3], px, pm[7], py, pm[11], pz), pm[15]);
            float p_w = frcp(fadd(hw, 0.0000001f));
            prx = fmul(hx, p_w);
            pry = fmul(hy, p_w);
            depth = tvz;
            // SIBR bounding-box cull
            if (px < p.boxmin[0] || py < p.boxmin[1] || pz < p.boxmin[2] || px > p.boxmax[0] || py > p.boxmax[1] ||
                pz > p.boxmax[2])
                alive = false;
        } else {
            // GSCuda.cu:302-309: glm mat4*vec4 = (m0*x + m1*y) + (m2*z + m3*w)
            float hx = fadd(fadd(fmul(pm[0], px), fmul(pm[4], py)), fadd(fmul(pm[8], pz), fmul(pm[12], pw)));
            float hy = fadd(fadd(fmul(pm[1], px), fmul(pm[5], py)), fadd(fmul(pm[9], pz), fmul(pm[13], pw)));
            float hz = fadd(fadd(fmul(pm[2], px), fmul(pm[6], py)), fadd(fmul(pm[10], pz), fmul(pm[14], pw)));
            float hw = fadd(fadd(fmul(pm[3], px), fmul(pm[7], py)), fadd(fmul(pm[11], pz), fmul(pm[15], pw)));
            float oow = frcp(fadd(0.001f, hw));
            prx = fmul(oow, hx);
            pry = fmul(oow, hy);
            float prz = fmul(oow, hz);
            if (prz < 0.0f || prz > 1.0f || prx < -1.3f || prx > 1.3f || pry < -1.3f || pry > 1.3f) alive = false;
            depth = prz;
            // computeCov2D's own transform: view * vec4(mean, 1.0f)   (GSCuda.cu:202)
            tvx = fadd(fadd(fmul(v[0], px), fmul(v[4], py)), fadd(fmul(v[8], pz), fmul(v[12], 1.0f)));
            tvy = fadd(fadd(fmul(v[1], px), fmul(v[5], py)), fadd(fmul(v[9], pz), fmul(v[13], 1.0f)));
            tvz = fadd(fadd(fmul(v[2], px), fmul(v[6], py)), fadd(fmul(v[10], pz), fmul(v[14], 1.0f)));
        }

        // The SH block of a Gaussian that passed the frustum test is needed ~200 instructions from now (after
        // the covariance math and the warp compaction): start pulling it into L2.  Only for centres on the screen
        // (GSR_PRE_PREFETCH_BOUND in NDC; a hint: an off-screen centre whose radius reaches the screen just misses) —
        // at 1.25 the Gaussians of the band around the screen, which are culled a moment later, cost ~35 MB of SH reads
        // per C2 frame (preprocess 0.170 -> 0.167 ms at 1.03, profiles/r02_ab.txt).  GSR_PRE_PREFETCH_MODE 0: both
        // 128-byte lines the 192-byte block touches; 1: only the line it owns entirely (the other 64 bytes share
        // their line with a neighbour that may be culled) — measured SLOWER, 0.177 ms: the demand loads of the
        // un-prefetched third then wait on DRAM; 2: no prefetch; 3 (needs GSR_PRE_SH_LANE=1): the owned line is
        // prefetched, the shared 64 bytes are pulled by a 4-byte load with a 64-byte L2 fetch hint — no gain
        // (0.177 vs 0.174 ms, profiles/r02l_ab_C2.txt).
#ifndef GSR_PRE_PREFETCH_BOUND
#define GSR_PRE_PREFETCH_BOUND 1.03f
#endif
#ifndef GSR_PRE_PREFETCH_MODE
#define GSR_PRE_PREFETCH_MODE 0
#endif
        if (!COMPAT && GSR_PRE_PREFETCH_MODE != 2 && alive && p.colors_precomp == nullptr &&
            fabsf(prx) < GSR_PRE_PREFETCH_BOUND && fabsf(pry) < GSR_PRE_PREFETCH_BOUND) {
            const char* shp = reinterpret_cast<const char*>(p.shs + (size_t)idx * p.M * 3);
            // (cp.async.bulk.prefetch.L2 of exactly the block's 192 bytes was tried instead of line prefetches:
            // preprocess 0.169 -> 0.190 ms, profiles/r01f_ab.txt)
            if (GSR_PRE_PREFETCH_MODE == 3 && p.M == 16) {
                // mode 3: the 128-byte line the block owns entirely is prefetched; the 64 bytes it shares a line with a
                // neighbour (who may be culled) are pulled by a 4-byte load with a 64-byte L2 fetch hint whose result
                // nobody reads — no line is fetched on behalf of a Gaussian that may not need it
                const uintptr_t a = reinterpret_cast<uintptr_t>(shp);
                const bool low = (a & 127u) == 0;  // block starts a line: owns [0,128) and the first half of the next
                asm volatile("prefetch.global.L2 [%0];" ::"l"(low ? shp : shp + 64));
                // (ptxas deletes a load nobody consumes: the word is coefficient 0 / 10.2 of this Gaussian and replaces the
                // copy the SH evaluation loads again)
                asm volatile("ld.global.nc.L2::64B.b32 %0, [%1];" : "=r"(pf_word) : "l"(low ? shp + 128 : shp));
                pf_low = low;
                pf_have = true;
            } else if (GSR_PRE_PREFETCH_MODE == 0 || p.M * 12 <= 128) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(shp));
                if (p.M * 12 > 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(shp + 128));
            } else {
                // the block starts on a 64-byte boundary (M = 16: 192 * idx): its first 128 bytes are a whole line when
                // the start is 128-aligned, otherwise its last 128 bytes are
                const uintptr_t a = reinterpret_cast<uintptr_t>(shp);
                asm volatile("prefetch.global.L2 [%0];" ::"l"((a & 127u) ? shp + (size_t)p.M * 12 - 128 : shp));
            }
        }

        float cov00 = 0.f, cov01 = 0.f, cov11 = 0.f;
        float o_depth = 0.f;
        float2 o_xy = make_float2(0.f, 0.f);
        float4 o_conic = make_float4(0.f, 0.f, 0.f, 0.f);
        if (alive) {
            // ---- 3D covariance ---------------------------------------------------------
            float c[6];
            if (p.cov3D_precomp) {
                const float* cp = p.cov3D_precomp + (size_t)idx * 6;
#pragma unroll
                for (int i = 0; i < 6; ++i) c[i] = __ldg(cp + i);
            } else {
                float sx, sy, sz;
                if (p.scales_stride == 3) {
                    sx = s_scales[3 * tid]; sy = s_scales[3 * tid + 1]; sz = s_scales[3 * tid + 2];
                } else {
                    float4 s4 = ldg_f4(p.scales + (size_t)idx * 4);
                    sx = s4.x; sy = s4.y; sz = s4.z;
                }
                const float4 q = GSR_PRE_HOIST ? q_early : ldg_f4(p.rotations + (size_t)idx * 4);
                const float s0 = fmul(p.scale_modifier, sx), s1 = fmul(p.scale_modifier, sy),
                            s2 = fmul(p.scale_modifier, sz);
                if (!COMPAT) {
                    const float r = q.x, x = q.y, y = q.z, z = q.w;
                    // R.c[col][row] exactly as the glm::mat3 constructor receives it
                    float R00 = fsub(1.f, fmul(2.f, fadd(fmul(y, y), fmul(z, z))));
                    float R01 = fmul(2.f, fsub(fmul(x, y), fmul(r, z)));
                    float R02 = fmul(2.f, fadd(fmul(x, z), fmul(r, y)));
                    float R10 = fmul(2.f, fadd(fmul(x, y), fmul(r, z)));
                    float R11 = fsub(1.f, fmul(2.f, fadd(fmul(x, x), fmul(z, z))));
                    float R12 = fmul(2.f, fsub(fmul(y, z), fmul(r, x)));
                    float R20 = fmul(2.f, fsub(fmul(x, z), fmul(r, y)));
                    float R21 = fmul(2.f, fadd(fmul(y, z), fmul(r, x)));
                    float R22 = fsub(1.f, fmul(2.f, fadd(fmul(x, x), fmul(y, y))));
                    // M = S * R  ->  M.c[c][k] = s_k * R.c[c][k]
                    float M00 = fmul(s0, R00), M01 = fmul(s1, R01), M02 = fmul(s2, R02);
                    float M10 = fmul(s0, R10), M11 = fmul(s1, R11), M12 = fmul(s2, R12);
                    float M20 = fmul(s0, R20), M21 = fmul(s1, R21), M22 = fmul(s2, R22);
                    // Sigma = M^T * M  ->  Sigma.c[c][r] = sum_k M.c[r][k] * M.c[c][k]
                    c[0] = dot3(M00, M00, M01, M01, M02, M02);
                    c[1] = dot3(M10, M00, M11, M01, M12, M02);
                    c[2] = dot3(M20, M00, M21, M01, M22, M02);
                    c[3] = dot3(M10, M10, M11, M11, M12, M12);
                    c[4] = dot3(M20, M10, M21, M11, M22, M12);
                    c[5] = dot3(M20, M20, M21, M21, M22, M22);
                } else {
                    // glm::normalize(vec4): v * (1/sqrt(dot)), dot pairwise   (GSCuda.cu:178)
                    float d = fadd(fadd(fmul(q.x, q.x), fmul(q.y, q.y)), fadd(fmul(q.z, q.z), fmul(q.w, q.w)));
                    float inv = frcp(fsqrt(d));
                    float qx = fmul(q.x, inv), qy = fmul(q.y, inv), qz = fmul(q.z, inv), qw = fmul(q.w, inv);
                    // quatToMat (GSCuda.cu:157-162): float inner sums, double outer 2.0* / -1.0
#define GSR_D1(sum) __double2float_rn(__dadd_rn(__dmul_rn(2.0, (double)(sum)), -1.0))
#define GSR_D2(val) __double2float_rn(__dmul_rn(2.0, (double)(val)))
                    float R00 = GSR_D1(fadd(fmul(qx, qx), fmul(qy, qy)));
                    float R01 = GSR_D2(fadd(fmul(qy, qz), fmul(qx, qw)));
                    float R02 = GSR_D2(fsub(fmul(qy, qw), fmul(qx, qz)));
                    float R10 = GSR_D2(fsub(fmul(qy, qz), fmul(qx, qw)));
                    float R11 = GSR_D1(fadd(fmul(qx, qx), fmul(qz, qz)));
                    float R12 = GSR_D2(fadd(fmul(qz, qw), fmul(qx, qy)));
                    float R20 = GSR_D2(fadd(fmul(qy, qw), fmul(qx, qz)));
                    float R21 = GSR_D2(fsub(fmul(qz, qw), fmul(qx, qy)));
                    float R22 = GSR_D1(fadd(fmul(qx, qx), fmul(qw, qw)));
#undef GSR_D1
#undef GSR_D2
                    // rs = R * S -> rs.c[c][r] = R.c[c][r] * s_c ; sigma = rs * rs^T ->
                    // sigma.c[c][r] = sum_k rs.c[k][r] * rs.c[k][c]
                    float a00 = fmul(R00, s0), a01 = fmul(R01, s0), a02 = fmul(R02, s0);
                    float a10 = fmul(R10, s1), a11 = fmul(R11, s1), a12 = fmul(R12, s1);
                    float a20 = fmul(R20, s2), a21 = fmul(R21, s2), a22 = fmul(R22, s2);
                    c[0] = dot3(a00, a00, a10, a10, a20, a20);  // sigma[0][0]
                    c[1] = dot3(a00, a01, a10, a11, a20, a21);  // sigma[1][0]: c=1, r=0
                    c[2] = dot3(a00, a02, a10, a12, a20, a22);  // sigma[2][0]
                    c[3] = dot3(a01, a01, a11, a11, a21, a21);  // sigma[1][1]
                    c[4] = dot3(a01, a02, a11, a12, a21, a22);  // sigma[2][1]: c=2, r=1
                    c[5] = dot3(a02, a02, a12, a12, a22, a22);  // sigma[2][2]
                }
                if (p.cov3D) {  // GSR_FLAG_LEAN_STATE drops it: nothing in the forward pass reads it back
                    float* co = p.cov3D + (size_t)idx * 6;
                    reinterpret_cast<float2*>(co)[0] = make_float2(c[0], c[1]);
                    reinterpret_cast<float2*>(co)[1] = make_float2(c[2], c[3]);
                    reinterpret_cast<float2*>(co)[2] = make_float2(c[4], c[5]);
                }
            }

            // ---- EWA projection (computeCov2D) -----------------------------------------
            const float limx = fmul(1.3f, p.tan_fovx), limy = fmul(1.3f, p.tan_fovy);
            const float txtz = fdiv(tvx, tvz), tytz = fdiv(tvy, tvz);
            float tx, ty;
            if (!COMPAT) {
                tx = fmul(fminf(limx, fmaxf(-limx, txtz)), tvz);
                ty = fmul(fminf(limy, fmaxf(-limy, tytz)), tvz);
            } else {
                tx = fmul(glm_min(limx, glm_max(-limx, txtz)), tvz);
                ty = fmul(glm_min(limy, glm_max(-limy, tytz)), tvz);
            }
            const float fx = COMPAT ? p.focal_y : p.focal_x;  // in-tree: single focalDist (GSCuda.cu:721)
            const float fy = p.focal_y;
            const float tz2 = fmul(tvz, tvz);
            const float j00 = fdiv(fx, tvz), j02 = fdiv(-fmul(fx, tx), tz2);
            const float j11 = fdiv(fy, tvz), j12 = fdiv(-fmul(fy, ty), tz2);
            float T0[3], T1[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                T0[r] = fadd(fmul(v[4 * r], j00), fmul(v[4 * r + 2], j02));
                T1[r] = fadd(fmul(v[4 * r + 1], j11), fmul(v[4 * r + 2], j12));
            }
            project_cov(T0, T1, c, cov00, cov01, cov11);
        }

        float det = fsub(fmul(cov00, cov11), fmul(cov01, cov01));
        if (alive && det == 0.0f) alive = false;

        if (alive) {
            const float det_inv = frcp(det);
            const float conx = fmul(cov11, det_inv), cony = fmul(-cov01, det_inv), conz = fmul(cov00, det_inv);
            const float mid = fmul(0.5f, fadd(cov00, cov11));
            float disc;
            if (!COMPAT) disc = fsqrt(fmaxf(0.1f, fsub(fmul(mid, mid), det)));
            else disc = fsqrt(glm_max(0.1f, fsub(fmul(mid, mid), det)));
            const float l1 = fadd(mid, disc), l2 = fsub(mid, disc);
            const float my_radius = ceilf(fmul(3.f, fsqrt(COMPAT ? glm_max(l1, l2) : fmaxf(l1, l2))));
            float ix, iy;
            if (!COMPAT) {
                ix = ndc2pix(prx, p.W);
                iy = ndc2pix(pry, p.H);
            } else {
                ix = fmul(fadd(fmul(prx, 0.5f), 0.5f), (float)p.W);  // GSCuda.cu:342
                iy = fmul(fadd(fmul(pry, 0.5f), 0.5f), (float)p.H);
            }
            int minx, miny, maxx, maxy;
            const int ri = __float2int_rz(my_radius);
            if (p.rects == nullptr) {
                get_rect(ix, iy, ri, ri, p.grid_x, p.grid_y, minx, miny, maxx, maxy);
            } else {
                const int rx = __float2int_rz(ceilf(fmul(3.f, fsqrt(cov00))));
                const int ry = COMPAT ? __float2int_rz(ceilf(fmul(3.0f, cov11)))  // GSCuda.cu:352 (no sqrt)
                                      : __float2int_rz(ceilf(fmul(3.f, fsqrt(cov11))));
                reinterpret_cast<int2*>(p.rects)[idx] = make_int2(rx, ry);
                get_rect(ix, iy, rx, ry, p.grid_x, p.grid_y, minx, miny, maxx, maxy);
            }
            const uint32_t area = (uint32_t)(maxx - minx) * (uint32_t)(maxy - miny);
            if (area != 0) {
                tiles = area;
                rec = make_uint2(((uint32_t)miny << 16) | (uint32_t)minx,
                                 ((uint32_t)(maxy - miny) << 16) | (uint32_t)(maxx - minx));
                radius_out = ri;
                o_depth = depth;
                o_xy = make_float2(ix, iy);
                o_conic = make_float4(conx, cony, conz, GSR_PRE_HOIST ? opac_early : __ldg(p.opacities + idx));
                if (p.colors_precomp == nullptr) {
                    if (!COMPAT) {
                        need_sh = true;
                        float dx = fsub(px, s_cam[32]), dy = fsub(py, s_cam[33]), dz = fsub(pz, s_cam[34]);
                        float len = fsqrt(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
                        dirx = fdiv(dx, len); diry = fdiv(dy, len); dirz = fdiv(dz, len);
                    } else {
                        const float* sh = p.shs + (size_t)idx * 48;  // GSCuda.cu:364-365
                        float* o = p.rgb + (size_t)idx * 3;
                        o[0] = fadd(0.5f, fmul(0.4f, __ldg(sh + 0)));
                        o[1] = fadd(0.5f, fmul(0.4f, __ldg(sh + 1)));
                        o[2] = fadd(0.5f, fmul(0.4f, __ldg(sh + 2)));
                    }
                }
            }
        }
        // depth / mean / conic are stored for EVERY Gaussian (zeros where nothing is emitted; the reference leaves
        // those entries stale): whole sectors are written, so the memory system never has to read a sector back to
        // merge a few surviving 4..16-byte records into it (measured: ~130 MB of DRAM reads per C2 frame)
        // depths doubles as the low half of the sort key (GSCuda.cu:466-471): Gaussians that emit nothing carry the
        // all-ones pattern there, which the depth sort drops (the reference leaves their entry stale)
        p.depths[idx] = tiles ? o_depth : __uint_as_float(0xffffffffu);
        reinterpret_cast<float2*>(p.means2D)[idx] = o_xy;
        reinterpret_cast<float4*>(p.conic_opacity)[idx] = o_conic;
        if (tiles == 0 && p.colors_precomp == nullptr) {
            float* o = p.rgb + (size_t)idx * 3;
            o[0] = 0.f; o[1] = 0.f; o[2] = 0.f;
            if (!COMPAT && p.clamped) {
                unsigned char* cl = p.clamped + (size_t)idx * 3;
                cl[0] = 0; cl[1] = 0; cl[2] = 0;
            }
        }
        if (p.radii) p.radii[idx] = radius_out;  // lean callers without a radii buffer: nothing reads internal_radii
        if (p.tiles_touched) p.tiles_touched[idx] = tiles;
        reinterpret_cast<uint2*>(p.tile_rects)[idx] = rec;
    }

    // ---- SH colour --------------------------------------------------------------------------------------
    // GSR_PRE_SH_LANE=1 (A/B, off): every lane evaluates the colour of ITS OWN Gaussian — 12 float4 loads of its 192-byte
    // block and 48 FMAs, basis factors formed in registers — ~170 warp instructions for 32 Gaussians where the compacting
    // 4-lane form below issues ~470 (ncu source page, r02h).  Built, bit-identical, measured SLOWER (profiles/r02l_ab_*.txt:
    // preprocess 0.164 -> 0.174 ms at C2, 0.283 -> 0.295 at C3): a warp-wide LDG.128 then touches 32 different lines where
    // the 4-lane form touches 16, and the kernel is bound by the L1 / LSU request rate and DRAM (84 % of the measured copy
    // bandwidth on the bytes it moves), not by the issue slots the shorter form saves.  Same summation tree:
    //   s_q = fma(b[4q+3],c[4q+3], fma(b[4q+2],c[4q+2], fma(b[4q+1],c[4q+1], b[4q]*c[4q])));  rgb = ((s0+s1)+(s2+s3))+0.5
#ifndef GSR_PRE_SH_LANE
#define GSR_PRE_SH_LANE 0
#endif
    if (!COMPAT && GSR_PRE_SH_LANE) {
        if (need_sh) {
            const int ncoef = min(p.M, (p.D + 1) * (p.D + 1));
            const float* sp = p.shs + (size_t)idx * p.M * 3;
            const float x = dirx, y = diry, z = dirz;
            const float xx = fmul(x, x), yy = fmul(y, y), zz = fmul(z, z);
            const float xy = fmul(x, y), yz = fmul(y, z), xz = fmul(x, z);
            float part[4][3];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int nk = max(0, min(4, ncoef - 4 * q));  // active coefficients of this group: 4q .. 4q+nk-1
                float s[12];
                if (nk == 4 && vec_sh) {
                    const float4 a = ldg_f4(sp + 12 * q), b = ldg_f4(sp + 12 * q + 4), cc = ldg_f4(sp + 12 * q + 8);
                    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
                    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
                    s[8] = cc.x; s[9] = cc.y; s[10] = cc.z; s[11] = cc.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 12; ++i) s[i] = (i < nk * 3) ? __ldg(sp + 12 * q + i) : 0.f;
                }
                if (GSR_PRE_PREFETCH_MODE == 3 && pf_have) {  // same bits as the word just loaded (see the prefetch above)
                    if (q == 0 && !pf_low) s[0] = __uint_as_float(pf_word);
                    if (q == 2 && pf_low) s[8] = __uint_as_float(pf_word);
                }
                float b0, b1, b2, b3;
                if (q == 0) {
                    b0 = SH_C0;
                    b1 = -fmul(SH_C1, y);
                    b2 = fmul(SH_C1, z);
                    b3 = -fmul(SH_C1, x);
                } else if (q == 1) {
                    b0 = fmul(SH_C2_0, xy);
                    b1 = fmul(SH_C2_1, yz);
                    b2 = fmul(SH_C2_2, fsub(fsub(fmul(2.0f, zz), xx), yy));
                    b3 = fmul(SH_C2_3, xz);
                } else if (q == 2) {
                    b0 = fmul(SH_C2_4, fsub(xx, yy));
                    b1 = fmul(fmul(SH_C3_0, y), fsub(fmul(3.0f, xx), yy));
                    b2 = fmul(fmul(SH_C3_1, xy), z);
                    b3 = fmul(fmul(SH_C3_2, y), fsub(fsub(fmul(4.0f, zz), xx), yy));
                } else {
                    b0 = fmul(fmul(SH_C3_3, z), fsub(fsub(fmul(2.0f, zz), fmul(3.0f, xx)), fmul(3.0f, yy)));
                    b1 = fmul(fmul(SH_C3_4, x), fsub(fsub(fmul(4.0f, zz), xx), yy));
                    b2 = fmul(fmul(SH_C3_5, z), fsub(xx, yy));
                    b3 = fmul(fmul(SH_C3_6, x), fsub(xx, fmul(3.0f, yy)));
                }
                if (nk < 1) b0 = 0.f;
                if (nk < 2) b1 = 0.f;
                if (nk < 3) b2 = 0.f;
                if (nk < 4) b3 = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    part[q][c] = __fmaf_rn(b3, s[9 + c], __fmaf_rn(b2, s[6 + c], __fmaf_rn(b1, s[3 + c], fmul(b0, s[c]))));
            }
            float* o = p.rgb + (size_t)idx * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = fadd(fadd(fadd(part[0][c], part[1][c]), fadd(part[2][c], part[3][c])), 0.5f);
                if (p.clamped) p.clamped[(size_t)idx * 3 + c] = v < 0.f;
                o[c] = fmaxf(v, 0.f);
            }
        }
    }

    // ---- SH colour (default): 4 lanes per surviving Gaussian -----------------------------------------------
    // Phase 1 (owner thread, all lanes busy, no divergence): the 16 basis factors of the view direction go to
    // shared memory.  Phase 2: lane (slot, q) owns coefficients 4q..4q+3 of survivor `slot` (3 coalesced float4
    // = 48 contiguous bytes, 192 B per Gaussian across the 4 lanes), chains them with FMAs, and a two-step
    // butterfly adds the four partial sums: rgb_c = ((s0 + s1) + (s2 + s3)) + 0.5 with
    // s_q = fma(b3,c3, fma(b2,c2, fma(b1,c1, b0*c0))).  oracle/gsr_oracle.cpp evaluates the same tree.
    if (!COMPAT && !GSR_PRE_SH_LANE) {
        const unsigned mask = __ballot_sync(0xffffffffu, need_sh);
        if (mask) {
            const int ncoef = min(p.M, (p.D + 1) * (p.D + 1));
            float4* bp = s_basis[warp];
            if (need_sh) {
                const int pos = __popc(mask & ((1u << lane) - 1u));
                s_queue[warp][pos] = idx;
                const float x = dirx, y = diry, z = dirz;
                const float xx = fmul(x, x), yy = fmul(y, y), zz = fmul(z, z);
                const float xy = fmul(x, y), yz = fmul(y, z), xz = fmul(x, z);
                float bf[16];
                bf[0] = SH_C0;
                bf[1] = -fmul(SH_C1, y);
                bf[2] = fmul(SH_C1, z);
                bf[3] = -fmul(SH_C1, x);
                bf[4] = fmul(SH_C2_0, xy);
                bf[5] = fmul(SH_C2_1, yz);
                bf[6] = fmul(SH_C2_2, fsub(fsub(fmul(2.0f, zz), xx), yy));
                bf[7] = fmul(SH_C2_3, xz);
                bf[8] = fmul(SH_C2_4, fsub(xx, yy));
                bf[9] = fmul(fmul(SH_C3_0, y), fsub(fmul(3.0f, xx), yy));
                bf[10] = fmul(fmul(SH_C3_1, xy), z);
                bf[11] = fmul(fmul(SH_C3_2, y), fsub(fsub(fmul(4.0f, zz), xx), yy));
                bf[12] = fmul(fmul(SH_C3_3, z), fsub(fsub(fmul(2.0f, zz), fmul(3.0f, xx)), fmul(3.0f, yy)));
                bf[13] = fmul(fmul(SH_C3_4, x), fsub(fsub(fmul(4.0f, zz), xx), yy));
                bf[14] = fmul(fmul(SH_C3_5, z), fsub(xx, yy));
                bf[15] = fmul(fmul(SH_C3_6, x), fsub(xx, fmul(3.0f, yy)));
#pragma unroll
                for (int k = 0; k < 16; ++k) bf[k] = (k < ncoef) ? bf[k] : 0.f;
                // 64 B per survivor, float4 slots XOR-swizzled so that both these stores and the phase-2 loads
                // are bank-conflict free
                const int sw = (pos >> 1) & 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) bp[pos * 4 + (j ^ sw)] = make_float4(bf[4 * j], bf[4 * j + 1], bf[4 * j + 2], bf[4 * j + 3]);
            }
            __syncwarp();
            const int n = __popc(mask);
            const int q = lane & 3;
            const int nk = max(0, min(4, ncoef - 4 * q));  // coefficients this lane owns: 4q .. 4q+nk-1
            // Software pipeline (GSR_PRE_SH_PIPE): the 48 bytes of the NEXT group of 8 survivors are requested before the
            // current group is evaluated, so a warp with 3-4 groups pays one exposed L2 round trip instead of one each.
#ifndef GSR_PRE_SH_PIPE
#define GSR_PRE_SH_PIPE 0
#endif
            float sn[12];
            int gnext = 0;
            auto fetch = [&](int slot_, float* dst, int& g_) {
                g_ = 0;
#pragma unroll
                for (int i = 0; i < 12; ++i) dst[i] = 0.f;
                if (slot_ < n) {
                    g_ = s_queue[warp][slot_];
                    const float* sp = p.shs + ((size_t)g_ * p.M + 4 * q) * 3;
                    if (nk == 4 && vec_sh) {
                        const float4 a = ldg_f4(sp), b = ldg_f4(sp + 4), cc = ldg_f4(sp + 8);
                        dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
                        dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
                        dst[8] = cc.x; dst[9] = cc.y; dst[10] = cc.z; dst[11] = cc.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 12; ++i) dst[i] = (i < nk * 3) ? __ldg(sp + i) : 0.f;
                    }
                }
            };
            if (GSR_PRE_SH_PIPE) fetch(lane >> 2, sn, gnext);
            for (int r0 = 0; r0 < n; r0 += 8) {
                const int slot = r0 + (lane >> 2);
                const bool act = slot < n;
                float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
                float s[12];
                int gidx;
                if (GSR_PRE_SH_PIPE) {
#pragma unroll
                    for (int i = 0; i < 12; ++i) s[i] = sn[i];
                    gidx = gnext;
                    if (r0 + 8 < n) fetch(slot + 8, sn, gnext);  // warp-uniform branch
                } else {
                    fetch(slot, s, gidx);
                }
                if (act) {
                    const float4 bq = bp[slot * 4 + (q ^ ((slot >> 1) & 3))];
                    acc0 = __fmaf_rn(bq.w, s[9], __fmaf_rn(bq.z, s[6], __fmaf_rn(bq.y, s[3], fmul(bq.x, s[0]))));
                    acc1 = __fmaf_rn(bq.w, s[10], __fmaf_rn(bq.z, s[7], __fmaf_rn(bq.y, s[4], fmul(bq.x, s[1]))));
                    acc2 = __fmaf_rn(bq.w, s[11], __fmaf_rn(bq.z, s[8], __fmaf_rn(bq.y, s[5], fmul(bq.x, s[2]))));
                }
                acc0 = fadd(acc0, __shfl_xor_sync(0xffffffffu, acc0, 1));
                acc1 = fadd(acc1, __shfl_xor_sync(0xffffffffu, acc1, 1));
                acc2 = fadd(acc2, __shfl_xor_sync(0xffffffffu, acc2, 1));
                acc0 = fadd(acc0, __shfl_xor_sync(0xffffffffu, acc0, 2));
                acc1 = fadd(acc1, __shfl_xor_sync(0xffffffffu, acc1, 2));
                acc2 = fadd(acc2, __shfl_xor_sync(0xffffffffu, acc2, 2));
                if (act && q < 3) {  // lane q writes channel q: 12 contiguous bytes per Gaussian across 3 lanes
                    const float v = fadd(q == 0 ? acc0 : (q == 1 ? acc1 : acc2), 0.5f);
                    if (p.clamped) p.clamped[(size_t)gidx * 3 + q] = v < 0.f;
                    p.rgb[(size_t)gidx * 3 + q] = fmaxf(v, 0.f);
                }
            }
        }
    }

    // ---- per-block partial sum for the tiles_touched scan --------------------------------
    // ... and of the (Gaussian, 8x8-tile bin) records the bin-expansion path will emit
    const uint2 crec = coarse_rect(rec);
    const uint32_t wsum = __reduce_add_sync(0xffffffffu, tiles);
    const uint32_t wsum2 = __reduce_add_sync(0xffffffffu, (crec.y >> 16) * (crec.y & 0xffffu));
    if (lane == 0) { s_wsum[warp] = wsum; s_wsum2[warp] = wsum2; }
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0, t2 = 0;
#pragma unroll
        for (int w = 0; w < PRE_THREADS / 32; ++w) { t += s_wsum[w]; t2 += s_wsum2[w]; }
        p.block_sums[blockIdx.x] = t;
        if (p.coarse_block_sums) p.coarse_block_sums[blockIdx.x] = t2;
    }
}

}  // namespace

int launch_preprocess(const PreprocessParams& p, bool compat, cudaStream_t s) {
    if (p.P <= 0) return 0;
    const int blocks = (p.P + PRE_THREADS - 1) / PRE_THREADS;
    const int vec_means = (p.means_stride == 3) && ((reinterpret_cast<uintptr_t>(p.means3D) & 15) == 0);
    const int vec_scales = p.scales && (p.scales_stride == 3) && ((reinterpret_cast<uintptr_t>(p.scales) & 15) == 0);
    const int vec_sh = p.shs && ((reinterpret_cast<uintptr_t>(p.shs) & 15) == 0) && ((p.M * 3) % 4 == 0);
    GSR_CARVEOUT(preprocess_kernel<true>, "PRE", -1);
    GSR_CARVEOUT(preprocess_kernel<false>, "PRE", -1);
    if (compat)
        preprocess_kernel<true><<<blocks, PRE_THREADS, 0, s>>>(p, vec_means, vec_scales, vec_sh);
    else
        preprocess_kernel<false><<<blocks, PRE_THREADS, 0, s>>>(p, vec_means, vec_scales, vec_sh);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return -(int)e;
    return 1;
}

}  // namespace gsr
