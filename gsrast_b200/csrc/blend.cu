// blend.cu — per-16x16-tile front-to-back alpha blending (sm_100a).
//
// Replaces renderCUDA / render (/root/reference/apps/gsrast/gscuda/GSCuda.cu:543-693).
// Semantics kept exactly: splats of a tile are visited in sorted order; a splat is skipped
// when power > 0 or alpha = min(0.99, opacity*exp(power)) < 1/255; a pixel stops (and does
// not blend the splat) when T*(1-alpha) < t_min; out = C + T*background; final_T and
// n_contrib (1-based index of the last blended splat) are recorded per pixel.
//
// Two kernels:
//   blend_simple_kernel  the reference's structure (256-splat shared-memory batches, one
//                        pixel per thread) with the block-wide early exit; kept for A/B.
//   blend_culled_kernel  default.  Each warp owns an 8x4-pixel sub-rectangle of the tile.
//                        While a batch is staged, the staging thread of every splat computes
//                        conservatively which of the 8 sub-rectangles the splat can reach with
//                        alpha >= 1/255 (bounding box of the alpha >= 1/255 ellipse, refined by the
//                        maximum of the concave quadratic `power` over the rectangle); each warp
//                        compacts its own survivors with warp ballots into a private index list,
//                        walks only those, and leaves the batch loop as soon as its 32 pixels
//                        have saturated (__all_sync), independently of the other warps.
// There is no dense contraction here, hence no tensor cores: the work is FP32 FMA + MUFU.EX2
// issue and shared-memory broadcast bandwidth.
#include "gsr_common.cuh"

namespace gsr {

namespace {

constexpr int BLEND_THREADS = TILE_X * TILE_Y;  // 256
constexpr int BATCH = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float ALPHA_MIN = 1.0f / 255.0f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLEND_THREADS) blend_simple_kernel(const BlendParams p) {
    __shared__ float4 s_a[BATCH];  // x, y, conic.x, conic.y
    __shared__ float4 s_b[BATCH];  // conic.z, opacity, r, g
    __shared__ float s_c[BATCH];   // b

    const int tile = blockIdx.x;
    const int tile_x = tile % p.grid_x, tile_y = tile / p.grid_x;
    const int tid = threadIdx.x;
    const int lx = tid & 15, ly = tid >> 4;
    const int pix_x = tile_x * TILE_X + lx, pix_y = tile_y * TILE_Y + ly;
    const bool inside = pix_x < p.W && pix_y < p.H;
    const float pixf_x = (float)pix_x, pixf_y = (float)pix_y;

    gsr_pdl_wait();
    const uint2 range = reinterpret_cast<const uint2*>(p.ranges)[tile];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + BATCH - 1) / BATCH;
    int todo = total;

    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t contributor = 0, last = 0;

    for (int r = 0; r < rounds; ++r, todo -= BATCH) {
        if (__syncthreads_count(done) == BLEND_THREADS) break;
        const int progress = r * BATCH + tid;
        if (progress < total) {
            const uint32_t id = __ldg(p.point_list + range.x + progress);
            const float2 xy = __ldg(reinterpret_cast<const float2*>(p.means2D) + id);
            const float4 co = __ldg(reinterpret_cast<const float4*>(p.conic_opacity) + id);
            const float* col = p.colors + (size_t)id * 3;
            s_a[tid] = make_float4(xy.x, xy.y, co.x, co.y);
            s_b[tid] = make_float4(co.z, co.w, __ldg(col), __ldg(col + 1));
            s_c[tid] = __ldg(col + 2);
        }
        __syncthreads();
        const int nb = min(BATCH, todo);
        for (int j = 0; !done && j < nb; ++j) {
            contributor++;
            const float4 a = s_a[j];
            const float4 b = s_b[j];
            const float dx = a.x - pixf_x, dy = a.y - pixf_y;
            const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, b.y * __expf(power));
            if (alpha < ALPHA_MIN) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < p.t_min) {
                done = true;
                continue;
            }
            const float w = alpha * T;
            C0 += b.z * w;
            C1 += b.w * w;
            C2 += s_c[j] * w;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t pix = (size_t)pix_y * p.W + pix_x;
        const size_t plane = (size_t)p.W * p.H;
        p.final_T[pix] = T;
        p.n_contrib[pix] = last;
        p.out_color[pix] = C0 + T * __ldg(p.background + 0);
        p.out_color[pix + plane] = C1 + T * __ldg(p.background + 1);
        p.out_color[pix + 2 * plane] = C2 + T * __ldg(p.background + 2);
    }
}

// -------------------------------------------------------------------------------------------
// Maximum of power(d) = -0.5*(a dx^2 + c dy^2) - b dx dy over the box dx in [x0,x1], dy in
// [y0,y1] for a positive-definite conic: it is 0 when the box contains the origin, otherwise it
// sits on the box edge nearest the origin along x or along y, at the 1-D optimum clamped to the
// edge.  mba = -b/a, mbc = -b/c.
__device__ __forceinline__ float max_power_in_box(float a, float b, float c, float mba, float mbc, float x0, float x1,
                                                  float y0, float y1) {
    const float ex = fminf(fmaxf(0.f, x0), x1);
    const float ey = fminf(fmaxf(0.f, y0), y1);
    const float dy1 = fminf(fmaxf(mbc * ex, y0), y1);
    const float dx2 = fminf(fmaxf(mba * ey, x0), x1);
    const float f1 = -0.5f * (a * ex * ex + c * dy1 * dy1) - b * ex * dy1;
    const float f2 = -0.5f * (a * dx2 * dx2 + c * ey * ey) - b * dx2 * ey;
    return fmaxf(f1, f2);
}

// Cull-bound arithmetic.  The bounds carry multiplicative and additive slack (1.0001 / 0.01 px / 1e-3 in the
// exponent), orders of magnitude above the 2-ulp error of the approximate units, and the 1-D optimum enters
// `power` only to second order — so MUFU-based sqrt and division replace the IEEE sequences
// (profiles/r01f_ab.txt: blend 0.356 -> 0.351 ms at C2).
#ifndef GSR_BLEND_FASTCULL
#define GSR_BLEND_FASTCULL 1
#endif
__device__ __forceinline__ float cull_sqrt(float x) {
#if GSR_BLEND_FASTCULL
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return sqrtf(x);
#endif
}
__device__ __forceinline__ float cull_div(float a, float b) {
#if GSR_BLEND_FASTCULL
    return __fdividef(a, b);
#else
    return a / b;
#endif
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// Staged splat record: 48 bytes, three 16-byte shared loads off one base address.
//   [0] x, y, a', b'      (conic pre-scaled by log2(e): power*log2(e) = a' dx^2 + c' dy^2 + b' dx dy)
//   [1] c', opacity, r, g
//   [2] b, -, -, -
#ifndef GSR_BLEND_BATCH
#define GSR_BLEND_BATCH 128   // splats staged per round by the culled kernel (<= BLEND_THREADS, multiple of 32)
#endif
constexpr int CBATCH = GSR_BLEND_BATCH;
static_assert(CBATCH <= BLEND_THREADS && CBATCH % 32 == 0, "one staging thread per splat of a batch");
#ifndef GSR_BLEND_CARVEOUT
#define GSR_BLEND_CARVEOUT 25   // 64 KB of shared memory: six 8.3 KB CTAs (+1 KB reserved each) fit, the rest stays L1 for the gathers
#endif
// GSR_BLEND_ABS16=1: the per-warp lists hold the 16-bit shared-window ADDRESS of a record instead of its offset, so
// the candidate loop loads the record straight off the list entry (one IADD less per trip).  Static shared memory of
// a non-cluster launch sits in the low 64 KB of the window; the kernel traps if that ever does not hold.
#ifndef GSR_BLEND_ABS16
#define GSR_BLEND_ABS16 1
#endif
#ifndef GSR_BLEND_MINB
#define GSR_BLEND_MINB 6   // 40 registers (12 B of spills outside the candidate loop): 48 resident warps, blend -2.6 %
#endif
// COUNT: the same kernel with work counters (gsr_stage_times.blend_counters, GSR_FLAG_BLEND_COUNT) — a separate
// instantiation for reporting; the default one carries none of it.
template <bool COUNT>
__global__ void __launch_bounds__(BLEND_THREADS, GSR_BLEND_MINB) blend_culled_kernel(const BlendParams p) {
    __shared__ float4 s_splat[CBATCH * 3];
    __shared__ unsigned char s_mask[CBATCH];                     // bit w: splat can reach warp w's 8x4 sub-rectangle
    __shared__ unsigned short s_list[BLEND_THREADS / 32][CBATCH]; // per warp: byte offsets (48 * staged index) of its candidates

    const int tile = p.tile_order ? (int)__ldg(p.tile_order + blockIdx.x) : (int)blockIdx.x;
    const int tile_x = tile % p.grid_x, tile_y = tile / p.grid_x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // warp w -> sub-rectangle (w&1, w>>1) of 8x4 pixels; lane -> (lane&7, lane>>3)
    const int lx = ((warp & 1) << 3) | (lane & 7), ly = ((warp >> 1) << 2) | (lane >> 3);
    const int pix_x = tile_x * TILE_X + lx, pix_y = tile_y * TILE_Y + ly;
    const bool inside = pix_x < p.W && pix_y < p.H;
    const float pixf_x = (float)pix_x, pixf_y = (float)pix_y;
    const float tile_x0 = (float)(tile_x * TILE_X), tile_y0 = (float)(tile_y * TILE_Y);
    const uint32_t lane_lt = (1u << lane) - 1u;
    unsigned short* my_list = s_list[warp];
    // 32-bit shared-window address of the staging buffer, formed once: the inner loop then issues plain
    // LDS [reg + imm] (a generic float4* made the compiler re-derive the window base inside the loop)
    uint32_t splat_base = (uint32_t)__cvta_generic_to_shared(s_splat);
    uint32_t list_base = (uint32_t)__cvta_generic_to_shared(my_list);
    float t_min = p.t_min;
    // opaque to the optimiser, so the three values live in registers instead of being re-materialised
    // (S2UR SR_CgaCtaId + ULEA + LDCU) in every trip of the inner loop
    float pixf_xo = pixf_x, pixf_yo = pixf_y;
    asm volatile("" : "+r"(splat_base), "+r"(list_base), "+f"(t_min), "+f"(pixf_xo), "+f"(pixf_yo));

#if GSR_BLEND_ABS16
    if ((splat_base + (uint32_t)(CBATCH * 48)) >> 16) __trap();
    const uint32_t rec_bias = splat_base;
#else
    constexpr uint32_t rec_bias = 0u;
#endif
    gsr_pdl_wait();
    const uint2 range = reinterpret_cast<const uint2*>(p.ranges)[tile];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + CBATCH - 1) / CBATCH;

    // pixels outside the image start "terminated" (negative T, see the inner loop)
    float T = inside ? 1.0f : -1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last = 0;
    bool warp_done = __all_sync(0xffffffffu, !inside);

    // Software pipeline: the gather of batch r+1 (sorted id -> mean, conic, colour: two dependent
    // memory round trips) is issued into registers before batch r is blended and lands while it runs.
    float2 n_xy = make_float2(0.f, 0.f);
    float4 n_co = make_float4(0.f, 0.f, 0.f, 0.f);
    float n_c0 = 0.f, n_c1 = 0.f, n_c2 = 0.f;
    const bool stager = tid < CBATCH;  // thread t stages splat t of every batch
    if (stager && tid < total) {
        const uint32_t id = __ldg(p.point_list + range.x + tid);
        n_xy = __ldg(reinterpret_cast<const float2*>(p.means2D) + id);
        n_co = __ldg(reinterpret_cast<const float4*>(p.conic_opacity) + id);
        const float* col = p.colors + (size_t)id * 3;
        n_c0 = __ldg(col); n_c1 = __ldg(col + 1); n_c2 = __ldg(col + 2);
    }

    for (int r = 0; r < rounds; ++r) {
        if (__syncthreads_and(warp_done)) break;
        if (COUNT && tid == 0) {
            atomicAdd(p.counters + 0, 1ull);
            atomicAdd(p.counters + 7, (unsigned long long)min(CBATCH, total - r * CBATCH));
        }
        const int progress = r * CBATCH + tid;
        uint32_t m = 0;
        const float2 xy = n_xy;
        const float4 co = n_co;
        const float cr = n_c0, cg = n_c1, cb = n_c2;
        if (stager && progress + CBATCH < total) {
            const uint32_t id = __ldg(p.point_list + range.x + progress + CBATCH);
            n_xy = __ldg(reinterpret_cast<const float2*>(p.means2D) + id);
            n_co = __ldg(reinterpret_cast<const float4*>(p.conic_opacity) + id);
            const float* col = p.colors + (size_t)id * 3;
            n_c0 = __ldg(col); n_c1 = __ldg(col + 1); n_c2 = __ldg(col + 2);
        }
        if (stager && progress < total) {
            const float a = co.x, b = co.y, c = co.z, o = co.w;
            s_splat[3 * tid + 0] = make_float4(xy.x, xy.y, -0.5f * LOG2E * a, -LOG2E * b);
            s_splat[3 * tid + 1] = make_float4(-0.5f * LOG2E * c, o, cr, cg);
            s_splat[3 * tid + 2] = make_float4(cb, 0.f, 0.f, 0.f);
            // ---- which sub-rectangles can see this splat with alpha >= 1/255 ? ----------
            // alpha >= 1/255  <=>  power >= -ln(255*o) =: thr.  First the axis-aligned bounding box of that
            // ellipse (half extents sqrt(-2 thr c/det), sqrt(-2 thr a/det)) picks candidate sub-rectangles,
            // then the exact maximum of `power` over each candidate decides.  Slack everywhere so float
            // rounding in the bounds can only add work, never drop a contributing splat.
            const float det = a * c - b * b;
            if (!(o >= ALPHA_MIN * 0.999f)) {
                m = 0;  // exp(power) <= 1  =>  alpha < 1/255 everywhere (also catches NaN opacity)
            } else if (!(a > 0.f && c > 0.f && det > 0.f) || !(fabsf(xy.x) < 1e7f) || !(fabsf(xy.y) < 1e7f)) {
                m = 0xffu;  // degenerate conic: no culling
            } else {
                const float thr = -__logf(255.0f * o) * 1.0001f - 1e-3f;
                const float sc = __fdividef(-2.0f * thr, det);
                const float ex = cull_sqrt(sc * c) * 1.0001f + 0.01f, ey = cull_sqrt(sc * a) * 1.0001f + 0.01f;
                const float bx1 = xy.x - tile_x0, by1 = xy.y - tile_y0;  // centre relative to the tile origin
                const float xlo = bx1 - ex, xhi = bx1 + ex, ylo = by1 - ey, yhi = by1 + ey;
                if (!(ex < 1e7f) || !(ey < 1e7f)) {
                    m = 0xffu;
                } else {
                    const uint32_t colm = ((xlo <= 7.f && xhi >= 0.f) ? 1u : 0u) | ((xlo <= 15.f && xhi >= 8.f) ? 2u : 0u);
                    uint32_t cand = 0;
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr)
                        if (ylo <= (float)(4 * rr + 3) && yhi >= (float)(4 * rr)) cand |= colm << (2 * rr);
                    if (cand) {
                        const float mba = cull_div(-b, a), mbc = cull_div(-b, c);
                        while (cand) {
                            const int w = __ffs(cand) - 1;
                            cand &= cand - 1;
                            // dx = x - px, px in [X, X+7]  ->  dx in [x - (X+7), x - X]
                            const float wx1 = bx1 - (float)((w & 1) << 3), wy1 = by1 - (float)((w >> 1) << 2);
                            if (max_power_in_box(a, b, c, mba, mbc, wx1 - 7.f, wx1, wy1 - 3.f, wy1) >= thr) m |= 1u << w;
                        }
                    }
                }
            }
        }
        if (stager) s_mask[tid] = (unsigned char)m;
        __syncthreads();

        if (!warp_done) {
            // this warp's candidates of the batch, in order
            int n = 0;
#pragma unroll
            for (int c0 = 0; c0 < CBATCH; c0 += 32) {
                const bool mine = (s_mask[c0 + lane] >> warp) & 1u;
                const unsigned bits = __ballot_sync(0xffffffffu, mine);
                if (mine) my_list[n + __popc(bits & lane_lt)] = (unsigned short)(rec_bias + (uint32_t)((c0 + lane) * 48));
                n += __popc(bits);
            }
            __syncwarp();
            if (COUNT && lane == 0) { atomicAdd(p.counters + 1, 1ull); atomicAdd(p.counters + 2, (unsigned long long)n); }
            uint32_t c_trips = 0, c_live = 0, c_cand = 0, c_blend = 0;
            // Branch-free inner loop.  A terminated pixel carries its final transmittance as a NEGATIVE T:
            // then w = alpha*T and T - w are negative, "T - w >= t_min" fails, nothing is blended, and
            // no separate `done` flag has to be tested.  Per candidate: LDS.U16, 2 LDS.128, 8 FP32 for the
            // exponent, MUFU.EX2, 4 FP32 + 4 FSETP for alpha / transmittance, then predicated: LDS, 3 FFMA, T, last.
            uint32_t last_off = 0xffffffffu;  // byte offset of the last splat blended in this batch
            for (int i0 = 0; i0 < n; i0 += 16) {
                const int i1 = min(n, i0 + 16);
#pragma unroll 2
                for (int i = i0; i < i1; ++i) {
                    const uint32_t off = lds16(list_base + 2u * (uint32_t)i);
#if GSR_BLEND_ABS16
                    const uint32_t rec = off;  // the entry is the record's address
#else
                    const uint32_t rec = splat_base + off;
#endif
                    const float4 a = lds128(rec);
                    const float4 b = lds128(rec + 16u);
                    const float dx = a.x - pixf_xo, dy = a.y - pixf_yo;
                    // the three terms are formed separately like the reference's (GSCuda.cu:634): a factored form
                    // saves one FMUL but rounds differently where they cancel (elongated splats far from the centre)
                    const float p2 = fmaf(a.z, dx * dx, fmaf(b.x, dy * dy, a.w * (dx * dy)));
                    const float alpha = fminf(0.99f, b.y * ex2_approx(p2));
                    const float w = alpha * T;
                    const float test_T = T - w;  // T*(1-alpha)
                    const bool cand = (p2 <= 0.0f) && (alpha >= ALPHA_MIN);
                    const bool pass = test_T >= t_min;
                    const bool ok = cand && pass;
                    if (COUNT) {
                        ++c_trips;
                        c_live += T > 0.0f;
                        c_cand += cand && T > 0.0f;
                        c_blend += ok;
                    }
                    if (ok) {
                        C0 = fmaf(b.z, w, C0);
                        C1 = fmaf(b.w, w, C1);
                        C2 = fmaf(lds32(rec + 32u), w, C2);
                        last_off = off;
                    }
                    // one select updates T for every candidate: blended -> T(1 - alpha); would fall below t_min -> stop,
                    // keeping the final T as -|T| (true again for every later candidate: idempotent)
                    if (cand) T = pass ? test_T : -fabsf(T);
                }
                if (__all_sync(0xffffffffu, T <= 0.0f)) {
                    warp_done = true;
                    break;
                }
            }
            if (last_off != 0xffffffffu) last = (uint32_t)(r * CBATCH + 1) + (last_off - rec_bias) / 48u;
            if (COUNT) {
                c_live = __reduce_add_sync(0xffffffffu, c_live);
                c_cand = __reduce_add_sync(0xffffffffu, c_cand);
                c_blend = __reduce_add_sync(0xffffffffu, c_blend);
                if (lane == 0) {
                    atomicAdd(p.counters + 3, (unsigned long long)c_trips);
                    atomicAdd(p.counters + 4, (unsigned long long)c_live);
                    atomicAdd(p.counters + 5, (unsigned long long)c_cand);
                    atomicAdd(p.counters + 6, (unsigned long long)c_blend);
                }
            }
        }
    }
    T = fabsf(T);
    if (inside) {
        const size_t pix = (size_t)pix_y * p.W + pix_x;
        const size_t plane = (size_t)p.W * p.H;
        p.final_T[pix] = T;
        p.n_contrib[pix] = last;
        p.out_color[pix] = fmaf(T, __ldg(p.background + 0), C0);
        p.out_color[pix + plane] = fmaf(T, __ldg(p.background + 1), C1);
        p.out_color[pix + 2 * plane] = fmaf(T, __ldg(p.background + 2), C2);
    }
}

// -------------------------------------------------------------------------------------------
// blend_pair_kernel: two pixels per thread on the packed FP32 pipe (FADD2 / FMUL2 / FFMA2 of sm_100, one issue slot
// for both pixels; a scalar operand is broadcast by the instruction's .F32 operand form, so per-splat coefficients
// need no duplication — tools/microbench_f32x2.cu measures the issue cost).  One CTA of 128 threads per 16x16 tile,
// warp w owns the 8x8 quadrant (w&1, w>>1), lane -> (lane&7, lane>>3) and (lane&7, (lane>>3)+4): the two pixels share
// dx, dx*dx and every per-splat load, and the per-pixel arithmetic of the exponent, alpha, the transmittance test and
// the colour accumulation is issued once for the pair.  A candidate trip costs ~38 issue slots for 64 pixels where the
// one-pixel kernel pays 28 for 32; the lists of an 8x8 quadrant are 0.61x as long as those of its two 8x4 halves
// together (tests/analysis/blend_cull_model.py), list building is done by 4 warps instead of 8, and all 128 threads
// stage (no idle staging warps).  Arithmetic per pixel is the one-pixel kernel's, operation for operation.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// GSR_PAIR_PREDCOL=1: colour accumulation as six predicated scalar FFMAs (FMA pipe) instead of two weight selects (ALU
// pipe) + three FFMA2 — one issue slot more per trip, two half-rate ALU instructions fewer; GSR_PAIR_UNROLL: trips per
// loop iteration.  Measured C2 / C5 blend (profiles/r02g_ab_*.txt): packed colour, unroll 2: 0.262 / 1.055 ms; unroll 4:
// 0.255 / 1.036; predicated colour: 0.251 / 1.018; both: 0.251 / 1.015.
#ifndef GSR_PAIR_PREDCOL
#define GSR_PAIR_PREDCOL 1
#endif
#ifndef GSR_PAIR_UNROLL
#define GSR_PAIR_UNROLL 4
#endif
constexpr int PAIR_UNROLL = GSR_PAIR_UNROLL;
constexpr int PAIR_THREADS = 128;
constexpr int PAIR_WARPS = PAIR_THREADS / 32;
// Measured (profiles/r02m_ab_*.txt, r02n_ab_*.txt; C2 / C5 / C3 blend ms): two barriers, 8 CTAs/SM 0.252 / 1.024 / 0.949;
// double-buffered at 8 CTAs/SM 0.264 / 1.018 / 0.984 (the second buffer costs the L1 the gathers live in); at 7 CTAs/SM
// (72 registers) 0.253 / 0.990 / 0.949 — the default; at 6: 0.265 / 1.033 / 1.000.  Two barriers at 7 and 6 CTAs/SM:
// 0.252 / 1.019 / 0.948 and 0.258 / 1.037 / 0.981.
#ifndef GSR_PAIR_DBUF
#define GSR_PAIR_DBUF 1
#endif
constexpr int PAIR_NBUF = GSR_PAIR_DBUF ? 2 : 1;
static_assert(PAIR_WARPS <= 32, "s_done_flags holds one bit per warp");
static_assert(CBATCH == PAIR_THREADS, "every thread of the pair kernel stages one splat per round");
#ifndef GSR_PAIR_MINB
#define GSR_PAIR_MINB (GSR_PAIR_DBUF ? 7 : 8)   // CTAs of 128 threads per SM.  Two-barrier form, measured C2 / C5 blend: 12 CTAs
                           // 0.290 / 1.173 ms, 10 CTAs 0.272 / 1.097, 8 CTAs (64 registers) 0.265 / 1.075 (profiles/r02b_ab_*.txt)
#endif
template <bool COUNT>
__global__ void __launch_bounds__(PAIR_THREADS, GSR_PAIR_MINB) blend_pair_kernel(const BlendParams p) {
    // GSR_PAIR_DBUF=1: two staging buffers and ONE barrier per round.  A warp that leaves its candidate walk early stages
    // its share of the next batch into the other buffer while the slower quadrants are still walking, instead of waiting
    // for them twice per round (ncu, r02h: 18 % of the stall samples sit on the round barrier, 12 % in the staging code
    // behind it); once every quadrant has reported done the staging is skipped, so the speculation costs nothing but
    // issue slots of warps that would otherwise wait.  Same rounds, same arithmetic, identical output.
    __shared__ float4 s_splat[PAIR_NBUF][CBATCH * 3];
    __shared__ unsigned char s_mask[PAIR_NBUF][CBATCH];       // bit w: splat can reach warp w's 8x8 quadrant
    __shared__ unsigned short s_list[PAIR_WARPS][CBATCH];     // per warp: shared-window addresses of its candidates' records
    __shared__ uint32_t s_done_flags;                         // DBUF: bit w set once warp w has finished all its pixels (atomics only)

    const int tile = (int)blockIdx.x;
    const int tile_x = tile % p.grid_x, tile_y = tile / p.grid_x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (GSR_PAIR_DBUF && tid == 0) s_done_flags = 0;
    const int lx = ((warp & 1) << 3) | (lane & 7), ly = ((warp >> 1) << 3) | (lane >> 3);
    const int pix_x = tile_x * TILE_X + lx, pix_y0 = tile_y * TILE_Y + ly, pix_y1 = pix_y0 + 4;
    const bool inside0 = pix_x < p.W && pix_y0 < p.H, inside1 = pix_x < p.W && pix_y1 < p.H;
    const float tile_x0 = (float)(tile_x * TILE_X), tile_y0 = (float)(tile_y * TILE_Y);
    const uint32_t lane_lt = (1u << lane) - 1u;
    unsigned short* my_list = s_list[warp];
    uint32_t splat_base = (uint32_t)__cvta_generic_to_shared(s_splat);
    uint32_t list_base = (uint32_t)__cvta_generic_to_shared(my_list);
    float t_min = p.t_min;
    float pixf_x = (float)pix_x;
    u64 npy2 = pk2(-(float)pix_y0, -(float)pix_y1);
    asm volatile("" : "+r"(splat_base), "+r"(list_base), "+f"(t_min), "+f"(pixf_x), "+l"(npy2));
    if ((splat_base + (uint32_t)(PAIR_NBUF * CBATCH * 48)) >> 16) __trap();  // the lists hold 16-bit shared-window addresses

    gsr_pdl_wait();
    const uint2 range = reinterpret_cast<const uint2*>(p.ranges)[tile];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + CBATCH - 1) / CBATCH;

    // pixels outside the image start "terminated" (negative T, see blend_culled_kernel)
    float T0 = inside0 ? 1.0f : -1.0f, T1 = inside1 ? 1.0f : -1.0f;
#if GSR_PAIR_PREDCOL
    float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f, c20 = 0.f, c21 = 0.f;  // channel x pixel
#else
    u64 C0 = 0ull, C1 = 0ull, C2 = 0ull;  // {pixel 0, pixel 1} per channel
#endif
    uint32_t last0 = 0, last1 = 0;
    bool warp_done = __all_sync(0xffffffffu, !inside0 && !inside1);

    float2 n_xy = make_float2(0.f, 0.f);
    float4 n_co = make_float4(0.f, 0.f, 0.f, 0.f);
    float n_c0 = 0.f, n_c1 = 0.f, n_c2 = 0.f;
    if (tid < total) {
        const uint32_t id = __ldg(p.point_list + range.x + tid);
        n_xy = __ldg(reinterpret_cast<const float2*>(p.means2D) + id);
        n_co = __ldg(reinterpret_cast<const float4*>(p.conic_opacity) + id);
        const float* col = p.colors + (size_t)id * 3;
        n_c0 = __ldg(col); n_c1 = __ldg(col + 1); n_c2 = __ldg(col + 2);
    }

    if (GSR_PAIR_DBUF) {
        __syncthreads();  // s_done_flags is cleared
        if (warp_done && lane == 0) atomicOr(&s_done_flags, 1u << warp);  // quadrant off the image
    }
    for (int r = 0; r < rounds; ++r) {
        const int buf = GSR_PAIR_DBUF ? (r & 1) : 0;
        float4* const sp = s_splat[buf];
        const uint32_t rec_bias = splat_base + (uint32_t)(buf * CBATCH * 48);
        if (!GSR_PAIR_DBUF) {
            if (__syncthreads_and(warp_done)) break;
        }
        if (COUNT && !GSR_PAIR_DBUF && tid == 0) {
            atomicAdd(p.counters + 0, 1ull);
            atomicAdd(p.counters + 7, (unsigned long long)min(CBATCH, total - r * CBATCH));
        }
        const int progress = r * CBATCH + tid;
        uint32_t m = 0;
        const float2 xy = n_xy;
        const float4 co = n_co;
        const float cr = n_c0, cg = n_c1, cb = n_c2;
        // DBUF: every quadrant has reported done -> the vote below ends the tile, nothing of this batch will be read
        bool stage = true;
        if (GSR_PAIR_DBUF) {  // a hint read without any ordering (shared-memory atomics on both sides): a stale value stages in vain
            uint32_t f = 0;
            if (lane == 0) f = atomicOr(&s_done_flags, 0u);
            stage = __shfl_sync(0xffffffffu, f, 0) != (1u << PAIR_WARPS) - 1u;
        }
        if (stage && progress + CBATCH < total) {
            const uint32_t id = __ldg(p.point_list + range.x + progress + CBATCH);
            n_xy = __ldg(reinterpret_cast<const float2*>(p.means2D) + id);
            n_co = __ldg(reinterpret_cast<const float4*>(p.conic_opacity) + id);
            const float* col = p.colors + (size_t)id * 3;
            n_c0 = __ldg(col); n_c1 = __ldg(col + 1); n_c2 = __ldg(col + 2);
        }
        if (stage && progress < total) {
            const float a = co.x, b = co.y, c = co.z, o = co.w;
            sp[3 * tid + 0] = make_float4(xy.x, xy.y, -0.5f * LOG2E * a, -LOG2E * b);
            sp[3 * tid + 1] = make_float4(-0.5f * LOG2E * c, o, cr, cg);
            sp[3 * tid + 2] = make_float4(cb, 0.f, 0.f, 0.f);
            // which 8x8 quadrants can see this splat with alpha >= 1/255 (same bounds as blend_culled_kernel)
            const float det = a * c - b * b;
            if (!(o >= ALPHA_MIN * 0.999f)) {
                m = 0;
            } else if (!(a > 0.f && c > 0.f && det > 0.f) || !(fabsf(xy.x) < 1e7f) || !(fabsf(xy.y) < 1e7f)) {
                m = 0xfu;
            } else {
                const float thr = -__logf(255.0f * o) * 1.0001f - 1e-3f;
                const float sc = __fdividef(-2.0f * thr, det);
                const float ex = cull_sqrt(sc * c) * 1.0001f + 0.01f, ey = cull_sqrt(sc * a) * 1.0001f + 0.01f;
                const float bx1 = xy.x - tile_x0, by1 = xy.y - tile_y0;
                const float xlo = bx1 - ex, xhi = bx1 + ex, ylo = by1 - ey, yhi = by1 + ey;
                if (!(ex < 1e7f) || !(ey < 1e7f)) {
                    m = 0xfu;
                } else {
                    const uint32_t colm = ((xlo <= 7.f && xhi >= 0.f) ? 1u : 0u) | ((xlo <= 15.f && xhi >= 8.f) ? 2u : 0u);
                    uint32_t cand = ((ylo <= 7.f && yhi >= 0.f) ? colm : 0u) | ((ylo <= 15.f && yhi >= 8.f) ? (colm << 2) : 0u);
                    if (cand) {
                        const float mba = cull_div(-b, a), mbc = cull_div(-b, c);
                        while (cand) {
                            const int w = __ffs(cand) - 1;
                            cand &= cand - 1;
                            const float wx1 = bx1 - (float)((w & 1) << 3), wy1 = by1 - (float)((w >> 1) << 3);
                            if (max_power_in_box(a, b, c, mba, mbc, wx1 - 7.f, wx1, wy1 - 7.f, wy1) >= thr) m |= 1u << w;
                        }
                    }
                }
            }
        }
        s_mask[buf][tid] = (unsigned char)m;
        if (GSR_PAIR_DBUF) {
            // the one barrier of the round: the batch is staged, and every warp has finished the walk of the previous
            // round (so the buffer staged NEXT round is free); the vote is the one the two-barrier form takes
            if (__syncthreads_and(warp_done)) break;
            if (COUNT && tid == 0) {
                atomicAdd(p.counters + 0, 1ull);
                atomicAdd(p.counters + 7, (unsigned long long)min(CBATCH, total - r * CBATCH));
            }
        } else {
            __syncthreads();
        }

        if (!warp_done) {
            int n = 0;
#pragma unroll
            for (int c0 = 0; c0 < CBATCH; c0 += 32) {
                const bool mine = (s_mask[buf][c0 + lane] >> warp) & 1u;
                const unsigned bits = __ballot_sync(0xffffffffu, mine);
                if (mine) my_list[n + __popc(bits & lane_lt)] = (unsigned short)(rec_bias + (uint32_t)((c0 + lane) * 48));
                n += __popc(bits);
            }
            __syncwarp();
            if (COUNT && lane == 0) { atomicAdd(p.counters + 1, 1ull); atomicAdd(p.counters + 2, (unsigned long long)n); }
            uint32_t c_trips = 0, c_live = 0, c_cand = 0, c_blend = 0;
            uint32_t last_off0 = 0xffffffffu, last_off1 = 0xffffffffu;
            for (int i0 = 0; i0 < n; i0 += 16) {
                const int i1 = min(n, i0 + 16);
#pragma unroll PAIR_UNROLL
                for (int i = i0; i < i1; ++i) {
                    const uint32_t rec = lds16(list_base + 2u * (uint32_t)i);
                    const float4 a = lds128(rec);        // x, y, a', b'
                    const float4 b = lds128(rec + 16u);  // c', o, r, g
                    const float cb3 = lds32(rec + 32u);
                    const float dx = a.x - pixf_x;
                    const u64 dy2 = add2(pk2(a.y, a.y), npy2);
                    const float dxdx = dx * dx;
                    // per pixel exactly blend_culled_kernel's fma(a', dx*dx, fma(c', dy*dy, b'*(dx*dy)))
                    const u64 bd2 = mul2(pk2(a.w, a.w), mul2(pk2(dx, dx), dy2));
                    const u64 u2 = fma2(pk2(b.x, b.x), mul2(dy2, dy2), bd2);
                    const u64 pw2 = fma2(pk2(a.z, a.z), pk2(dxdx, dxdx), u2);
                    float p0, p1;
                    unpk2(pw2, p0, p1);
                    const u64 al2 = mul2(pk2(b.y, b.y), pk2(ex2_approx(p0), ex2_approx(p1)));
                    float al0, al1;
                    unpk2(al2, al0, al1);
                    al0 = fminf(0.99f, al0);
                    al1 = fminf(0.99f, al1);
                    const u64 T2 = pk2(T0, T1);
                    const u64 w2 = mul2(pk2(al0, al1), T2);
                    const u64 tt2 = sub2(T2, w2);  // T*(1-alpha) as T - alpha*T
                    float w0, w1, tt0, tt1;
                    unpk2(w2, w0, w1);
                    unpk2(tt2, tt0, tt1);
                    // The decision tail of a pixel.  The ALU pipe (FSETP / FSEL / FMNMX / SEL, half rate) is the busiest
                    // pipe of this kernel (ncu: 58 %), so the tail is written with THREE compares per pixel: `ok` takes
                    // `cand` as its predicate input, and under `cand` the transmittance select needs `ok` only (ok ==
                    // pass there).  Tried and dropped: T' = T - wm as a packed FADD2 with a predicated
                    // -|T| for the stop (ptxas keeps a select: 42 slots per trip instead of 39); an inline-PTX
                    // two-destination setp (ptxas turns it back into FSETP + PLOP3 + selects); `last` copied by a
                    // predicated FMUL x*1 on the FMA pipe (ptxas hoists the multiply and selects); a warp-uniform branch
                    // around the two FMNMX for splats with opacity <= 0.99 (if-converted to predicated FMNMX).
                    const bool cand0 = (p0 <= 0.0f) && (al0 >= ALPHA_MIN), cand1 = (p1 <= 0.0f) && (al1 >= ALPHA_MIN);
                    const bool ok0 = cand0 && (tt0 >= t_min), ok1 = cand1 && (tt1 >= t_min);
                    if (COUNT) {
                        ++c_trips;
                        c_live += (T0 > 0.0f) + (T1 > 0.0f);
                        c_cand += (cand0 && T0 > 0.0f) + (cand1 && T1 > 0.0f);
                        c_blend += ok0 + ok1;
                    }
                    if (ok0) last_off0 = rec;
                    if (ok1) last_off1 = rec;
                    // blended -> T(1 - alpha); would fall below t_min -> stop, keeping the final T as -|T|
                    if (cand0) T0 = ok0 ? tt0 : -fabsf(T0);
                    if (cand1) T1 = ok1 ? tt1 : -fabsf(T1);
#if GSR_PAIR_PREDCOL
                    if (ok0) { c00 = fmaf(b.z, w0, c00); c10 = fmaf(b.w, w0, c10); c20 = fmaf(cb3, w0, c20); }
                    if (ok1) { c01 = fmaf(b.z, w1, c01); c11 = fmaf(b.w, w1, c11); c21 = fmaf(cb3, w1, c21); }
#else
                    // a pixel that does not blend this splat adds +0 * colour (its weight is selected to zero)
                    const u64 wm2 = pk2(ok0 ? w0 : 0.0f, ok1 ? w1 : 0.0f);
                    C0 = fma2(pk2(b.z, b.z), wm2, C0);
                    C1 = fma2(pk2(b.w, b.w), wm2, C1);
                    C2 = fma2(pk2(cb3, cb3), wm2, C2);
#endif
                }
                if (__all_sync(0xffffffffu, T0 <= 0.0f && T1 <= 0.0f)) {
                    warp_done = true;
                    break;
                }
            }
            if (last_off0 != 0xffffffffu) last0 = (uint32_t)(r * CBATCH + 1) + (last_off0 - rec_bias) / 48u;
            if (last_off1 != 0xffffffffu) last1 = (uint32_t)(r * CBATCH + 1) + (last_off1 - rec_bias) / 48u;
            if (GSR_PAIR_DBUF && warp_done && lane == 0) atomicOr(&s_done_flags, 1u << warp);
            if (COUNT) {
                c_live = __reduce_add_sync(0xffffffffu, c_live);
                c_cand = __reduce_add_sync(0xffffffffu, c_cand);
                c_blend = __reduce_add_sync(0xffffffffu, c_blend);
                if (lane == 0) {
                    atomicAdd(p.counters + 3, (unsigned long long)c_trips * 2ull);  // in 32-pixel units, like the one-pixel kernel
                    atomicAdd(p.counters + 4, (unsigned long long)c_live);
                    atomicAdd(p.counters + 5, (unsigned long long)c_cand);
                    atomicAdd(p.counters + 6, (unsigned long long)c_blend);
                }
            }
        }
    }
    T0 = fabsf(T0);
    T1 = fabsf(T1);
#if !GSR_PAIR_PREDCOL
    float c00, c01, c10, c11, c20, c21;
    unpk2(C0, c00, c01);
    unpk2(C1, c10, c11);
    unpk2(C2, c20, c21);
#endif
    const size_t plane = (size_t)p.W * p.H;
    const float bg0 = __ldg(p.background + 0), bg1 = __ldg(p.background + 1), bg2 = __ldg(p.background + 2);
    if (inside0) {
        const size_t pix = (size_t)pix_y0 * p.W + pix_x;
        p.final_T[pix] = T0;
        p.n_contrib[pix] = last0;
        p.out_color[pix] = fmaf(T0, bg0, c00);
        p.out_color[pix + plane] = fmaf(T0, bg1, c10);
        p.out_color[pix + 2 * plane] = fmaf(T0, bg2, c20);
    }
    if (inside1) {
        const size_t pix = (size_t)pix_y1 * p.W + pix_x;
        p.final_T[pix] = T1;
        p.n_contrib[pix] = last1;
        p.out_color[pix] = fmaf(T1, bg0, c01);
        p.out_color[pix + plane] = fmaf(T1, bg1, c11);
        p.out_color[pix + 2 * plane] = fmaf(T1, bg2, c21);
    }
}

__global__ void fill_background_kernel(int n, const float* __restrict__ background, float* __restrict__ out_color,
                                       float* __restrict__ final_T, uint32_t* __restrict__ n_contrib) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    final_T[i] = 1.0f;
    n_contrib[i] = 0;
    out_color[i] = background[0];
    out_color[i + (size_t)n] = background[1];
    out_color[i + 2 * (size_t)n] = background[2];
}

}  // namespace

// Which blend kernel is the default: GSR_BLEND_PAIR (compile time), overridden by the environment variable
// GSR_BLEND_PAIR=0/1 (A/B runs on one build).
#ifndef GSR_BLEND_PAIR
#define GSR_BLEND_PAIR 1
#endif
#ifndef GSR_PAIR_CARVEOUT
#define GSR_PAIR_CARVEOUT (GSR_PAIR_DBUF ? 50 : 44)   // seven 13.6 KB CTAs (+1 KB reserved each): 102 KB; eight 7.3 KB ones: 66 KB
#endif
static bool blend_use_pair() {
    static const int v = [] {
        const char* e = getenv("GSR_BLEND_PAIR");
        return e ? atoi(e) : GSR_BLEND_PAIR;
    }();
    return v != 0;
}

int launch_blend(const BlendParams& p, bool simple, cudaStream_t s, bool one_pixel) {
    const int tiles = p.grid_x * p.grid_y;
    if (tiles <= 0) return 0;
    cudaError_t e;
    if (simple)
        e = launch_pdl(blend_simple_kernel, dim3(tiles), dim3(BLEND_THREADS), 0, s, p);
    else if (!one_pixel && blend_use_pair()) {
        if (p.counters) {
            GSR_CARVEOUT(blend_pair_kernel<true>, "BLEND", GSR_PAIR_CARVEOUT);
            e = launch_pdl(blend_pair_kernel<true>, dim3(tiles), dim3(PAIR_THREADS), 0, s, p);
        } else {
            GSR_CARVEOUT(blend_pair_kernel<false>, "BLEND", GSR_PAIR_CARVEOUT);
            e = launch_pdl(blend_pair_kernel<false>, dim3(tiles), dim3(PAIR_THREADS), 0, s, p);
        }
    } else if (p.counters) {
        GSR_CARVEOUT(blend_culled_kernel<true>, "BLEND", GSR_BLEND_CARVEOUT);
        e = launch_pdl(blend_culled_kernel<true>, dim3(tiles), dim3(BLEND_THREADS), 0, s, p);
    } else {
        GSR_CARVEOUT(blend_culled_kernel<false>, "BLEND", GSR_BLEND_CARVEOUT);
        e = launch_pdl(blend_culled_kernel<false>, dim3(tiles), dim3(BLEND_THREADS), 0, s, p);
    }
    return e == cudaSuccess ? 1 : -(int)e;
}

int launch_fill_background(int W, int H, const float* background, float* out_color, float* final_T,
                           uint32_t* n_contrib, cudaStream_t s) {
    const int n = W * H;
    fill_background_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, background, out_color, final_T, n_contrib);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

}  // namespace gsr
