// gsr_oracle.cpp — CPU ORACLE for the Gaussian-splat forward rasterizer.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under oracle/ is linked, imported or
// executed by the product (gsrast_b200/, include/).  Only tests/, the smoke check in
// __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.
//
// PARITY PINNING STATUS
//   MODE_GSRAST   PINNED.  tests/golden/gsrast_ref_*.npz are outputs of the reference's own in-tree
//                 rasterizer (apps/gsrast/gscuda/GSCuda.cu compiled unmodified into oracle/_ref and run on
//                 a B200, see tests/golden/README.md); tests/test_golden.py requires this oracle to match
//                 them bit for bit (radii, tile counts, offsets, sorted keys/values, ranges, float scratch)
//                 and the image within tolerance.
//   MODE_CONTRACT **parity unpinned** — the rasterizer the contract names
//                 (graphdeco-inria/diff-gaussian-rasterization @ 59f5f77e, submodule
//                 deps/diff-gaussian-rasterization) is an empty directory in /root/reference
//                 (.SUBMODULES.json:9-15) and the reference ships no tests, goldens or fixtures for it
//                 (SURVEY.md §4).  It is restated from SURVEY.md Appendix A; everything the two modes
//                 share (getRect, scan, key packing, stable sort, ranges, blend loop) is pinned through
//                 MODE_GSRAST.
//
// What it restates (all file:line relative to /root/reference):
//   forward orchestration        apps/gsrast/gscuda/GSCuda.cu:695-811
//   preprocess                   GSCuda.cu:261-375    (+ contract deltas)
//   quatToMat / computeCov3D     GSCuda.cu:157-195
//   computeCov2D                 GSCuda.cu:197-231
//   getRect (radius / rect)      GSCuda.cu:237-259
//   InclusiveSum                 GSCuda.cu:771
//   duplicateWithKeys            GSCuda.cu:422-475
//   getHigherMsb                 GSCuda.cu:481-502
//   SortPairs (stable LSD radix) GSCuda.cu:794-797
//   identifyTileRanges           GSCuda.cu:504-538
//   renderCUDA                   GSCuda.cu:543-677
//
// Arithmetic rules: every float operation is an individually rounded IEEE-754 binary32
// operation in the association order the source expression has (this file must be built
// with -ffp-contract=off and without -ffast-math).  GLM operators are restated with GLM's
// own association (mat3*mat3 and mat3*vec3 left-to-right, mat4*vec4 pairwise, vec4 dot
// pairwise).  float->int conversions saturate like CUDA's cvt.rzi.s32.f32 (NaN -> 0).
//
// Two modes:
//   MODE_CONTRACT (flags == 0): CudaRasterizer::Rasterizer::forward semantics as named by
//       BASELINE.json north_star (SIBR-era 29-argument form).
//   MODE_GSRAST   (flags & 1) : bit-for-bit restatement of the in-tree gscuda::forward
//       including its divergences (NDC cull, NDC-z depth keys, DC-only colour, vec4
//       strides, T<0.001 termination, y-extent without sqrt, R==1 range quirk, stale image
//       on R==0).

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

constexpr int BLOCK_X = 16;  // GSCuda.cu:20-21, Config.hpp:47-48
constexpr int BLOCK_Y = 16;

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
constexpr float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                            -1.0925484305920792f, 0.5462742152960396f};
constexpr float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                            0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                            -0.5900435899266435f};

// CUDA's float -> int32 conversion (cvt.rzi.s32.f32): truncates, saturates, NaN -> 0.
inline int f2i(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return std::numeric_limits<int>::max();
    if (v <= -2147483648.0f) return std::numeric_limits<int>::min();
    return (int)v;
}

struct V3 { float x, y, z; };

// Column-major 3x3 with GLM semantics: m.c[col][row].
struct M3 {
    float c[3][3];
};
inline M3 m3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) {
    M3 m;
    m.c[0][0] = a0; m.c[0][1] = a1; m.c[0][2] = a2;
    m.c[1][0] = b0; m.c[1][1] = b1; m.c[1][2] = b2;
    m.c[2][0] = c0; m.c[2][1] = c1; m.c[2][2] = c2;
    return m;
}
// glm::operator*(mat3, mat3): Result[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]
inline M3 mul(const M3& A, const M3& B) {
    M3 R;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            float p0 = A.c[0][r] * B.c[c][0];
            float p1 = A.c[1][r] * B.c[c][1];
            float p2 = A.c[2][r] * B.c[c][2];
            float s = p0 + p1;
            R.c[c][r] = s + p2;
        }
    return R;
}
inline M3 transpose(const M3& A) {
    M3 R;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) R.c[c][r] = A.c[r][c];
    return R;
}

// upstream auxiliary.h transformPoint4x3 / 4x4: left-to-right, column-major float[16].
inline V3 transformPoint4x3(const V3& p, const float* m) {
    V3 o;
    o.x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12];
    o.y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13];
    o.z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14];
    return o;
}
inline void transformPoint4x4(const V3& p, const float* m, float out[4]) {
    for (int r = 0; r < 4; ++r) out[r] = ((m[r] * p.x + m[4 + r] * p.y) + m[8 + r] * p.z) + m[12 + r];
}
// glm::operator*(mat4, vec4): (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w)    (GSCuda.cu:303,202)
inline void glmMat4Vec4(const float* m, const float v[4], float out[4]) {
    for (int r = 0; r < 4; ++r) {
        float a = m[r] * v[0];
        float b = m[4 + r] * v[1];
        float c = m[8 + r] * v[2];
        float d = m[12 + r] * v[3];
        float ab = a + b;
        float cd = c + d;
        out[r] = ab + cd;
    }
}

// getRect, both variants (GSCuda.cu:237-259): extent ex/ey are ints converted to float by
// the usual arithmetic conversions; division by 16 then truncation toward zero.
inline void getRect(float px, float py, int ex, int ey, int gx, int gy, uint32_t& minx, uint32_t& miny,
                    uint32_t& maxx, uint32_t& maxy) {
    minx = (uint32_t)std::min(gx, std::max(0, f2i((px - (float)ex) / (float)BLOCK_X)));
    miny = (uint32_t)std::min(gy, std::max(0, f2i((py - (float)ey) / (float)BLOCK_Y)));
    maxx = (uint32_t)std::min(gx, std::max(0, f2i((((px + (float)ex) + (float)BLOCK_X) - 1.0f) / (float)BLOCK_X)));
    maxy = (uint32_t)std::min(gy, std::max(0, f2i((((py + (float)ey) + (float)BLOCK_Y) - 1.0f) / (float)BLOCK_Y)));
}

template <class F>
void parallel_for(int64_t n, int threads, F f) {
    if (threads <= 1 || n < 2) {
        f(0, 0, n);
        return;
    }
    std::vector<std::thread> th;
    int64_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        int64_t b = std::min(n, t * chunk), e = std::min(n, b + chunk);
        th.emplace_back([=]() { f(t, b, e); });
    }
    for (auto& x : th) x.join();
}

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

struct GsrOracleIn {
    int32_t P, D, M, W, H;
    int32_t means_stride;   // floats between consecutive means (3 contract, 4 GSRast vec4)
    int32_t scales_stride;  // 3 or 4
    int32_t flags;          // bit0: MODE_GSRAST
    int32_t prefiltered;
    int32_t use_rects;      // rects != nullptr in the call (SIBR fast-culling variant)
    int32_t threads;
    float scale_modifier, tan_fovx, tan_fovy;
    const float* background;      // [3]
    const float* means3D;
    const float* shs;             // [P][M][3] contract; raw 48-float block in MODE_GSRAST
    const float* colors_precomp;  // [P][3] or null
    const float* opacities;       // [P]
    const float* scales;
    const float* rotations;       // [P][4] (r,x,y,z)
    const float* cov3D_precomp;   // [P][6] or null
    const float* viewmatrix;      // [16] column-major
    const float* projmatrix;      // [16]
    const float* cam_pos;         // [3]
    const float* boxmin;          // [3] or null
    const float* boxmax;          // [3] or null
};

struct GsrOracleOut {
    // caller-allocated, per Gaussian
    float* depths;            // [P]
    uint8_t* clamped;         // [3P]
    int32_t* radii;           // [P]
    float* means2D;           // [2P]
    float* cov3D;             // [6P]
    float* conic_opacity;     // [4P]
    float* rgb;               // [3P]
    uint32_t* tiles_touched;  // [P]
    uint32_t* point_offsets;  // [P]
    int32_t* rects;           // [2P]
    // caller-allocated, per tile / pixel
    uint32_t* ranges;     // [2T]
    uint32_t* n_contrib;  // [W*H]
    float* final_T;       // [W*H]
    float* out_color;     // [3*W*H] planar
    // oracle-allocated (free with gsr_oracle_free)
    uint64_t* keys_unsorted;
    uint32_t* values_unsorted;
    uint64_t* keys;
    uint32_t* values;
    int64_t num_rendered;
    int64_t pairs_evaluated;  // pixel-splat pairs visited by the blend loop
    double t_preprocess, t_scan, t_duplicate, t_sort, t_ranges, t_blend, t_total;
};

// GSCuda.cu:481-502
uint32_t gsr_oracle_get_higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb)
            msb += step;
        else
            msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

void gsr_oracle_free(GsrOracleOut* o) {
    free(o->keys_unsorted); o->keys_unsorted = nullptr;
    free(o->values_unsorted); o->values_unsorted = nullptr;
    free(o->keys); o->keys = nullptr;
    free(o->values); o->values = nullptr;
}

}  // extern "C"

namespace {

// ---------------------------------------------------------------------------------------
// MODE_CONTRACT preprocess — CudaRasterizer forward.cu preprocessCUDA as restated in
// SURVEY.md Appendix A; same step structure as GSCuda.cu:261-375.
// ---------------------------------------------------------------------------------------
void computeCov3D_contract(const float* s, float mod, const float* q, float* cov3D) {
    // S = diag(mod*s); R from (r,x,y,z) = (q[0],q[1],q[2],q[3]) WITHOUT normalising.
    M3 S = m3(mod * s[0], 0, 0, 0, mod * s[1], 0, 0, 0, mod * s[2]);
    float r = q[0], x = q[1], y = q[2], z = q[3];
    M3 R = m3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
              2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
              2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    M3 Mm = mul(S, R);
    M3 Sigma = mul(transpose(Mm), Mm);
    cov3D[0] = Sigma.c[0][0];
    cov3D[1] = Sigma.c[0][1];
    cov3D[2] = Sigma.c[0][2];
    cov3D[3] = Sigma.c[1][1];
    cov3D[4] = Sigma.c[1][2];
    cov3D[5] = Sigma.c[2][2];
}

void computeCov2D_contract(const V3& mean, float fx, float fy, float tan_fovx, float tan_fovy, const float* cov3D,
                           const float* v, float cov[3]) {
    V3 t = transformPoint4x3(mean, v);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
    t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
    M3 J = m3(fx / t.z, 0.0f, -(fx * t.x) / (t.z * t.z), 0.0f, fy / t.z, -(fy * t.y) / (t.z * t.z), 0, 0, 0);
    M3 Wm = m3(v[0], v[4], v[8], v[1], v[5], v[9], v[2], v[6], v[10]);
    M3 T = mul(Wm, J);
    M3 Vrk = m3(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    M3 c = mul(mul(transpose(T), transpose(Vrk)), T);
    c.c[0][0] += 0.3f;
    c.c[1][1] += 0.3f;
    cov[0] = c.c[0][0];
    cov[1] = c.c[0][1];
    cov[2] = c.c[1][1];
}

// upstream auxiliary.h ndc2Pix: double literals -> evaluated in double, rounded to float once.
inline float ndc2Pix(float v, int S) { return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5); }

void computeColorFromSH(int idx, int deg, int max_coeffs, const V3& pos, const float* campos, const float* shs,
                        uint8_t* clamped, float* rgb_out) {
    V3 dir = {pos.x - campos[0], pos.y - campos[1], pos.z - campos[2]};
    float len = sqrtf((dir.x * dir.x + dir.y * dir.y) + dir.z * dir.z);
    dir.x = dir.x / len;
    dir.y = dir.y / len;
    dir.z = dir.z / len;
    const float* sh = shs + (size_t)idx * max_coeffs * 3;
    float res[3];
    float x = dir.x, y = dir.y, z = dir.z;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    // Basis factors exactly as upstream forward.cu computeColorFromSH writes them (sign folded in).
    float bf[16] = {SH_C0,
                    -(SH_C1 * y),
                    SH_C1 * z,
                    -(SH_C1 * x),
                    SH_C2[0] * xy,
                    SH_C2[1] * yz,
                    SH_C2[2] * ((2.0f * zz - xx) - yy),
                    SH_C2[3] * xz,
                    SH_C2[4] * (xx - yy),
                    (SH_C3[0] * y) * (3.0f * xx - yy),
                    (SH_C3[1] * xy) * z,
                    (SH_C3[2] * y) * ((4.0f * zz - xx) - yy),
                    (SH_C3[3] * z) * ((2.0f * zz - 3.0f * xx) - 3.0f * yy),
                    (SH_C3[4] * x) * ((4.0f * zz - xx) - yy),
                    (SH_C3[5] * z) * (xx - yy),
                    (SH_C3[6] * x) * (xx - 3.0f * yy)};
    // Summation order.  Upstream writes one left-to-right chain per channel, which nvcc (default -fmad=true)
    // contracts into FMAs in an order of its own choosing, so no association is canonical for the contract;
    // the colours only feed the image (tolerance 1/255).  This restatement fixes the order the sm_100a kernel
    // uses: four groups of four coefficients, each an FMA chain, added pairwise:
    //   s_q = fma(b[4q+3],c[4q+3], fma(b[4q+2],c[4q+2], fma(b[4q+1],c[4q+1], b[4q]*c[4q])))
    //   rgb = ((s0 + s1) + (s2 + s3)) + 0.5            (coefficients beyond the active degree count as 0)
    // tests/test_oracle_kat.py (KAT-5) holds it to the FP64 real SH basis.
    const int ncoef = std::min(max_coeffs, (deg + 1) * (deg + 1));
    for (int k = 0; k < 16; ++k)
        if (k >= ncoef) bf[k] = 0.0f;
    for (int c = 0; c < 3; ++c) {
        auto S = [&](int k) { return k < ncoef ? sh[k * 3 + c] : 0.0f; };
        float part[4];
        for (int q = 0; q < 4; ++q) {
            float acc = bf[4 * q] * S(4 * q);
            acc = std::fmaf(bf[4 * q + 1], S(4 * q + 1), acc);
            acc = std::fmaf(bf[4 * q + 2], S(4 * q + 2), acc);
            acc = std::fmaf(bf[4 * q + 3], S(4 * q + 3), acc);
            part[q] = acc;
        }
        float result = (part[0] + part[1]) + (part[2] + part[3]);
        result += 0.5f;
        res[c] = result;
    }
    for (int c = 0; c < 3; ++c) {
        clamped[3 * idx + c] = (res[c] < 0);
        rgb_out[3 * idx + c] = fmaxf(res[c], 0.0f);
    }
}

void preprocess_contract(const GsrOracleIn& in, GsrOracleOut& o, int64_t b, int64_t e) {
    const int gx = (in.W + BLOCK_X - 1) / BLOCK_X, gy = (in.H + BLOCK_Y - 1) / BLOCK_Y;
    const float focal_y = in.H / (2.0f * in.tan_fovy);
    const float focal_x = in.W / (2.0f * in.tan_fovx);
    const float fmax_ = std::numeric_limits<float>::max();
    float bmin[3] = {-fmax_, -fmax_, -fmax_}, bmax[3] = {fmax_, fmax_, fmax_};
    if (in.boxmin) memcpy(bmin, in.boxmin, 12);
    if (in.boxmax) memcpy(bmax, in.boxmax, 12);
    for (int64_t idx = b; idx < e; ++idx) {
        o.radii[idx] = 0;
        o.tiles_touched[idx] = 0;
        const float* mp = in.means3D + idx * in.means_stride;
        V3 p_orig = {mp[0], mp[1], mp[2]};
        // in_frustum (auxiliary.h): near-plane test on view-space z only
        V3 p_view = transformPoint4x3(p_orig, in.viewmatrix);
        if (p_view.z <= 0.2f) continue;  // (prefiltered would __trap() upstream; see gsr_forward)
        float p_hom[4];
        transformPoint4x4(p_orig, in.projmatrix, p_hom);
        float p_w = 1.0f / (p_hom[3] + 0.0000001f);
        float p_proj[3] = {p_hom[0] * p_w, p_hom[1] * p_w, p_hom[2] * p_w};
        // SIBR bounding-box cull
        if (p_orig.x < bmin[0] || p_orig.y < bmin[1] || p_orig.z < bmin[2] || p_orig.x > bmax[0] ||
            p_orig.y > bmax[1] || p_orig.z > bmax[2])
            continue;
        const float* cov3D;
        if (in.cov3D_precomp) {
            cov3D = in.cov3D_precomp + idx * 6;
        } else {
            computeCov3D_contract(in.scales + idx * in.scales_stride, in.scale_modifier, in.rotations + idx * 4,
                                  o.cov3D + idx * 6);
            cov3D = o.cov3D + idx * 6;
        }
        float cov[3];
        computeCov2D_contract(p_orig, focal_x, focal_y, in.tan_fovx, in.tan_fovy, cov3D, in.viewmatrix, cov);
        float det = cov[0] * cov[2] - cov[1] * cov[1];
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float conic[3] = {cov[2] * det_inv, -cov[1] * det_inv, cov[0] * det_inv};
        float mid = 0.5f * (cov[0] + cov[2]);
        float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
        float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
        float px = ndc2Pix(p_proj[0], in.W), py = ndc2Pix(p_proj[1], in.H);
        uint32_t minx, miny, maxx, maxy;
        if (!in.use_rects) {
            int r = f2i(my_radius);
            getRect(px, py, r, r, gx, gy, minx, miny, maxx, maxy);
        } else {
            int rx = f2i(ceilf(3.f * sqrtf(cov[0]))), ry = f2i(ceilf(3.f * sqrtf(cov[2])));
            o.rects[2 * idx] = rx;
            o.rects[2 * idx + 1] = ry;
            getRect(px, py, rx, ry, gx, gy, minx, miny, maxx, maxy);
        }
        if ((maxx - minx) * (maxy - miny) == 0) continue;
        if (!in.colors_precomp)
            computeColorFromSH((int)idx, in.D, in.M, p_orig, in.cam_pos, in.shs, o.clamped, o.rgb);
        o.depths[idx] = p_view.z;
        o.radii[idx] = f2i(my_radius);
        o.means2D[2 * idx] = px;
        o.means2D[2 * idx + 1] = py;
        o.conic_opacity[4 * idx] = conic[0];
        o.conic_opacity[4 * idx + 1] = conic[1];
        o.conic_opacity[4 * idx + 2] = conic[2];
        o.conic_opacity[4 * idx + 3] = in.opacities[idx];
        o.tiles_touched[idx] = (maxy - miny) * (maxx - minx);
    }
}

// ---------------------------------------------------------------------------------------
// MODE_GSRAST preprocess — literal restatement of GSCuda.cu:157-375.
// ---------------------------------------------------------------------------------------
inline float glmMin(float a, float b) { return (b < a) ? b : a; }  // glm::min
inline float glmMax(float a, float b) { return (a < b) ? b : a; }  // glm::max

void computeCov3D_gsrast(const float* s, float mod, const float* rot, float* cov3D) {
    // GSCuda.cu:168-195.  glm::normalize(vec4) = v * (1/sqrt(dot)), dot pairwise.
    M3 S = m3(mod * s[0], 0, 0, 0, mod * s[1], 0, 0, 0, mod * s[2]);
    float d = (rot[0] * rot[0] + rot[1] * rot[1]) + (rot[2] * rot[2] + rot[3] * rot[3]);
    float inv = 1.0f / sqrtf(d);
    float qx = rot[0] * inv, qy = rot[1] * inv, qz = rot[2] * inv, qw = rot[3] * inv;
    // GSCuda.cu:157-162: double literals promote the outer multiply/subtract to double.
    auto D1 = [](float sum) { return (float)(2.0 * (double)sum - 1.0); };
    auto D2 = [](float v) { return (float)(2.0 * (double)v); };
    M3 R = m3(D1(qx * qx + qy * qy), D2(qy * qz + qx * qw), D2(qy * qw - qx * qz),  //
              D2(qy * qz - qx * qw), D1(qx * qx + qz * qz), D2(qz * qw + qx * qy),  //
              D2(qy * qw + qx * qz), D2(qz * qw - qx * qy), D1(qx * qx + qw * qw));
    M3 rs = mul(R, S);
    M3 sigma = mul(rs, transpose(rs));
    cov3D[0] = sigma.c[0][0];
    cov3D[1] = sigma.c[1][0];
    cov3D[2] = sigma.c[2][0];
    cov3D[3] = sigma.c[1][1];
    cov3D[4] = sigma.c[2][1];
    cov3D[5] = sigma.c[2][2];
}

void computeCov2D_gsrast(const float mean[3], float focal, float tan_fovx, float tan_fovy, const float* cov3D,
                         const float* v, float cov[3]) {
    // GSCuda.cu:197-231
    float m4[4] = {mean[0], mean[1], mean[2], 1.0f}, t4[4];
    glmMat4Vec4(v, m4, t4);
    float tx = t4[0], ty = t4[1], tz = t4[2];
    float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
    float txtz = tx / tz, tytz = ty / tz;
    tx = glmMin(limx, glmMax(-limx, txtz)) * tz;
    ty = glmMin(limy, glmMax(-limy, tytz)) * tz;
    M3 J = m3(focal / tz, 0.0f, (-focal * tx) / (tz * tz), 0.0f, focal / tz, (-focal * ty) / (tz * tz), 0, 0, 0);
    // mat3(transpose(mat4 view)): column c = (v[c], v[4+c], v[8+c])
    M3 Wm = m3(v[0], v[4], v[8], v[1], v[5], v[9], v[2], v[6], v[10]);
    M3 T = mul(Wm, J);
    M3 Vrk = m3(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    M3 c = mul(mul(transpose(T), Vrk), T);
    c.c[0][0] += 0.3f;
    c.c[1][1] += 0.3f;
    cov[0] = c.c[0][0];
    cov[1] = c.c[0][1];
    cov[2] = c.c[1][1];
}

void preprocess_gsrast(const GsrOracleIn& in, GsrOracleOut& o, int64_t b, int64_t e) {
    const int gx = (in.W + BLOCK_X - 1) / BLOCK_X, gy = (in.H + BLOCK_Y - 1) / BLOCK_Y;
    const float focal = in.H / (2.0f * in.tan_fovy);  // GSCuda.cu:721
    for (int64_t idx = b; idx < e; ++idx) {
        o.radii[idx] = 0;
        o.tiles_touched[idx] = 0;
        const float* mp = in.means3D + idx * in.means_stride;
        // GSCuda.cu:303 multiplies the stored vec4 (w comes from the buffer; 1 when stride is 3)
        float m4[4] = {mp[0], mp[1], mp[2], in.means_stride >= 4 ? mp[3] : 1.0f}, ph[4];
        glmMat4Vec4(in.projmatrix, m4, ph);
        float oneOverW = 1.0f / (0.001f + ph[3]);
        float prx = oneOverW * ph[0], pry = oneOverW * ph[1], prz = oneOverW * ph[2];
        if (prz < 0.0f || prz > 1.0f || prx < -1.3f || prx > 1.3f || pry < -1.3f || pry > 1.3f) continue;
        const float* cov3D;
        if (in.cov3D_precomp) {
            cov3D = in.cov3D_precomp + idx * 6;
        } else {
            computeCov3D_gsrast(in.scales + idx * in.scales_stride, in.scale_modifier, in.rotations + idx * 4,
                                o.cov3D + idx * 6);
            cov3D = o.cov3D + idx * 6;
        }
        float cov[3];
        computeCov2D_gsrast(mp, focal, in.tan_fovx, in.tan_fovy, cov3D, in.viewmatrix, cov);
        float det = cov[0] * cov[2] - cov[1] * cov[1];
        if (det == 0.0f) continue;
        float detInv = 1.0f / det;
        float conic[3] = {cov[2] * detInv, -cov[1] * detInv, cov[0] * detInv};
        float mid = 0.5f * (cov[0] + cov[2]);
        float lambda1 = mid + sqrtf(glmMax(0.1f, mid * mid - det));
        float lambda2 = mid - sqrtf(glmMax(0.1f, mid * mid - det));
        float myRadius = ceilf(3.0f * sqrtf(glmMax(lambda1, lambda2)));
        float px = (prx * 0.5f + 0.5f) * (float)in.W;
        float py = (pry * 0.5f + 0.5f) * (float)in.H;
        uint32_t minx, miny, maxx, maxy;
        if (!in.use_rects) {
            int r = f2i(myRadius);
            getRect(px, py, r, r, gx, gy, minx, miny, maxx, maxy);
        } else {
            // GSCuda.cu:352 — y extent has no sqrt (in-tree divergence, kept in this mode)
            int rx = f2i(ceilf(3.0f * sqrtf(cov[0]))), ry = f2i(ceilf(3.0f * cov[2]));
            o.rects[2 * idx] = rx;
            o.rects[2 * idx + 1] = ry;
            getRect(px, py, rx, ry, gx, gy, minx, miny, maxx, maxy);
        }
        if ((maxx - minx) * (maxy - miny) == 0) continue;
        if (!in.colors_precomp) {
            const float* sh = in.shs + (size_t)idx * 48;  // GSCuda.cu:364-365
            for (int c = 0; c < 3; ++c) o.rgb[3 * idx + c] = 0.5f + 0.4f * sh[c];
        }
        o.depths[idx] = prz;
        o.radii[idx] = f2i(myRadius);
        o.means2D[2 * idx] = px;
        o.means2D[2 * idx + 1] = py;
        o.conic_opacity[4 * idx] = conic[0];
        o.conic_opacity[4 * idx + 1] = conic[1];
        o.conic_opacity[4 * idx + 2] = conic[2];
        o.conic_opacity[4 * idx + 3] = in.opacities[idx];
        o.tiles_touched[idx] = (maxx - minx) * (maxy - miny);
    }
}

// Stable LSD radix sort of (u64 key, u32 value) over bits [0, end_bit) — the observable
// behaviour of cub::DeviceRadixSort::SortPairs at GSCuda.cu:794-797.  Result lands in
// (keys_out, vals_out); (keys_in, vals_in) are clobbered.
void radix_sort_pairs(uint64_t* keys_in, uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out, int64_t n,
                      int end_bit, int threads) {
    const int RB = 8, NB = 1 << RB;
    int passes = (end_bit + RB - 1) / RB;
    threads = std::max(1, threads);
    std::vector<int64_t> hist((size_t)threads * NB);
    uint64_t* ka = keys_in; uint32_t* va = vals_in;
    uint64_t* kb = keys_out; uint32_t* vb = vals_out;
    for (int p = 0; p < passes; ++p) {
        int shift = p * RB;
        uint64_t mask = (uint64_t)NB - 1;
        if (shift + RB > end_bit) mask = ((uint64_t)1 << (end_bit - shift)) - 1;
        std::fill(hist.begin(), hist.end(), 0);
        parallel_for(n, threads, [&](int t, int64_t b, int64_t e) {
            int64_t* h = &hist[(size_t)t * NB];
            for (int64_t i = b; i < e; ++i) h[(ka[i] >> shift) & mask]++;
        });
        int64_t run = 0;
        for (int d = 0; d < NB; ++d)
            for (int t = 0; t < threads; ++t) {
                int64_t c = hist[(size_t)t * NB + d];
                hist[(size_t)t * NB + d] = run;
                run += c;
            }
        parallel_for(n, threads, [&](int t, int64_t b, int64_t e) {
            int64_t* h = &hist[(size_t)t * NB];
            for (int64_t i = b; i < e; ++i) {
                int64_t dst = h[(ka[i] >> shift) & mask]++;
                kb[dst] = ka[i];
                vb[dst] = va[i];
            }
        });
        std::swap(ka, kb);
        std::swap(va, vb);
    }
    if (ka != keys_out) {  // even number of passes: result sits in the input buffers
        memcpy(keys_out, ka, (size_t)n * 8);
        memcpy(vals_out, va, (size_t)n * 4);
    }
}

}  // namespace

extern "C" {

// Stand-alone entry points for known-answer tests -------------------------------------
void gsr_oracle_get_rect(float px, float py, int ex, int ey, int gx, int gy, uint32_t* out4) {
    getRect(px, py, ex, ey, gx, gy, out4[0], out4[1], out4[2], out4[3]);
}

void gsr_oracle_sort_pairs(uint64_t* keys_in, uint32_t* vals_in, uint64_t* keys_out, uint32_t* vals_out, int64_t n,
                           int end_bit, int threads) {
    radix_sort_pairs(keys_in, vals_in, keys_out, vals_out, n, end_bit, threads);
}

// identifyTileRanges (GSCuda.cu:504-538).  compat: in-tree placement of the last-element
// close inside the else-branch (R==1 never closes); contract: tested unconditionally.
void gsr_oracle_identify_tile_ranges(int64_t R, const uint64_t* keys, uint32_t* ranges, int compat) {
    for (int64_t idx = 0; idx < R; ++idx) {
        uint32_t cur = (uint32_t)(keys[idx] >> 32);
        if (idx == 0) {
            ranges[2 * cur] = 0;
        } else {
            uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
            if (prev != cur) {
                ranges[2 * prev + 1] = (uint32_t)idx;
                ranges[2 * cur] = (uint32_t)idx;
            }
            if (compat && idx == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
        }
        if (!compat && idx == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
    }
}

// renderCUDA (GSCuda.cu:543-677) for tiles [tb, te) — one pixel at a time, splats in
// sorted order.  T-termination threshold: 0.0001f contract, 0.001f in-tree (GSCuda.cu:653).
static int64_t blend_tiles(const GsrOracleIn& in, GsrOracleOut& o, const float* colors, int tile, float t_min) {
    const int W = in.W, H = in.H;
    const int gx = (W + BLOCK_X - 1) / BLOCK_X;
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t rb = o.ranges[2 * tile], re = o.ranges[2 * tile + 1];
    const int n = (re > rb) ? (int)(re - rb) : 0;
    int64_t evaluated = 0;
    // stage the tile's splats contiguously (the role the shared-memory batches play)
    std::vector<float> st((size_t)n * 9);
    for (int j = 0; j < n; ++j) {
        uint32_t id = o.values[rb + j];
        float* s = &st[(size_t)j * 9];
        s[0] = o.means2D[2 * id]; s[1] = o.means2D[2 * id + 1];
        s[2] = o.conic_opacity[4 * id]; s[3] = o.conic_opacity[4 * id + 1];
        s[4] = o.conic_opacity[4 * id + 2]; s[5] = o.conic_opacity[4 * id + 3];
        s[6] = colors[3 * id]; s[7] = colors[3 * id + 1]; s[8] = colors[3 * id + 2];
    }
    for (int ly = 0; ly < BLOCK_Y; ++ly)
        for (int lx = 0; lx < BLOCK_X; ++lx) {
            int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
            if (pxi >= W || pyi >= H) continue;
            float pixx = (float)pxi, pixy = (float)pyi;
            float T = 1.0f, C[3] = {0, 0, 0};
            uint32_t contributor = 0, last = 0;
            for (int j = 0; j < n; ++j) {
                contributor++;
                const float* s = &st[(size_t)j * 9];
                float dx = s[0] - pixx, dy = s[1] - pixy;
                float power = -0.5f * ((s[2] * dx) * dx + (s[4] * dy) * dy) - (s[3] * dx) * dy;
                if (power > 0.0f) continue;
                // Exact shortcut, CPU speed only: for opacity <= 1, exp(power) < exp(-6) = 0.00248 < 1/255,
                // so the alpha test below would skip this pair anyway.
                if (power < -6.0f && s[5] <= 1.0f) continue;
                float alpha = fminf(0.99f, s[5] * expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                float test_T = T * (1.0f - alpha);
                if (test_T < t_min) break;  // done = true: nothing later is visited
                C[0] += (s[6] * alpha) * T;  // colour * alpha * T, left to right (GSCuda.cu:661)
                C[1] += (s[7] * alpha) * T;
                C[2] += (s[8] * alpha) * T;
                T = test_T;
                last = contributor;
            }
            evaluated += contributor;
            int pix = pyi * W + pxi;
            o.final_T[pix] = T;
            o.n_contrib[pix] = last;
            for (int c = 0; c < 3; ++c) o.out_color[(size_t)c * W * H + pix] = C[c] + T * in.background[c];
        }
    return evaluated;
}

// Full forward.  Returns num_rendered (>= 0) or a negative error.
int64_t gsr_oracle_forward(const GsrOracleIn* pin, GsrOracleOut* po) {
    const GsrOracleIn& in = *pin;
    GsrOracleOut& o = *po;
    const bool compat = in.flags & 1;
    const int threads = std::max(1, in.threads);
    const int P = in.P, W = in.W, H = in.H;
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const int T = gx * gy;
    o.keys_unsorted = o.keys = nullptr;
    o.values_unsorted = o.values = nullptr;
    o.num_rendered = 0;
    o.pairs_evaluated = 0;
    double t0 = now();

    // 1. preprocess (GSCuda.cu:744-768)
    parallel_for(P, threads, [&](int, int64_t b, int64_t e) {
        if (compat)
            preprocess_gsrast(in, o, b, e);
        else
            preprocess_contract(in, o, b, e);
    });
    double t1 = now();

    // 2. inclusive scan, u32 wrap-around like the device scan (GSCuda.cu:771)
    uint32_t run = 0;
    for (int i = 0; i < P; ++i) {
        run += o.tiles_touched[i];
        o.point_offsets[i] = run;
    }
    const int64_t R = P > 0 ? (int64_t)o.point_offsets[P - 1] : 0;
    o.num_rendered = R;
    double t2 = now();

    if (R == 0) {
        // GSCuda.cu:775-778 returns with out_color untouched (MODE_GSRAST); the contract renders
        // the background and clears the per-pixel state.
        if (!compat) {
            memset(o.ranges, 0, sizeof(uint32_t) * 2 * T);
            for (int i = 0; i < W * H; ++i) {
                o.final_T[i] = 1.0f;
                o.n_contrib[i] = 0;
                for (int c = 0; c < 3; ++c) o.out_color[(size_t)c * W * H + i] = 0.0f + 1.0f * in.background[c];
            }
        }
        o.t_preprocess = t1 - t0; o.t_scan = t2 - t1;
        o.t_duplicate = o.t_sort = o.t_ranges = o.t_blend = 0;
        o.t_total = now() - t0;
        return 0;
    }

    o.keys_unsorted = (uint64_t*)malloc((size_t)R * 8);
    o.keys = (uint64_t*)malloc((size_t)R * 8);
    o.values_unsorted = (uint32_t*)malloc((size_t)R * 4);
    o.values = (uint32_t*)malloc((size_t)R * 4);
    if (!o.keys_unsorted || !o.keys || !o.values_unsorted || !o.values) return -2;

    // 3. duplicateWithKeys (GSCuda.cu:422-475)
    parallel_for(P, threads, [&](int, int64_t b, int64_t e) {
        for (int64_t idx = b; idx < e; ++idx) {
            if (o.radii[idx] <= 0) continue;
            uint32_t off = (idx == 0) ? 0 : o.point_offsets[idx - 1];
            uint32_t minx, miny, maxx, maxy;
            if (!in.use_rects)
                getRect(o.means2D[2 * idx], o.means2D[2 * idx + 1], o.radii[idx], o.radii[idx], gx, gy, minx, miny,
                        maxx, maxy);
            else
                getRect(o.means2D[2 * idx], o.means2D[2 * idx + 1], o.rects[2 * idx], o.rects[2 * idx + 1], gx, gy,
                        minx, miny, maxx, maxy);
            uint32_t dbits;
            memcpy(&dbits, &o.depths[idx], 4);
            for (uint32_t y = miny; y < maxy; ++y)
                for (uint32_t x = minx; x < maxx; ++x) {
                    uint64_t key = (uint64_t)(y * (uint32_t)gx + x);
                    key <<= 32;
                    key |= dbits;
                    o.keys_unsorted[off] = key;
                    o.values_unsorted[off] = (uint32_t)idx;
                    off++;
                }
        }
    });
    const double t_dup_end = now();

    // 4. sort (GSCuda.cu:791-797).  The sort clobbers its input, so sort a scratch copy and keep
    //    keys_unsorted / values_unsorted observable.
    double t3;
    {
        int bit = (int)gsr_oracle_get_higher_msb((uint32_t)T);
        uint64_t* ktmp = (uint64_t*)malloc((size_t)R * 8);
        uint32_t* vtmp = (uint32_t*)malloc((size_t)R * 4);
        if (!ktmp || !vtmp) return -2;
        memcpy(ktmp, o.keys_unsorted, (size_t)R * 8);
        memcpy(vtmp, o.values_unsorted, (size_t)R * 4);
        t3 = now();  // scratch copies are not part of the sort
        radix_sort_pairs(ktmp, vtmp, o.keys, o.values, R, 32 + bit, threads);
        free(ktmp);
        free(vtmp);
    }
    double t4 = now();

    // 5. ranges (GSCuda.cu:800-801)
    memset(o.ranges, 0, sizeof(uint32_t) * 2 * T);
    gsr_oracle_identify_tile_ranges(R, o.keys, o.ranges, compat ? 1 : 0);
    double t5 = now();

    // 6. blend (GSCuda.cu:803-810)
    const float* colors = in.colors_precomp ? in.colors_precomp : o.rgb;
    const float t_min = compat ? 0.001f : 0.0001f;
    std::atomic<int> next(0);
    std::atomic<int64_t> evaluated(0);
    parallel_for(threads, threads, [&](int, int64_t, int64_t) {
        int64_t ev = 0;
        for (;;) {
            int tile = next.fetch_add(1);
            if (tile >= T) break;
            ev += blend_tiles(in, o, colors, tile, t_min);
        }
        evaluated += ev;
    });
    double t6 = now();
    o.pairs_evaluated = evaluated;
    o.t_preprocess = t1 - t0; o.t_scan = t2 - t1; o.t_duplicate = t_dup_end - t2; o.t_sort = t4 - t3;
    o.t_ranges = t5 - t4; o.t_blend = t6 - t5; o.t_total = t6 - t0;
    return R;
}

int gsr_oracle_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
