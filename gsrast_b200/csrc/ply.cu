// ply.cu — host-side scene staging: the 62-float "RichPoint" .ply the viewer loads, straight into the
// contract layout the rasterizer consumes.  No device code; lives in this library so a host in any language
// gets the loader through the same C ABI.
//
// Mirrors SplatData (apps/gsrast/SplatData.{hpp,cpp} of the reference):
//   record layout   position[3] normal[3] shs[48] opacity scale[3] rotation[4]   SplatData.hpp:17-25
//   header          the vertex count is the third token of the THIRD header line; the body starts after the
//                   line "end_header"                                             SplatData.cpp:126-145
//   short body      "Reader is EOF?" -> invalid                                   SplatData.cpp:146-152
//   activation      scale = exp(s); rotation = normalize(q); opacity = sigmoid(o) SplatData.cpp:8-11,48-58
// and closes the layout gap the viewer leaves open for SH degree > 0: the file stores
// f_dc[3] then f_rest[c*15 + k-1]; the rasterizer contract wants sh[k][c].
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "gsr_common.cuh"

namespace {

constexpr int PLY_FLOATS = 62;

// reads one header line (without the terminator); false at end of file
bool read_line(FILE* f, std::string& out) {
    out.clear();
    int c;
    bool any = false;
    while ((c = fgetc(f)) != EOF) {
        any = true;
        if (c == '\n') break;
        out.push_back((char)c);
    }
    if (!out.empty() && out.back() == '\r') out.pop_back();
    return any;
}

// header -> vertex count; leaves the file positioned at the first record
int parse_header(FILE* f, long long* count) {
    std::string line;
    for (int i = 0; i < 3; ++i)
        if (!read_line(f, line)) return GSR_ERR_PLY_FORMAT;
    char a[64], b[64];
    long long n = -1;
    if (sscanf(line.c_str(), "%63s %63s %lld", a, b, &n) != 3 || n < 0) return GSR_ERR_PLY_FORMAT;
    for (;;) {
        if (!read_line(f, line)) return GSR_ERR_PLY_FORMAT;  // no end_header
        if (line == "end_header") break;
    }
    *count = n;
    return 0;
}

inline float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }  // SplatData.cpp:8-11

}  // namespace

extern "C" {

int gsr_ply_count(const char* path, int* num_gaussians) {
    if (!path || !num_gaussians) return GSR_ERR_INVALID_ARG;
    FILE* f = fopen(path, "rb");
    if (!f) return GSR_ERR_PLY_OPEN;
    long long n = 0;
    int rc = parse_header(f, &n);
    fclose(f);
    if (rc < 0) return rc;
    if (n > 0x7fffffffLL) return GSR_ERR_PLY_FORMAT;
    *num_gaussians = (int)n;
    return 0;
}

int gsr_ply_load(const char* path, int capacity, float* means3D, float* scales, float* rotations, float* opacities,
                 float* shs, float* bbox_min_max, float* center) {
    if (!path || capacity < 0) return GSR_ERR_INVALID_ARG;
    FILE* f = fopen(path, "rb");
    if (!f) return GSR_ERR_PLY_OPEN;
    long long n = 0;
    int rc = parse_header(f, &n);
    if (rc < 0) { fclose(f); return rc; }
    if (n > (long long)capacity) { fclose(f); return GSR_ERR_INVALID_ARG; }
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    double sum[3] = {0.0, 0.0, 0.0};
    constexpr size_t CHUNK = 16384;  // records per read: 4 MB
    std::vector<float> buf(CHUNK * PLY_FLOATS);
    for (long long base = 0; base < n; base += (long long)CHUNK) {
        const size_t want = (size_t)((n - base < (long long)CHUNK) ? (n - base) : (long long)CHUNK);
        if (fread(buf.data(), sizeof(float) * PLY_FLOATS, want, f) != want) {
            fclose(f);
            return GSR_ERR_PLY_TRUNCATED;  // "Reader is EOF?"
        }
        for (size_t j = 0; j < want; ++j) {
            const float* r = buf.data() + j * PLY_FLOATS;
            const size_t i = (size_t)base + j;
            for (int c = 0; c < 3; ++c) {
                lo[c] = fminf(lo[c], r[c]);
                hi[c] = fmaxf(hi[c], r[c]);
                sum[c] += r[c];
            }
            if (means3D) { means3D[3 * i] = r[0]; means3D[3 * i + 1] = r[1]; means3D[3 * i + 2] = r[2]; }
            if (shs) {
                float* o = shs + i * 48;
                const float* s = r + 6;
                o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
                for (int k = 1; k < 16; ++k)
                    for (int c = 0; c < 3; ++c) o[3 * k + c] = s[3 + c * 15 + (k - 1)];
            }
            if (opacities) opacities[i] = sigmoidf(r[54]);
            if (scales) { scales[3 * i] = expf(r[55]); scales[3 * i + 1] = expf(r[56]); scales[3 * i + 2] = expf(r[57]); }
            if (rotations) {
                const float* q = r + 58;
                const float inv = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);  // glm::normalize
                for (int c = 0; c < 4; ++c) rotations[4 * i + c] = q[c] * inv;
            }
        }
    }
    fclose(f);
    if (bbox_min_max && n > 0)
        for (int c = 0; c < 3; ++c) { bbox_min_max[c] = lo[c]; bbox_min_max[3 + c] = hi[c]; }
    if (center)  // SplatData.cpp: mean of the positions
        for (int c = 0; c < 3; ++c) center[c] = n > 0 ? (float)(sum[c] / (double)n) : 0.0f;
    return (int)n;
}

}  // extern "C"
