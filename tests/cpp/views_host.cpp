// views_host.cpp — a C++ host that renders a batch of camera views of a .ply scene the way a GSRast-side tool would,
// entirely through the C ABI (include/gsrast_b200.h): the native .ply staging (gsr_ply_count / gsr_ply_load, the
// SplatData replacement — /root/reference/apps/gsrast/SplatData.cpp:114-156, 48-58), the resident-scene renderer
// (gsr_renderer_create, the GSGaussians replacement — GSGaussians.cpp:44-153) and the 8-bit frame delivery
// (gsr_renderer_render_host_u8).  tests/test_gpu_cpp_dropin.py checks its frames against the ctypes path.
//
//   views_host <scene.ply> <cameras.bin> <out.bin>
// cameras.bin: int32 n, W, H ; float tanx, tany ; float cam[n][36] (view[16] proj[16] cam_pos[3] pad)
// out.bin:     int32 P, n ; int32 num_rendered[n] ; uint8 frames[n][3][H][W]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include <gsrast_b200.h>

#define CK(x)                                                                     \
    do {                                                                          \
        int rc_ = (x);                                                            \
        if (rc_ < 0) {                                                            \
            fprintf(stderr, "%s -> %d (%s)\n", #x, rc_, gsr_error_string(rc_));   \
            return 1;                                                             \
        }                                                                         \
    } while (0)

template <typename T>
static T* upload(const std::vector<T>& h) {
    T* d = nullptr;
    cudaMalloc(&d, h.size() * sizeof(T) + 16);
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    int P = 0;
    CK(gsr_ply_count(argv[1], &P));
    std::vector<float> means((size_t)P * 3), scales((size_t)P * 3), rot((size_t)P * 4), opac(P), shs((size_t)P * 48);
    float bbox[6], center[3];
    CK(gsr_ply_load(argv[1], P, means.data(), scales.data(), rot.data(), opac.data(), shs.data(), bbox, center));

    FILE* f = fopen(argv[2], "rb");
    if (!f) return 2;
    int hdr[3];
    float tan[2];
    if (fread(hdr, sizeof(int), 3, f) != 3 || fread(tan, sizeof(float), 2, f) != 2) return 2;
    const int n = hdr[0], W = hdr[1], H = hdr[2];
    std::vector<float> cams((size_t)n * 36);
    if (fread(cams.data(), sizeof(float), cams.size(), f) != cams.size()) return 2;
    fclose(f);

    const std::vector<float> bg = {0.f, 0.f, 0.f};
    float *d_means = upload(means), *d_scales = upload(scales), *d_rot = upload(rot), *d_opac = upload(opac),
          *d_shs = upload(shs), *d_bg = upload(bg);
    void* r = gsr_renderer_create(P, 3, 16, d_means, d_shs, nullptr, d_opac, d_scales, d_rot, d_bg, 1.0f, W, H, nullptr, 0);
    if (!r) { fprintf(stderr, "gsr_renderer_create failed\n"); return 1; }
    const size_t frame = (size_t)3 * W * H;
    unsigned char* frames = nullptr;
    cudaHostAlloc(&frames, frame * n, cudaHostAllocDefault);  // pinned: the copies overlap the next view's render
    std::vector<int> nr(n);
    CK(gsr_renderer_render_host_u8(r, cams.data(), n, tan[0], tan[1], frames, nr.data()));
    gsr_renderer_destroy(r);

    FILE* o = fopen(argv[3], "wb");
    if (!o) return 2;
    const int head[2] = {P, n};
    fwrite(head, sizeof(int), 2, o);
    fwrite(nr.data(), sizeof(int), n, o);
    fwrite(frames, 1, frame * n, o);
    fclose(o);
    printf("views_host: P=%d views=%d first num_rendered=%d bbox x [%g, %g]\n", P, n, n ? nr[0] : 0, bbox[0], bbox[3]);
    cudaFreeHost(frames);
    return 0;
}
