// interop_host.cpp — the hand-off INTEGRATION.md option C promises, driven the way GSGaussians::draw would drive it
// (/root/reference/apps/gsrast/GSGaussians.cpp:155-212 with the mapped GL buffer of CudaBuffer.cpp:64-77,100-112):
//   * the caller owns a NON-default stream and the output buffer (stand-in for the mapped interop buffer);
//   * work the caller queued on its stream BEFORE the render (here: a long chain of memsets that ends by filling the
//     output buffer with 0xff) must be finished before any frame is written;
//   * work the caller queues AFTER the render call returns (a D2H copy on the same stream) must see finished frames —
//     without any cudaDeviceSynchronize / cudaStreamSynchronize between the two (the reference syncs the whole device
//     after every forward call, CudaBuffer.hpp:8-12).
// The frames are compared bit for bit with the same views rendered through the synchronous host-delivery call.
// Also reads the Inspector's fields through gsr_renderer_map_geometry_state on a GSR_FLAG_KEEP_STATE renderer
// (GSGaussians.cpp:214-219, Inspector.cpp:174-188): sum(tiles_touched) of the lane's last view == its num_rendered.
//
//   interop_host <scene.ply> <cameras.bin>
// cameras.bin: int32 n, W, H ; float tanx, tany ; float cam[n][36]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <gsrast_b200.h>

#define CK(x)                                                                   \
    do {                                                                        \
        int rc_ = (x);                                                          \
        if (rc_ < 0) {                                                          \
            fprintf(stderr, "%s -> %d (%s)\n", #x, rc_, gsr_error_string(rc_)); \
            return 1;                                                           \
        }                                                                       \
    } while (0)
#define CU(x)                                                                        \
    do {                                                                             \
        cudaError_t e_ = (x);                                                        \
        if (e_ != cudaSuccess) {                                                     \
            fprintf(stderr, "%s -> %s\n", #x, cudaGetErrorString(e_));               \
            return 1;                                                                \
        }                                                                            \
    } while (0)

template <typename T>
static T* upload(const std::vector<T>& h) {
    T* d = nullptr;
    cudaMalloc(&d, h.size() * sizeof(T) + 16);
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    int P = 0;
    CK(gsr_ply_count(argv[1], &P));
    std::vector<float> means((size_t)P * 3), scales((size_t)P * 3), rot((size_t)P * 4), opac(P), shs((size_t)P * 48);
    CK(gsr_ply_load(argv[1], P, means.data(), scales.data(), rot.data(), opac.data(), shs.data(), nullptr, nullptr));
    FILE* f = fopen(argv[2], "rb");
    if (!f) return 2;
    int hdr[3];
    float tan[2];
    if (fread(hdr, sizeof(int), 3, f) != 3 || fread(tan, sizeof(float), 2, f) != 2) return 2;
    const int n = hdr[0], W = hdr[1], H = hdr[2];
    std::vector<float> cams((size_t)n * 36);
    if (fread(cams.data(), sizeof(float), cams.size(), f) != cams.size()) return 2;
    fclose(f);

    const std::vector<float> bg = {0.1f, 0.2f, 0.3f};
    float *d_means = upload(means), *d_scales = upload(scales), *d_rot = upload(rot), *d_opac = upload(opac),
          *d_shs = upload(shs), *d_bg = upload(bg);
    const size_t frame = (size_t)3 * W * H;

    // reference frames: the synchronous host-delivery call on its own renderer
    std::vector<float> want(frame * n);
    std::vector<int> nr_want(n);
    {
        void* r0 = gsr_renderer_create(P, 3, 16, d_means, d_shs, nullptr, d_opac, d_scales, d_rot, d_bg, 1.0f, W, H, nullptr, 0);
        if (!r0) return 1;
        CK(gsr_renderer_render_host(r0, cams.data(), n, tan[0], tan[1], want.data(), nr_want.data()));
        gsr_renderer_destroy(r0);
    }

    cudaStream_t us;
    CU(cudaStreamCreateWithFlags(&us, cudaStreamNonBlocking));
    float* ext = nullptr;  // externally owned output (the viewer's mapped interop buffer)
    CU(cudaMalloc(&ext, frame * n * sizeof(float)));
    char* busy = nullptr;
    const size_t busy_bytes = (size_t)1 << 30;
    CU(cudaMalloc(&busy, busy_bytes));
    float* got = nullptr;
    CU(cudaHostAlloc(&got, frame * n * sizeof(float), cudaHostAllocDefault));
    void* r = gsr_renderer_create(P, 3, 16, d_means, d_shs, nullptr, d_opac, d_scales, d_rot, d_bg, 1.0f, W, H, us,
                                  GSR_FLAG_KEEP_STATE);
    if (!r) return 1;
    std::vector<int> nr(n);
    int bad_rounds = 0;
    for (int round = 0; round < 3; ++round) {
        // caller's earlier work on ITS stream: ~10 ms of memsets, the last one trashes the output buffer
        for (int i = 0; i < 24; ++i) CU(cudaMemsetAsync(busy, i, busy_bytes, us));
        CU(cudaMemsetAsync(ext, 0xff, frame * n * sizeof(float), us));
        CK(gsr_renderer_render(r, cams.data(), n, tan[0], tan[1], ext, nr.data(), nullptr));
        // caller's later work, same stream, no synchronisation in between
        CU(cudaMemcpyAsync(got, ext, frame * n * sizeof(float), cudaMemcpyDeviceToHost, us));
        CU(cudaStreamSynchronize(us));
        if (memcmp(got, want.data(), frame * n * sizeof(float)) != 0) ++bad_rounds;
        for (int v = 0; v < n; ++v)
            if (nr[v] != nr_want[v]) ++bad_rounds;
    }

    // Inspector accessor over the renderer's private scratch
    const int lanes = gsr_renderer_num_lanes();
    int last_on_lane0 = -1;
    for (int v = 0; v < n; ++v)
        if (v % lanes == 0) last_on_lane0 = v;
    gsr_geometry_state g;
    CK(gsr_renderer_map_geometry_state(r, 0, &g));
    std::vector<uint32_t> tt(P), off(P);
    std::vector<int> radii(P);
    CU(cudaMemcpy(tt.data(), g.tiles_touched, (size_t)P * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(off.data(), g.point_offsets, (size_t)P * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(radii.data(), g.internal_radii, (size_t)P * 4, cudaMemcpyDeviceToHost));
    unsigned long long sum = 0;
    int vis = 0, scan_bad = 0;
    for (int i = 0; i < P; ++i) {
        sum += tt[i];
        vis += radii[i] > 0;
        if (off[i] != (uint32_t)sum) ++scan_bad;
        if ((tt[i] > 0) != (radii[i] > 0)) ++scan_bad;
    }
    const bool state_ok = last_on_lane0 >= 0 && sum == (unsigned long long)nr[last_on_lane0] && scan_bad == 0 && vis > 0;
    gsr_renderer_destroy(r);
    printf("interop_host: views=%d bad_rounds=%d state_ok=%d visible=%d sum_tiles=%llu num_rendered=%d\n", n, bad_rounds,
           (int)state_ok, vis, sum, last_on_lane0 >= 0 ? nr[last_on_lane0] : -1);
    return (bad_rounds == 0 && state_ok) ? 0 : 1;
}
