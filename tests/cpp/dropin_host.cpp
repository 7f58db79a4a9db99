// dropin_host.cpp — a GL-free stand-in for GSRast's host side of the splat draw path, used by
// tests/test_gpu_cpp_dropin.py to prove the C++ boundary: it owns the device buffers and the
// three grow-only scratch allocators exactly like GSGaussians does
// (/root/reference/apps/gsrast/GSGaussians.cpp:27-42 resizeFunctional, :109-153 uploads,
// :155-212 draw) and calls FORWARD through the header-only shims in include/.
//
//   dropin_host <scene.bin> <out.bin> <mode>     mode: contract | gscuda
// scene.bin: int32 P,W,H,D,M ; float tanx,tany ; float bg[3] view[16] proj[16] campos[3] ;
//            then means, scales, rotations, opacities, shs as float arrays in the mode's layout.
// out.bin:   int32 num_rendered, int32 calls[3] ; float out_color[3*H*W]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include <gscuda_dropin.h>
#include <rasterizer.h>

static int g_calls[3] = {0, 0, 0};

static std::function<char*(size_t)> resizeFunctional(void** ptr, size_t& S, int which) {
    return [ptr, &S, which](size_t N) {
        g_calls[which]++;
        if (N > S) {
            if (*ptr) cudaFree(*ptr);
            cudaMalloc(ptr, 2 * N);
            S = 2 * N;
        }
        return reinterpret_cast<char*>(*ptr);
    };
}

template <typename T>
static T* upload(const std::vector<T>& h) {
    T* d = nullptr;
    cudaMalloc(&d, h.size() * sizeof(T) + 16);
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

template <typename T>
static std::vector<T> rd(FILE* f, size_t n) {
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) {
        fprintf(stderr, "short read\n");
        exit(2);
    }
    return v;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const bool gsc = strcmp(argv[3], "gscuda") == 0;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    auto hdr = rd<int>(f, 5);
    const int P = hdr[0], W = hdr[1], H = hdr[2], D = hdr[3], M = hdr[4];
    auto tans = rd<float>(f, 2);
    auto bg = rd<float>(f, 3);
    auto view = rd<float>(f, 16);
    auto proj = rd<float>(f, 16);
    auto campos = rd<float>(f, 3);
    const int ms = gsc ? 4 : 3;
    auto means = rd<float>(f, (size_t)P * ms);
    auto scales = rd<float>(f, (size_t)P * ms);
    auto rots = rd<float>(f, (size_t)P * 4);
    auto opac = rd<float>(f, (size_t)P);
    auto shs = rd<float>(f, (size_t)P * 48);
    fclose(f);

    float *d_bg = upload(bg), *d_view = upload(view), *d_proj = upload(proj), *d_cam = upload(campos);
    float *d_means = upload(means), *d_scales = upload(scales), *d_rots = upload(rots), *d_opac = upload(opac),
          *d_shs = upload(shs);
    float* d_out = nullptr;
    cudaMalloc(&d_out, sizeof(float) * 3 * W * H);
    cudaMemset(d_out, 0, sizeof(float) * 3 * W * H);
    int* d_rects = nullptr;
    cudaMalloc(&d_rects, sizeof(int) * 2 * (P + 1));

    void *geomPtr = nullptr, *binningPtr = nullptr, *imgPtr = nullptr;
    size_t allocdGeom = 0, allocdBinning = 0, allocdImg = 0;
    auto geomFunc = resizeFunctional(&geomPtr, allocdGeom, 0);
    auto binningFunc = resizeFunctional(&binningPtr, allocdBinning, 1);
    auto imgFunc = resizeFunctional(&imgPtr, allocdImg, 2);

    int R = -1;
    for (int frame = 0; frame < 2; ++frame) {  // second frame: allocators must not grow
        if (gsc)
            R = gscuda::forward(geomFunc, binningFunc, imgFunc, P, 3, 16, d_bg, W, H, d_means, d_shs, nullptr, d_opac,
                                d_scales, 1.0f, d_rots, nullptr, d_view, d_proj, d_cam, tans[0], tans[1], false, d_out,
                                nullptr, d_rects, nullptr, nullptr);
        else
            R = CudaRasterizer::Rasterizer::forward(geomFunc, binningFunc, imgFunc, P, D, M, d_bg, W, H, d_means, d_shs,
                                                    nullptr, d_opac, d_scales, 1.0f, d_rots, nullptr, d_view, d_proj,
                                                    d_cam, tans[0], tans[1], false, d_out, nullptr, d_rects);
        // CHECK_CUDA_ERROR of the reference (CudaBuffer.hpp:8-12)
        cudaDeviceSynchronize();
        if (cudaPeekAtLastError() != cudaSuccess) {
            fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
            return 3;
        }
        if (R < 0) {
            fprintf(stderr, "forward failed: %s\n", gsr_error_string(R));
            return 4;
        }
    }
    // the Inspector's path: recover the geometry fields from the raw chunk (GSGaussians.cpp:214-219)
    char* chunk = reinterpret_cast<char*>(geomPtr);
    gscuda::gs::GeometryState gs = gscuda::gs::GeometryState::fromChunk(chunk, P);
    int radius0 = 0;
    if (P > 0) cudaMemcpy(&radius0, gs.internalRadii, sizeof(int), cudaMemcpyDeviceToHost);

    std::vector<float> out((size_t)3 * W * H);
    cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 2;
    int head[5] = {R, g_calls[0], g_calls[1], g_calls[2], radius0};
    fwrite(head, sizeof(int), 5, o);
    fwrite(out.data(), sizeof(float), out.size(), o);
    fclose(o);
    printf("dropin_host: mode=%s P=%d R=%d calls=%d/%d/%d\n", argv[3], P, R, g_calls[0], g_calls[1], g_calls[2]);
    return 0;
}
