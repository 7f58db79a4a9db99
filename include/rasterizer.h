// rasterizer.h — header-only C++ shim: CudaRasterizer::Rasterizer::forward on top of the C ABI
// of libgsrast_b200.so.
//
// This is the include GSRast's splat draw path names for the non-gscuda branch
// (apps/gsrast/GSGaussians.cpp:20-22: `#include <rasterizer.h>`,
// `#define FORWARD CudaRasterizer::Rasterizer::forward`).  The argument list is the one used at
// the call site GSGaussians.cpp:179-206.  Each std::function<char*(size_t)> is trampolined
// through the (function pointer, user pointer) pair the C ABI takes.
#pragma once

#include <cstddef>
#include <functional>

#include "gsrast_b200.h"

namespace CudaRasterizer {

class Rasterizer {
    using BufferFn = std::function<char*(size_t)>;

    static char* trampoline(size_t bytes, void* user) {
        try {
            return (*static_cast<BufferFn*>(user))(bytes);
        } catch (...) {
            return nullptr;  // never unwind through the C boundary
        }
    }

public:
    // Returns num_rendered (>= 0) or a negative gsrast_b200 / CUDA error code.
    static int forward(BufferFn geometryBuffer, BufferFn binningBuffer, BufferFn imageBuffer, const int P, int D,
                       int M, const float* background, const int width, int height, const float* means3D,
                       const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
                       const float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const float* viewmatrix, const float* projmatrix, const float* cam_pos, const float tan_fovx,
                       float tan_fovy, const bool prefiltered, float* out_color, int* radii = nullptr,
                       int* rects = nullptr, float* boxmin = nullptr, float* boxmax = nullptr,
                       void* stream = nullptr) {
        return gsr_forward(&trampoline, &geometryBuffer, &trampoline, &binningBuffer, &trampoline, &imageBuffer, P, D,
                           M, background, width, height, means3D, shs, colors_precomp, opacities, scales,
                           scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx,
                           tan_fovy, prefiltered ? 1 : 0, out_color, radii, rects, boxmin, boxmax, stream);
    }
};

}  // namespace CudaRasterizer
