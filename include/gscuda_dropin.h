// gscuda_dropin.h — header-only C++ shim giving GSRast its own entry point back:
// gscuda::forward (apps/gsrast/gscuda/GSCuda.cuh:103-126, defined at GSCuda.cu:695-811) with the
// semantics and buffer layouts the viewer uses today (vec4 positions/scales, raw-PLY SH block,
// in-tree cull / depth / colour rules), executed by libgsrast_b200.so.
//
// Use: in apps/gsrast/GSGaussians.cpp replace `#include <GSCuda.cuh>` by this header and link
// gsrast_b200 instead of the gscuda static library.  mapGeometryState (GSGaussians.cpp:214-219)
// maps onto gscuda::gs::GeometryState::fromChunk below.
#pragma once

#include <cstddef>
#include <cstdint>
#include <functional>

#include "gsrast_b200.h"

namespace gscuda {

namespace detail {
using BufferFn = std::function<char*(size_t)>;
inline char* trampoline(size_t bytes, void* user) {
    try {
        return (*static_cast<BufferFn*>(user))(bytes);
    } catch (...) {
        return nullptr;
    }
}
}  // namespace detail

inline int forward(detail::BufferFn geometryBuffer, detail::BufferFn binningBuffer, detail::BufferFn imageBuffer,
                   int numGaussians, int shDims, int M, const float* background, int width, int height,
                   const float* means3D, const float* shs, const float* colorsPrecomp, const float* opacities,
                   const float* scales, float scaleModifier, const float* rotations, const float* cov3DPrecomp,
                   const float* viewMatrix, const float* projMatrix, const float* camPos, float tanFOVx, float tanFOVy,
                   bool prefiltered, float* outColor, int* radii, int* rects, float* boxMin, float* boxMax) {
    return gsr_forward_gscuda(&detail::trampoline, &geometryBuffer, &detail::trampoline, &binningBuffer,
                              &detail::trampoline, &imageBuffer, numGaussians, shDims, M, background, width, height,
                              means3D, shs, colorsPrecomp, opacities, scales, scaleModifier, rotations, cov3DPrecomp,
                              viewMatrix, projMatrix, camPos, tanFOVx, tanFOVy, prefiltered ? 1 : 0, outColor, radii,
                              rects, boxMin, boxMax, nullptr);
}

template <typename T>
size_t required(int num);

namespace gs {

// Field access for the Inspector (apps/gsrast/Inspector.cpp:174-188).  Same member names as
// AuxBuffer.cuh:38-54; element types are plain floats/ints (glm::vec2/vec3/vec4 are layout-compatible).
struct GeometryState {
    uint32_t* tilesTouched;
    float* depths;
    bool* clamped;
    int* internalRadii;
    float* means2D;       // vec2[P]
    float* cov3D;         // float[6P]
    float* conicOpacity;  // vec4[P]
    float* rgb;           // vec3[P]
    uint32_t* pointOffsets;

    static GeometryState fromChunk(char*& chunk, int numGaussians) {
        gsr_geometry_state g;
        size_t used = gsr_geometry_state_map(chunk, numGaussians, &g);
        GeometryState s;
        s.tilesTouched = g.tiles_touched;
        s.depths = g.depths;
        s.clamped = reinterpret_cast<bool*>(g.clamped);
        s.internalRadii = g.internal_radii;
        s.means2D = g.means2D;
        s.cov3D = g.cov3D;
        s.conicOpacity = g.conic_opacity;
        s.rgb = g.rgb;
        s.pointOffsets = g.point_offsets;
        chunk += used;
        return s;
    }
};

}  // namespace gs

template <>
inline size_t required<gs::GeometryState>(int num) {
    return gsr_geometry_state_required(num);
}

}  // namespace gscuda
