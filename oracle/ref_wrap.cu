// ref_wrap.cu — C entry point around the reference's OWN gscuda::forward, compiled from
// /root/reference/apps/gsrast/gscuda/*.cu where those files lie (oracle/Makefile, target `ref`).
// TEST INFRASTRUCTURE ONLY: gives tests/ and bench.py an executable copy of the in-tree
// rasterizer (GSCuda.cu:695-811) for GSRast-mode parity fixtures and the "reference built for
// sm_100" GPU baseline.  Nothing in the product links this.
#include <GSCuda.cuh>
#include <AuxBuffer.cuh>

#include <cstddef>
#include <functional>

extern "C" {

typedef char* (*ref_alloc_fn)(size_t bytes, void* user);

void gscuda_ref_forward(ref_alloc_fn ga, void* gu, ref_alloc_fn ba, void* bu, ref_alloc_fn ia, void* iu, int P, int D,
                        int M, const float* background, int width, int height, const float* means3D,
                        const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
                        float scale_modifier, const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                        float tan_fovy, int prefiltered, float* out_color, int* radii, int* rects, float* boxmin,
                        float* boxmax) {
    std::function<char*(size_t)> g = [=](size_t n) { return ga(n, gu); };
    std::function<char*(size_t)> b = [=](size_t n) { return ba(n, bu); };
    std::function<char*(size_t)> i = [=](size_t n) { return ia(n, iu); };
    gscuda::forward(g, b, i, P, D, M, background, width, height, means3D, shs, colors_precomp, opacities, scales,
                    scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy,
                    prefiltered != 0, out_color, radii, rects, boxmin, boxmax);
}

// field offsets of the reference's own chunk layouts, so tests can read its scratch buffers
struct gscuda_ref_geometry { size_t tilesTouched, depths, clamped, internalRadii, means2D, cov3D, conicOpacity, rgb, pointOffsets, total; };
struct gscuda_ref_binning { size_t keysUnsorted, keys, valuesUnsorted, values, total; };
struct gscuda_ref_image { size_t ranges, nContrib, accumAlpha, total; };

void gscuda_ref_geometry_layout(char* base, int P, gscuda_ref_geometry* o) {
    char* c = base;
    gscuda::gs::GeometryState s = gscuda::gs::GeometryState::fromChunk(c, P);
    o->tilesTouched = (char*)s.tilesTouched - base; o->depths = (char*)s.depths - base; o->clamped = (char*)s.clamped - base;
    o->internalRadii = (char*)s.internalRadii - base; o->means2D = (char*)s.means2D - base; o->cov3D = (char*)s.cov3D - base;
    o->conicOpacity = (char*)s.conicOpacity - base; o->rgb = (char*)s.rgb - base; o->pointOffsets = (char*)s.pointOffsets - base;
    o->total = c - base;
}
void gscuda_ref_binning_layout(char* base, int R, gscuda_ref_binning* o) {
    char* c = base;
    gscuda::gs::BinningState s = gscuda::gs::BinningState::fromChunk(c, R);
    o->keysUnsorted = (char*)s.pointListKeysUnsorted - base; o->keys = (char*)s.pointListKeys - base;
    o->valuesUnsorted = (char*)s.pointListUnsorted - base; o->values = (char*)s.pointList - base; o->total = c - base;
}
void gscuda_ref_image_layout(char* base, int N, gscuda_ref_image* o) {
    char* c = base;
    gscuda::gs::ImageState s = gscuda::gs::ImageState::fromChunk(c, N);
    o->ranges = (char*)s.ranges - base; o->nContrib = (char*)s.nContrib - base; o->accumAlpha = (char*)s.accumAlpha - base;
    o->total = c - base;
}

}  // extern "C"
