// views.cu — resident-scene renderer for streams / batches of camera views, and the one-shot
// repack of GSRast's viewer buffers into the contract layout.
//
// gsr_renderer mirrors the host object that owns the splat draw path in the reference,
// GSGaussians (/root/reference/apps/gsrast/GSGaussians.{hpp,cpp}):
//   constructor + configureFromSplatData (GSGaussians.cpp:44-153)  -> gsr_renderer_create
//   resizeFunctional grow-only scratch (GSGaussians.cpp:27-42)     -> Lane::{geom,binning,img}
//   draw(): upload view/proj/camPos, call forward (:155-212)       -> gsr_renderer_render*
// What is B200-first: two lanes (stream + scratch + pinned readback slot each) alternate
// views, so the 4-byte num_rendered round trip of view k+1 hides behind the sort/blend of
// view k, camera matrices travel as one 144-byte pinned async copy instead of three blocking
// cudaMemcpy + cudaDeviceSynchronize (CudaBuffer.hpp:42-49), and frames can stream back to
// pinned host memory while the next view renders.  A single frame is never split.
#include <cstring>
#include <new>

#include "gsr_common.cuh"

namespace gsr {

namespace {

#ifndef GSR_LANES
#define GSR_LANES 2
#endif
constexpr int LANES = GSR_LANES;
constexpr int CAM_FLOATS = 36;  // view[16] proj[16] cam_pos[3] pad[1]

struct Chunk {
    void* ptr = nullptr;
    size_t size = 0;
    cudaStream_t stream = nullptr;  // the lane's stream: the only one that ever touches the chunk
};

// resizeFunctional (GSGaussians.cpp:27-42): grow-only, over-allocates 2x, same base otherwise.  The reference grows
// with cudaFree + cudaMalloc, i.e. a device-wide synchronisation in the middle of the frame; here the chunk is
// stream-ordered memory of its lane (cudaFreeAsync / cudaMallocAsync), so growing one lane's chunk never stalls the
// other lane.
char* chunk_alloc(size_t n, void* user) {
    Chunk* c = static_cast<Chunk*>(user);
    if (n > c->size) {
        if (c->ptr) cudaFreeAsync(c->ptr, c->stream);
        c->ptr = nullptr;
        c->size = 0;
        if (cudaMallocAsync(&c->ptr, 2 * n, c->stream) != cudaSuccess) {
            c->ptr = nullptr;
            return nullptr;
        }
        c->size = 2 * n;
    }
    return static_cast<char*>(c->ptr);
}

// Every entry point runs on the device the renderer was created on, whatever the caller's current device is.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

struct Lane {
    cudaStream_t stream = nullptr;
    Chunk geom, binning, img;
    HostSlot slot;
    float* cam_dev = nullptr;   // [CAM_FLOATS]
    float* cam_host = nullptr;  // pinned
    // host-output rendering only: two frame buffers per lane, drained by the lane's own copy stream, so the
    // lane renders its next view while the previous frame travels to the host
    float* frame_dev[2] = {nullptr, nullptr};  // [3*H*W] each
    unsigned char* frame_u8[2] = {nullptr, nullptr};  // [3*H*W] each, 8-bit delivery only
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t rendered[2] = {nullptr, nullptr};
    cudaEvent_t copied[2] = {nullptr, nullptr};
    bool copy_pending[2] = {false, false};
    int frame_ix = 0;
    int* rects = nullptr;       // [2*P], GSRast-compat only: the viewer always passes _rects (GSGaussians.cpp:137,204)
    cudaEvent_t cam_free = nullptr;  // cam_host may be overwritten once this fired
    cudaEvent_t done = nullptr;
    bool cam_pending = false;
};

}  // namespace

struct Renderer {
    int device = 0;
    int P, D, M, W, H;
    const float *means3D, *shs, *colors_precomp, *opacities, *scales, *rotations, *background;
    float scale_modifier;
    unsigned flags;
    cudaStream_t user_stream;
    Lane lane[LANES];
    cudaEvent_t ready = nullptr;
    gsr_stage_times last;
    bool want_times = false;
};

namespace {

int lane_init(Lane& l) {
    GSR_CUDA_TRY(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
    l.geom.stream = l.binning.stream = l.img.stream = l.stream;
    GSR_CUDA_TRY(cudaMalloc(&l.cam_dev, CAM_FLOATS * sizeof(float)));
    GSR_CUDA_TRY(cudaHostAlloc(&l.cam_host, CAM_FLOATS * sizeof(float), cudaHostAllocDefault));
    GSR_CUDA_TRY(cudaEventCreateWithFlags(&l.cam_free, cudaEventDisableTiming));
    GSR_CUDA_TRY(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
    return ensure_slot(l.slot);
}

void lane_destroy(Lane& l) {
    if (l.geom.ptr) cudaFreeAsync(l.geom.ptr, l.stream);
    if (l.binning.ptr) cudaFreeAsync(l.binning.ptr, l.stream);
    if (l.img.ptr) cudaFreeAsync(l.img.ptr, l.stream);
    if (l.stream) cudaStreamSynchronize(l.stream);
    if (l.cam_dev) cudaFree(l.cam_dev);
    if (l.cam_host) cudaFreeHost(l.cam_host);
    if (l.copy_stream) cudaStreamSynchronize(l.copy_stream);
    for (int b = 0; b < 2; ++b) {
        if (l.frame_dev[b]) cudaFree(l.frame_dev[b]);
        if (l.frame_u8[b]) cudaFree(l.frame_u8[b]);
        if (l.rendered[b]) cudaEventDestroy(l.rendered[b]);
        if (l.copied[b]) cudaEventDestroy(l.copied[b]);
    }
    if (l.copy_stream) cudaStreamDestroy(l.copy_stream);
    if (l.rects) cudaFree(l.rects);
    if (l.cam_free) cudaEventDestroy(l.cam_free);
    if (l.done) cudaEventDestroy(l.done);
    release_slot(l.slot);
    if (l.stream) cudaStreamDestroy(l.stream);
}

// One view on one lane: camera H2D + forward, all on the lane's stream.
int render_one(Renderer* r, Lane& l, const float* cam36, float tan_fovx, float tan_fovy, float* out_dev,
               gsr_stage_times* times) {
    if (l.cam_pending) {  // previous upload from cam_host must have been consumed
        GSR_CUDA_TRY(cudaEventSynchronize(l.cam_free));
        l.cam_pending = false;
    }
    memcpy(l.cam_host, cam36, CAM_FLOATS * sizeof(float));
    GSR_CUDA_TRY(cudaMemcpyAsync(l.cam_dev, l.cam_host, CAM_FLOATS * sizeof(float), cudaMemcpyHostToDevice, l.stream));
    GSR_CUDA_TRY(cudaEventRecord(l.cam_free, l.stream));
    l.cam_pending = true;

    gsr_forward_args a;
    memset(&a, 0, sizeof(a));
    a.geometry_alloc = chunk_alloc; a.geometry_user = &l.geom;
    a.binning_alloc = chunk_alloc;  a.binning_user = &l.binning;
    a.image_alloc = chunk_alloc;    a.image_user = &l.img;
    a.P = r->P; a.D = r->D; a.M = r->M; a.background = r->background; a.width = r->W; a.height = r->H;
    const bool compat = (r->flags & GSR_FLAG_GSRAST_COMPAT) != 0;
    a.means3D = r->means3D; a.means_stride = compat ? 4 : 3;
    a.shs = r->shs; a.colors_precomp = r->colors_precomp; a.opacities = r->opacities;
    a.scales = r->scales; a.scales_stride = compat ? 4 : 3;
    a.scale_modifier = r->scale_modifier; a.rotations = r->rotations;
    a.viewmatrix = l.cam_dev; a.projmatrix = l.cam_dev + 16; a.cam_pos = l.cam_dev + 32;
    a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
    a.out_color = out_dev;
    if (compat) {
        if (!l.rects && r->P > 0) GSR_CUDA_TRY(cudaMalloc(&l.rects, (size_t)r->P * 2 * sizeof(int)));
        a.rects = l.rects;
    }
    a.stream = l.stream;
    // the lane's scratch is private: unless the renderer was created with GSR_FLAG_KEEP_STATE (the Inspector's panel,
    // gsr_renderer_map_geometry_state) nobody can read cov3D / clamped / tiles_touched / point_offsets, so they are skipped
    a.flags = r->flags & ~GSR_FLAG_KEEP_STATE;
    if (!(r->flags & GSR_FLAG_KEEP_STATE)) a.flags |= GSR_FLAG_LEAN_STATE;
    a.timings = times;
    return forward_impl(&a, &l.slot);
}

// float frame -> 8 bits per channel, same planar [3][H][W] layout: round(clamp(x, 0, 1) * 255), what the viewer's RGBA8
// framebuffer and its screenshot path (apps/gsrast/Inspector.cpp:222-257) end up holding.  16 bytes in, 4 bytes out per thread.
__global__ void quantize_u8_kernel(const size_t n4, const float4* __restrict__ in, uchar4* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = in[i];
    out[i] = make_uchar4((unsigned char)__float2int_rn(__saturatef(v.x) * 255.0f),
                         (unsigned char)__float2int_rn(__saturatef(v.y) * 255.0f),
                         (unsigned char)__float2int_rn(__saturatef(v.z) * 255.0f),
                         (unsigned char)__float2int_rn(__saturatef(v.w) * 255.0f));
}
__global__ void quantize_u8_tail_kernel(const size_t first, const size_t n, const float* __restrict__ in,
                                        unsigned char* __restrict__ out) {
    const size_t i = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned char)__float2int_rn(__saturatef(in[i]) * 255.0f);
}
int launch_quantize_u8(const float* in, unsigned char* out, size_t n, cudaStream_t s) {
    const size_t n4 = n / 4;
    if (n4) quantize_u8_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(n4, reinterpret_cast<const float4*>(in),
                                                                           reinterpret_cast<uchar4*>(out));
    if (n4 * 4 < n) quantize_u8_tail_kernel<<<1, 32, 0, s>>>(n4 * 4, n, in, out);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

// vec4 means/scales -> float3; PLY-order SH (f_dc[3], f_rest[c*15+k-1]) -> [16][3] interleaved.
// (apps/gsrast/SplatData.hpp:17-25, SplatData.cpp:48-58, GSGaussians.cpp:121-134)
__global__ void repack_kernel(int P, const float4* __restrict__ means4, const float4* __restrict__ scales4,
                              const float* __restrict__ shs_raw, float* __restrict__ means3,
                              float* __restrict__ scales3, float* __restrict__ shs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    if (means4 && means3) {
        const float4 m = means4[i];
        means3[3 * i] = m.x; means3[3 * i + 1] = m.y; means3[3 * i + 2] = m.z;
    }
    if (scales4 && scales3) {
        const float4 s = scales4[i];
        scales3[3 * i] = s.x; scales3[3 * i + 1] = s.y; scales3[3 * i + 2] = s.z;
    }
    if (shs_raw && shs) {
        const float* in = shs_raw + (size_t)i * 48;
        float* out = shs + (size_t)i * 48;
        out[0] = in[0]; out[1] = in[1]; out[2] = in[2];
        for (int k = 1; k < 16; ++k)
            for (int c = 0; c < 3; ++c) out[k * 3 + c] = in[3 + c * 15 + (k - 1)];
    }
}

}  // namespace

}  // namespace gsr

using namespace gsr;

extern "C" {

void gsr_renderer_destroy(void* h) {
    Renderer* r = static_cast<Renderer*>(h);
    if (!r) return;
    DeviceGuard guard(r->device);
    for (auto& l : r->lane) lane_destroy(l);
    if (r->ready) cudaEventDestroy(r->ready);
    delete r;
}

void* gsr_renderer_create(int P, int D, int M, const float* means3D, const float* shs, const float* colors_precomp,
                          const float* opacities, const float* scales, const float* rotations,
                          const float* background, float scale_modifier, int width, int height, void* stream,
                          unsigned flags) {
    if (P < 0 || width <= 0 || height <= 0 || !background) return nullptr;
    Renderer* r = new (std::nothrow) Renderer();
    if (!r) return nullptr;
    r->P = P; r->D = D; r->M = M; r->W = width; r->H = height;
    r->means3D = means3D; r->shs = shs; r->colors_precomp = colors_precomp; r->opacities = opacities;
    r->scales = scales; r->rotations = rotations; r->background = background;
    r->scale_modifier = scale_modifier; r->flags = flags;
    r->user_stream = static_cast<cudaStream_t>(stream);
    memset(&r->last, 0, sizeof(r->last));
    if (cudaGetDevice(&r->device) != cudaSuccess) {
        delete r;
        return nullptr;
    }
    bool ok = cudaEventCreateWithFlags(&r->ready, cudaEventDisableTiming) == cudaSuccess;
    for (auto& l : r->lane) ok = ok && lane_init(l) == 0;
    if (!ok) {
        gsr_renderer_destroy(r);
        return nullptr;
    }
    return r;
}

// cameras: HOST float[n_views][36] (view[16], proj[16], cam_pos[3], pad); out_color: DEVICE
// float[n_views][3][H][W]; num_rendered: HOST int[n_views] or NULL.  Work is ordered after
// everything already queued on the renderer's stream and that stream waits for the frames.
int gsr_renderer_render(void* h, const float* cameras, int n_views, float tan_fovx, float tan_fovy, float* out_color,
                        int* num_rendered, void* timings) {
    Renderer* r = static_cast<Renderer*>(h);
    if (!r || !cameras || !out_color || n_views < 0) return GSR_ERR_INVALID_ARG;
    DeviceGuard guard(r->device);
    GSR_CUDA_TRY(cudaEventRecord(r->ready, r->user_stream));
    for (auto& l : r->lane) GSR_CUDA_TRY(cudaStreamWaitEvent(l.stream, r->ready, 0));
    const size_t frame = (size_t)3 * r->W * r->H;
    long long total = 0;
    for (int v = 0; v < n_views; ++v) {
        // stage timings need the whole pipeline on one lane, otherwise alternate
        Lane& l = r->lane[timings ? 0 : (v % LANES)];
        int rc = render_one(r, l, cameras + (size_t)v * CAM_FLOATS, tan_fovx, tan_fovy, out_color + (size_t)v * frame,
                            timings ? &r->last : nullptr);
        if (rc < 0) return rc;
        if (num_rendered) num_rendered[v] = rc;
        total += rc;
    }
    for (auto& l : r->lane) {
        GSR_CUDA_TRY(cudaEventRecord(l.done, l.stream));
        GSR_CUDA_TRY(cudaStreamWaitEvent(r->user_stream, l.done, 0));
    }
    if (timings) memcpy(timings, &r->last, sizeof(r->last));
    return (int)(total > 0x7fffffffLL ? 0x7fffffff : total);
}

// Same, but every frame is copied to HOST memory out_color[n_views][3][H][W] (pinned memory
// makes the copy asynchronous) while the next view renders; returns after all frames landed.
// u8: frames are quantised on the device first and 8-bit frames travel (a quarter of the bytes).
static int render_host_impl(Renderer* r, const float* cameras, int n_views, float tan_fovx, float tan_fovy,
                            void* out_host, bool u8, int* num_rendered) {
    if (!r || !cameras || !out_host || n_views < 0) return GSR_ERR_INVALID_ARG;
    DeviceGuard guard(r->device);
    const size_t frame = (size_t)3 * r->W * r->H;
    long long total = 0;
    int err = 0;
    // a failing CUDA call ends the loop; EVERY way out then drains both lanes, so no copy into the caller's buffer is
    // still in flight when this function returns
#define GSR_HOST_TRY(expr)                                    \
    if (err == 0) {                                           \
        cudaError_t e_ = (expr);                              \
        if (e_ != cudaSuccess) err = -(int)e_;                \
    }
    for (auto& l : r->lane) {
        if (!l.copy_stream) GSR_HOST_TRY(cudaStreamCreateWithFlags(&l.copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            if (!l.frame_dev[b]) GSR_HOST_TRY(cudaMalloc(&l.frame_dev[b], frame * sizeof(float)));
            if (u8 && !l.frame_u8[b]) GSR_HOST_TRY(cudaMalloc(&l.frame_u8[b], frame));
            if (!l.rendered[b]) GSR_HOST_TRY(cudaEventCreateWithFlags(&l.rendered[b], cudaEventDisableTiming));
            if (!l.copied[b]) GSR_HOST_TRY(cudaEventCreateWithFlags(&l.copied[b], cudaEventDisableTiming));
        }
    }
    GSR_HOST_TRY(cudaEventRecord(r->ready, r->user_stream));
    for (auto& l : r->lane) GSR_HOST_TRY(cudaStreamWaitEvent(l.stream, r->ready, 0));
    for (int v = 0; v < n_views && err == 0; ++v) {
        Lane& l = r->lane[v % LANES];
        const int b = l.frame_ix;
        l.frame_ix ^= 1;
        // the frame buffer is free again once the copy issued from it two views ago has drained
        if (l.copy_pending[b]) GSR_HOST_TRY(cudaStreamWaitEvent(l.stream, l.copied[b], 0));
        if (err) break;
        int rc = render_one(r, l, cameras + (size_t)v * CAM_FLOATS, tan_fovx, tan_fovy, l.frame_dev[b], nullptr);
        if (rc < 0) { err = rc; break; }
        if (u8) {
            int rq = launch_quantize_u8(l.frame_dev[b], l.frame_u8[b], frame, l.stream);
            if (rq < 0) { err = rq; break; }
        }
        GSR_HOST_TRY(cudaEventRecord(l.rendered[b], l.stream));
        GSR_HOST_TRY(cudaStreamWaitEvent(l.copy_stream, l.rendered[b], 0));
        if (u8) {
            GSR_HOST_TRY(cudaMemcpyAsync(static_cast<unsigned char*>(out_host) + (size_t)v * frame, l.frame_u8[b], frame,
                                         cudaMemcpyDeviceToHost, l.copy_stream));
        } else {
            GSR_HOST_TRY(cudaMemcpyAsync(static_cast<float*>(out_host) + (size_t)v * frame, l.frame_dev[b],
                                         frame * sizeof(float), cudaMemcpyDeviceToHost, l.copy_stream));
        }
        GSR_HOST_TRY(cudaEventRecord(l.copied[b], l.copy_stream));
        if (err == 0) l.copy_pending[b] = true;
        if (num_rendered) num_rendered[v] = rc;
        total += rc;
    }
#undef GSR_HOST_TRY
    for (auto& l : r->lane) {
        cudaError_t e1 = l.copy_stream ? cudaStreamSynchronize(l.copy_stream) : cudaSuccess;
        cudaError_t e2 = cudaStreamSynchronize(l.stream);
        if (err == 0 && e1 != cudaSuccess) err = -(int)e1;
        if (err == 0 && e2 != cudaSuccess) err = -(int)e2;
        l.copy_pending[0] = l.copy_pending[1] = false;
        const int ea = take_async_error(l.slot);  // both lanes are idle: every watchdog of this batch has reported
        if (err == 0 && ea < 0) err = ea;
    }
    if (err) return err;
    return (int)(total > 0x7fffffffLL ? 0x7fffffff : total);
}

int gsr_renderer_render_host(void* h, const float* cameras, int n_views, float tan_fovx, float tan_fovy,
                             float* out_color_host, int* num_rendered) {
    return render_host_impl(static_cast<Renderer*>(h), cameras, n_views, tan_fovx, tan_fovy, out_color_host, false,
                            num_rendered);
}

int gsr_renderer_render_host_u8(void* h, const float* cameras, int n_views, float tan_fovx, float tan_fovy,
                                unsigned char* out_color_host, int* num_rendered) {
    return render_host_impl(static_cast<Renderer*>(h), cameras, n_views, tan_fovx, tan_fovy, out_color_host, true,
                            num_rendered);
}

int gsr_frames_to_u8(const float* frames, unsigned char* out, size_t n_values, void* stream) {
    if (!frames || !out) return GSR_ERR_INVALID_ARG;
    return launch_quantize_u8(frames, out, n_values, static_cast<cudaStream_t>(stream));
}

int gsr_renderer_last_times(void* h, gsr_stage_times* out) {
    Renderer* r = static_cast<Renderer*>(h);
    if (!r || !out) return GSR_ERR_INVALID_ARG;
    *out = r->last;
    return 0;
}

int gsr_renderer_num_lanes(void) { return LANES; }

void* gsr_pinned_alloc(size_t bytes, unsigned flags) {
    void* p = nullptr;
    unsigned f = cudaHostAllocDefault;
    if (flags & GSR_PINNED_WRITE_COMBINED) f |= cudaHostAllocWriteCombined;
    if (flags & GSR_PINNED_PORTABLE) f |= cudaHostAllocPortable;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, f) != cudaSuccess) return nullptr;
    return p;
}

void gsr_pinned_free(void* p) {
    if (p) cudaFreeHost(p);
}

// GSGaussians::mapGeometryState (GSGaussians.cpp:214-219): fromChunk over the lane's geometry chunk.
int gsr_renderer_map_geometry_state(void* h, int lane, gsr_geometry_state* out) {
    Renderer* r = static_cast<Renderer*>(h);
    if (!r || !out || lane < 0 || lane >= LANES || !r->lane[lane].geom.ptr) return GSR_ERR_INVALID_ARG;
    gsr_geometry_state_map(static_cast<char*>(r->lane[lane].geom.ptr), r->P, out);
    return 0;
}

int gsr_repack_gsrast_scene(int P, const float* means4, const float* scales4, const float* shs_raw, float* means3,
                            float* scales3, float* shs, void* stream) {
    if (P <= 0) return 0;
    repack_kernel<<<(P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        P, reinterpret_cast<const float4*>(means4), reinterpret_cast<const float4*>(scales4), shs_raw, means3, scales3,
        shs);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

}  // extern "C"
