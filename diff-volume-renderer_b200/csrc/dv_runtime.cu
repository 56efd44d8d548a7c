// dv_runtime.cu -- contexts, plans, fields: the host-side objects behind hp.h.
//
// Validation and defaulting follow the reference statement by statement
// (reference hotpath/src/runtime/hp_runtime.cpp:15-146, 259-374) so that a
// caller sees the same hp_status for the same arguments.  What differs is what
// the objects hold: a context names a GPU and a stream, a field owns a copy of
// its grid in HBM.
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <cuda.h>   // types of the green-context driver API; the entry points are looked up at run time

#include "dv_objects.h"

namespace dv {

namespace {
thread_local std::string g_last_error;
}

void set_last_error(const std::string& what) { g_last_error = what; }
const char* last_error_text() { return g_last_error.c_str(); }

hp_status cuda_fail(cudaError_t err, const char* what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorString(err);
    cudaGetLastError();  // clear the sticky-free error state
    switch (err) {
        case cudaErrorMemoryAllocation: return HP_STATUS_OUT_OF_MEMORY;
        case cudaErrorNoDevice:
        case cudaErrorInsufficientDriver:
        case cudaErrorInvalidDevice: return HP_STATUS_UNSUPPORTED;
        default: return HP_STATUS_INTERNAL_ERROR;
    }
}

// ---- SM reservation (hpx_ctx_ext2.reserve_sms): a green context over all but `reserve` SMs ------------------------
namespace {
struct GreenApi {
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                          unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    bool ok = false;
};

template <typename F>
bool driver_fn(const char* name, F& out) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    out = reinterpret_cast<F>(p);
    return true;
}

const GreenApi& green_api() {
    static const GreenApi api = [] {
        GreenApi a;
        a.ok = driver_fn("cuDeviceGet", a.DeviceGet) && driver_fn("cuDeviceGetDevResource", a.DeviceGetDevResource) &&
               driver_fn("cuDevSmResourceSplitByCount", a.DevSmResourceSplitByCount) &&
               driver_fn("cuDevResourceGenerateDesc", a.DevResourceGenerateDesc) && driver_fn("cuGreenCtxCreate", a.GreenCtxCreate) &&
               driver_fn("cuGreenCtxDestroy", a.GreenCtxDestroy) && driver_fn("cuGreenCtxStreamCreate", a.GreenCtxStreamCreate);
        return a;
    }();
    return api;
}

// Creates ctx->green_ctx over (total - reserve) SMs (rounded DOWN to the partition granularity so that at least `reserve`
// SMs stay free) and ctx->stream inside it.
hp_status create_green_stream(const hp_ctx* ctx) {
    const GreenApi& g = green_api();
    if (!g.ok) {
        set_last_error("SM reservation needs the green-context driver API (CUDA 12.4+)");
        return HP_STATUS_UNSUPPORTED;
    }
    DV_CUDA(cudaFree(nullptr));   // the primary context exists
    CUdevice dev;
    CUdevResource all{}, part{}, rest{};
    if (g.DeviceGet(&dev, ctx->device) != CUDA_SUCCESS || g.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) {
        set_last_error("cuDeviceGetDevResource failed");
        return HP_STATUS_INTERNAL_ERROR;
    }
    const unsigned int total = all.sm.smCount;
    ctx->total_sms = total;
    if (ctx->reserve_sms >= total) {
        set_last_error("reserve_sms leaves no SM for the context");
        return HP_STATUS_INVALID_ARGUMENT;
    }
    // the split rounds a request UP to the partition granularity (8 SMs on sm_90+): start from the largest multiple of 8
    // that leaves `reserve` SMs free and step down until the partition really does
    unsigned int want = (total - ctx->reserve_sms) / 8u * 8u;
    CUresult rc = CUDA_ERROR_INVALID_VALUE;
    while (want >= 8u) {
        unsigned int groups = 1;
        rc = g.DevSmResourceSplitByCount(&part, &groups, &all, &rest, 0u, want);
        if (rc != CUDA_SUCCESS || groups != 1) break;
        if (part.sm.smCount + ctx->reserve_sms <= total) break;
        want -= 8u;
    }
    if (rc != CUDA_SUCCESS || part.sm.smCount == 0 || total - part.sm.smCount < ctx->reserve_sms) {
        set_last_error("cuDevSmResourceSplitByCount could not set " + std::to_string(ctx->reserve_sms) + " SMs aside (driver error " +
                       std::to_string(static_cast<int>(rc)) + ")");
        return HP_STATUS_UNSUPPORTED;
    }
    CUdevResourceDesc desc = nullptr;
    CUgreenCtx green = nullptr;
    if ((rc = g.DevResourceGenerateDesc(&desc, &part, 1)) != CUDA_SUCCESS ||
        (rc = g.GreenCtxCreate(&green, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM)) != CUDA_SUCCESS) {
        set_last_error("cuGreenCtxCreate failed with driver error " + std::to_string(static_cast<int>(rc)));
        return HP_STATUS_INTERNAL_ERROR;
    }
    CUstream stream = nullptr;
    if ((rc = g.GreenCtxStreamCreate(&stream, green, CU_STREAM_NON_BLOCKING, 0)) != CUDA_SUCCESS) {
        g.GreenCtxDestroy(green);
        set_last_error("cuGreenCtxStreamCreate failed with driver error " + std::to_string(static_cast<int>(rc)));
        return HP_STATUS_INTERNAL_ERROR;
    }
    ctx->green_ctx = green;
    ctx->stream = reinterpret_cast<cudaStream_t>(stream);
    ctx->owns_stream = true;
    ctx->usable_sms = part.sm.smCount;
    return HP_STATUS_SUCCESS;
}
}  // namespace

static void destroy_device_state(const hp_ctx* ctx) {
    if (ctx->owns_stream && ctx->stream != nullptr) cudaStreamDestroy(ctx->stream);
    if (ctx->green_ctx != nullptr && green_api().ok) green_api().GreenCtxDestroy(static_cast<CUgreenCtx>(ctx->green_ctx));
    ctx->green_ctx = nullptr;
    cudaFree(ctx->d_status);
    cudaFree(ctx->d_total);
    if (ctx->h_status != nullptr) cudaFreeHost(ctx->h_status);
    if (ctx->h_total != nullptr) cudaFreeHost(ctx->h_total);
    for (cudaEvent_t& e : ctx->marks) {
        if (e != nullptr) cudaEventDestroy(e);
        e = nullptr;
    }
    ctx->stream = nullptr;
    ctx->owns_stream = false;
    ctx->d_status = nullptr; ctx->d_total = nullptr; ctx->h_status = nullptr; ctx->h_total = nullptr;
}

hp_status ensure_device(const hp_ctx* ctx) {
    if (ctx == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (ctx->failed) return HP_STATUS_UNSUPPORTED;
    if (!ctx->ready) {
        int count = 0;
        const cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count <= 0) {
            ctx->failed = true;
            cudaGetLastError();
            set_last_error("no usable CUDA device: this library has no CPU compute path");
            return HP_STATUS_UNSUPPORTED;
        }
        int dev = ctx->requested_ordinal;
        if (dev < 0) {
            if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
        }
        if (dev >= count) {
            ctx->failed = true;
            set_last_error("requested CUDA device ordinal does not exist");
            return HP_STATUS_UNSUPPORTED;
        }
        ctx->device = dev;
        DV_CUDA(cudaSetDevice(dev));
        // all or nothing: a failure half way rolls back what was created, so that a retry starts clean (no leak)
        cudaError_t err = cudaSuccess;
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) ctx->total_sms = ctx->usable_sms = static_cast<uint32_t>(sms);
        if (ctx->reserve_sms != 0) {
            if (ctx->has_user_stream) {
                set_last_error("hpx_ctx_ext2.reserve_sms needs a library-owned stream (stream must be NULL)");
                return HP_STATUS_INVALID_ARGUMENT;
            }
            const hp_status gs = create_green_stream(ctx);
            if (gs != HP_STATUS_SUCCESS) {
                destroy_device_state(ctx);
                return gs;
            }
        } else if (ctx->has_user_stream) {
            ctx->stream = ctx->user_stream;
            ctx->owns_stream = false;
        } else {
            err = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
            ctx->owns_stream = err == cudaSuccess;
        }
        if (err == cudaSuccess) err = cudaMalloc(&ctx->d_status, sizeof(uint32_t));
        if (err == cudaSuccess) err = cudaMalloc(&ctx->d_total, sizeof(unsigned long long));
        if (err == cudaSuccess) err = cudaMallocHost(&ctx->h_status, sizeof(uint32_t));
        if (err == cudaSuccess) err = cudaMallocHost(&ctx->h_total, sizeof(unsigned long long));
        if (err != cudaSuccess) {
            destroy_device_state(ctx);
            return cuda_fail(err, "context device state");
        }
        ctx->ready = true;
        return HP_STATUS_SUCCESS;
    }
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != ctx->device) DV_CUDA(cudaSetDevice(ctx->device));
    return HP_STATUS_SUCCESS;
}

hp_status DeviceScope::enter(const hp_ctx* ctx) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) {
        cudaGetLastError();
        cur = -1;
    }
    const hp_status st = ensure_device(ctx);
    if (st == HP_STATUS_SUCCESS && cur >= 0 && cur != ctx->device) {
        prev = cur;
        switched = true;
    }
    return st;
}

DeviceScope::~DeviceScope() {
    if (switched) cudaSetDevice(prev);
}

const hp_ctx* ctx_retain(const hp_ctx* ctx) {
    if (ctx != nullptr) ctx->refs.fetch_add(1, std::memory_order_relaxed);
    return ctx;
}

void ctx_unref(const hp_ctx* ctx) {
    if (ctx == nullptr || ctx->refs.fetch_sub(1, std::memory_order_acq_rel) != 1) return;
    if (ctx->ready) {
        DeviceScope scope;
        scope.enter(ctx);
        cudaStreamSynchronize(ctx->stream);
        destroy_device_state(ctx);
    }
    delete ctx;
}

const hp_plan* plan_retain(const hp_plan* plan) {
    if (plan != nullptr) plan->refs.fetch_add(1, std::memory_order_relaxed);
    return plan;
}

void plan_unref(const hp_plan* plan) {
    if (plan == nullptr || plan->refs.fetch_sub(1, std::memory_order_acq_rel) != 1) return;
    ctx_unref(plan->ctx);
    delete plan;
}

// ---- plan ---------------------------------------------------------------------
hp_status resolve_plan_desc(hp_plan_desc* p) {
    if (p->width == 0 || p->height == 0) return HP_STATUS_INVALID_ARGUMENT;
    if (!(p->t_far > p->t_near)) return HP_STATUS_INVALID_ARGUMENT;

    hp_camera_desc& cam = p->camera;
    if (cam.model != HP_CAMERA_PINHOLE && cam.model != HP_CAMERA_ORTHOGRAPHIC) cam.model = HP_CAMERA_PINHOLE;
    const bool no_intrinsics = std::all_of(cam.K, cam.K + 9, [](float v) { return v == 0.0f; });
    if (no_intrinsics) {
        cam.K[0] = cam.K[4] = cam.K[8] = 1.0f;
        cam.K[2] = static_cast<float>(p->width) * 0.5f;
        cam.K[5] = static_cast<float>(p->height) * 0.5f;
    }
    if (cam.K[0] == 0.0f) cam.K[0] = 1.0f;
    if (cam.K[4] == 0.0f) cam.K[4] = 1.0f;
    const bool no_pose = std::all_of(cam.c2w, cam.c2w + 12, [](float v) { return v == 0.0f; });
    if (no_pose) cam.c2w[0] = cam.c2w[5] = cam.c2w[10] = 1.0f;
    if (cam.model == HP_CAMERA_ORTHOGRAPHIC && cam.ortho_scale <= 0.0f) cam.ortho_scale = 1.0f;

    hp_roi_desc& roi = p->roi;
    if (roi.width == 0 || roi.height == 0) roi = hp_roi_desc{0, 0, p->width, p->height};
    // 32-bit sums on purpose: identical acceptance to the reference (hp_runtime.cpp:107)
    if (roi.x + roi.width > p->width || roi.y + roi.height > p->height) return HP_STATUS_INVALID_ARGUMENT;
    const uint64_t roi_rays = static_cast<uint64_t>(roi.width) * roi.height;
    if (p->max_rays == 0U) p->max_rays = static_cast<uint32_t>(std::min<uint64_t>(roi_rays, UINT32_MAX));
    if (roi_rays > p->max_rays) return HP_STATUS_INVALID_ARGUMENT;

    hp_sampling_desc& s = p->sampling;
    if (!(s.dt > 0.0f)) {
        const float span = p->t_far - p->t_near;
        const float fallback = span > 0.0f ? span / 64.0f : 1.0f;
        s.dt = fallback > 0.0f ? fallback : 1.0f;
    }
    if (s.max_steps == 0U) s.max_steps = 64U;
    if (s.mode != HP_SAMPLING_FIXED && s.mode != HP_SAMPLING_STRATIFIED) s.mode = HP_SAMPLING_FIXED;

    if (p->max_samples == 0U) {
        const uint64_t want = static_cast<uint64_t>(p->max_rays) * s.max_steps;
        const uint64_t capped = std::min<uint64_t>(want, UINT32_MAX);
        p->max_samples = capped == 0 ? p->max_rays : static_cast<uint32_t>(capped);
    }
    if (p->max_samples < p->max_rays) return HP_STATUS_INVALID_ARGUMENT;
    return HP_STATUS_SUCCESS;
}

// Host replay of the marching loop's emit / skip / stop decisions
// (reference samp_cpu.cpp:222-244).  This file is compiled without FMA
// contraction, so the float arithmetic matches the device code bit for bit.
void emitted_samples(const hp_plan_desc& d, uint32_t* count, bool* gap_free) {
    uint32_t n = 0;
    bool gaps = false, skipped = false;
    const float tn = d.t_near, tf = d.t_far, dts = d.sampling.dt;
    if (tf > tn) {
        for (uint32_t step = 0; step < d.sampling.max_steps; ++step) {
            volatile float base = tn + static_cast<float>(step) * dts;
            if (base >= tf) break;
            volatile float end = std::min(base + dts, tf);
            volatile float dta = end - base;
            if (!(dta > 0.0f)) { skipped = true; continue; }
            if (skipped) gaps = true;
            ++n;
        }
    }
    *count = n;
    *gap_free = !gaps;
}

CameraParams camera_params(const hp_camera_desc& cam) {
    CameraParams c{};
    c.fx = cam.K[0]; c.fy = cam.K[4]; c.cx = cam.K[2]; c.cy = cam.K[5];
    c.r00 = cam.c2w[0]; c.r01 = cam.c2w[1]; c.r02 = cam.c2w[2];
    c.r10 = cam.c2w[4]; c.r11 = cam.c2w[5]; c.r12 = cam.c2w[6];
    c.r20 = cam.c2w[8]; c.r21 = cam.c2w[9]; c.r22 = cam.c2w[10];
    c.ox = cam.c2w[3]; c.oy = cam.c2w[7]; c.oz = cam.c2w[11];
    c.ortho = cam.model == HP_CAMERA_ORTHOGRAPHIC ? 1u : 0u;
    return c;
}

FrameParams frame_params_from_plan(const hp_plan& plan) {
    const hp_plan_desc& d = plan.desc;
    FrameParams p{};
    p.cam = camera_params(d.camera);
    p.march.t_near = d.t_near;
    p.march.t_far = d.t_far;
    p.march.dt = d.sampling.dt;
    p.march.max_steps = d.sampling.max_steps;
    p.march.stratified = d.sampling.mode == HP_SAMPLING_STRATIFIED ? 1u : 0u;
    p.march.uniform_count = plan.uniform_count;
    p.march.seed = d.seed;
    p.march.ray_index_base = 0;
    p.roi = RoiParams{d.roi.x, d.roi.y, d.roi.width, d.roi.height, d.width, d.height, 1u, 0u, 0u};
    return p;
}

// ---- tensors ---------------------------------------------------------------------
void shape_tensor(hp_tensor& t, hp_dtype dtype, hp_memspace ms, uint32_t rank, int64_t d0, int64_t d1, int64_t d2) {
    t.dtype = dtype;
    t.memspace = ms;
    t.rank = rank;
    const int64_t dims[3] = {d0, d1, d2};
    int64_t stride = 1;
    for (int i = static_cast<int>(rank) - 1; i >= 0; --i) {
        t.shape[i] = dims[i];
        t.stride[i] = stride;
        stride *= dims[i];
    }
}

void* Bump::take(size_t bytes, size_t alignment) {
    if (bytes == 0 || ptr == nullptr) return nullptr;
    const uintptr_t cur = reinterpret_cast<uintptr_t>(ptr);
    const uintptr_t aligned = (cur + (alignment - 1)) & ~static_cast<uintptr_t>(alignment - 1);
    const size_t pad = static_cast<size_t>(aligned - cur);
    if (pad > remaining || bytes > remaining - pad) return nullptr;
    void* out = ptr + pad;
    ptr += pad + bytes;
    remaining -= pad + bytes;
    return out;
}

DeviceScratch::~DeviceScratch() {
    for (void* p : ptrs) cudaFree(p);
}

void* DeviceScratch::take(size_t bytes) {
    void* p = nullptr;
    const cudaError_t e = cudaMalloc(&p, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMalloc(scratch)");
        return nullptr;
    }
    ptrs.push_back(p);
    return p;
}

static GridParams grid_params(const hp_field* f) {
    GridParams g{};
    if (f == nullptr) return g;
    g.data = f->d_data;
    g.nx = f->nx; g.ny = f->ny; g.nz = f->nz; g.channels = f->stride;   // GridParams::channels is the voxel stride
    g.linear = f->interp == HP_INTERP_LINEAR ? 1u : 0u;
    g.clamp = f->oob == HP_OOB_CLAMP ? 1u : 0u;
    g.present = 1u;
    return g;
}

FieldPair field_pair(const hp_field* fs, const hp_field* fc) {
    FieldPair p{};
    p.sigma = grid_params(fs);
    p.color = grid_params(fc);
    p.packed = nullptr;
    return p;
}

ScatterParams scatter_params(const hpx_grid& g) {
    ScatterParams sp{};
    sp.grad = reinterpret_cast<float4*>(g.d_grad);
    sp.nx = g.nx; sp.ny = g.ny; sp.nz = g.nz;
    sp.nearest = g.linear ? 0u : 1u;
    sp.clamp = g.clamp ? 1u : 0u;
    sp.unit_bbox = 1u;
    for (int i = 0; i < 3; ++i) {
        sp.bmin[i] = g.bmin[i];
        sp.bmax[i] = g.bmax[i];
        if (g.bmin[i] != 0.0f || g.bmax[i] != 1.0f) sp.unit_bbox = 0u;
    }
    sp.box_nx = g.nx; sp.box_ny = g.ny; sp.box_nz = g.nz;
    if (g.gsx != 0) {
        sp.box_sx = g.gsx; sp.box_sy = g.gsy; sp.box_sz = g.gsz;
    } else {
        sp.box_sx = 1u;
        sp.box_sy = static_cast<uint32_t>(g.nx);
        sp.box_sz = static_cast<uint32_t>(g.nx) * static_cast<uint32_t>(g.ny);
    }
    return sp;
}

}  // namespace dv

using namespace dv;

// =============================================================================
// extern "C": version, context, plan, field
// =============================================================================
extern "C" {

HP_API hp_version hp_get_version(void) { DV_RANGE("hp_get_version"); return hp_version{HP_VERSION_MAJOR, HP_VERSION_MINOR, HP_VERSION_PATCH}; }

HP_API const char* hpx_last_error(void) { DV_RANGE("hpx_last_error"); return dv::last_error_text(); }

// reference hp_runtime.cpp:15-31
HP_API hp_status hp_ctx_create(const hp_ctx_desc* desc, hp_ctx** out_ctx) {
    DV_RANGE("hp_ctx_create");
    if (out_ctx == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    hp_ctx* ctx = new (std::nothrow) hp_ctx();
    if (ctx == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    if (desc != nullptr) ctx->desc = *desc;
    if (ctx->desc.preferred_device != nullptr) {
        // the caller's string may be a temporary (reference src/core/context.cpp:36): keep a copy
        ctx->device_name = ctx->desc.preferred_device;
        ctx->desc.preferred_device = ctx->device_name.c_str();
        const size_t colon = ctx->device_name.find(':');
        if (colon != std::string::npos) {
            char* end = nullptr;
            const long v = std::strtol(ctx->device_name.c_str() + colon + 1, &end, 10);
            if (end != ctx->device_name.c_str() + colon + 1 && v >= 0) ctx->requested_ordinal = static_cast<int>(v);
        }
    }
    if (ctx->desc.reserved != nullptr) {
        const auto* ext = static_cast<const hpx_ctx_ext*>(ctx->desc.reserved);
        if (ext->magic == HPX_CTX_EXT_MAGIC || ext->magic == HPX_CTX_EXT2_MAGIC) {
            if (ext->device_ordinal >= 0) ctx->requested_ordinal = ext->device_ordinal;
            if (ext->stream != nullptr) {
                ctx->user_stream = static_cast<cudaStream_t>(ext->stream);
                ctx->has_user_stream = true;
            }
            if (ext->magic == HPX_CTX_EXT2_MAGIC) ctx->reserve_sms = static_cast<const hpx_ctx_ext2*>(ctx->desc.reserved)->reserve_sms;
        }
        ctx->desc.reserved = nullptr;  // not ours to hand back later
    }
    *out_ctx = ctx;
    return HP_STATUS_SUCCESS;
}

HP_API void hp_ctx_release(hp_ctx* ctx) { DV_RANGE("hp_ctx_release"); ctx_unref(ctx); }   // objects created from it keep it alive (hp_ctx::refs)

HP_API hp_status hp_ctx_get_desc(const hp_ctx* ctx, hp_ctx_desc* out_desc) {
    DV_RANGE("hp_ctx_get_desc");
    if (ctx == nullptr || out_desc == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    *out_desc = ctx->desc;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_ctx_synchronize(const hp_ctx* ctx) {
    DV_RANGE("hpx_ctx_synchronize");
    DV_ENTER(ctx);
    DV_CUDA(cudaStreamSynchronize(ctx->stream));
    return HP_STATUS_SUCCESS;
}

// Device-side stage timing for callers without a CUDA runtime of their own (dvren::Renderer's RenderStats): record an
// event on the context's stream under a small integer name, read the elapsed time between two of them later.
HP_API hp_status hpx_ctx_mark(const hp_ctx* ctx, uint32_t slot) {
    DV_RANGE("hpx_ctx_mark");
    if (slot >= 16) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(ctx);
    if (ctx->marks[slot] == nullptr) DV_CUDA(cudaEventCreate(&ctx->marks[slot]));
    DV_CUDA(cudaEventRecord(ctx->marks[slot], ctx->stream));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_ctx_elapsed_ms(const hp_ctx* ctx, uint32_t slot_begin, uint32_t slot_end, float* out_ms) {
    DV_RANGE("hpx_ctx_elapsed_ms");
    if (slot_begin >= 16 || slot_end >= 16 || out_ms == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(ctx);
    if (ctx->marks[slot_begin] == nullptr || ctx->marks[slot_end] == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_CUDA(cudaEventSynchronize(ctx->marks[slot_end]));
    DV_CUDA(cudaEventElapsedTime(out_ms, ctx->marks[slot_begin], ctx->marks[slot_end]));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_copy_to_host(const hp_ctx* ctx, void* host_dst, const void* device_src, size_t bytes) {
    DV_RANGE("hpx_copy_to_host");
    if (host_dst == nullptr || device_src == nullptr) return bytes == 0 ? HP_STATUS_SUCCESS : HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(ctx);
    DV_CUDA(cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DV_CUDA(cudaStreamSynchronize(ctx->stream));
    return HP_STATUS_SUCCESS;
}

// Page-lock a caller-owned host range so that copies to / from it are direct DMA (dvren::Renderer registers result
// vectors it sees repeatedly).  Returns UNSUPPORTED when the range cannot be registered; the caller just keeps going.
HP_API hp_status hpx_host_register(const hp_ctx* ctx, void* host_ptr, size_t bytes) {
    DV_RANGE("hpx_host_register");
    if (host_ptr == nullptr || bytes == 0) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(ctx);
    const cudaError_t e = cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaHostRegister");
        return HP_STATUS_UNSUPPORTED;
    }
    return HP_STATUS_SUCCESS;
}

HP_API void hpx_host_unregister(const hp_ctx* ctx, void* host_ptr) {
    DV_RANGE("hpx_host_unregister");
    if (host_ptr == nullptr || ctx == nullptr || !ctx->ready) return;
    DeviceScope scope;
    if (scope.enter(ctx) != HP_STATUS_SUCCESS) return;
    cudaStreamSynchronize(ctx->stream);
    if (cudaHostUnregister(host_ptr) != cudaSuccess) cudaGetLastError();
}

HP_API hp_status hpx_device_alloc(const hp_ctx* ctx, size_t bytes, void** out_device_ptr) {
    DV_RANGE("hpx_device_alloc");
    if (out_device_ptr == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(ctx);
    DV_CUDA(cudaMalloc(out_device_ptr, bytes ? bytes : 16));
    return HP_STATUS_SUCCESS;
}

HP_API void hpx_device_free(const hp_ctx* ctx, void* device_ptr) {
    DV_RANGE("hpx_device_free");
    if (device_ptr == nullptr || ctx == nullptr || !ctx->ready) return;
    DeviceScope scope;
    if (scope.enter(ctx) != HP_STATUS_SUCCESS) return;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(device_ptr);
}

HP_API hp_status hpx_copy_to_device(const hp_ctx* ctx, void* device_dst, const void* host_src, size_t bytes) {
    DV_RANGE("hpx_copy_to_device");
    if (device_dst == nullptr || host_src == nullptr) return bytes == 0 ? HP_STATUS_SUCCESS : HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(ctx);
    DV_CUDA(cudaMemcpyAsync(device_dst, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    DV_CUDA(cudaStreamSynchronize(ctx->stream));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_ctx_sm_counts(const hp_ctx* ctx, uint32_t* out_usable, uint32_t* out_total) {
    DV_RANGE("hpx_ctx_sm_counts");
    DV_ENTER(ctx);
    if (out_usable != nullptr) *out_usable = ctx->usable_sms;
    if (out_total != nullptr) *out_total = ctx->total_sms;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_ctx_device(const hp_ctx* ctx, int32_t* out_ordinal, void** out_stream) {
    DV_RANGE("hpx_ctx_device");
    DV_ENTER(ctx);
    if (out_ordinal != nullptr) *out_ordinal = ctx->device;
    if (out_stream != nullptr) *out_stream = ctx->stream;
    return HP_STATUS_SUCCESS;
}

// reference hp_runtime.cpp:45-146
HP_API hp_status hp_plan_create(const hp_ctx* ctx, const hp_plan_desc* desc, hp_plan** out_plan) {
    DV_RANGE("hp_plan_create");
    if (ctx == nullptr || desc == nullptr || out_plan == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    hp_plan* plan = new (std::nothrow) hp_plan();
    if (plan == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    plan->desc = *desc;
    const hp_status st = resolve_plan_desc(&plan->desc);
    if (st != HP_STATUS_SUCCESS) {
        delete plan;
        return st;
    }
    plan->ctx = ctx_retain(ctx);
    emitted_samples(plan->desc, &plan->uniform_count, &plan->gap_free);
    *out_plan = plan;
    return HP_STATUS_SUCCESS;
}

HP_API void hp_plan_release(hp_plan* plan) { DV_RANGE("hp_plan_release"); plan_unref(plan); }

HP_API hp_status hp_plan_get_desc(const hp_plan* plan, hp_plan_desc* out_desc) {
    DV_RANGE("hp_plan_get_desc");
    if (plan == nullptr || out_desc == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    *out_desc = plan->desc;
    return HP_STATUS_SUCCESS;
}

// reference hp_runtime.cpp:259-339 -- same checks, same order; then the upload
static hp_status create_dense_field(const hp_ctx* ctx, const hp_tensor* grid, uint32_t interp, uint32_t oob,
                                    bool color, hp_field** out_field) {
    if (ctx == nullptr || grid == nullptr || out_field == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (grid->dtype != HP_DTYPE_F32) return HP_STATUS_INVALID_ARGUMENT;
    // HOST tensors are uploaded; DEVICE tensors (additive) are copied device to device
    if (grid->memspace != HP_MEMSPACE_HOST && grid->memspace != HP_MEMSPACE_DEVICE) return HP_STATUS_UNSUPPORTED;
    if (color) {
        if (grid->rank < 4 || grid->rank > 8 || grid->shape[grid->rank - 1] != 3) return HP_STATUS_INVALID_ARGUMENT;
    } else {
        if (grid->rank < 3 || grid->rank > 8) return HP_STATUS_INVALID_ARGUMENT;
    }
    for (uint32_t i = 0; i < grid->rank; ++i) {
        if (grid->shape[i] <= 0) return HP_STATUS_INVALID_ARGUMENT;
    }
    hp_field* f = new (std::nothrow) hp_field();
    if (f == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    f->kind = color ? FieldKind::kDenseColor : FieldKind::kDenseSigma;
    f->ctx = ctx;
    f->source = *grid;
    f->interp = interp == static_cast<uint32_t>(HP_INTERP_NEAREST) ? HP_INTERP_NEAREST : HP_INTERP_LINEAR;
    f->oob = oob == static_cast<uint32_t>(HP_OOB_CLAMP) ? HP_OOB_CLAMP : HP_OOB_ZERO;
    // the reference reads shape[0..2] (= nz, ny, nx) and, for colour, shape[3] as the channel stride
    f->nz = static_cast<int32_t>(grid->shape[0]);
    f->ny = static_cast<int32_t>(grid->shape[1]);
    f->nx = static_cast<int32_t>(grid->shape[2]);
    f->channels = color ? static_cast<int32_t>(grid->shape[3]) : 1;
    f->stride = f->channels;
    if (color && f->channels < 3) {
        delete f;
        return HP_STATUS_INVALID_ARGUMENT;
    }
    if (grid->data == nullptr) {
        // the reference accepts the handle and fails at the first query; fail early instead
        delete f;
        return HP_STATUS_INVALID_ARGUMENT;
    }
    DeviceScope scope;
    hp_status st = scope.enter(ctx);
    if (st != HP_STATUS_SUCCESS) {
        delete f;
        return st;
    }
    const size_t elems = static_cast<size_t>(f->nx) * f->ny * f->nz * f->channels;
    cudaError_t e = cudaMalloc(&f->d_data, elems * sizeof(float));
    if (e == cudaSuccess) {
        e = cudaMemcpyAsync(f->d_data, grid->data, elems * sizeof(float),
                            grid->memspace == HP_MEMSPACE_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                            ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        st = cuda_fail(e, "field upload");
        cudaFree(f->d_data);
        delete f;
        return st;
    }
    ctx_retain(ctx);
    *out_field = f;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hp_field_create_grid_sigma(const hp_ctx* ctx, const hp_tensor* grid, uint32_t interp, uint32_t oob,
                                            hp_field** out_field) {
    DV_RANGE("hp_field_create_grid_sigma");
    return create_dense_field(ctx, grid, interp, oob, false, out_field);
}

HP_API hp_status hp_field_create_grid_color(const hp_ctx* ctx, const hp_tensor* grid, uint32_t interp, uint32_t oob,
                                            hp_field** out_field) {
    DV_RANGE("hp_field_create_grid_color");
    return create_dense_field(ctx, grid, interp, oob, true, out_field);
}

// The reference's hash-MLP field is a fixed-size toy outside the dense-grid hot
// path; this library does not implement it.
HP_API hp_status hp_field_create_hash_mlp(const hp_ctx* ctx, const hp_tensor* params, hp_field** out_field) {
    DV_RANGE("hp_field_create_hash_mlp");
    if (ctx == nullptr || params == nullptr || out_field == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    *out_field = nullptr;
    return HP_STATUS_UNSUPPORTED;
}

HP_API void hp_field_release(hp_field* field) {
    DV_RANGE("hp_field_release");
    if (field == nullptr) return;
    if (field->ctx != nullptr && field->ctx->ready) {
        DeviceScope scope;
        scope.enter(field->ctx);
        if (field->alias_of != nullptr) {   // a view of a packed grid: tell the grid it is gone
            auto& v = field->alias_of->views;
            v.erase(std::remove(v.begin(), v.end(), field), v.end());
        }
        if (field->owns_data) cudaFree(field->d_data);
    }
    ctx_unref(field->ctx);
    delete field;
}

}  // extern "C"
