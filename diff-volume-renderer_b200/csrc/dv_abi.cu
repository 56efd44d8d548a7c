// dv_abi.cu -- the compute entry points of hp.h on the GPU.
//
// Each entry point dispatches on one tensor's memspace like the reference
// (reference hp_runtime.cpp:171,196,220,245,388; api_ray.cpp:25):
//   DEVICE  buffers are used in place; outputs with .data == NULL come out of
//           the caller's device workspace (4-byte bump allocation, reference
//           order), hp_diff excepted (see there).
//   HOST    inputs are staged into HBM, the same kernels run, results are
//           copied back into the caller's buffers / host workspace.  There is
//           no CPU compute path.
// All entry points are synchronous with respect to the host, as the reference's
// are (blocking cudaMemcpy / cudaDeviceSynchronize in its CUDA files).
#include <algorithm>
#include <cstring>
#include <new>

#include "dv_objects.h"

using namespace dv;

namespace {

constexpr hp_memspace kDev = HP_MEMSPACE_DEVICE;
constexpr hp_memspace kHost = HP_MEMSPACE_HOST;

size_t dim0(const hp_tensor& t) { return t.rank >= 1 && t.shape[0] > 0 ? static_cast<size_t>(t.shape[0]) : 0; }

// reference samp_cpu.cpp:48-59
size_t infer_ray_count(const hp_rays_t* rays, const hp_plan_desc& d) {
    if (rays->t_near.rank >= 1 && rays->t_near.shape[0] > 0) return static_cast<size_t>(rays->t_near.shape[0]);
    if (rays->origins.rank >= 2 && rays->origins.shape[0] > 0) return static_cast<size_t>(rays->origins.shape[0]);
    return static_cast<size_t>(d.roi.width) * d.roi.height;
}

hp_status sync_stream(const hp_ctx* ctx) {
    DV_CUDA(cudaStreamSynchronize(ctx->stream));
    return HP_STATUS_SUCCESS;
}

hp_status h2d(const hp_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return HP_STATUS_SUCCESS;
    DV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return HP_STATUS_SUCCESS;
}

hp_status d2h(const hp_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return HP_STATUS_SUCCESS;
    DV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return HP_STATUS_SUCCESS;
}

#define DV_TRY(expr)                                     \
    do {                                                 \
        const hp_status dv_st__ = (expr);                \
        if (dv_st__ != HP_STATUS_SUCCESS) return dv_st__; \
    } while (0)

// Reads and clears the device status word (after a stream sync).
hp_status fetch_status(const hp_ctx* ctx, uint32_t* out) {
    DV_CUDA(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    DV_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = *ctx->h_status;
    return HP_STATUS_SUCCESS;
}

hp_status clear_status(const hp_ctx* ctx) {
    DV_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(uint32_t), ctx->stream));
    return HP_STATUS_SUCCESS;
}

// ---- tensor bundle shaping (what the reference's configure_* helpers write) ---
void shape_rays(hp_rays_t* r, size_t n, hp_memspace ms) {
    shape_tensor(r->origins, HP_DTYPE_F32, ms, 2, static_cast<int64_t>(n), 3);
    shape_tensor(r->directions, HP_DTYPE_F32, ms, 2, static_cast<int64_t>(n), 3);
    shape_tensor(r->t_near, HP_DTYPE_F32, ms, 1, static_cast<int64_t>(n));
    shape_tensor(r->t_far, HP_DTYPE_F32, ms, 1, static_cast<int64_t>(n));
    shape_tensor(r->pixel_ids, HP_DTYPE_U32, ms, 1, static_cast<int64_t>(n));
}

void shape_samp(hp_samp_t* s, size_t m, size_t n_rays, hp_memspace ms) {
    shape_tensor(s->positions, HP_DTYPE_F32, ms, 2, static_cast<int64_t>(m), 3);
    shape_tensor(s->dt, HP_DTYPE_F32, ms, 1, static_cast<int64_t>(m));
    shape_tensor(s->sigma, HP_DTYPE_F32, ms, 1, static_cast<int64_t>(m));
    shape_tensor(s->color, HP_DTYPE_F32, ms, 2, static_cast<int64_t>(m), 3);
    shape_tensor(s->ray_offset, HP_DTYPE_U32, ms, 1, static_cast<int64_t>(n_rays + 1));
}

void shape_intl(hp_intl_t* t, size_t n_rays, size_t m, hp_memspace ms) {
    shape_tensor(t->radiance, HP_DTYPE_F32, ms, 2, static_cast<int64_t>(n_rays), 3);
    shape_tensor(t->transmittance, HP_DTYPE_F32, ms, 1, static_cast<int64_t>(n_rays));
    shape_tensor(t->opacity, HP_DTYPE_F32, ms, 1, static_cast<int64_t>(n_rays));
    shape_tensor(t->depth, HP_DTYPE_F32, ms, 1, static_cast<int64_t>(n_rays));
    shape_tensor(t->aux, HP_DTYPE_F32, ms, 2, static_cast<int64_t>(m), 4);
}

void shape_img(hp_img_t* g, size_t w, size_t h, hp_memspace ms) {
    shape_tensor(g->image, HP_DTYPE_F32, ms, 3, static_cast<int64_t>(h), static_cast<int64_t>(w), 3);
    shape_tensor(g->trans, HP_DTYPE_F32, ms, 2, static_cast<int64_t>(h), static_cast<int64_t>(w));
    shape_tensor(g->opacity, HP_DTYPE_F32, ms, 2, static_cast<int64_t>(h), static_cast<int64_t>(w));
    shape_tensor(g->depth, HP_DTYPE_F32, ms, 2, static_cast<int64_t>(h), static_cast<int64_t>(w));
    shape_tensor(g->hitmask, HP_DTYPE_U32, ms, 2, static_cast<int64_t>(h), static_cast<int64_t>(w));
}

void shape_grads(hp_grads_t* g, size_t m, hp_memspace ms) {
    shape_tensor(g->sigma, HP_DTYPE_F32, ms, 1, static_cast<int64_t>(m));
    shape_tensor(g->color, HP_DTYPE_F32, ms, 2, static_cast<int64_t>(m), 3);
    shape_tensor(g->camera, HP_DTYPE_F32, ms, 2, 3, 4);
}

// ---- workspace hand-out in the reference's order -------------------------------
// reference ray_cpu.cpp:84-116
hp_status alloc_rays(hp_rays_t* r, size_t n, Bump& ws) {
    if (n == 0) return HP_STATUS_SUCCESS;
    if (!r->origins.data) r->origins.data = ws.take(n * 12);
    if (!r->directions.data) r->directions.data = ws.take(n * 12);
    if (!r->t_near.data) r->t_near.data = ws.take(n * 4);
    if (!r->t_far.data) r->t_far.data = ws.take(n * 4);
    if (!r->pixel_ids.data) r->pixel_ids.data = ws.take(n * 4);
    if (!r->origins.data || !r->directions.data || !r->t_near.data || !r->t_far.data || !r->pixel_ids.data)
        return HP_STATUS_OUT_OF_MEMORY;
    return HP_STATUS_SUCCESS;
}

// reference samp_cpu.cpp:111-136 (capacity-sized, ray_offset last)
hp_status alloc_samp(hp_samp_t* s, size_t capacity, size_t n_rays, Bump& ws) {
    if (!s->positions.data) s->positions.data = ws.take(capacity * 12);
    if (!s->dt.data) s->dt.data = ws.take(capacity * 4);
    if (!s->sigma.data) s->sigma.data = ws.take(capacity * 4);
    if (!s->color.data) s->color.data = ws.take(capacity * 12);
    if (!s->ray_offset.data) s->ray_offset.data = ws.take((n_rays + 1) * 4);
    if (capacity > 0 && (!s->positions.data || !s->dt.data || !s->sigma.data || !s->color.data))
        return HP_STATUS_OUT_OF_MEMORY;
    if (!s->ray_offset.data) return HP_STATUS_OUT_OF_MEMORY;
    return HP_STATUS_SUCCESS;
}

// reference int_cpu.cpp:66-88
hp_status alloc_intl(hp_intl_t* t, size_t n_rays, size_t m, Bump& ws) {
    if (!t->radiance.data && n_rays) t->radiance.data = ws.take(n_rays * 12);
    if (!t->transmittance.data && n_rays) t->transmittance.data = ws.take(n_rays * 4);
    if (!t->opacity.data && n_rays) t->opacity.data = ws.take(n_rays * 4);
    if (!t->depth.data && n_rays) t->depth.data = ws.take(n_rays * 4);
    if (!t->aux.data && m) t->aux.data = ws.take(m * 16);
    if (n_rays && (!t->radiance.data || !t->transmittance.data || !t->opacity.data || !t->depth.data))
        return HP_STATUS_OUT_OF_MEMORY;
    if (m && !t->aux.data) return HP_STATUS_OUT_OF_MEMORY;
    return HP_STATUS_SUCCESS;
}

// reference img_cpu.cpp:72-96
hp_status alloc_img(hp_img_t* g, size_t pixels, Bump& ws) {
    if (pixels == 0) return HP_STATUS_SUCCESS;
    if (!g->image.data) g->image.data = ws.take(pixels * 12);
    if (!g->trans.data) g->trans.data = ws.take(pixels * 4);
    if (!g->opacity.data) g->opacity.data = ws.take(pixels * 4);
    if (!g->depth.data) g->depth.data = ws.take(pixels * 4);
    if (!g->hitmask.data) g->hitmask.data = ws.take(pixels * 4);
    if (!g->image.data || !g->trans.data || !g->opacity.data || !g->depth.data || !g->hitmask.data)
        return HP_STATUS_OUT_OF_MEMORY;
    return HP_STATUS_SUCCESS;
}

RayArrays ray_arrays(const hp_rays_t& r) {
    RayArrays a;
    a.origins = static_cast<float*>(r.origins.data);
    a.directions = static_cast<float*>(r.directions.data);
    a.t_near = static_cast<float*>(r.t_near.data);
    a.t_far = static_cast<float*>(r.t_far.data);
    a.pixel_ids = static_cast<uint32_t*>(r.pixel_ids.data);
    return a;
}

SampleArrays sample_arrays(const hp_samp_t& s) {
    SampleArrays a;
    a.positions = static_cast<float*>(s.positions.data);
    a.dt = static_cast<float*>(s.dt.data);
    a.ray_offset = static_cast<uint32_t*>(s.ray_offset.data);
    a.sigma = static_cast<float*>(s.sigma.data);
    a.color = static_cast<float*>(s.color.data);
    return a;
}

IntegralArrays integral_arrays(const hp_intl_t& t) {
    IntegralArrays a;
    a.radiance = static_cast<float*>(t.radiance.data);
    a.transmittance = static_cast<float*>(t.transmittance.data);
    a.opacity = static_cast<float*>(t.opacity.data);
    a.depth = static_cast<float*>(t.depth.data);
    a.aux = static_cast<float*>(t.aux.data);
    return a;
}

ImagePlanes image_planes(const hp_img_t& g) {
    ImagePlanes a;
    a.image = static_cast<float*>(g.image.data);
    a.trans = static_cast<float*>(g.trans.data);
    a.opacity = static_cast<float*>(g.opacity.data);
    a.depth = static_cast<float*>(g.depth.data);
    a.hitmask = static_cast<uint32_t*>(g.hitmask.data);
    return a;
}

bool field_ok(const hp_field* f, FieldKind kind) { return f == nullptr || (f->kind == kind && f->d_data != nullptr); }

MarchParams march_params(const hp_plan* plan) {
    FrameParams p = frame_params_from_plan(*plan);
    return p.march;
}

// Device-side sampler core shared by hp_samp and hp_samp_int_fused.
// Phase 1: per-ray counts + offsets (into `offsets`), total to the host.
hp_status sampler_offsets(const hp_plan* plan, const RayArrays& rays, size_t n_rays, uint32_t* offsets,
                          unsigned long long* out_total) {
    const hp_ctx* ctx = plan->ctx;
    DeviceScratch scratch;
    void* tmp = scratch.take(scan_scratch_bytes(static_cast<uint32_t>(n_rays)));
    if (tmp == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    DV_CUDA(launch_count_and_scan(ctx->stream, march_params(plan), rays, static_cast<uint32_t>(n_rays), offsets,
                                  ctx->d_total, tmp));
    DV_CUDA(cudaMemcpyAsync(ctx->h_total, ctx->d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            ctx->stream));
    DV_CUDA(cudaStreamSynchronize(ctx->stream));
    *out_total = *ctx->h_total;
    return HP_STATUS_SUCCESS;
}

}  // namespace

// =============================================================================
// hp_ray  (reference api_ray.cpp:16-38, ray_cpu.cpp:122-229, ray_cuda.cu:178-268)
// =============================================================================
extern "C" HP_API hp_status hp_ray(const hp_plan* plan, const hp_rays_t* override_or_null, hp_rays_t* rays, void* ws,
                                   size_t ws_bytes) {
    DV_RANGE("hp_ray");
    if (rays == nullptr || plan == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    const hp_memspace ms = rays->origins.memspace == kDev ? kDev : kHost;
    const hp_plan_desc& d = plan->desc;
    const uint64_t n64 = static_cast<uint64_t>(d.roi.width) * d.roi.height;
    if (n64 > d.max_rays) return HP_STATUS_INVALID_ARGUMENT;
    const size_t n = static_cast<size_t>(n64);
    shape_rays(rays, n, ms);
    Bump bump(ws, ws_bytes);
    const hp_status alloc = alloc_rays(rays, n, bump);
    if (alloc != HP_STATUS_SUCCESS) {
        // the reference's CUDA path takes no workspace and calls missing buffers invalid (ray_cuda.cu:213-217)
        return (ms == kDev && ws == nullptr) ? HP_STATUS_INVALID_ARGUMENT : alloc;
    }
    if (n == 0) return HP_STATUS_SUCCESS;

    if (override_or_null != nullptr) {
        const hp_rays_t& o = *override_or_null;
        if (ms == kHost) {
            if (o.origins.memspace != kHost || o.directions.memspace != kHost || o.t_near.memspace != kHost ||
                o.t_far.memspace != kHost || o.pixel_ids.memspace != kHost)
                return HP_STATUS_UNSUPPORTED;
        }
        if (!o.origins.data || !o.directions.data || !o.t_near.data || !o.t_far.data || !o.pixel_ids.data)
            return HP_STATUS_INVALID_ARGUMENT;
        if (ms == kHost) {  // a plain copy, no arithmetic (ray_cpu.cpp:36-42)
            std::memcpy(rays->origins.data, o.origins.data, n * 12);
            std::memcpy(rays->directions.data, o.directions.data, n * 12);
            std::memcpy(rays->t_near.data, o.t_near.data, n * 4);
            std::memcpy(rays->t_far.data, o.t_far.data, n * 4);
            std::memcpy(rays->pixel_ids.data, o.pixel_ids.data, n * 4);
            return HP_STATUS_SUCCESS;
        }
        DV_ENTER(plan->ctx);
        cudaStream_t s = plan->ctx->stream;
        DV_CUDA(cudaMemcpyAsync(rays->origins.data, o.origins.data, n * 12, cudaMemcpyDefault, s));
        DV_CUDA(cudaMemcpyAsync(rays->directions.data, o.directions.data, n * 12, cudaMemcpyDefault, s));
        DV_CUDA(cudaMemcpyAsync(rays->t_near.data, o.t_near.data, n * 4, cudaMemcpyDefault, s));
        DV_CUDA(cudaMemcpyAsync(rays->t_far.data, o.t_far.data, n * 4, cudaMemcpyDefault, s));
        DV_CUDA(cudaMemcpyAsync(rays->pixel_ids.data, o.pixel_ids.data, n * 4, cudaMemcpyDefault, s));
        return sync_stream(plan->ctx);
    }

    DV_ENTER(plan->ctx);
    const hp_ctx* ctx = plan->ctx;
    const FrameParams fp = frame_params_from_plan(*plan);
    if (ms == kDev) {
        DV_CUDA(launch_rays(ctx->stream, fp, ray_arrays(*rays), static_cast<uint32_t>(n)));
        return sync_stream(ctx);
    }
    DeviceScratch scratch;
    RayArrays dev;
    dev.origins = static_cast<float*>(scratch.take(n * 12));
    dev.directions = static_cast<float*>(scratch.take(n * 12));
    dev.t_near = static_cast<float*>(scratch.take(n * 4));
    dev.t_far = static_cast<float*>(scratch.take(n * 4));
    dev.pixel_ids = static_cast<uint32_t*>(scratch.take(n * 4));
    if (!dev.origins || !dev.directions || !dev.t_near || !dev.t_far || !dev.pixel_ids) return HP_STATUS_OUT_OF_MEMORY;
    DV_CUDA(launch_rays(ctx->stream, fp, dev, static_cast<uint32_t>(n)));
    DV_TRY(d2h(ctx, rays->origins.data, dev.origins, n * 12));
    DV_TRY(d2h(ctx, rays->directions.data, dev.directions, n * 12));
    DV_TRY(d2h(ctx, rays->t_near.data, dev.t_near, n * 4));
    DV_TRY(d2h(ctx, rays->t_far.data, dev.t_far, n * 4));
    DV_TRY(d2h(ctx, rays->pixel_ids.data, dev.pixel_ids, n * 4));
    return sync_stream(ctx);
}

// =============================================================================
// hp_samp / hp_samp_int_fused
// (reference hp_runtime.cpp:160-185,376-400; samp_cpu.cpp:151-313,433-548;
//  samp_int_fused.cpp:11-76; samp_int_fused.cu:10-62)
// =============================================================================
namespace {

hp_status sampler_entry(const hp_plan* plan, const hp_field* fs, const hp_field* fc, const hp_rays_t* rays,
                        hp_samp_t* samp, hp_intl_t* intl /* null: hp_samp */, void* ws, size_t ws_bytes) {
    const bool fused = intl != nullptr;
    if (fused && (ws == nullptr || ws_bytes == 0)) return HP_STATUS_INVALID_ARGUMENT;
    if (fs == nullptr && fc == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    const hp_memspace ms = rays->origins.memspace == kDev ? kDev : kHost;
    auto vec3_ok = [&](const hp_tensor& t) {
        if (ms == kHost) return t.memspace == kHost && t.rank != 0;   // samp_cpu.cpp:37-45
        return t.memspace == kDev && t.dtype == HP_DTYPE_F32 && t.rank >= 2 && t.shape[1] == 3;  // :406-411
    };
    auto vec1_ok = [&](const hp_tensor& t) {
        if (ms == kHost) return t.memspace == kHost && t.rank != 0;
        return t.memspace == kDev && t.dtype == HP_DTYPE_F32 && t.rank >= 1;
    };
    if (!vec3_ok(rays->origins) || !vec3_ok(rays->directions) || !vec1_ok(rays->t_near) || !vec1_ok(rays->t_far))
        return HP_STATUS_INVALID_ARGUMENT;
    if (!field_ok(fs, FieldKind::kDenseSigma) || !field_ok(fc, FieldKind::kDenseColor))
        return HP_STATUS_INVALID_ARGUMENT;

    const hp_plan_desc& d = plan->desc;
    const size_t n_rays = infer_ray_count(rays, d);
    if (n_rays > d.max_rays) return HP_STATUS_INVALID_ARGUMENT;
    const size_t capacity = d.max_samples;
    if (capacity == 0 && n_rays > 0) return HP_STATUS_INVALID_ARGUMENT;

    shape_samp(samp, capacity, n_rays, ms);
    Bump bump(ws, ws_bytes);
    DV_TRY(alloc_samp(samp, capacity, n_rays, bump));
    if (!rays->origins.data || !rays->directions.data || !rays->t_near.data || !rays->t_far.data)
        return HP_STATUS_INVALID_ARGUMENT;

    DV_ENTER(plan->ctx);
    const hp_ctx* ctx = plan->ctx;
    DeviceScratch scratch;
    const MarchParams mp = march_params(plan);
    FieldPair fields = field_pair(fs, fc);

    // rays on the device
    RayArrays dr = ray_arrays(*rays);
    if (ms == kHost) {
        dr.origins = static_cast<float*>(scratch.take(n_rays * 12));
        dr.directions = static_cast<float*>(scratch.take(n_rays * 12));
        dr.t_near = static_cast<float*>(scratch.take(n_rays * 4));
        dr.t_far = static_cast<float*>(scratch.take(n_rays * 4));
        dr.pixel_ids = nullptr;
        if (!dr.origins || !dr.directions || !dr.t_near || !dr.t_far) return HP_STATUS_OUT_OF_MEMORY;
        DV_TRY(h2d(ctx, dr.origins, rays->origins.data, n_rays * 12));
        DV_TRY(h2d(ctx, dr.directions, rays->directions.data, n_rays * 12));
        DV_TRY(h2d(ctx, dr.t_near, rays->t_near.data, n_rays * 4));
        DV_TRY(h2d(ctx, dr.t_far, rays->t_far.data, n_rays * 4));
    }

    // phase 1: offsets and the total, which the host needs for the capacity check and the shapes
    uint32_t* d_offsets = ms == kDev ? static_cast<uint32_t*>(samp->ray_offset.data)
                                     : static_cast<uint32_t*>(scratch.take((n_rays + 1) * 4));
    if (d_offsets == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    unsigned long long total = 0;
    DV_TRY(sampler_offsets(plan, dr, n_rays, d_offsets, &total));
    if (total > capacity) return HP_STATUS_INVALID_ARGUMENT;   // samp_cpu.cpp:245-247
    const size_t m = static_cast<size_t>(total);

    // where the integrator's outputs go (fused only)
    IntegralArrays di{};
    char* intl_base_host = nullptr;
    if (fused) {
        const char* base = static_cast<const char*>(ws);
        const char* top = base;
        auto bump_top = [&](const void* p, size_t bytes) {
            if (p != nullptr && bytes > 0) {
                const char* e = static_cast<const char*>(p) + bytes;
                if (e > top) top = e;
            }
        };
        // CPU flavour measures the used part (samp_int_fused.cpp:37-56), CUDA flavour the capacity (.cu:46-50)
        const size_t cnt = ms == kHost ? m : capacity;
        bump_top(samp->positions.data, cnt * 12);
        bump_top(samp->dt.data, cnt * 4);
        bump_top(samp->sigma.data, cnt * 4);
        bump_top(samp->color.data, cnt * 12);
        bump_top(samp->ray_offset.data, (n_rays + 1) * 4);
        const size_t used = static_cast<size_t>(top - base);
        if (used > ws_bytes) return HP_STATUS_INTERNAL_ERROR;
        const size_t rest = ws_bytes - used;
        if (ms == kHost && rest == 0) return HP_STATUS_OUT_OF_MEMORY;
        intl_base_host = const_cast<char*>(base) + used;
        shape_intl(intl, n_rays, m, ms);
        Bump ibump(intl_base_host, rest);
        DV_TRY(alloc_intl(intl, n_rays, m, ibump));
    }

    // phase 2: fill
    SampleArrays ds = sample_arrays(*samp);
    if (ms == kHost) {
        ds.positions = static_cast<float*>(scratch.take(m * 12));
        ds.dt = static_cast<float*>(scratch.take(m * 4));
        ds.sigma = static_cast<float*>(scratch.take(m * 4));
        ds.color = static_cast<float*>(scratch.take(m * 12));
        if (!ds.positions || !ds.dt || !ds.sigma || !ds.color) return HP_STATUS_OUT_OF_MEMORY;
    }
    ds.ray_offset = d_offsets;
    if (fused) {
        di = integral_arrays(*intl);
        if (ms == kHost) {
            di.radiance = static_cast<float*>(scratch.take(n_rays * 12));
            di.transmittance = static_cast<float*>(scratch.take(n_rays * 4));
            di.opacity = static_cast<float*>(scratch.take(n_rays * 4));
            di.depth = static_cast<float*>(scratch.take(n_rays * 4));
            di.aux = static_cast<float*>(scratch.take(m * 16));
            if (!di.radiance || !di.transmittance || !di.opacity || !di.depth || !di.aux) return HP_STATUS_OUT_OF_MEMORY;
        }
    }
    DV_CUDA(launch_sample(ctx->stream, mp, d.t_near, d.t_far, fields, dr, static_cast<uint32_t>(n_rays), ds, fused, di));

    if (ms == kHost) {
        DV_TRY(d2h(ctx, samp->positions.data, ds.positions, m * 12));
        DV_TRY(d2h(ctx, samp->dt.data, ds.dt, m * 4));
        DV_TRY(d2h(ctx, samp->sigma.data, ds.sigma, m * 4));
        DV_TRY(d2h(ctx, samp->color.data, ds.color, m * 12));
        DV_TRY(d2h(ctx, samp->ray_offset.data, d_offsets, (n_rays + 1) * 4));
        if (fused) {
            DV_TRY(d2h(ctx, intl->radiance.data, di.radiance, n_rays * 12));
            DV_TRY(d2h(ctx, intl->transmittance.data, di.transmittance, n_rays * 4));
            DV_TRY(d2h(ctx, intl->opacity.data, di.opacity, n_rays * 4));
            DV_TRY(d2h(ctx, intl->depth.data, di.depth, n_rays * 4));
            DV_TRY(d2h(ctx, intl->aux.data, di.aux, m * 16));
        }
    }
    DV_TRY(sync_stream(ctx));

    const int64_t m64 = static_cast<int64_t>(m);
    samp->positions.shape[0] = m64;
    samp->dt.shape[0] = m64;
    samp->sigma.shape[0] = m64;
    samp->color.shape[0] = m64;
    samp->ray_offset.shape[0] = static_cast<int64_t>(n_rays + 1);
    return HP_STATUS_SUCCESS;
}

}  // namespace

extern "C" HP_API hp_status hp_samp(const hp_plan* plan, const hp_field* fs, const hp_field* fc, const hp_rays_t* rays,
                                    hp_samp_t* samp, void* ws, size_t ws_bytes) {
    DV_RANGE("hp_samp");
    if (plan == nullptr || rays == nullptr || samp == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    return sampler_entry(plan, fs, fc, rays, samp, nullptr, ws, ws_bytes);
}

extern "C" HP_API hp_status hp_samp_int_fused(const hp_plan* plan, const hp_field* fs, const hp_field* fc,
                                              const hp_rays_t* rays, hp_samp_t* samp, hp_intl_t* intl, void* ws,
                                              size_t ws_bytes) {
    DV_RANGE("hp_samp_int_fused");
    if (plan == nullptr || rays == nullptr || samp == nullptr || intl == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    return sampler_entry(plan, fs, fc, rays, samp, intl, ws, ws_bytes);
}

// =============================================================================
// hp_int  (reference hp_runtime.cpp:187-208, int_cpu.cpp:115-230, int_cuda.cu:116-232)
// =============================================================================
extern "C" HP_API hp_status hp_int(const hp_plan* plan, const hp_samp_t* samp, hp_intl_t* intl, void* ws,
                                   size_t ws_bytes) {
    DV_RANGE("hp_int");
    if (plan == nullptr || samp == nullptr || intl == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    const hp_memspace ms = samp->sigma.memspace == kDev ? kDev : kHost;
    if (samp->dt.memspace != ms || samp->sigma.memspace != ms || samp->ray_offset.memspace != ms ||
        samp->color.memspace != ms)
        return HP_STATUS_INVALID_ARGUMENT;
    if (samp->dt.dtype != HP_DTYPE_F32 || samp->sigma.dtype != HP_DTYPE_F32 || samp->color.dtype != HP_DTYPE_F32 ||
        samp->ray_offset.dtype != HP_DTYPE_U32)
        return HP_STATUS_INVALID_ARGUMENT;
    const size_t m = samp->dt.rank >= 1 ? static_cast<size_t>(samp->dt.shape[0]) : 0;
    const size_t n_rays = samp->ray_offset.rank >= 1 && samp->ray_offset.shape[0] > 0
                              ? static_cast<size_t>(samp->ray_offset.shape[0] - 1) : 0;
    if (m > plan->desc.max_samples || n_rays > plan->desc.max_rays) return HP_STATUS_INVALID_ARGUMENT;
    if ((m > 0 && (!samp->dt.data || !samp->sigma.data || !samp->color.data)) ||
        (n_rays > 0 && !samp->ray_offset.data))
        return HP_STATUS_INVALID_ARGUMENT;
    shape_intl(intl, n_rays, m, ms);
    Bump bump(ws, ws_bytes);
    DV_TRY(alloc_intl(intl, n_rays, m, bump));
    if (n_rays == 0) return HP_STATUS_SUCCESS;

    DV_ENTER(plan->ctx);
    const hp_ctx* ctx = plan->ctx;
    DeviceScratch scratch;
    SampleArrays ds = sample_arrays(*samp);
    IntegralArrays di = integral_arrays(*intl);
    if (ms == kHost) {
        ds.positions = nullptr;
        ds.dt = static_cast<float*>(scratch.take(m * 4));
        ds.sigma = static_cast<float*>(scratch.take(m * 4));
        ds.color = static_cast<float*>(scratch.take(m * 12));
        ds.ray_offset = static_cast<uint32_t*>(scratch.take((n_rays + 1) * 4));
        di.radiance = static_cast<float*>(scratch.take(n_rays * 12));
        di.transmittance = static_cast<float*>(scratch.take(n_rays * 4));
        di.opacity = static_cast<float*>(scratch.take(n_rays * 4));
        di.depth = static_cast<float*>(scratch.take(n_rays * 4));
        di.aux = static_cast<float*>(scratch.take(m * 16));
        if (!ds.dt || !ds.sigma || !ds.color || !ds.ray_offset || !di.radiance || !di.transmittance || !di.opacity ||
            !di.depth || !di.aux)
            return HP_STATUS_OUT_OF_MEMORY;
        DV_TRY(h2d(ctx, ds.dt, samp->dt.data, m * 4));
        DV_TRY(h2d(ctx, ds.sigma, samp->sigma.data, m * 4));
        DV_TRY(h2d(ctx, ds.color, samp->color.data, m * 12));
        DV_TRY(h2d(ctx, ds.ray_offset, samp->ray_offset.data, (n_rays + 1) * 4));
    }
    DV_TRY(clear_status(ctx));
    DV_CUDA(launch_integrate(ctx->stream, plan->desc.t_near, plan->desc.t_far, ds, static_cast<uint32_t>(n_rays),
                             static_cast<uint32_t>(m), di, ctx->d_status));
    if (ms == kHost) {
        DV_TRY(d2h(ctx, intl->radiance.data, di.radiance, n_rays * 12));
        DV_TRY(d2h(ctx, intl->transmittance.data, di.transmittance, n_rays * 4));
        DV_TRY(d2h(ctx, intl->opacity.data, di.opacity, n_rays * 4));
        DV_TRY(d2h(ctx, intl->depth.data, di.depth, n_rays * 4));
        DV_TRY(d2h(ctx, intl->aux.data, di.aux, m * 16));
    }
    uint32_t status = 0;
    DV_TRY(fetch_status(ctx, &status));
    if (status & kErrBadOffsets) return HP_STATUS_INVALID_ARGUMENT;   // int_cpu.cpp:176-178
    intl->aux.shape[0] = static_cast<int64_t>(m);
    return HP_STATUS_SUCCESS;
}

// =============================================================================
// hp_img  (reference hp_runtime.cpp:210-232, img_cpu.cpp:110-188, img_cuda.cu:112-210)
// =============================================================================
extern "C" HP_API hp_status hp_img(const hp_plan* plan, const hp_intl_t* intl, const hp_rays_t* rays, hp_img_t* img,
                                   void* ws, size_t ws_bytes) {
    DV_RANGE("hp_img");
    if (plan == nullptr || intl == nullptr || rays == nullptr || img == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    const hp_memspace ms = rays->pixel_ids.memspace == kDev ? kDev : kHost;
    if (intl->radiance.memspace != ms || intl->transmittance.memspace != ms || intl->opacity.memspace != ms ||
        intl->depth.memspace != ms)
        return HP_STATUS_INVALID_ARGUMENT;
    if (rays->pixel_ids.dtype != HP_DTYPE_U32 || intl->radiance.dtype != HP_DTYPE_F32 ||
        intl->transmittance.dtype != HP_DTYPE_F32 || intl->opacity.dtype != HP_DTYPE_F32 ||
        intl->depth.dtype != HP_DTYPE_F32)
        return HP_STATUS_INVALID_ARGUMENT;
    const size_t n_rays = intl->transmittance.rank >= 1 ? static_cast<size_t>(intl->transmittance.shape[0]) : 0;
    const size_t w = plan->desc.width, h = plan->desc.height, pixels = w * h;
    shape_img(img, w, h, ms);
    Bump bump(ws, ws_bytes);
    DV_TRY(alloc_img(img, pixels, bump));
    if (n_rays > 0 && (!intl->radiance.data || !intl->transmittance.data || !intl->opacity.data || !intl->depth.data))
        return HP_STATUS_INVALID_ARGUMENT;

    DV_ENTER(plan->ctx);
    const hp_ctx* ctx = plan->ctx;
    DeviceScratch scratch;
    ImagePlanes dimg = image_planes(*img);
    IntegralArrays di = integral_arrays(*intl);
    const uint32_t* d_pix = static_cast<const uint32_t*>(rays->pixel_ids.data);
    if (ms == kHost) {
        dimg.image = static_cast<float*>(scratch.take(pixels * 12));
        dimg.trans = static_cast<float*>(scratch.take(pixels * 4));
        dimg.opacity = static_cast<float*>(scratch.take(pixels * 4));
        dimg.depth = static_cast<float*>(scratch.take(pixels * 4));
        dimg.hitmask = static_cast<uint32_t*>(scratch.take(pixels * 4));
        di.radiance = static_cast<float*>(scratch.take(n_rays * 12));
        di.transmittance = static_cast<float*>(scratch.take(n_rays * 4));
        di.opacity = static_cast<float*>(scratch.take(n_rays * 4));
        di.depth = static_cast<float*>(scratch.take(n_rays * 4));
        if (!dimg.image || !dimg.trans || !dimg.opacity || !dimg.depth || !dimg.hitmask || !di.radiance ||
            !di.transmittance || !di.opacity || !di.depth)
            return HP_STATUS_OUT_OF_MEMORY;
        DV_TRY(h2d(ctx, di.radiance, intl->radiance.data, n_rays * 12));
        DV_TRY(h2d(ctx, di.transmittance, intl->transmittance.data, n_rays * 4));
        DV_TRY(h2d(ctx, di.opacity, intl->opacity.data, n_rays * 4));
        DV_TRY(h2d(ctx, di.depth, intl->depth.data, n_rays * 4));
        if (d_pix != nullptr) {
            uint32_t* p = static_cast<uint32_t*>(scratch.take(n_rays * 4));
            if (!p) return HP_STATUS_OUT_OF_MEMORY;
            DV_TRY(h2d(ctx, p, rays->pixel_ids.data, n_rays * 4));
            d_pix = p;
        }
    }
    DV_TRY(clear_status(ctx));
    DV_CUDA(launch_background(ctx->stream, dimg, pixels, plan->desc.t_far));
    DV_CUDA(launch_compose(ctx->stream, dimg, pixels, d_pix, di, static_cast<uint32_t>(n_rays), ctx->d_status));
    uint32_t status = 0;
    DV_TRY(fetch_status(ctx, &status));
    if (status & kErrBadPixel) return HP_STATUS_INVALID_ARGUMENT;   // img_cpu.cpp:156-158
    if (status & kFlagDuplicatePixel) {
        // repeated pixel ids: redo in ray order (first hit writes, later hits accumulate)
        DV_CUDA(launch_background(ctx->stream, dimg, pixels, plan->desc.t_far));
        DV_CUDA(launch_compose_sequential(ctx->stream, dimg, pixels, d_pix, di, static_cast<uint32_t>(n_rays)));
    }
    if (ms == kHost) {
        DV_TRY(d2h(ctx, img->image.data, dimg.image, pixels * 12));
        DV_TRY(d2h(ctx, img->trans.data, dimg.trans, pixels * 4));
        DV_TRY(d2h(ctx, img->opacity.data, dimg.opacity, pixels * 4));
        DV_TRY(d2h(ctx, img->depth.data, dimg.depth, pixels * 4));
        DV_TRY(d2h(ctx, img->hitmask.data, dimg.hitmask, pixels * 4));
    }
    return sync_stream(ctx);
}

// =============================================================================
// hp_diff  (reference hp_runtime.cpp:234-257, diff_cpu.cpp:89-198, diff_cuda.cu:69-224)
// =============================================================================
extern "C" HP_API hp_status hp_diff(const hp_plan* plan, const hp_tensor* dL_dI, const hp_samp_t* samp,
                                    const hp_intl_t* intl, hp_grads_t* grads, void* ws, size_t ws_bytes) {
    DV_RANGE("hp_diff");
    if (plan == nullptr || dL_dI == nullptr || samp == nullptr || intl == nullptr || grads == nullptr)
        return HP_STATUS_INVALID_ARGUMENT;
    const hp_memspace ms = dL_dI->memspace == kDev ? kDev : kHost;
    if (samp->dt.memspace != ms || samp->sigma.memspace != ms || samp->color.memspace != ms ||
        samp->ray_offset.memspace != ms)
        return HP_STATUS_INVALID_ARGUMENT;
    if (ms == kHost) {  // diff_cpu.cpp:106-109
        if (intl->radiance.memspace != kHost || intl->transmittance.memspace != kHost ||
            intl->opacity.memspace != kHost || intl->depth.memspace != kHost)
            return HP_STATUS_INVALID_ARGUMENT;
    } else if (intl->aux.memspace != kDev) {  // diff_cuda.cu:86
        return HP_STATUS_INVALID_ARGUMENT;
    }
    if (dL_dI->dtype != HP_DTYPE_F32 || samp->dt.dtype != HP_DTYPE_F32 || samp->sigma.dtype != HP_DTYPE_F32 ||
        samp->color.dtype != HP_DTYPE_F32 || samp->ray_offset.dtype != HP_DTYPE_U32 || intl->aux.dtype != HP_DTYPE_F32)
        return HP_STATUS_INVALID_ARGUMENT;
    const size_t m = samp->dt.rank >= 1 ? static_cast<size_t>(samp->dt.shape[0]) : 0;
    const size_t n_rays = samp->ray_offset.rank >= 1 && samp->ray_offset.shape[0] > 0
                              ? static_cast<size_t>(samp->ray_offset.shape[0] - 1) : 0;
    if (m > plan->desc.max_samples || n_rays > plan->desc.max_rays) return HP_STATUS_INVALID_ARGUMENT;
    if (ms == kDev && (m == 0 || n_rays == 0)) return HP_STATUS_SUCCESS;   // diff_cuda.cu:112-114

    dv::DeviceScope dv_scope__;
    if (ms == kDev || m > 0) DV_TRY(dv_scope__.enter(plan->ctx));
    const hp_ctx* ctx = plan->ctx;

    float *g_sigma = nullptr, *g_color = nullptr, *g_camera = nullptr;
    DeviceScratch scratch;
    if (ms == kHost) {
        shape_grads(grads, m, kHost);
        Bump bump(ws, ws_bytes);   // diff_cpu.cpp:53-67
        if (!grads->sigma.data && m) grads->sigma.data = bump.take(m * 4);
        if (!grads->color.data && m) grads->color.data = bump.take(m * 12);
        if (!grads->camera.data) grads->camera.data = bump.take(12 * 4);
        if ((m && (!grads->sigma.data || !grads->color.data)) || !grads->camera.data) return HP_STATUS_OUT_OF_MEMORY;
        std::memset(grads->camera.data, 0, 12 * sizeof(float));   // camera gradients are zero in this ABI (:73-74)
        if (m) {
            std::memset(grads->sigma.data, 0, m * 4);
            std::memset(grads->color.data, 0, m * 12);
        }
        if (m == 0 || n_rays == 0) return HP_STATUS_SUCCESS;
    } else {
        // The reference ignores `ws`, cudaMallocs the three outputs and leaves the cudaFree to the caller
        // (diff_cuda.cu:116-167).  Kept for buffers the caller did not provide; provided buffers are used.
        void* p = nullptr;
        if (!grads->sigma.data) { DV_CUDA(cudaMalloc(&p, m * 4)); grads->sigma.data = p; }
        if (!grads->color.data) { DV_CUDA(cudaMalloc(&p, m * 12)); grads->color.data = p; }
        if (!grads->camera.data) { DV_CUDA(cudaMalloc(&p, 12 * 4)); grads->camera.data = p; }
        shape_grads(grads, m, kDev);
    }
    if (intl->aux.data == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (dL_dI->rank < 2 || dL_dI->shape[0] != static_cast<int64_t>(n_rays) || dL_dI->shape[1] < 3)
        return HP_STATUS_INVALID_ARGUMENT;
    const int64_t stride_ray = dL_dI->stride[0], stride_c = dL_dI->stride[1];

    SampleArrays ds = sample_arrays(*samp);
    const float* d_aux = static_cast<const float*>(intl->aux.data);
    const float* d_g = static_cast<const float*>(dL_dI->data);
    int64_t sr = stride_ray, sc = stride_c;
    if (ms == kHost) {
        ds.positions = nullptr;
        ds.sigma = nullptr;
        ds.dt = static_cast<float*>(scratch.take(m * 4));
        ds.color = static_cast<float*>(scratch.take(m * 12));
        ds.ray_offset = static_cast<uint32_t*>(scratch.take((n_rays + 1) * 4));
        float* aux = static_cast<float*>(scratch.take(m * 16));
        float* g = static_cast<float*>(scratch.take(n_rays * 12));
        g_sigma = static_cast<float*>(scratch.take(m * 4));
        g_color = static_cast<float*>(scratch.take(m * 12));
        if (!ds.dt || !ds.color || !ds.ray_offset || !aux || !g || !g_sigma || !g_color) return HP_STATUS_OUT_OF_MEMORY;
        DV_TRY(h2d(ctx, ds.dt, samp->dt.data, m * 4));
        DV_TRY(h2d(ctx, ds.color, samp->color.data, m * 12));
        DV_TRY(h2d(ctx, ds.ray_offset, samp->ray_offset.data, (n_rays + 1) * 4));
        DV_TRY(h2d(ctx, aux, intl->aux.data, m * 16));
        // gather the strided host gradient into a dense (rays,3) block
        std::vector<float> dense(n_rays * 3);
        const float* hg = static_cast<const float*>(dL_dI->data);
        for (size_t r = 0; r < n_rays; ++r)
            for (int c = 0; c < 3; ++c) dense[3 * r + c] = hg[static_cast<int64_t>(r) * stride_ray + c * stride_c];
        DV_TRY(h2d(ctx, g, dense.data(), n_rays * 12));
        DV_TRY(sync_stream(ctx));   // `dense` leaves scope below
        d_aux = aux;
        d_g = g;
        sr = 3;
        sc = 1;
    } else {
        g_sigma = static_cast<float*>(grads->sigma.data);
        g_color = static_cast<float*>(grads->color.data);
        g_camera = static_cast<float*>(grads->camera.data);
        DV_CUDA(cudaMemsetAsync(g_camera, 0, 12 * 4, ctx->stream));
    }
    DV_CUDA(cudaMemsetAsync(g_sigma, 0, m * 4, ctx->stream));
    DV_CUDA(cudaMemsetAsync(g_color, 0, m * 12, ctx->stream));
    DV_TRY(clear_status(ctx));
    DV_CUDA(launch_diff(ctx->stream, d_g, sr, sc, ds, d_aux, static_cast<uint32_t>(n_rays), static_cast<uint32_t>(m),
                        g_sigma, g_color, ctx->d_status));
    if (ms == kHost) {
        DV_TRY(d2h(ctx, grads->sigma.data, g_sigma, m * 4));
        DV_TRY(d2h(ctx, grads->color.data, g_color, m * 12));
    }
    uint32_t status = 0;
    DV_TRY(fetch_status(ctx, &status));
    if (status & kErrBadOffsets) return HP_STATUS_INVALID_ARGUMENT;   // diff_cpu.cpp:159-161
    return HP_STATUS_SUCCESS;
}
