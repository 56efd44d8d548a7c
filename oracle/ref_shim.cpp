// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" door into the UNMODIFIED reference's C++ surface, compiled
// together with the reference's own sources (where they lie under
// /root/reference) into oracle/_ref/libdvren_ref.so by oracle/Makefile.  It
// contains no algorithm: it only forwards to dvren::DenseGridField and
// dvren::Renderer (reference include/dvren/fields/dense_grid.hpp:24-75,
// include/dvren/render/renderer.hpp:68-149) so Python can reach them via ctypes.
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <new>
#include <span>
#include <vector>

#include "dvren/core/context.hpp"
#include "dvren/core/plan.hpp"
#include "dvren/core/tensor_utils.hpp"
#include "dvren/fields/dense_grid.hpp"
#include "dvren/render/renderer.hpp"

namespace {

dvren::PlanDescriptor ToDescriptor(const hp_plan_desc& d) {
    dvren::PlanDescriptor p{};
    p.width = d.width;
    p.height = d.height;
    p.t_near = d.t_near;
    p.t_far = d.t_far;
    p.sampling.dt = d.sampling.dt;
    p.sampling.max_steps = d.sampling.max_steps;
    p.sampling.mode = d.sampling.mode == HP_SAMPLING_STRATIFIED ? dvren::SamplingMode::kStratified
                                                                : dvren::SamplingMode::kFixed;
    if (d.roi.width != 0 && d.roi.height != 0) {
        p.roi = dvren::Roi{d.roi.x, d.roi.y, d.roi.width, d.roi.height};
    }
    p.max_rays = d.max_rays;
    p.max_samples = d.max_samples;
    p.seed = d.seed;
    p.camera.model = d.camera.model == HP_CAMERA_ORTHOGRAPHIC ? dvren::CameraModel::kOrthographic
                                                              : dvren::CameraModel::kPinhole;
    std::memcpy(p.camera.K.data(), d.camera.K, sizeof(d.camera.K));
    std::memcpy(p.camera.c2w.data(), d.camera.c2w, sizeof(d.camera.c2w));
    p.camera.ortho_scale = d.camera.ortho_scale;
    return p;
}

double MsSince(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

extern "C" {

// DenseGridField::AccumulateSampleGradients on a zero-valued grid of `res`.
int ref_scatter(const int32_t res[3], const float bbox_min[3], const float bbox_max[3],
                uint32_t interp, uint32_t oob, size_t n_samples, const float* positions,
                const float* grad_sigma, const float* grad_color, float* sigma_grad,
                float* color_grad) {
    dvren::Context ctx;
    if (!dvren::Context::Create({}, ctx).ok()) return -1;
    dvren::DenseGridConfig cfg{};
    cfg.resolution = {res[0], res[1], res[2]};
    const size_t voxels = static_cast<size_t>(res[0]) * res[1] * res[2];
    cfg.sigma.assign(voxels, 0.0f);
    cfg.color.assign(voxels * 3, 0.0f);
    cfg.bbox_min = {bbox_min[0], bbox_min[1], bbox_min[2]};
    cfg.bbox_max = {bbox_max[0], bbox_max[1], bbox_max[2]};
    cfg.interp = static_cast<hp_interp_mode>(interp);
    cfg.oob = static_cast<hp_oob_policy>(oob);
    dvren::DenseGridField field;
    if (!dvren::DenseGridField::Create(ctx, cfg, field).ok()) return -2;
    hp_samp_t samp{};
    samp.positions.data = const_cast<float*>(positions);
    samp.positions.dtype = HP_DTYPE_F32;
    samp.positions.memspace = HP_MEMSPACE_HOST;
    samp.positions.rank = 2;
    samp.positions.shape[0] = static_cast<int64_t>(n_samples);
    samp.positions.shape[1] = 3;
    samp.positions.stride[0] = 3;
    samp.positions.stride[1] = 1;
    samp.dt.dtype = HP_DTYPE_F32;
    samp.dt.memspace = HP_MEMSPACE_HOST;
    field.ZeroGradients();
    const dvren::Status st = field.AccumulateSampleGradients(
        samp, std::span<const float>(grad_sigma, n_samples),
        std::span<const float>(grad_color, n_samples * 3));
    if (!st.ok()) return -3;
    std::memcpy(sigma_grad, field.sigma_gradients().data(), voxels * sizeof(float));
    std::memcpy(color_grad, field.color_gradients().data(), voxels * 3 * sizeof(float));
    return 0;
}

// Context -> Plan -> DenseGridField -> Renderer::Forward [-> Renderer::Backward].
// Per-pixel outputs are full frames; dL_dI is (rays,3) per ray or NULL.
// Returns 0 or a negative stage code; timings are the reference calls only.
int ref_render(const hp_plan_desc* desc, const int32_t res[3], const float* sigma,
               const float* color, const float bbox_min[3], const float bbox_max[3],
               uint32_t interp, uint32_t oob, int use_fused, const float* dL_dI, float* image,
               float* trans, float* opacity, float* depth, uint32_t* hitmask, float* sigma_grad,
               float* color_grad, float* camera_grad12, uint64_t* ray_count,
               uint64_t* sample_count, double* forward_ms, double* backward_ms) {
    dvren::Context ctx;
    if (!dvren::Context::Create({}, ctx).ok()) return -1;
    dvren::Plan plan;
    if (!dvren::Plan::Create(ctx, ToDescriptor(*desc), plan).ok()) return -2;
    dvren::DenseGridConfig cfg{};
    cfg.resolution = {res[0], res[1], res[2]};
    const size_t voxels = static_cast<size_t>(res[0]) * res[1] * res[2];
    cfg.sigma.assign(sigma, sigma + voxels);
    cfg.color.assign(color, color + voxels * 3);
    cfg.bbox_min = {bbox_min[0], bbox_min[1], bbox_min[2]};
    cfg.bbox_max = {bbox_max[0], bbox_max[1], bbox_max[2]};
    cfg.interp = static_cast<hp_interp_mode>(interp);
    cfg.oob = static_cast<hp_oob_policy>(oob);
    dvren::DenseGridField field;
    if (!dvren::DenseGridField::Create(ctx, cfg, field).ok()) return -3;
    dvren::RenderOptions opt{};
    opt.use_fused_path = use_fused != 0;
    dvren::Renderer renderer(ctx, plan, opt);
    dvren::ForwardResult fwd;
    auto t0 = std::chrono::steady_clock::now();
    if (!renderer.Forward(field, fwd).ok()) return -4;
    if (forward_ms) *forward_ms = MsSince(t0);
    if (ray_count) *ray_count = fwd.ray_count;
    if (sample_count) *sample_count = fwd.sample_count;
    if (image) std::memcpy(image, fwd.image.data(), fwd.image.size() * sizeof(float));
    if (trans) std::memcpy(trans, fwd.transmittance.data(), fwd.transmittance.size() * sizeof(float));
    if (opacity) std::memcpy(opacity, fwd.opacity.data(), fwd.opacity.size() * sizeof(float));
    if (depth) std::memcpy(depth, fwd.depth.data(), fwd.depth.size() * sizeof(float));
    if (hitmask) std::memcpy(hitmask, fwd.hitmask.data(), fwd.hitmask.size() * sizeof(uint32_t));
    if (dL_dI != nullptr) {
        dvren::BackwardResult bwd;
        t0 = std::chrono::steady_clock::now();
        if (!renderer.Backward(field, std::span<const float>(dL_dI, fwd.ray_count * 3), bwd).ok()) return -5;
        if (backward_ms) *backward_ms = MsSince(t0);
        if (sigma_grad) std::memcpy(sigma_grad, bwd.sigma.data(), bwd.sigma.size() * sizeof(float));
        if (color_grad) std::memcpy(color_grad, bwd.color.data(), bwd.color.size() * sizeof(float));
        if (camera_grad12) std::memcpy(camera_grad12, bwd.camera.data(), 12 * sizeof(float));
    }
    return 0;
}

// ---- bench.py --impl reference / cpu_baseline: the hot-path calls themselves, per host thread ----------------------
// One worker per thread: hp.h fields that VIEW the caller's shared grid arrays (reference hp_runtime.cpp:259-339 keeps a
// view, so the threads share one copy of the values), plus a DenseGridField that receives the gradient scatter.  A run is
// the reference's own call sequence for one ROI band -- hp_ray -> hp_samp_int_fused -> hp_img, then hp_diff ->
// DenseGridField::AccumulateSampleGradients (reference src/render/renderer.cpp:259-365,415-427) -- timed WITHOUT
// Renderer's whole-grid bookkeeping (ZeroGradients and the V-sized result copies, renderer.cpp:424,441-442): that cost is
// per call, not per sample, and would dominate a bounded band of a 512^3 grid while it vanishes in the full frame.
struct RefWorker {
    dvren::Context ctx;
    hp_field* fs{nullptr};
    hp_field* fc{nullptr};
    dvren::DenseGridField scatter;
    std::vector<std::byte> ws, ws_img, ws_diff;
    std::vector<float> rays_f;
    std::vector<uint32_t> rays_u;
};

void* ref_worker_create(const int32_t res[3], const float* sigma, const float* color, uint32_t interp, uint32_t oob) {
    RefWorker* w = new (std::nothrow) RefWorker();
    if (w == nullptr) return nullptr;
    if (!dvren::Context::Create({}, w->ctx).ok()) { delete w; return nullptr; }
    const int64_t nx = res[0], ny = res[1], nz = res[2];
    hp_tensor st = dvren::MakeHostTensor(static_cast<void*>(const_cast<float*>(sigma)), HP_DTYPE_F32, {nz, ny, nx});
    hp_tensor ct = dvren::MakeHostTensor(static_cast<void*>(const_cast<float*>(color)), HP_DTYPE_F32, {nz, ny, nx, 3});
    if (hp_field_create_grid_sigma(w->ctx.handle(), &st, interp, oob, &w->fs) != HP_STATUS_SUCCESS ||
        hp_field_create_grid_color(w->ctx.handle(), &ct, interp, oob, &w->fc) != HP_STATUS_SUCCESS) {
        delete w;
        return nullptr;
    }
    dvren::DenseGridConfig cfg{};
    cfg.resolution = {res[0], res[1], res[2]};
    const size_t voxels = static_cast<size_t>(nx) * ny * nz;
    cfg.sigma.assign(voxels, 0.0f);   // the scatter never reads the values (dense_grid.cpp:171-309)
    cfg.color.assign(voxels * 3, 0.0f);
    cfg.interp = static_cast<hp_interp_mode>(interp);
    cfg.oob = static_cast<hp_oob_policy>(oob);
    if (!dvren::DenseGridField::Create(w->ctx, cfg, w->scatter).ok()) { delete w; return nullptr; }
    w->scatter.ZeroGradients();
    return w;
}

void ref_worker_destroy(void* handle) {
    RefWorker* w = static_cast<RefWorker*>(handle);
    if (w == nullptr) return;
    if (w->fs) hp_field_release(w->fs);
    if (w->fc) hp_field_release(w->fc);
    delete w;
}

int ref_worker_run(void* handle, const hp_plan_desc* desc, const float* dL_dI, uint64_t* sample_count,
                   double* forward_ms, double* backward_ms) {
    RefWorker* w = static_cast<RefWorker*>(handle);
    if (w == nullptr || desc == nullptr) return -1;
    hp_plan* plan = nullptr;
    if (hp_plan_create(w->ctx.handle(), desc, &plan) != HP_STATUS_SUCCESS) return -2;
    hp_plan_desc d{};
    hp_plan_get_desc(plan, &d);
    const size_t n = static_cast<size_t>(d.roi.width) * d.roi.height, cap = d.max_samples;
    const size_t pixels = static_cast<size_t>(d.width) * d.height;
    w->rays_f.resize(n * 8);
    w->rays_u.resize(n);
    w->ws.resize(cap * 32 + (n + 1) * 4 + n * 24 + cap * 16 + 256);
    w->ws_img.resize(pixels * 28 + 256);
    hp_rays_t rays{};
    rays.origins.data = w->rays_f.data();
    rays.directions.data = w->rays_f.data() + n * 3;
    rays.t_near.data = w->rays_f.data() + n * 6;
    rays.t_far.data = w->rays_f.data() + n * 7;
    rays.pixel_ids.data = w->rays_u.data();
    for (hp_tensor* t : {&rays.origins, &rays.directions, &rays.t_near, &rays.t_far, &rays.pixel_ids}) t->memspace = HP_MEMSPACE_HOST;
    hp_samp_t samp{};
    hp_intl_t intl{};
    hp_img_t img{};
    int rc = 0;
    auto t0 = std::chrono::steady_clock::now();
    if (hp_ray(plan, nullptr, &rays, nullptr, 0) != HP_STATUS_SUCCESS) rc = -3;
    if (rc == 0 && hp_samp_int_fused(plan, w->fs, w->fc, &rays, &samp, &intl, w->ws.data(), w->ws.size()) != HP_STATUS_SUCCESS) rc = -4;
    if (rc == 0 && hp_img(plan, &intl, &rays, &img, w->ws_img.data(), w->ws_img.size()) != HP_STATUS_SUCCESS) rc = -5;
    if (forward_ms) *forward_ms = MsSince(t0);
    const size_t m = rc == 0 && samp.dt.rank >= 1 ? static_cast<size_t>(samp.dt.shape[0]) : 0;
    if (sample_count) *sample_count = m;
    if (rc == 0 && dL_dI != nullptr) {
        w->ws_diff.resize(m * 16 + 256);
        hp_tensor g = dvren::MakeHostTensor(static_cast<void*>(const_cast<float*>(dL_dI)), HP_DTYPE_F32, {static_cast<int64_t>(n), 3});
        hp_grads_t grads{};
        t0 = std::chrono::steady_clock::now();
        if (hp_diff(plan, &g, &samp, &intl, &grads, w->ws_diff.data(), w->ws_diff.size()) != HP_STATUS_SUCCESS) rc = -6;
        if (rc == 0 && !w->scatter.AccumulateSampleGradients(samp, std::span<const float>(static_cast<const float*>(grads.sigma.data), m),
                                                             std::span<const float>(static_cast<const float*>(grads.color.data), m * 3)).ok())
            rc = -7;
        if (backward_ms) *backward_ms = MsSince(t0);
    }
    hp_plan_release(plan);
    return rc;
}

}  // extern "C"
