// renderer.cpp -- dvren::Renderer with device-resident internals (reference src/render/renderer.cpp).
//
//   fused  (use_fused_path, default)  hpx_forward / hpx_backward on one hpx_frame: no per-sample
//          buffers at all; with enable_graph the forward is a replayed CUDA graph.
//   staged (use_fused_path = false)   hp_ray -> hp_samp -> hp_int -> hp_img and hp_diff + grid scatter
//          on DEVICE tensors with capacity-sized workspaces, i.e. the reference's staged call
//          sequence (renderer.cpp:259-365,415-427) executed by the materialising GPU kernels.
// Either way the only host traffic is what ForwardResult / BackwardResult hold.
#include "dvren/render/renderer.hpp"

#include <chrono>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace dvren {

namespace {

using Clock = std::chrono::steady_clock;

double MsSince(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

hp_tensor DeviceTensor(void* data) {
    hp_tensor t{};
    t.data = data;
    t.memspace = HP_MEMSPACE_DEVICE;
    t.dtype = HP_DTYPE_F32;
    return t;
}

Status Fail(hp_status st, const char* what) {
    return Status::FromHotpath(st, std::string(what) + " failed: " + hpx_last_error());
}

// event slots of hpx_ctx_mark (GPU-side stage times for RenderStats)
enum : uint32_t { kMarkBegin = 0, kMarkRays = 1, kMarkSample = 2, kMarkIntegrate = 3, kMarkCompose = 4,
                  kMarkBwdBegin = 5, kMarkBwdKernels = 6, kMarkBwdRead = 7 };

double GpuMs(const hp_ctx* ctx, uint32_t a, uint32_t b) {
    float ms = 0.0f;
    return hpx_ctx_elapsed_ms(ctx, a, b, &ms) == HP_STATUS_SUCCESS ? static_cast<double>(ms) : 0.0;
}

}  // namespace

struct Renderer::Impl {
    const hp_ctx* ctx{nullptr};
    hpx_frame* frame{nullptr};
    bool graph_captured{false};
    const hpx_grid* graph_grid{nullptr};
    // staged path: device buffers sized from the plan's capacities
    void* d_rays{nullptr};
    void* d_ws{nullptr};
    void* d_img{nullptr};
    void* d_grads{nullptr};
    void* d_dl{nullptr};
    size_t rays_bytes{0}, ws_bytes{0}, img_bytes{0}, grads_bytes{0}, dl_bytes{0};
    hp_rays_t rays{};
    hp_samp_t samp{};
    hp_intl_t intl{};

    // RenderOptions::pin_result_buffers: result vectors the caller hands in again and again (a training loop reuses
    // its ForwardResult / BackwardResult) are page-locked the second time they are seen: from then on the read-back is
    // direct DMA into the vector.  Never done unasked: a vector freed while registered (a ForwardResult local to a
    // frame loop, reference tests/render/test_smoke_animation.cpp:324) would leave a stale registration behind.
    std::vector<std::pair<void*, size_t>> seen, pinned;

    ~Impl() {
        for (auto& r : pinned) hpx_host_unregister(ctx, r.first);
        if (frame) hpx_frame_release(frame);
        for (void* p : {d_rays, d_ws, d_img, d_grads, d_dl}) hpx_device_free(ctx, p);
    }

    bool pin_enabled{false};   // RenderOptions::pin_result_buffers

    void PinIfRepeated(void* ptr, size_t bytes) {
        if (!pin_enabled || ptr == nullptr || bytes < (size_t(1) << 20)) return;   // small results are not worth a registration
        for (auto it = pinned.begin(); it != pinned.end(); ++it) {
            if (it->first == ptr && it->second >= bytes) return;
            if (it->first == ptr) {   // same address, grown: register afresh
                hpx_host_unregister(ctx, ptr);
                pinned.erase(it);
                break;
            }
        }
        for (auto& r : seen) {
            if (r.first == ptr && r.second == bytes) {
                if (hpx_host_register(ctx, ptr, bytes) == HP_STATUS_SUCCESS) pinned.emplace_back(ptr, bytes);
                return;
            }
        }
        if (seen.size() >= 16) seen.erase(seen.begin());
        seen.emplace_back(ptr, bytes);
    }

    Status Ensure(void*& ptr, size_t& have, size_t want) {
        if (ptr != nullptr && have >= want) return Status::Ok();
        hpx_device_free(ctx, ptr);
        ptr = nullptr;
        const hp_status st = hpx_device_alloc(ctx, want, &ptr);
        if (st != HP_STATUS_SUCCESS) return Fail(st, "device allocation");
        have = want;
        return Status::Ok();
    }
};

Renderer::Renderer(const Context& ctx, const Plan& plan, RenderOptions options)
    : ctx_(&ctx), plan_(&plan), options_(options), impl_(new Impl()) {
    impl_->ctx = ctx.handle();
    impl_->pin_enabled = options.pin_result_buffers;
}

Renderer::~Renderer() { delete impl_; }

Status Renderer::EnsureFrame() {
    if (impl_->frame != nullptr) return Status::Ok();
    const hp_status st = hpx_frame_create(plan_->handle(), &impl_->frame);
    if (st != HP_STATUS_SUCCESS) return Fail(st, "hpx_frame_create");
    return Status::Ok();
}

Status Renderer::ForwardFused(const DenseGridField& field, RenderStats& stats) {
    Status st = EnsureFrame();
    if (!st.ok()) return st;
    hp_status hs = hpx_ctx_mark(impl_->ctx, kMarkBegin);
    if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hpx_ctx_mark");
    if (options_.enable_graph) {
        if (!impl_->graph_captured || impl_->graph_grid != field.device_grid()) {
            hs = hpx_frame_capture(impl_->frame, field.device_grid(), 0);
            if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hpx_frame_capture");
            impl_->graph_captured = true;
            impl_->graph_grid = field.device_grid();
            stats.notes.emplace_back("graph_forward_captured");
        }
        hs = hpx_frame_replay(impl_->frame);
    } else {
        hs = hpx_forward(impl_->frame, field.device_grid());
    }
    if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hpx_forward");
    hpx_ctx_mark(impl_->ctx, kMarkSample);
    hpx_counts counts{};
    hs = hpx_frame_counts(impl_->frame, &counts);   // synchronises the stream
    if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hpx_frame_counts");
    // ray generation, marching, integration and image composition are ONE kernel: its GPU time (CUDA events) is
    // reported as sample_ms; ray_ms / integrate_ms stay 0 on the fused path
    stats.sample_ms = GpuMs(impl_->ctx, kMarkBegin, kMarkSample);
    stats.notes.emplace_back("timing=cuda_events");
    last_ray_count_ = static_cast<size_t>(counts.rays);
    last_sample_count_ = static_cast<size_t>(counts.samples);
    live_samples_ = static_cast<size_t>(counts.live_samples);
    return Status::Ok();
}

Status Renderer::ForwardStaged(const DenseGridField& field, RenderStats& stats) {
    const hp_plan_desc& d = plan_->descriptor();
    const size_t n = static_cast<size_t>(d.roi.width) * d.roi.height;
    const size_t cap = d.max_samples;
    const size_t pixels = static_cast<size_t>(d.width) * d.height;
    Status st = impl_->Ensure(impl_->d_rays, impl_->rays_bytes, n * 36);
    if (!st.ok()) return st;
    // samples (capacity sized, reference samp_cpu.cpp:111-136) then the integrator's outputs
    st = impl_->Ensure(impl_->d_ws, impl_->ws_bytes, cap * 32 + (n + 1) * 4 + n * 24 + cap * 16 + 64);
    if (!st.ok()) return st;
    st = impl_->Ensure(impl_->d_img, impl_->img_bytes, pixels * 28);
    if (!st.ok()) return st;

    char* rb = static_cast<char*>(impl_->d_rays);
    impl_->rays = hp_rays_t{};
    impl_->rays.origins = DeviceTensor(rb);
    impl_->rays.directions = DeviceTensor(rb + n * 12);
    impl_->rays.t_near = DeviceTensor(rb + n * 24);
    impl_->rays.t_far = DeviceTensor(rb + n * 28);
    impl_->rays.pixel_ids = DeviceTensor(rb + n * 32);
    hpx_ctx_mark(impl_->ctx, kMarkBegin);
    hp_status hs = hp_ray(plan_->handle(), nullptr, &impl_->rays, nullptr, 0);
    if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hp_ray");
    hpx_ctx_mark(impl_->ctx, kMarkRays);
    last_ray_count_ = impl_->rays.t_near.rank >= 1 ? static_cast<size_t>(impl_->rays.t_near.shape[0]) : 0;

    const size_t samp_bytes = cap * 32 + (n + 1) * 4;
    impl_->samp = hp_samp_t{};
    impl_->intl = hp_intl_t{};
    hs = hp_samp(plan_->handle(), field.sigma_field(), field.color_field(), &impl_->rays, &impl_->samp, impl_->d_ws,
                 samp_bytes);
    if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hp_samp");
    hpx_ctx_mark(impl_->ctx, kMarkSample);
    hs = hp_int(plan_->handle(), &impl_->samp, &impl_->intl, static_cast<char*>(impl_->d_ws) + samp_bytes,
                impl_->ws_bytes - samp_bytes);
    if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hp_int");
    hpx_ctx_mark(impl_->ctx, kMarkIntegrate);
    stats.ray_ms = GpuMs(impl_->ctx, kMarkBegin, kMarkRays);
    stats.sample_ms = GpuMs(impl_->ctx, kMarkRays, kMarkSample);
    stats.integrate_ms = GpuMs(impl_->ctx, kMarkSample, kMarkIntegrate);
    stats.notes.emplace_back("timing=cuda_events");
    last_sample_count_ = impl_->samp.dt.rank >= 1 ? static_cast<size_t>(impl_->samp.dt.shape[0]) : 0;
    live_samples_ = 0;   // not tracked by the staged entry points
    return Status::Ok();
}

Status Renderer::Forward(const DenseGridField& field, ForwardResult& out) {
    if (!field.valid() || field.device_grid() == nullptr) return Status(StatusCode::kInvalidArgument, "field is invalid");
    if (plan_ == nullptr || ctx_ == nullptr || !plan_->valid())
        return Status(StatusCode::kInvalidArgument, "renderer is not bound to a context/plan");
    RenderStats stats{};
    const auto total0 = Clock::now();
    const bool fused = options_.use_fused_path;
    if (options_.enable_graph) stats.notes.emplace_back("graph_capture_enabled");
    stats.notes.emplace_back(fused ? "forward_mode=fused" : "forward_mode=staged");

    const hp_plan_desc& d = plan_->descriptor();
    const size_t pixels = static_cast<size_t>(d.width) * d.height;
    Status st = fused ? ForwardFused(field, stats) : ForwardStaged(field, stats);
    if (!st.ok()) return st;
    last_forward_staged_ = !fused;
    if (!fused && options_.enable_graph) {
        // keep the lean frame around too: the graph path owns its workspace (accounting, test_core.cpp:163)
        st = EnsureFrame();
        if (!st.ok()) return st;
    }
    if (last_ray_count_ == 0) {
        out = ForwardResult{};
        stats.total_ms = MsSince(total0);
        out.stats = stats;
        return Status::Ok();
    }

    out.image.resize(pixels * 3);
    out.transmittance.resize(pixels);
    out.opacity.resize(pixels);
    out.depth.resize(pixels);
    out.hitmask.resize(pixels);
    impl_->PinIfRepeated(out.image.data(), pixels * 12);
    impl_->PinIfRepeated(out.transmittance.data(), pixels * 4);
    impl_->PinIfRepeated(out.opacity.data(), pixels * 4);
    impl_->PinIfRepeated(out.depth.data(), pixels * 4);
    impl_->PinIfRepeated(out.hitmask.data(), pixels * 4);
    hpx_ctx_mark(impl_->ctx, kMarkIntegrate);   // start of the compose stage (re-recorded: the forward stages were read above)
    if (fused) {
        const hp_status hs = hpx_frame_read(impl_->frame, out.image.data(), out.transmittance.data(),
                                            out.opacity.data(), out.depth.data(), out.hitmask.data());
        if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hpx_frame_read");
    } else {
        char* ib = static_cast<char*>(impl_->d_img);
        hp_img_t img{};
        img.image = DeviceTensor(ib);
        img.trans = DeviceTensor(ib + pixels * 12);
        img.opacity = DeviceTensor(ib + pixels * 16);
        img.depth = DeviceTensor(ib + pixels * 20);
        img.hitmask = DeviceTensor(ib + pixels * 24);
        hp_status hs = hp_img(plan_->handle(), &impl_->intl, &impl_->rays, &img, nullptr, 0);
        if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hp_img");
        const hp_ctx* c = ctx_->handle();
        hs = hpx_copy_to_host(c, out.image.data(), img.image.data, pixels * 12);
        if (hs == HP_STATUS_SUCCESS) hs = hpx_copy_to_host(c, out.transmittance.data(), img.trans.data, pixels * 4);
        if (hs == HP_STATUS_SUCCESS) hs = hpx_copy_to_host(c, out.opacity.data(), img.opacity.data, pixels * 4);
        if (hs == HP_STATUS_SUCCESS) hs = hpx_copy_to_host(c, out.depth.data(), img.depth.data, pixels * 4);
        if (hs == HP_STATUS_SUCCESS) hs = hpx_copy_to_host(c, out.hitmask.data(), img.hitmask.data, pixels * 4);
        if (hs != HP_STATUS_SUCCESS) return Fail(hs, "image read-back");
    }
    hpx_ctx_mark(impl_->ctx, kMarkCompose);
    stats.compose_ms = GpuMs(impl_->ctx, kMarkIntegrate, kMarkCompose);   // staged: hp_img + read-back; fused: the read-back
    out.ray_count = last_ray_count_;
    out.sample_count = last_sample_count_;
    stats.total_ms = MsSince(total0);
    out.stats = std::move(stats);
    return Status::Ok();
}

Status Renderer::BackwardStaged(DenseGridField& field, std::span<const float> dL_dI) {
    const size_t n = last_ray_count_, m = last_sample_count_;
    Status st = impl_->Ensure(impl_->d_dl, impl_->dl_bytes, n * 12);
    if (!st.ok()) return st;
    st = impl_->Ensure(impl_->d_grads, impl_->grads_bytes, m * 16 + 64);
    if (!st.ok()) return st;
    const hp_ctx* c = ctx_->handle();
    hp_status hs = hpx_copy_to_device(c, impl_->d_dl, dL_dI.data(), n * 12);
    if (hs != HP_STATUS_SUCCESS) return Fail(hs, "dL/dI upload");
    hp_tensor g = DeviceTensor(impl_->d_dl);
    g.rank = 2;
    g.shape[0] = static_cast<int64_t>(n);
    g.shape[1] = 3;
    g.stride[0] = 3;
    g.stride[1] = 1;
    char* gb = static_cast<char*>(impl_->d_grads);
    hp_grads_t grads{};
    grads.sigma = DeviceTensor(gb);
    grads.color = DeviceTensor(gb + m * 4);
    grads.camera = DeviceTensor(gb + m * 16);
    hs = hp_diff(plan_->handle(), &g, &impl_->samp, &impl_->intl, &grads, nullptr, 0);
    if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hp_diff");
    field.ZeroGradients();
    hs = hpx_grid_accumulate_samples(field.device_grid(), static_cast<const float*>(impl_->samp.positions.data),
                                     static_cast<const float*>(grads.sigma.data),
                                     static_cast<const float*>(grads.color.data), m, HP_MEMSPACE_DEVICE);
    if (hs != HP_STATUS_SUCCESS) return Fail(hs, "gradient scatter");
    return Status::Ok();
}

Status Renderer::Backward(DenseGridField& field, std::span<const float> dL_dI, BackwardResult& out) {
    if (!field.valid() || field.device_grid() == nullptr) return Status(StatusCode::kInvalidArgument, "field is invalid");
    if (last_ray_count_ == 0 || last_sample_count_ == 0)
        return Status(StatusCode::kInvalidArgument, "forward pass not executed or produced zero samples");
    if (dL_dI.size() != last_ray_count_ * 3) return Status(StatusCode::kInvalidArgument, "dL/dI size mismatch");

    const auto total0 = Clock::now();
    backward_stats_ = RenderStats{};
    // full-grid copies are what the reference's API returns (renderer.cpp:441-442): un-interleaved on the device and
    // read STRAIGHT into the caller's vectors (no intermediate host copy; page-locked once the vectors repeat)
    const size_t voxels = field.voxel_count();
    out.sigma.resize(voxels);
    out.color.resize(voxels * 3);
    impl_->PinIfRepeated(out.sigma.data(), voxels * 4);
    impl_->PinIfRepeated(out.color.data(), voxels * 12);
    std::array<float, 16> cam16{};
    bool read_back_done = false;
    hpx_ctx_mark(impl_->ctx, kMarkBwdBegin);
    if (last_forward_staged_) {
        const Status st = BackwardStaged(field, dL_dI);
        if (!st.ok()) return st;
    } else {
        uint32_t flags = HPX_BACKWARD_GRID | HPX_BACKWARD_ZERO;
        if (options_.camera_gradients) flags |= HPX_BACKWARD_CAMERA;
        if (impl_->pin_enabled) {
            // page-locked result vectors: the read-back runs UNDER the backward kernel (hpx_backward_streamed)
            hp_status hs = hpx_backward_streamed(impl_->frame, field.device_grid(), dL_dI.data(), HP_MEMSPACE_HOST, flags, out.sigma.data(),
                                                 out.color.data(), cam16.data());
            if (hs == HP_STATUS_SUCCESS) hs = hpx_ctx_synchronize(impl_->ctx);
            if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hpx_backward_streamed");
            read_back_done = true;
            backward_stats_.notes.emplace_back("gradient_read_back=streamed_under_the_backward_kernel");
        } else {
            const hp_status hs = hpx_backward(impl_->frame, field.device_grid(), dL_dI.data(), HP_MEMSPACE_HOST, flags);
            if (hs != HP_STATUS_SUCCESS) return Fail(hs, "hpx_backward");
        }
    }
    hpx_ctx_mark(impl_->ctx, kMarkBwdKernels);
    field.MarkGradientsStale();   // the field's own host mirrors are refreshed only if somebody asks for them
    if (!read_back_done) {
        const hp_status rs = hpx_grid_read_grad(field.device_grid(), out.sigma.data(), out.color.data(), cam16.data(), HP_MEMSPACE_HOST);
        if (rs != HP_STATUS_SUCCESS) return Fail(rs, "hpx_grid_read_grad");
    }
    hpx_ctx_mark(impl_->ctx, kMarkBwdRead);
    out.camera.fill(0.0f);
    intrinsics_grad_.fill(0.0f);
    if (options_.camera_gradients && !last_forward_staged_) {
        for (int i = 0; i < 12; ++i) out.camera[static_cast<size_t>(i)] = cam16[static_cast<size_t>(i)];
        for (int i = 0; i < 4; ++i) intrinsics_grad_[static_cast<size_t>(i)] = cam16[static_cast<size_t>(12 + i)];
    }
    out.sample_count = last_sample_count_;
    backward_stats_.sample_ms = GpuMs(impl_->ctx, kMarkBwdBegin, kMarkBwdKernels);
    backward_stats_.compose_ms = GpuMs(impl_->ctx, kMarkBwdKernels, kMarkBwdRead);
    backward_stats_.total_ms = MsSince(total0);
    backward_stats_.notes.emplace_back(last_forward_staged_ ? "backward_mode=staged" : "backward_mode=fused");
    backward_stats_.notes.emplace_back("timing=cuda_events");
    return Status::Ok();
}

WorkspaceInfo Renderer::workspace_info() const {
    WorkspaceInfo info{};
    const hp_plan_desc& d = plan_->descriptor();
    const size_t pixels = static_cast<size_t>(d.width) * d.height;
    info.ray_buffer_bytes = impl_->rays_bytes;
    info.sample_buffer_bytes = impl_->ws_bytes;
    info.integration_buffer_bytes = 0;   // carved out of the sample workspace, like the reference's fused call
    info.image_buffer_bytes = impl_->img_bytes + (impl_->frame ? pixels * 28 : 0);
    info.gradient_buffer_bytes = impl_->grads_bytes + impl_->dl_bytes;
    const size_t frame_bytes = impl_->frame ? hpx_frame_bytes(impl_->frame) : 0;
    info.workspace_buffer_bytes = frame_bytes > pixels * 28 ? frame_bytes - pixels * 28 : 0;
    return info;
}

}  // namespace dvren
