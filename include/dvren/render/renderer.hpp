// dvren/render/renderer.hpp -- forward / backward renderer over a Plan and a DenseGridField.
// Public interface of the reference (include/dvren/render/renderer.hpp:17-149).  Internals are
// device resident: the fused path is one hpx_forward / hpx_backward pair on a per-plan hpx_frame,
// the staged path drives hp_ray / hp_samp / hp_int / hp_img / hp_diff on DEVICE tensors, and
// enable_graph replays a captured CUDA graph.  Results cross to the host only in the vectors of
// ForwardResult / BackwardResult, as the reference's API demands.
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <span>
#include <string>
#include <vector>

#include "dvren/core/context.hpp"
#include "dvren/core/plan.hpp"
#include "dvren/core/status.hpp"
#include "dvren/fields/dense_grid.hpp"

namespace dvren {

struct RenderOptions {
    bool use_fused_path{true};
    bool enable_graph{false};
    bool capture_stats{true};
    // additive: fill BackwardResult::camera with d/d c2w (the reference leaves it zero,
    // src/render/renderer.cpp:408,443) and expose d/d {fx,fy,cx,cy} via intrinsics_gradient()
    bool camera_gradients{false};
    // additive: page-lock the caller's result vectors (ForwardResult / BackwardResult members of >= 1 MiB) the second
    // time the same buffer is handed in, so that the read-back is direct DMA into them.  OPT-IN, because the renderer
    // cannot see a vector being freed: only for callers that reuse ONE set of result objects and keep them alive (and
    // un-resized) for as long as the Renderer lives -- a training loop.  Off: results are copied into pageable memory.
    bool pin_result_buffers{false};
};

struct WorkspaceInfo {
    size_t ray_buffer_bytes{0};
    size_t sample_buffer_bytes{0};
    size_t integration_buffer_bytes{0};
    size_t image_buffer_bytes{0};
    size_t gradient_buffer_bytes{0};
    size_t workspace_buffer_bytes{0};

    [[nodiscard]] size_t total_bytes() const {
        return ray_buffer_bytes + sample_buffer_bytes + integration_buffer_bytes + image_buffer_bytes +
               gradient_buffer_bytes + workspace_buffer_bytes;
    }
};

struct RenderStats {
    double total_ms{0.0};
    double ray_ms{0.0};
    double sample_ms{0.0};
    double integrate_ms{0.0};
    double compose_ms{0.0};
    std::vector<std::string> notes;
};

struct ForwardResult {
    std::vector<float> image;          // [H][W][3]
    std::vector<float> transmittance;  // [H][W]
    std::vector<float> opacity;        // [H][W]
    std::vector<float> depth;          // [H][W]
    std::vector<uint32_t> hitmask;     // [H][W]
    size_t ray_count{0};
    size_t sample_count{0};
    RenderStats stats;
};

struct BackwardResult {
    std::vector<float> sigma;          // [V]
    std::vector<float> color;          // [V][3]
    std::array<float, 12> camera{};    // d/d c2w (zero unless RenderOptions::camera_gradients)
    size_t sample_count{0};
};

class Renderer {
public:
    Renderer(const Context& ctx, const Plan& plan, RenderOptions options = {});
    ~Renderer();
    Renderer(const Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;

    Status Forward(const DenseGridField& field, ForwardResult& out);
    // dL_dI: (ray_count, 3) per ray in plan order (not H x W x 3), as in the reference.
    Status Backward(DenseGridField& field, std::span<const float> dL_dI, BackwardResult& out);

    [[nodiscard]] WorkspaceInfo workspace_info() const;
    [[nodiscard]] const RenderOptions& options() const { return options_; }

    // ---- additive -------------------------------------------------------------------------
    [[nodiscard]] std::array<float, 4> intrinsics_gradient() const { return intrinsics_grad_; }
    [[nodiscard]] size_t live_sample_count() const { return live_samples_; }
    // Stage times of the last Backward (the reference's BackwardResult carries no stats): sample_ms = adjoint + grid
    // scatter kernels (GPU time, CUDA events), compose_ms = gradient read-back, total_ms = host wall clock.
    [[nodiscard]] const RenderStats& backward_stats() const { return backward_stats_; }

private:
    struct Impl;
    Status EnsureFrame();
    Status ForwardFused(const DenseGridField& field, RenderStats& stats);
    Status ForwardStaged(const DenseGridField& field, RenderStats& stats);
    Status BackwardStaged(DenseGridField& field, std::span<const float> dL_dI);

    const Context* ctx_{nullptr};
    const Plan* plan_{nullptr};
    RenderOptions options_{};
    Impl* impl_{nullptr};
    size_t last_ray_count_{0};
    size_t last_sample_count_{0};
    size_t live_samples_{0};
    bool last_forward_staged_{false};
    std::array<float, 4> intrinsics_grad_{};
    RenderStats backward_stats_{};
};

}  // namespace dvren
