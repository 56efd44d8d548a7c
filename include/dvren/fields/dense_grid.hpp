// dvren/fields/dense_grid.hpp -- dense sigma + RGB grid with its gradient accumulators.
// Public interface of the reference (include/dvren/fields/dense_grid.hpp:13-75).  Here the values
// live in HBM: two hp_field handles (raw arrays, for the hp.h entry points) plus one packed
// {r,g,b,sigma} hpx_grid (for the fused path) that also owns the packed gradient grid.  The
// gradient vectors returned by sigma_gradients() / color_gradients() are host mirrors, refreshed
// from the device on demand.
#pragma once

#include <array>
#include <cstdint>
#include <span>
#include <vector>

#include "dvren/core/context.hpp"
#include "dvren/core/status.hpp"
#include "hotpath/hp_b200.h"

namespace dvren {

struct DenseGridConfig {
    std::array<int32_t, 3> resolution{1, 1, 1};          // {nx, ny, nz}
    std::vector<float> sigma;                             // [nz][ny][nx]
    std::vector<float> color;                             // [nz][ny][nx][3]
    std::array<float, 3> bbox_min{0.0f, 0.0f, 0.0f};      // used by the gradient scatter only,
    std::array<float, 3> bbox_max{1.0f, 1.0f, 1.0f};      // like the reference (SURVEY finding 8)
    hp_interp_mode interp{HP_INTERP_LINEAR};
    hp_oob_policy oob{HP_OOB_ZERO};
};

class DenseGridField {
public:
    DenseGridField() = default;
    ~DenseGridField();
    DenseGridField(DenseGridField&& other) noexcept;
    DenseGridField& operator=(DenseGridField&& other) noexcept;
    DenseGridField(const DenseGridField&) = delete;
    DenseGridField& operator=(const DenseGridField&) = delete;

    static Status Create(const Context& ctx, const DenseGridConfig& config, DenseGridField& out);

    [[nodiscard]] bool valid() const { return sigma_field_ != nullptr && color_field_ != nullptr; }
    [[nodiscard]] const hp_field* sigma_field() const { return sigma_field_; }
    [[nodiscard]] hp_field* sigma_field() { return sigma_field_; }
    [[nodiscard]] const hp_field* color_field() const { return color_field_; }
    [[nodiscard]] hp_field* color_field() { return color_field_; }
    [[nodiscard]] std::array<int32_t, 3> resolution() const { return resolution_; }
    [[nodiscard]] std::array<float, 3> bbox_min() const { return bbox_min_; }
    [[nodiscard]] std::array<float, 3> bbox_max() const { return bbox_max_; }
    [[nodiscard]] hp_interp_mode interpolation() const { return interp_; }
    [[nodiscard]] hp_oob_policy oob_policy() const { return oob_; }

    void ZeroGradients();
    // Scatter per-sample gradients (HOST views, as hp_diff returns them) into the grid gradients.
    Status AccumulateSampleGradients(const hp_samp_t& samples, std::span<const float> grad_sigma,
                                     std::span<const float> grad_color);

    [[nodiscard]] const std::vector<float>& sigma_gradients() const;
    [[nodiscard]] const std::vector<float>& color_gradients() const;
    [[nodiscard]] size_t voxel_count() const {
        return static_cast<size_t>(resolution_[0]) * static_cast<size_t>(resolution_[1]) *
               static_cast<size_t>(resolution_[2]);
    }

    // ---- additive (device residency) -------------------------------------------------------
    [[nodiscard]] hpx_grid* device_grid() const { return grid_; }
    // Replace the grid values in place (parameter update) without re-creating the field.
    Status UpdateValues(std::span<const float> sigma, std::span<const float> color);
    // Called by Renderer after it wrote gradients on the device.
    void MarkGradientsStale() const { mirrors_stale_ = true; }

private:
    void Release();
    void RefreshMirrors() const;

    hp_field* sigma_field_{nullptr};
    hp_field* color_field_{nullptr};
    hpx_grid* grid_{nullptr};
    std::array<int32_t, 3> resolution_{1, 1, 1};
    std::array<float, 3> bbox_min_{0.0f, 0.0f, 0.0f};
    std::array<float, 3> bbox_max_{1.0f, 1.0f, 1.0f};
    hp_interp_mode interp_{HP_INTERP_LINEAR};
    hp_oob_policy oob_{HP_OOB_ZERO};
    mutable std::vector<float> sigma_grad_;
    mutable std::vector<float> color_grad_;
    mutable std::array<float, 16> camera_grad_{};
    mutable bool mirrors_stale_{false};

    friend class Renderer;
};

}  // namespace dvren
