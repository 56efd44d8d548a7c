// dense_grid.cpp -- DenseGridField on the GPU (reference src/fields/dense_grid.cpp).
// Create uploads the grid (two hp_field handles + one packed hpx_grid); gradients accumulate in the
// packed device gradient grid and are mirrored to host vectors only when somebody asks for them.
#include "dvren/fields/dense_grid.hpp"

#include <utility>

#include "dvren/core/tensor_utils.hpp"

namespace dvren {

DenseGridField::~DenseGridField() { Release(); }

DenseGridField::DenseGridField(DenseGridField&& o) noexcept { *this = std::move(o); }

DenseGridField& DenseGridField::operator=(DenseGridField&& o) noexcept {
    if (this == &o) return *this;
    Release();
    sigma_field_ = std::exchange(o.sigma_field_, nullptr);
    color_field_ = std::exchange(o.color_field_, nullptr);
    grid_ = std::exchange(o.grid_, nullptr);
    resolution_ = o.resolution_;
    bbox_min_ = o.bbox_min_;
    bbox_max_ = o.bbox_max_;
    interp_ = o.interp_;
    oob_ = o.oob_;
    sigma_grad_ = std::move(o.sigma_grad_);
    color_grad_ = std::move(o.color_grad_);
    camera_grad_ = o.camera_grad_;
    mirrors_stale_ = o.mirrors_stale_;
    return *this;
}

void DenseGridField::Release() {
    if (sigma_field_ != nullptr) hp_field_release(sigma_field_);   // the views first, then the grid they look into
    if (color_field_ != nullptr) hp_field_release(color_field_);
    if (grid_ != nullptr) hpx_grid_release(grid_);
    grid_ = nullptr;
    sigma_field_ = color_field_ = nullptr;
    sigma_grad_.clear();
    color_grad_.clear();
}

Status DenseGridField::Create(const Context& ctx, const DenseGridConfig& config, DenseGridField& out) {
    if (!ctx.valid()) return Status(StatusCode::kInvalidArgument, "context is invalid");
    const auto& r = config.resolution;
    if (r[0] <= 0 || r[1] <= 0 || r[2] <= 0) return Status(StatusCode::kInvalidArgument, "resolution must be positive");
    const int64_t nx = r[0], ny = r[1], nz = r[2];
    const int64_t voxels = nx * ny * nz;
    if (static_cast<int64_t>(config.sigma.size()) != voxels)
        return Status(StatusCode::kInvalidArgument, "sigma data size mismatch");
    if (static_cast<int64_t>(config.color.size()) != voxels * 3)
        return Status(StatusCode::kInvalidArgument, "color data size mismatch");

    DenseGridField f;
    // the library snapshots the values into HBM, so no host copy has to outlive this call
    hp_tensor st = MakeHostTensor(const_cast<float*>(config.sigma.data()), HP_DTYPE_F32, {nz, ny, nx});
    hp_tensor ct = MakeHostTensor(const_cast<float*>(config.color.data()), HP_DTYPE_F32, {nz, ny, nx, 3});
    hp_status hs = hp_field_create_grid_sigma(ctx.handle(), &st, static_cast<uint32_t>(config.interp),
                                              static_cast<uint32_t>(config.oob), &f.sigma_field_);
    if (hs != HP_STATUS_SUCCESS || f.sigma_field_ == nullptr)
        return Status::FromHotpath(hs, std::string("hp_field_create_grid_sigma failed: ") + hpx_last_error());
    hs = hp_field_create_grid_color(ctx.handle(), &ct, static_cast<uint32_t>(config.interp),
                                    static_cast<uint32_t>(config.oob), &f.color_field_);
    if (hs != HP_STATUS_SUCCESS || f.color_field_ == nullptr)
        return Status::FromHotpath(hs, std::string("hp_field_create_grid_color failed: ") + hpx_last_error());
    hs = hpx_grid_create(ctx.handle(), f.sigma_field_, f.color_field_, config.bbox_min.data(), config.bbox_max.data(),
                         &f.grid_);
    if (hs != HP_STATUS_SUCCESS) return Status::FromHotpath(hs, std::string("hpx_grid_create failed: ") + hpx_last_error());
    // one copy of the values in HBM: the two fields become views of the packed grid, so UpdateValues reaches the staged
    // hp_samp / hp_graph paths (sigma_field() / color_field()) as well as the fused one
    hs = hpx_grid_adopt_fields(f.grid_, f.sigma_field_, f.color_field_);
    if (hs != HP_STATUS_SUCCESS) return Status::FromHotpath(hs, std::string("hpx_grid_adopt_fields failed: ") + hpx_last_error());
    // empty-space skipping: a field is created once per scene (or per animation frame), so the occupancy bits are built
    // here; they only take effect when at least one brick in twenty is empty (dense volumes keep the plain kernels).
    // Skipped samples contribute exactly nothing, results do not change (hp_b200.h: hpx_grid_build_occupancy).
    {
        float empty_sigma = 0.0f, empty_all = 0.0f;
        if (hpx_grid_build_occupancy(f.grid_, 1, &empty_sigma, &empty_all) == HP_STATUS_SUCCESS && empty_sigma < 0.05f)
            hpx_grid_set_occupancy(f.grid_, 0);
    }
    f.resolution_ = config.resolution;
    f.bbox_min_ = config.bbox_min;
    f.bbox_max_ = config.bbox_max;
    f.interp_ = config.interp == HP_INTERP_NEAREST ? HP_INTERP_NEAREST : HP_INTERP_LINEAR;
    f.oob_ = config.oob == HP_OOB_CLAMP ? HP_OOB_CLAMP : HP_OOB_ZERO;
    // host mirrors of the gradients are created on first use (sigma_gradients() / color_gradients()): a 512^3 field
    // would otherwise pin down 2.1 GB of host memory nobody may ever read
    f.mirrors_stale_ = true;
    out = std::move(f);
    return Status::Ok();
}

void DenseGridField::ZeroGradients() {
    if (grid_ != nullptr) hpx_grid_zero_grad(grid_);
    camera_grad_.fill(0.0f);
    if (sigma_grad_.empty() && color_grad_.empty()) {   // mirrors not materialised yet: the device grid is the truth
        mirrors_stale_ = true;
        return;
    }
    std::fill(sigma_grad_.begin(), sigma_grad_.end(), 0.0f);
    std::fill(color_grad_.begin(), color_grad_.end(), 0.0f);
    mirrors_stale_ = false;
}

Status DenseGridField::AccumulateSampleGradients(const hp_samp_t& samples, std::span<const float> grad_sigma,
                                                 std::span<const float> grad_color) {
    if (samples.positions.memspace != HP_MEMSPACE_HOST || samples.positions.dtype != HP_DTYPE_F32 ||
        samples.dt.memspace != HP_MEMSPACE_HOST || samples.dt.dtype != HP_DTYPE_F32)
        return Status(StatusCode::kInvalidArgument, "samples must reside on host");
    if (samples.positions.rank != 2 || samples.positions.shape[1] != 3)
        return Status(StatusCode::kInvalidArgument, "sample positions must have shape (M,3)");
    const size_t count = static_cast<size_t>(samples.positions.shape[0]);
    if (grad_sigma.size() != count) return Status(StatusCode::kInvalidArgument, "grad_sigma size mismatch");
    if (grad_color.size() != count * 3) return Status(StatusCode::kInvalidArgument, "grad_color size mismatch");
    if (voxel_count() == 0 || grid_ == nullptr) return Status(StatusCode::kInvalidArgument, "voxel count is zero");
    const hp_status hs = hpx_grid_accumulate_samples(grid_, static_cast<const float*>(samples.positions.data),
                                                     grad_sigma.data(), grad_color.data(), count, HP_MEMSPACE_HOST);
    if (hs != HP_STATUS_SUCCESS) return Status::FromHotpath(hs, std::string("gradient scatter failed: ") + hpx_last_error());
    mirrors_stale_ = true;
    return Status::Ok();
}

void DenseGridField::RefreshMirrors() const {
    if (!mirrors_stale_ || grid_ == nullptr) return;
    sigma_grad_.resize(voxel_count());
    color_grad_.resize(voxel_count() * 3);
    hpx_grid_read_grad(grid_, sigma_grad_.data(), color_grad_.data(), camera_grad_.data(), HP_MEMSPACE_HOST);
    mirrors_stale_ = false;
}

const std::vector<float>& DenseGridField::sigma_gradients() const {
    RefreshMirrors();
    return sigma_grad_;
}

const std::vector<float>& DenseGridField::color_gradients() const {
    RefreshMirrors();
    return color_grad_;
}

Status DenseGridField::UpdateValues(std::span<const float> sigma, std::span<const float> color) {
    if (grid_ == nullptr) return Status(StatusCode::kInvalidArgument, "field is invalid");
    if ((!sigma.empty() && sigma.size() != voxel_count()) || (!color.empty() && color.size() != voxel_count() * 3))
        return Status(StatusCode::kInvalidArgument, "value size mismatch");
    const hp_status hs = hpx_grid_update(grid_, sigma.empty() ? nullptr : sigma.data(),
                                         color.empty() ? nullptr : color.data(), HP_MEMSPACE_HOST);
    if (hs != HP_STATUS_SUCCESS) return Status::FromHotpath(hs, std::string("hpx_grid_update failed: ") + hpx_last_error());
    return Status::Ok();
}

}  // namespace dvren
