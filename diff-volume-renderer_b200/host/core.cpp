// core.cpp -- Status, Context, Plan, MakeHostTensor of the dvren C++ surface
// (reference src/core/{status,context,plan,tensor_utils}.cpp), written against the C ABI only.
#include <cstring>
#include <utility>

#include "dvren/core/context.hpp"
#include "dvren/core/plan.hpp"
#include "dvren/core/status.hpp"
#include "dvren/core/tensor_utils.hpp"

namespace dvren {

// ---- Status -------------------------------------------------------------------------------
Status Status::FromHotpath(hp_status code, std::string message) {
    static constexpr StatusCode kMap[] = {StatusCode::kOk, StatusCode::kInvalidArgument, StatusCode::kOutOfMemory,
                                          StatusCode::kNotImplemented, StatusCode::kUnsupported,
                                          StatusCode::kInternalError};
    const auto raw = static_cast<unsigned>(code);
    return Status(raw < 6 ? kMap[raw] : StatusCode::kInternalError, std::move(message));
}

std::string Status::ToString() const {
    static constexpr const char* kNames[] = {"ok", "invalid_argument", "out_of_memory", "not_implemented",
                                             "unsupported", "internal_error"};
    if (ok()) return "ok";
    const std::string name = kNames[static_cast<int>(code_)];
    return message_.empty() ? name : name + ": " + message_;
}

// ---- Context ------------------------------------------------------------------------------
Context::~Context() { Reset(nullptr, {}); }

Context::Context(Context&& other) noexcept : ctx_(std::exchange(other.ctx_, nullptr)), desc_(other.desc_) {}

Context& Context::operator=(Context&& other) noexcept {
    if (this != &other) {
        Reset(std::exchange(other.ctx_, nullptr), other.desc_);
    }
    return *this;
}

void Context::Reset(hp_ctx* ctx, const hp_ctx_desc& desc) {
    if (ctx_ != nullptr) hp_ctx_release(ctx_);
    ctx_ = ctx;
    desc_ = desc;
}

Status Context::Create(const ContextOptions& options, Context& out) {
    hp_ctx_desc desc{};
    desc.flags = options.flags;
    // the library copies the string, so a temporary ContextOptions is fine
    desc.preferred_device = options.preferred_device.empty() ? nullptr : options.preferred_device.c_str();
    hp_ctx* raw = nullptr;
    const hp_status st = hp_ctx_create(&desc, &raw);
    if (st != HP_STATUS_SUCCESS || raw == nullptr) return Status::FromHotpath(st, "hp_ctx_create failed");
    hp_ctx_desc actual{};
    const hp_status got = hp_ctx_get_desc(raw, &actual);
    if (got != HP_STATUS_SUCCESS) {
        hp_ctx_release(raw);
        return Status::FromHotpath(got, "hp_ctx_get_desc failed");
    }
    out.Reset(raw, actual);
    return Status::Ok();
}

// ---- Plan ---------------------------------------------------------------------------------
Plan::~Plan() { Reset(nullptr, {}); }

Plan::Plan(Plan&& other) noexcept : plan_(std::exchange(other.plan_, nullptr)), desc_(other.desc_) {}

Plan& Plan::operator=(Plan&& other) noexcept {
    if (this != &other) Reset(std::exchange(other.plan_, nullptr), other.desc_);
    return *this;
}

void Plan::Reset(hp_plan* plan, const hp_plan_desc& desc) {
    if (plan_ != nullptr) hp_plan_release(plan_);
    plan_ = plan;
    desc_ = desc;
}

Status Plan::Create(const Context& ctx, const PlanDescriptor& d, Plan& out) {
    if (!ctx.valid()) return Status(StatusCode::kInvalidArgument, "context is invalid");
    hp_plan_desc p{};
    p.width = d.width;
    p.height = d.height;
    p.t_near = d.t_near;
    p.t_far = d.t_far;
    p.max_rays = d.max_rays;
    p.max_samples = d.max_samples;
    p.seed = d.seed;
    p.sampling.dt = d.sampling.dt;
    p.sampling.max_steps = d.sampling.max_steps;
    p.sampling.mode = d.sampling.mode == SamplingMode::kStratified ? HP_SAMPLING_STRATIFIED : HP_SAMPLING_FIXED;
    p.camera.model = d.camera.model == CameraModel::kOrthographic ? HP_CAMERA_ORTHOGRAPHIC : HP_CAMERA_PINHOLE;
    p.camera.ortho_scale = d.camera.ortho_scale;
    std::memcpy(p.camera.K, d.camera.K.data(), sizeof(p.camera.K));
    std::memcpy(p.camera.c2w, d.camera.c2w.data(), sizeof(p.camera.c2w));
    if (d.roi.has_value()) p.roi = hp_roi_desc{d.roi->x, d.roi->y, d.roi->width, d.roi->height};

    hp_plan* raw = nullptr;
    const hp_status st = hp_plan_create(ctx.handle(), &p, &raw);
    if (st != HP_STATUS_SUCCESS || raw == nullptr) return Status::FromHotpath(st, "hp_plan_create failed");
    hp_plan_desc resolved{};
    const hp_status got = hp_plan_get_desc(raw, &resolved);
    if (got != HP_STATUS_SUCCESS) {
        hp_plan_release(raw);
        return Status::FromHotpath(got, "hp_plan_get_desc failed");
    }
    out.Reset(raw, resolved);
    return Status::Ok();
}

// ---- tensors ------------------------------------------------------------------------------
hp_tensor MakeHostTensor(void* data, hp_dtype dtype, const std::vector<int64_t>& shape) {
    hp_tensor t{};
    t.data = data;
    t.dtype = dtype;
    t.memspace = HP_MEMSPACE_HOST;
    t.rank = static_cast<uint32_t>(shape.size() < 8 ? shape.size() : 8);
    int64_t stride = 1;
    for (int i = static_cast<int>(t.rank) - 1; i >= 0; --i) {
        t.shape[i] = shape[static_cast<size_t>(i)];
        t.stride[i] = stride;
        stride *= t.shape[i];
    }
    return t;
}

hp_tensor MakeHostTensor(void* data, hp_dtype dtype, std::initializer_list<int64_t> shape) {
    return MakeHostTensor(data, dtype, std::vector<int64_t>(shape));
}

}  // namespace dvren
