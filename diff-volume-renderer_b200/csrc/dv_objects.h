// dv_objects.h -- the opaque handles behind hp.h / hp_b200.h and small host helpers.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "dv_lean.h"
#include "dv_staged.h"
#include "dv_types.h"
#include "hotpath/hp.h"
#include "hotpath/hp_b200.h"

// ---- handles ---------------------------------------------------------------
// Context: which GPU, which stream.  Device state is created on first use so
// that hp_ctx_create / hp_plan_create (pure host logic, reference
// hp_runtime.cpp:15-146) work on machines without a GPU; every compute entry
// point fails with HP_STATUS_UNSUPPORTED there.
struct hp_ctx {
    // Plans, fields, grids and frames keep their context alive: hp_ctx_release only drops the caller's reference, the
    // stream and the device buffers go away with the last object (the reference's fields never touch their context, so
    // releasing it first is legal there: hp_runtime.cpp:33-36).
    mutable std::atomic<int> refs{1};
    hp_ctx_desc desc{};
    std::string device_name;      // owned copy of desc.preferred_device
    hp_version version{HP_VERSION_MAJOR, HP_VERSION_MINOR, HP_VERSION_PATCH};
    int requested_ordinal = -1;
    cudaStream_t user_stream = nullptr;
    bool has_user_stream = false;
    uint32_t reserve_sms = 0;                        // hpx_ctx_ext2: SMs kept out of this context's green context
    mutable void* green_ctx = nullptr;               // CUgreenCtx owning the SMs the stream may use
    mutable uint32_t usable_sms = 0, total_sms = 0;
    // lazily initialised
    mutable bool ready = false;
    mutable bool failed = false;
    mutable int device = 0;
    mutable cudaStream_t stream = nullptr;
    mutable bool owns_stream = false;
    mutable uint32_t* d_status = nullptr;            // device status word for kernels
    mutable unsigned long long* d_total = nullptr;   // device sample total
    mutable uint32_t* h_status = nullptr;            // pinned mirrors
    mutable unsigned long long* h_total = nullptr;
    mutable cudaEvent_t marks[16] = {};              // hpx_ctx_mark / hpx_ctx_elapsed_ms: device-side stage timing
};

struct hp_plan {
    mutable std::atomic<int> refs{1};   // frames and graphs keep their plan alive
    hp_plan_desc desc{};
    const hp_ctx* ctx = nullptr;
    uint32_t uniform_count = 0;   // samples a generated ray emits
    bool gap_free = true;         // emitted samples are steps 0..uniform_count-1
};

enum class FieldKind : uint32_t { kDenseSigma = 0, kDenseColor = 1 };

// Dense-grid field.  The reference keeps a view of the caller's HOST tensor
// (hp_runtime.cpp:259-339); this library snapshots it into HBM at creation.
struct hp_field {
    FieldKind kind = FieldKind::kDenseSigma;
    const hp_ctx* ctx = nullptr;
    hp_tensor source{};           // the caller's view, kept for hp_plan-style introspection
    hp_interp_mode interp = HP_INTERP_LINEAR;
    hp_oob_policy oob = HP_OOB_ZERO;
    int32_t nx = 0, ny = 0, nz = 0, channels = 1;
    int32_t stride = 1;           // floats between consecutive voxels of d_data (= channels for an own snapshot)
    float* d_data = nullptr;      // [nz][ny][nx] x stride floats; channel c of voxel v at d_data[v * stride + c]
    bool owns_data = true;
    // hpx_grid_adopt_fields: the field no longer owns a snapshot but views the packed {r,g,b,sigma} grid (stride 4),
    // so that hpx_grid_update is seen by the staged hp_samp path as well -- one copy of the values in HBM.
    struct hpx_grid* alias_of = nullptr;
};

struct hpx_grid {
    const hp_ctx* ctx = nullptr;
    int32_t nx = 0, ny = 0, nz = 0;
    bool linear = true, clamp = false;
    float bmin[3] = {0.f, 0.f, 0.f}, bmax[3] = {1.f, 1.f, 1.f};
    float4* d_values = nullptr;   // [V] {r,g,b,sigma}; nullptr while the grid is stored as halfs
    void* d_half = nullptr;       // [V] dv::HalfVoxel (hpx_grid_set_storage(HPX_STORAGE_F16)): 8 B per voxel instead of 16
    float* d_grad = nullptr;      // [4V + 16]: packed gradient grid, then camera gradient
    float* d_unpacked = nullptr;  // [V + 3V] staging for un-interleaved read-back (lazily allocated)
    size_t unpacked_voxels = 0;
    size_t voxels = 0;
    std::vector<hp_field*> views;   // fields adopted by hpx_grid_adopt_fields (detached again when the grid goes away)
    // deterministic backward (HPX_BACKWARD_DETERMINISTIC): 64-bit fixed-point shadow of the gradient grid, lazily allocated
    unsigned long long* d_fixed = nullptr;   // [4V], all zero between backward passes
    float* d_fixed_meta = nullptr;           // {bits max|grid value|, bits max|dL/dI|, 1/quantum, quantum}
    bool value_max_stale = true;
    // empty-space skipping (hpx_grid_build_occupancy): 2 bits per brick of 8^3 cells, see dv::PackedGrid::occ
    uint32_t* d_occ = nullptr;
    size_t occ_words = 0;
    unsigned int* d_occ_counts = nullptr;
    bool occ_ready = false, occ_enabled = false;
    // element strides of the gradient block (hpx_grid_set_grad_layout); default: x fastest, z slowest like the values
    int grad_slow_axis = 2;
    uint32_t gsx = 0, gsy = 0, gsz = 0;
};

struct hpx_frame {
    const hp_plan* plan = nullptr;
    const hp_ctx* ctx = nullptr;
    dv::FrameParams h_params{};
    dv::FrameParams* d_params = nullptr;
    dv::FrameParams* h_pinned = nullptr;
    bool params_dirty = true;
    dv::LeanBuffers buf{};
    float* d_dL_dI = nullptr;     // [rays][3]
    double* d_cam_partials = nullptr;
    unsigned int* d_box_miss = nullptr;   // contributions hpx_backward_box had to drop (must stay 0)
    unsigned int* d_group_done = nullptr; // [16] completion counters of hpx_backward_signalled
    size_t device_bytes = 0;
    uint64_t rays = 0, samples = 0;
    bool forward_done = false;
    // hpx_backward_streamed: row groups, slab runs and the copy stream for this frame / grid / camera (dv_comm.cu)
    void* stream_plan = nullptr;
    void (*stream_plan_free)(void*) = nullptr;
    // captured graph
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    std::vector<void*> allocations;
};

namespace dv {

// NVTX range around every exported entry point (a timeline tool shows the hp.h / hp_b200.h calls by
// name; header-only NVTX v3: a no-op unless a profiler is attached).
struct ApiRange {
    explicit ApiRange(const char* name) { nvtxRangePushA(name); }
    ~ApiRange() { nvtxRangePop(); }
    ApiRange(const ApiRange&) = delete;
    ApiRange& operator=(const ApiRange&) = delete;
};
#define DV_RANGE(name) dv::ApiRange dv_api_range__(name)

// ---- errors ----------------------------------------------------------------
void set_last_error(const std::string& what);
const char* last_error_text();
hp_status cuda_fail(cudaError_t err, const char* what);  // records text, maps to hp_status

#define DV_CUDA(call)                                                   \
    do {                                                                \
        const cudaError_t dv_err__ = (call);                            \
        if (dv_err__ != cudaSuccess) return dv::cuda_fail(dv_err__, #call); \
    } while (0)

// Makes ctx's device current and creates its stream on first use.
hp_status ensure_device(const hp_ctx* ctx);

// Entry-point scope: ensure_device + the caller's current device restored on exit (a library call must not change it).
struct DeviceScope {
    int prev = -1;
    bool switched = false;
    hp_status enter(const hp_ctx* ctx);
    ~DeviceScope();
};
#define DV_ENTER(ctx)                                                        \
    dv::DeviceScope dv_scope__;                                              \
    do {                                                                     \
        const hp_status dv_enter_st__ = dv_scope__.enter(ctx);               \
        if (dv_enter_st__ != HP_STATUS_SUCCESS) return dv_enter_st__;        \
    } while (0)

// Reference counting of contexts and plans (see hp_ctx::refs).
const hp_ctx* ctx_retain(const hp_ctx* ctx);
void ctx_unref(const hp_ctx* ctx);
const hp_plan* plan_retain(const hp_plan* plan);
void plan_unref(const hp_plan* plan);

// ---- plan helpers -----------------------------------------------------------
hp_status resolve_plan_desc(hp_plan_desc* desc);   // reference hp_runtime.cpp:54-142
void emitted_samples(const hp_plan_desc& desc, uint32_t* count, bool* gap_free);
FrameParams frame_params_from_plan(const hp_plan& plan);
CameraParams camera_params(const hp_camera_desc& cam);

// ---- tensors ----------------------------------------------------------------
void shape_tensor(hp_tensor& t, hp_dtype dtype, hp_memspace ms, uint32_t rank, int64_t d0, int64_t d1 = 0,
                  int64_t d2 = 0);

// Bump allocator over the caller's workspace (reference workspace.hpp:6-33).
struct Bump {
    char* ptr;
    size_t remaining;
    Bump(void* base, size_t bytes) : ptr(static_cast<char*>(base)), remaining(base ? bytes : 0) {}
    void* take(size_t bytes, size_t alignment = 4);
};

// Scratch device allocations freed on scope exit.
struct DeviceScratch {
    std::vector<void*> ptrs;
    ~DeviceScratch();
    void* take(size_t bytes);   // nullptr on failure (error text recorded)
};

FieldPair field_pair(const hp_field* fs, const hp_field* fc);
ScatterParams scatter_params(const hpx_grid& grid);
// Slabs [lo, hi) of the gradient block (current layout) -> un-interleaved on `stream` and copied into the caller's HOST
// arrays in the reference layout (sigma_grad[V], color_grad[3V]; either may be null), at their own positions there.
// color_stream / ev (optional): the colour copy goes to a second stream (ordered behind the un-interleave kernel by `ev`),
// so that the two pitched copies of a slab range can run on two copy engines at once.
hp_status grid_slabs_to_host(hpx_grid* grid, cudaStream_t stream, int32_t lo, int32_t hi, float* sigma_host, float* color_host,
                             cudaStream_t color_stream = nullptr, cudaEvent_t ev = nullptr);
// Per slab of the grid's CURRENT gradient layout: rows [lo, hi) inside the slab that the frame's backward can touch
// (lo >= hi: none).  out_lo / out_hi: [slabs] ints.  Blocks until done.
hp_status frame_slab_rows(hpx_frame* frame, const hpx_grid* grid, int32_t* out_lo, int32_t* out_hi);
// hpx_frame_bounds of the image rows [row0, row0 + rows) of the frame's ROI (blocks until done).
hp_status frame_rows_bounds(hpx_frame* frame, const hpx_grid* grid, uint32_t row0, uint32_t rows, int32_t out_box[6]);

}  // namespace dv
