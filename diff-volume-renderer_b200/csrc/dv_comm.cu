// dv_comm.cu -- multi-GPU behind the C ABI (hp_b200.h): an NCCL communicator per context and the sharded frame step.
//
// The reference has no multi-device support (single process, default stream).  Rays are independent,
// so the path shards with NO forward collective; the one exchange is the sum of the packed gradient block
// [4 V grid floats | 16 camera floats].
//
//   hpx_comm          one rank of an NCCL communicator bound to an hp_ctx (NCCL is loaded with dlopen: the library has no
//                     link-time dependency on it and every other entry point works without it)
//   hpx_grid_allreduce_grad   plain data parallelism: whole-block all-reduce behind whatever the context's stream holds
//   hpx_shard         ONE frame rendered by all ranks (strong scaling): every rank marches the CTA tile rows t with
//                     t % world == rank (equal mix of short and long rays), the gradient block is laid out with the world
//                     axis the image rows advance along as its slowest axis, and the backward launch signals per row group
//                     (hpx_backward_signalled) so that a high-priority side stream all-reduces, IN PLACE, the slabs a
//                     finished group leaves behind while later rows are still rendering.  With hpx_ctx_ext2.reserve_sms
//                     the rendering kernels keep a few SMs free, so the collective's CTAs start at once instead of
//                     waiting for the rendering launch to drain (profiles/README.md, round 1: that wait exposed the
//                     whole 2.1 GB all-reduce at 8 GPUs).
#include <dlfcn.h>
#include <nccl.h>   // types and enumerators only; every function is resolved with dlsym

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <set>
#include <vector>

#include "dv_objects.h"

using namespace dv;

#define DV_TRY(expr)                                     \
    do {                                                 \
        const hp_status dv_st__ = (expr);                \
        if (dv_st__ != HP_STATUS_SUCCESS) return dv_st__; \
    } while (0)

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;
};

const NcclApi& nccl() {
    static const NcclApi api = [] {
        NcclApi a;
        // a process that already carries NCCL (PyTorch) hands back that copy: dlopen matches the SONAME
        void* h = nullptr;
        const char* env = std::getenv("DVREN_NCCL_LIBRARY");
        for (const char* name : {env, "libnccl.so.2", "libnccl.so"}) {
            if (name == nullptr || *name == '\0') continue;
            h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (h != nullptr) break;
        }
        if (h == nullptr) {
            a.why = "libnccl.so.2 not found (set DVREN_NCCL_LIBRARY)";
            return a;
        }
        auto sym = [&](const char* n) { return dlsym(h, n); };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommInitRankConfig = reinterpret_cast<decltype(a.CommInitRankConfig)>(sym("ncclCommInitRankConfig"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd && a.GetErrorString;
        if (!a.ok) a.why = "the NCCL library lacks a required symbol";
        return a;
    }();
    return api;
}

hp_status nccl_fail(ncclResult_t r, const char* what) {
    set_last_error(std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
    return HP_STATUS_INTERNAL_ERROR;
}

#define DV_NCCL(call)                                                     \
    do {                                                                  \
        const ncclResult_t dv_nr__ = (call);                              \
        if (dv_nr__ != ncclSuccess) return nccl_fail(dv_nr__, #call);     \
    } while (0)

typedef int (*wait_value_fn)(CUstream_st*, unsigned long long, uint32_t, unsigned int);   // cuStreamWaitValue32

wait_value_fn wait_value() {
    static wait_value_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<wait_value_fn>(p);
    }();
    return fn;
}

}  // namespace

struct hpx_comm {
    const hp_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t side = nullptr;        // collectives run here: highest priority, primary context (all SMs visible)
    cudaEvent_t ev_main = nullptr, ev_side = nullptr;
};

struct hpx_shard {
    hpx_comm* comm = nullptr;
    hpx_grid* grid = nullptr;
    hp_plan* plan = nullptr;            // the FULL frame's plan (own handle)
    hpx_frame* frame = nullptr;         // this rank's interleaved tile rows of it
    int slow_axis = 2;
    size_t slab_floats = 0;
    int32_t n_slabs = 0;
    std::vector<uint32_t> group_end_rows;                            // owned tile rows, cumulative
    std::vector<std::vector<std::pair<int32_t, int32_t>>> runs;      // per group: slab runs that are final once it is done
    std::vector<std::pair<int32_t, int32_t>> ranges;                 // per group: slabs it can touch ([lo, hi), lo >= hi: none)
    std::vector<uint32_t> group_rows;                                // image rows per group
    cudaEvent_t ev_zero = nullptr;
    bool reduce = true;
};

namespace {

// Contiguous row bands of the plan's ROI whose heights follow `weights`, cut on multiples of `align` rows.
struct RowBand { uint32_t y0, rows; };

std::vector<RowBand> weighted_bands(const hp_plan_desc& d, const std::vector<float>& weights, uint32_t align) {
    const uint32_t h = d.roi.height;
    const uint32_t units = (h + align - 1) / align;
    double total = 0.0;
    for (float w : weights) total += w;
    std::vector<uint32_t> cuts{0};
    double acc = 0.0;
    for (size_t i = 0; i + 1 < weights.size(); ++i) {
        acc += weights[i];
        const uint32_t c = static_cast<uint32_t>(std::llround(units * acc / total));
        cuts.push_back(std::min(units, std::max(cuts.back(), c)));
    }
    cuts.push_back(units);
    std::vector<RowBand> out;
    for (size_t i = 0; i < weights.size(); ++i) {
        const uint32_t r0 = std::min(cuts[i] * align, h), r1 = std::min(cuts[i + 1] * align, h);
        out.push_back(RowBand{d.roi.y + r0, r1 - r0});
    }
    return out;
}

// Per group, the contiguous runs of slabs that are FINAL once that group is done: touched by it or an earlier group and
// by no later one.  Every touched slab appears in exactly one run, so reducing the runs reduces the gradient once.
std::vector<std::vector<std::pair<int32_t, int32_t>>> final_slab_runs(const std::vector<std::pair<int32_t, int32_t>>& ranges) {
    std::vector<std::vector<std::pair<int32_t, int32_t>>> out;
    std::set<int32_t> done;
    for (size_t g = 0; g < ranges.size(); ++g) {
        std::set<int32_t> touched, later;
        for (size_t i = 0; i <= g; ++i)
            for (int32_t s = ranges[i].first; s < ranges[i].second; ++s) touched.insert(s);
        for (size_t i = g + 1; i < ranges.size(); ++i)
            for (int32_t s = ranges[i].first; s < ranges[i].second; ++s) later.insert(s);
        std::vector<std::pair<int32_t, int32_t>> runs;
        for (int32_t s : touched) {
            if (later.count(s) || done.count(s)) continue;
            done.insert(s);
            if (!runs.empty() && runs.back().second == s) runs.back().second = s + 1;
            else runs.emplace_back(s, s + 1);
        }
        out.push_back(std::move(runs));
    }
    return out;
}

}  // namespace

extern "C" {

HP_API hp_status hpx_comm_unique_id(uint8_t out_id[HPX_COMM_ID_BYTES]) {
    DV_RANGE("hpx_comm_unique_id");
    if (out_id == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    static_assert(sizeof(ncclUniqueId) == HPX_COMM_ID_BYTES, "hp_b200.h: HPX_COMM_ID_BYTES");
    if (!nccl().ok) {
        set_last_error("NCCL is not available: " + nccl().why);
        return HP_STATUS_UNSUPPORTED;
    }
    ncclUniqueId id;
    DV_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(out_id, &id, sizeof(id));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_comm_create(const hp_ctx* ctx, const uint8_t id[HPX_COMM_ID_BYTES], int32_t rank, int32_t world,
                                 int32_t max_ctas, hpx_comm** out_comm) {
    DV_RANGE("hpx_comm_create");
    if (ctx == nullptr || out_comm == nullptr || world < 1 || rank < 0 || rank >= world) return HP_STATUS_INVALID_ARGUMENT;
    if (world > 1 && id == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(ctx);
    hpx_comm* c = new (std::nothrow) hpx_comm();
    if (c == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    c->ctx = ctx_retain(ctx);
    c->rank = rank;
    c->world = world;
    auto fail = [&](hp_status st) {
        hpx_comm_release(c);
        return st;
    };
    int lo = 0, hi = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically lowest = highest priority
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming);
    if (e != cudaSuccess) return fail(cuda_fail(e, "communicator streams"));
    if (world > 1) {
        if (!nccl().ok) {
            set_last_error("NCCL is not available: " + nccl().why);
            return fail(HP_STATUS_UNSUPPORTED);
        }
        ncclUniqueId uid;
        std::memcpy(&uid, id, sizeof(uid));
        ncclResult_t r;
        if (max_ctas > 0 && nccl().CommInitRankConfig != nullptr) {
            ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
            cfg.maxCTAs = max_ctas;   // no more CTAs than the SMs the rendering context leaves free
            r = nccl().CommInitRankConfig(&c->comm, world, uid, rank, &cfg);
        } else {
            r = nccl().CommInitRank(&c->comm, world, uid, rank);
        }
        if (r != ncclSuccess) return fail(nccl_fail(r, "ncclCommInitRank"));
    }
    *out_comm = c;
    return HP_STATUS_SUCCESS;
}

HP_API void hpx_comm_release(hpx_comm* c) {
    DV_RANGE("hpx_comm_release");
    if (c == nullptr) return;
    if (c->ctx != nullptr && c->ctx->ready) {
        DeviceScope scope;
        scope.enter(c->ctx);
        if (c->side != nullptr) cudaStreamSynchronize(c->side);
        if (c->comm != nullptr && nccl().ok) nccl().CommDestroy(c->comm);
        if (c->side != nullptr) cudaStreamDestroy(c->side);
        if (c->ev_main != nullptr) cudaEventDestroy(c->ev_main);
        if (c->ev_side != nullptr) cudaEventDestroy(c->ev_side);
    }
    ctx_unref(c->ctx);
    delete c;
}

HP_API hp_status hpx_comm_info(const hpx_comm* c, int32_t* out_rank, int32_t* out_world, int32_t* out_nccl_version) {
    DV_RANGE("hpx_comm_info");
    if (c == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (out_rank) *out_rank = c->rank;
    if (out_world) *out_world = c->world;
    if (out_nccl_version) {
        int v = 0;
        if (nccl().ok && nccl().GetVersion) nccl().GetVersion(&v);
        *out_nccl_version = v;
    }
    return HP_STATUS_SUCCESS;
}

// In-place sum over the ranks of `floats` floats at device_buf, ordered after everything already enqueued on the context's
// stream; the context's stream continues only when the sum is there.
HP_API hp_status hpx_comm_allreduce(hpx_comm* c, float* device_buf, size_t floats) {
    DV_RANGE("hpx_comm_allreduce");
    if (c == nullptr || (device_buf == nullptr && floats != 0)) return HP_STATUS_INVALID_ARGUMENT;
    if (c->world == 1 || floats == 0) return HP_STATUS_SUCCESS;
    DV_ENTER(c->ctx);
    DV_CUDA(cudaEventRecord(c->ev_main, c->ctx->stream));
    DV_CUDA(cudaStreamWaitEvent(c->side, c->ev_main, 0));
    DV_NCCL(nccl().AllReduce(device_buf, device_buf, floats, ncclFloat32, ncclSum, c->comm, c->side));
    DV_CUDA(cudaEventRecord(c->ev_side, c->side));
    DV_CUDA(cudaStreamWaitEvent(c->ctx->stream, c->ev_side, 0));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_grid_allreduce_grad(hpx_comm* c, hpx_grid* g) {
    DV_RANGE("hpx_grid_allreduce_grad");
    if (c == nullptr || g == nullptr || c->ctx != g->ctx) return HP_STATUS_INVALID_ARGUMENT;
    float* buf = nullptr;
    size_t floats = 0;
    DV_TRY(hpx_grid_grad_buffer(g, &buf, &floats));
    return hpx_comm_allreduce(c, buf, floats);
}

// ---- one frame over all ranks ----------------------------------------------------------------------------------
HP_API hp_status hpx_shard_create(hpx_comm* c, const hp_plan* full_plan, hpx_grid* g, const float* group_weights,
                                  uint32_t n_groups, hpx_shard** out_shard) {
    DV_RANGE("hpx_shard_create");
    if (c == nullptr || full_plan == nullptr || g == nullptr || out_shard == nullptr || n_groups == 0 || n_groups > 16)
        return HP_STATUS_INVALID_ARGUMENT;
    if (full_plan->ctx != c->ctx || g->ctx != c->ctx) {
        set_last_error("communicator, plan and grid must come from one context");
        return HP_STATUS_INVALID_ARGUMENT;
    }
    if (!g->linear || g->clamp || scatter_params(*g).unit_bbox == 0u) {
        set_last_error("hpx_shard needs a linear OOB-zero field whose scatter box is the unit cube (hpx_frame_bounds)");
        return HP_STATUS_UNSUPPORTED;
    }
    DV_ENTER(c->ctx);
    hpx_shard* s = new (std::nothrow) hpx_shard();
    if (s == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    s->comm = c;
    s->grid = g;
    auto fail = [&](hp_status st) {
        hpx_shard_release(s);
        return st;
    };
    const hp_plan_desc& d = full_plan->desc;
    // world axis the image rows advance along = the camera's "down" vector (second column of c2w's rotation)
    const float down[3] = {std::fabs(d.camera.c2w[1]), std::fabs(d.camera.c2w[5]), std::fabs(d.camera.c2w[9])};
    s->slow_axis = down[0] > down[1] ? (down[0] > down[2] ? 0 : 2) : (down[1] >= down[2] ? 1 : 2);
    hp_status st = hpx_grid_set_grad_layout(g, s->slow_axis, &s->slab_floats, &s->n_slabs);
    if (st != HP_STATUS_SUCCESS) return fail(st);
    hp_plan_desc copy = d;
    st = hp_plan_create(c->ctx, &copy, &s->plan);
    if (st == HP_STATUS_SUCCESS) st = hpx_frame_create(s->plan, &s->frame);
    if (st == HP_STATUS_SUCCESS) st = hpx_frame_set_interleave(s->frame, static_cast<uint32_t>(c->world), static_cast<uint32_t>(c->rank));
    if (st != HP_STATUS_SUCCESS) return fail(st);
    if (cudaEventCreateWithFlags(&s->ev_zero, cudaEventDisableTiming) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "event"));

    std::vector<float> weights(n_groups, 1.0f);
    if (group_weights != nullptr) weights.assign(group_weights, group_weights + n_groups);
    const uint32_t tile_rows_px = kTileH * kWarpsY;
    const uint32_t stride = static_cast<uint32_t>(c->world), phase = static_cast<uint32_t>(c->rank);
    uint32_t owned = 0;
    for (const RowBand& b : weighted_bands(d, weights, tile_rows_px * stride)) {
        if (b.rows == 0) continue;
        // slabs the WHOLE band (all ranks' tile rows) can touch: a probe frame over the band with every tile row
        hp_plan_desc bd = d;
        bd.roi.y = b.y0;
        bd.roi.height = b.rows;
        bd.max_rays = 0;
        bd.max_samples = 0;
        hp_plan* probe_plan = nullptr;
        hpx_frame* probe = nullptr;
        int32_t box[6] = {0, 0, 0, 0, 0, 0};
        st = hp_plan_create(c->ctx, &bd, &probe_plan);
        if (st == HP_STATUS_SUCCESS) st = hpx_frame_create(probe_plan, &probe);
        if (st == HP_STATUS_SUCCESS) st = hpx_frame_bounds(probe, g, box);
        hpx_frame_release(probe);
        hp_plan_release(probe_plan);
        if (st != HP_STATUS_SUCCESS) return fail(st);
        const int32_t lo = box[s->slow_axis], n = box[3 + s->slow_axis];
        s->ranges.emplace_back(n > 0 ? lo : 0, n > 0 ? lo + n : 0);
        const uint32_t tile_rows = (b.rows + tile_rows_px - 1) / tile_rows_px;
        owned += tile_rows > phase ? (tile_rows - phase + stride - 1) / stride : 0u;
        s->group_end_rows.push_back(owned);
        s->group_rows.push_back(b.rows);
    }
    if (s->group_end_rows.empty()) return fail(HP_STATUS_INVALID_ARGUMENT);
    s->runs = final_slab_runs(s->ranges);
    *out_shard = s;
    return HP_STATUS_SUCCESS;
}

HP_API void hpx_shard_release(hpx_shard* s) {
    DV_RANGE("hpx_shard_release");
    if (s == nullptr) return;
    if (s->comm != nullptr && s->comm->ctx != nullptr && s->comm->ctx->ready) {
        DeviceScope scope;
        scope.enter(s->comm->ctx);
        cudaStreamSynchronize(s->comm->ctx->stream);
        if (s->comm->side != nullptr) cudaStreamSynchronize(s->comm->side);
        if (s->ev_zero != nullptr) cudaEventDestroy(s->ev_zero);
    }
    hpx_frame_release(s->frame);
    hp_plan_release(s->plan);
    if (s->grid != nullptr) hpx_grid_set_grad_layout(s->grid, 2, nullptr, nullptr);   // back to the default order
    delete s;
}

HP_API hp_status hpx_shard_frame(hpx_shard* s, hpx_frame** out_frame) {
    DV_RANGE("hpx_shard_frame");
    if (s == nullptr || out_frame == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    *out_frame = s->frame;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_shard_set_reduce(hpx_shard* s, int32_t enabled) {
    DV_RANGE("hpx_shard_set_reduce");
    if (s == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    s->reduce = enabled != 0;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_shard_layout(const hpx_shard* s, int32_t* out_slow_axis, uint32_t* out_groups, uint32_t* out_group_rows,
                                  int32_t* out_slab_ranges) {
    DV_RANGE("hpx_shard_layout");
    if (s == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (out_slow_axis) *out_slow_axis = s->slow_axis;
    if (out_groups) *out_groups = static_cast<uint32_t>(s->group_end_rows.size());
    for (size_t i = 0; i < s->group_rows.size(); ++i) {
        if (out_group_rows) out_group_rows[i] = s->group_rows[i];
        if (out_slab_ranges) {
            out_slab_ranges[2 * i] = s->ranges[i].first;
            out_slab_ranges[2 * i + 1] = s->ranges[i].second;
        }
    }
    return HP_STATUS_SUCCESS;
}

// One step: [zero the gradient block] -> forward of this rank's tile rows -> ONE signalled backward launch -> per row group,
// on the side stream: wait for the group's counter, all-reduce in place the slabs it finished.  Leaves the summed
// gradient of all ranks in the grid's gradient block (slab order of hpx_grid_set_grad_layout; hpx_grid_read_grad returns
// the reference order).  dL_dI: DEVICE pointer, (rays of the WHOLE frame, 3).
HP_API hp_status hpx_shard_step(hpx_shard* s, const float* dL_dI_device, uint32_t flags) {
    DV_RANGE("hpx_shard_step");
    if (s == nullptr || dL_dI_device == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    hpx_comm* c = s->comm;
    DV_ENTER(c->ctx);
    cudaStream_t main = c->ctx->stream, side = c->side;
    const bool collect = c->world > 1 && s->reduce;
    if (flags & HPX_BACKWARD_ZERO) {
        // the gradient block is cleared on the side stream WHILE the forward kernel runs (a bandwidth-bound memset next to
        // an L1-bound kernel): 2.1 GB at 512^3 would otherwise add its full 0.35 ms to every step
        float* buf = nullptr;
        size_t floats = 0;
        DV_TRY(hpx_grid_grad_buffer(s->grid, &buf, &floats));
        DV_CUDA(cudaEventRecord(c->ev_main, main));
        DV_CUDA(cudaStreamWaitEvent(side, c->ev_main, 0));
        DV_CUDA(cudaMemsetAsync(buf, 0, floats * sizeof(float), side));
        DV_CUDA(cudaEventRecord(s->ev_zero, side));
    }
    DV_TRY(hpx_forward(s->frame, s->grid));
    if (flags & HPX_BACKWARD_ZERO) DV_CUDA(cudaStreamWaitEvent(main, s->ev_zero, 0));
    uint32_t* counters = nullptr;
    DV_TRY(hpx_frame_reset_group_counters(s->frame, &counters));
    DV_CUDA(cudaEventRecord(c->ev_main, main));   // counters cleared: the previous step's counts cannot satisfy the waits
    uint32_t expected[16] = {};
    const uint32_t n_groups = static_cast<uint32_t>(s->group_end_rows.size());
    DV_TRY(hpx_backward_signalled(s->frame, s->grid, dL_dI_device, HP_MEMSPACE_DEVICE, flags & ~HPX_BACKWARD_ZERO,
                                  s->group_end_rows.data(), n_groups, &counters, expected));
    if (!collect) return HP_STATUS_SUCCESS;
    if (wait_value() == nullptr) {
        set_last_error("cuStreamWaitValue32 is not available from this driver");
        return HP_STATUS_UNSUPPORTED;
    }
    float* block = nullptr;
    size_t floats = 0;
    DV_TRY(hpx_grid_grad_buffer(s->grid, &block, &floats));
    DV_CUDA(cudaStreamWaitEvent(side, c->ev_main, 0));
    for (uint32_t g = 0; g < n_groups; ++g) {
        const int rc = wait_value()(side, reinterpret_cast<unsigned long long>(counters + g), expected[g], 0u /* GEQ */);
        if (rc != 0) {
            set_last_error("cuStreamWaitValue32 failed with driver error " + std::to_string(rc));
            return HP_STATUS_INTERNAL_ERROR;
        }
        const bool last = g + 1 == n_groups;
        if (s->runs[g].empty() && !last) continue;
        DV_NCCL(nccl().GroupStart());
        for (const auto& run : s->runs[g]) {
            float* p = block + static_cast<size_t>(run.first) * s->slab_floats;
            const ncclResult_t r = nccl().AllReduce(p, p, static_cast<size_t>(run.second - run.first) * s->slab_floats, ncclFloat32,
                                                    ncclSum, c->comm, side);
            if (r != ncclSuccess) {
                nccl().GroupEnd();
                return nccl_fail(r, "ncclAllReduce(slabs)");
            }
        }
        if (last) {   // the camera gradients ride with the last group
            float* cam = block + floats - 16;
            const ncclResult_t r = nccl().AllReduce(cam, cam, 16, ncclFloat32, ncclSum, c->comm, side);
            if (r != ncclSuccess) {
                nccl().GroupEnd();
                return nccl_fail(r, "ncclAllReduce(camera)");
            }
        }
        DV_NCCL(nccl().GroupEnd());
    }
    DV_CUDA(cudaEventRecord(c->ev_side, side));
    DV_CUDA(cudaStreamWaitEvent(main, c->ev_side, 0));
    return HP_STATUS_SUCCESS;
}

}  // extern "C"
