// dv_comm.cu -- multi-GPU behind the C ABI (hp_b200.h) and the streamed gradient read-back.
//
// The reference has no multi-device support (single process, default stream).  Rays are independent,
// so the path shards with NO forward collective; the one exchange is the sum of the packed gradient block
// [4 V grid floats | 16 camera floats].
//
//   hpx_comm          one rank of an NCCL communicator bound to an hp_ctx (NCCL is loaded with dlopen: the library has no
//                     link-time dependency on it and every other entry point works without it)
//   hpx_grid_allreduce_grad   data parallelism over views: whole-block all-reduce behind whatever the context's stream holds
//   hpx_shard (bands) ONE frame rendered by all ranks (strong scaling), the default: contiguous row bands cut for equal
//                     marching work and re-cut from measured time, gradient block in slabs along the world axis the image
//                     rows advance along, so that a band's backward touches one slab wedge; every slab has an owner, and
//                     after a stream-ordered cross-GPU barrier each owner PULLS the rows of its slabs its neighbours' rays
//                     can reach straight out of their gradient blocks (peer access / CUDA IPC over NVLink) and adds them
//                     in rank order (peer_reduce_kernel).  Result: reduce-scatter (owned) or, with one more gather kernel,
//                     the whole sum on every rank (replicated).  NCCL send/recv + broadcast is the fallback path.
//   hpx_shard (interleaved)   round 1's design, kept for comparison: tile rows t % world == rank, ONE backward launch that
//                     signals per row group, slab all-reduces on a priority stream behind it; optional SM reservation
//                     through a green context (hpx_ctx_ext2.reserve_sms).  Measured slower than the bands at every N.
//   hpx_backward_streamed     single GPU: the same per-row-group signals drive a copy stream that un-interleaves and
//                     copies finished slabs to the caller's HOST arrays while later rows still render.
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>   // types and enumerators only; every function is resolved with dlsym

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
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
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
        a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
        a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
        a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.Send && a.Recv && a.Broadcast && a.AllGather && a.GroupStart &&
               a.GroupEnd && a.GetErrorString;
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
    // ---- band mode (hpx_shard_create_bands) ----
    bool bands = false;
    int result = HPX_SHARD_RESULT_REPLICATED;
    std::vector<uint32_t> band_row0, band_rows;                     // per rank: first image row (inside the ROI), rows
    std::vector<std::pair<int32_t, int32_t>> wedges;                // per rank: slabs its rows can touch [lo, hi)
    std::vector<int32_t> cuts;                                      // world + 1: rank r owns slabs [cuts[r], cuts[r + 1])
    struct Xfer { int peer; int32_t lo, hi; size_t staging_off; };  // slab range, offset (floats) in `staging`
    std::vector<Xfer> sends, recvs;
    float* staging = nullptr;
    int32_t hull_lo = 0, hull_hi = 0;                               // slabs ANY rank can touch
    size_t dl_offset_floats = 0;                                    // this rank's rows inside the frame's dL/dI
    int tile_order = 0;                                             // hpx_frame_set_row_order of this rank's band (best_tile_order)
    bool order_tuned = false;                                       // ... chosen by measurement (hpx_shard_tune_order)
    // direct exchange: every rank's gradient block mapped into this GPU's address space (same process: peer access;
    // other processes: CUDA IPC), read by this rank's own kernels over NVLink
    bool direct = false;
    std::vector<float*> peer_block;                                 // per rank (own entry = own block)
    std::vector<void*> ipc_opened;
    int* d_flag = nullptr;                                          // 1 int: payload of the cross-GPU barrier
    // per rank and slab: the rows [lo, hi) inside the slab that rank's band can touch (frame_slab_rows); the direct exchange
    // moves only those rows of a slab
    std::vector<std::vector<int2>> slab_rows;                       // [world][n_slabs]
    int2* d_slab_rows = nullptr;                                    // the same on the device, rank-major
    int32_t rows_per_slab = 0;
    size_t row_floats = 0;
    // measured rebalancing (hpx_shard_rebalance)
    hp_plan_desc full_desc{};
    FrameParams full_params{};
    std::vector<double> unit_cost;                                  // estimated work per CTA tile row, rescaled by measurements
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;                   // around this rank's forward + backward
    bool timed = false;
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
typedef std::vector<std::pair<int32_t, int32_t>> SlabIntervals;

std::vector<SlabIntervals> final_slab_runs(const std::vector<SlabIntervals>& ranges) {
    std::vector<SlabIntervals> out;
    std::set<int32_t> done;
    for (size_t g = 0; g < ranges.size(); ++g) {
        std::set<int32_t> touched, later;
        for (size_t i = 0; i <= g; ++i)
            for (const auto& r : ranges[i])
                for (int32_t s = r.first; s < r.second; ++s) touched.insert(s);
        for (size_t i = g + 1; i < ranges.size(); ++i)
            for (const auto& r : ranges[i])
                for (int32_t s = r.first; s < r.second; ++s) later.insert(s);
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

std::vector<SlabIntervals> final_slab_runs(const std::vector<std::pair<int32_t, int32_t>>& ranges) {
    std::vector<SlabIntervals> sets;
    for (const auto& r : ranges) sets.push_back(SlabIntervals{r});
    return final_slab_runs(sets);
}

// ---- band mode helpers ------------------------------------------------------------------------------------------
// dst[i] += src[i] (float4 lanes; counts are multiples of 4 floats: a slab is nx * ny * 4 floats)
__global__ void add_slabs_kernel(float4* __restrict__ dst, const float4* __restrict__ src, size_t n4) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float4 a = dst[i];
        const float4 b = __ldg(src + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        dst[i] = a;
    }
}


// ---- exchange kernels: ONE launch per phase, all peers at once ---------------------------------------------------------
// Reduce phase: the owner walks the slabs it owns; for every element it adds, in ascending rank order (the sum never
// depends on timing), the partial sums of every peer whose wedge contains the slab.  The loads from the different peers
// of an element are independent and issued back to back, so all NVLink ports that feed this GPU are busy at once, and
// the owner's own block is read and written once instead of once per peer.
constexpr int kMaxPeers = 16;
struct PullTable {
    const float4* src[kMaxPeers];    // peer block (mapped into this GPU's address space)
    const int2* rows[kMaxPeers];     // per slab: the rows [x, y) of it that peer's band can touch (only those are read)
    int32_t lo[kMaxPeers], hi[kMaxPeers];   // slabs of that peer's wedge inside this rank's owned range
    int32_t n;
    int32_t first_slab;              // blockIdx.y = 0
    uint32_t row4;                   // float4 per row
};

__global__ void __launch_bounds__(256) peer_reduce_kernel(float4* __restrict__ block, PullTable t, size_t slab4) {
    const int32_t y = t.first_slab + static_cast<int32_t>(blockIdx.y);
    uint32_t mask = 0;
    int32_t r_lo[kMaxPeers], r_hi[kMaxPeers], u_lo = INT_MAX, u_hi = 0;
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
        r_lo[p] = 0; r_hi[p] = 0;
        if (p < t.n && y >= t.lo[p] && y < t.hi[p]) {
            const int2 r = __ldg(t.rows[p] + y);
            if (r.y > r.x) {
                mask |= 1u << p;
                r_lo[p] = r.x; r_hi[p] = r.y;
                u_lo = min(u_lo, r.x); u_hi = max(u_hi, r.y);
            }
        }
    }
    if (mask == 0u) return;
    const size_t base = static_cast<size_t>(y) * slab4;
    const size_t begin = static_cast<size_t>(u_lo) * t.row4, end = static_cast<size_t>(u_hi) * t.row4;
    for (size_t i = begin + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < end; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int32_t row = static_cast<int32_t>(i / t.row4);
        uint32_t m = 0;
        float4 v[kMaxPeers];
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
            if (p < t.n && ((mask >> p) & 1u) && row >= r_lo[p] && row < r_hi[p]) { v[p] = t.src[p][base + i]; m |= 1u << p; }
        if (m == 0u) continue;
        float4 a = block[base + i];
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
            if ((m >> p) & 1u) { a.x += v[p].x; a.y += v[p].y; a.z += v[p].z; a.w += v[p].w; }
        block[base + i] = a;
    }
}

// Gather phase (replicated result): every slab of the touched hull this rank does not own is copied from its owner.
struct OwnerTable {
    const float4* src[kMaxPeers];   // by rank
    int32_t cuts[kMaxPeers + 1];
    int32_t world, me, first_slab, rotate;
};

__global__ void __launch_bounds__(256) peer_gather_kernel(float4* __restrict__ block, OwnerTable t, size_t slab4) {
    // CTAs are dispatched in blockIdx order: every rank starts behind its OWN slabs, so that at any time the ranks read from
    // different owners (rank r from r + 1, r + 2, ...) instead of all draining one owner's outbound port after the other
    const int32_t y = t.first_slab + static_cast<int32_t>((blockIdx.y + static_cast<uint32_t>(t.rotate)) % gridDim.y);
    int owner = 0;
    while (owner + 1 < t.world && y >= t.cuts[owner + 1]) ++owner;
    if (owner == t.me) return;
    const float4* __restrict__ src = t.src[owner] + static_cast<size_t>(y) * slab4;
    float4* __restrict__ dst = block + static_cast<size_t>(y) * slab4;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    for (; i + 3 * stride < slab4; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = src[i + k * stride];
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[i + k * stride] = v[k];
    }
    for (; i < slab4; i += stride) dst[i] = src[i];
}

struct PeerInfo {
    cudaIpcMemHandle_t handle;
    unsigned long long ptr;
    long long pid;
    int device;
    int ok;
};

// Maps every rank's gradient block into this GPU's address space.  Collective: all ranks call it; returns with
// s->direct == true on ALL ranks or on none (any rank that cannot map a peer vetoes).
hp_status map_peer_blocks(hpx_shard* s) {
    hpx_comm* c = s->comm;
    const int world = c->world, me = c->rank;
    cudaStream_t main = c->ctx->stream;
    float* block = nullptr;
    size_t floats = 0;
    DV_TRY(hpx_grid_grad_buffer(s->grid, &block, &floats));
    DV_CUDA(cudaMalloc(&s->d_flag, sizeof(int)));
    DV_CUDA(cudaMemsetAsync(s->d_flag, 0, sizeof(int), main));
    s->peer_block.assign(static_cast<size_t>(world), nullptr);
    s->peer_block[static_cast<size_t>(me)] = block;
    if (world == 1) return HP_STATUS_SUCCESS;
    const char* env = std::getenv("DVREN_SHARD_EXCHANGE");
    int ok = (env != nullptr && std::strcmp(env, "nccl") == 0) || world > kMaxPeers ? 0 : 1;
    PeerInfo mine{};
    mine.ptr = reinterpret_cast<unsigned long long>(block);
    mine.pid = static_cast<long long>(getpid());
    mine.device = c->ctx->device;
    if (ok && cudaIpcGetMemHandle(&mine.handle, block) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    mine.ok = ok;
    PeerInfo* d_all = nullptr;
    DV_CUDA(cudaMalloc(&d_all, sizeof(PeerInfo) * world));
    std::vector<PeerInfo> all(static_cast<size_t>(world));
    hp_status st = HP_STATUS_SUCCESS;
    do {
        if (cudaMemcpyAsync(d_all + me, &mine, sizeof(PeerInfo), cudaMemcpyHostToDevice, main) != cudaSuccess) { st = cuda_fail(cudaGetLastError(), "peer info"); break; }
        const ncclResult_t r = nccl().AllGather(d_all + me, d_all, sizeof(PeerInfo), ncclChar, c->comm, main);
        if (r != ncclSuccess) { st = nccl_fail(r, "ncclAllGather(peer info)"); break; }
        if (cudaMemcpyAsync(all.data(), d_all, sizeof(PeerInfo) * world, cudaMemcpyDeviceToHost, main) != cudaSuccess ||
            cudaStreamSynchronize(main) != cudaSuccess) { st = cuda_fail(cudaGetLastError(), "peer info"); break; }
    } while (false);
    cudaFree(d_all);
    if (st != HP_STATUS_SUCCESS) return st;
    for (int r = 0; r < world && ok; ++r) {
        if (r == me) continue;
        if (!all[r].ok) { ok = 0; break; }
        if (all[r].pid == mine.pid) {   // another thread of this process: plain peer access
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c->ctx->device, all[r].device) != cudaSuccess || !can) { cudaGetLastError(); ok = 0; break; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ok = 0; break; }
            cudaGetLastError();
            s->peer_block[static_cast<size_t>(r)] = reinterpret_cast<float*>(all[r].ptr);
        } else {
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            s->ipc_opened.push_back(p);
            s->peer_block[static_cast<size_t>(r)] = static_cast<float*>(p);
        }
    }
    // unanimous or not at all
    int* d_ok = s->d_flag;
    DV_CUDA(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, main));
    DV_NCCL(nccl().AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, c->comm, main));
    int all_ok = 0;
    DV_CUDA(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, main));
    DV_CUDA(cudaStreamSynchronize(main));
    s->direct = all_ok != 0;
    return HP_STATUS_SUCCESS;
}

// Cross-GPU barrier in stream order: returns (on the stream) once every rank's stream has reached it.
hp_status stream_barrier(hpx_shard* s) {
    hpx_comm* c = s->comm;
    DV_NCCL(nccl().AllReduce(s->d_flag, s->d_flag, 1, ncclInt32, ncclMax, c->comm, c->ctx->stream));
    return HP_STATUS_SUCCESS;
}

// Marching work of one image row, in samples: steps of its rays that lie inside the unit cube (what the kernels spend
// their time on; steps outside are skipped by range) plus a small constant per ray.  Host restatement of make_ray /
// cube_interval on every 4th pixel -- an ESTIMATE that only places the band boundaries, no result depends on it.
double pixel_work(const FrameParams& p, uint32_t px, uint32_t gy) {
    const CameraParams& c = p.cam;
    const MarchParams& m = p.march;
    const float t_last = std::min(m.t_far, m.t_near + static_cast<float>(m.uniform_count) * m.dt);
    float qx = (static_cast<float>(px) + 0.5f - c.cx) / c.fx, qy = (static_cast<float>(gy) + 0.5f - c.cy) / c.fy;
    if (c.ortho) qx = qy = 0.0f;
    const float v[3] = {c.r00 * qx + c.r01 * qy + c.r02, c.r10 * qx + c.r11 * qy + c.r12, c.r20 * qx + c.r21 * qy + c.r22};
    const float len = std::sqrt(std::max(v[0] * v[0] + v[1] * v[1] + v[2] * v[2], 1e-30f));
    const float o[3] = {c.ox, c.oy, c.oz};
    float t_in = m.t_near, t_out = t_last;
    for (int i = 0; i < 3 && t_in < t_out; ++i) {
        const float d = v[i] / len;
        if (std::fabs(d) > 1e-12f) {
            const float a = (0.0f - o[i]) / d, b = (1.0f - o[i]) / d;
            t_in = std::max(t_in, std::min(a, b));
            t_out = std::min(t_out, std::max(a, b));
        } else if (o[i] < 0.0f || o[i] > 1.0f) {
            t_out = t_in;
        }
    }
    return 8.0 + (t_out > t_in && m.dt > 0.0f ? static_cast<double>((t_out - t_in) / m.dt) : 0.0);
}

double row_work(const FrameParams& p, uint32_t gy, uint32_t x0, uint32_t w) {
    double work = 0.0;
    uint32_t n = 0;
    for (uint32_t px = x0; px < x0 + w; px += 4, ++n) work += pixel_work(p, px, gy);
    return n != 0 ? work * static_cast<double>(w) / n : 0.0;
}

// Which dispatch order (hpx_frame_set_row_order) ends the launches of the band [y0, y0 + rows) of the frame soonest: the
// CTAs of a launch are handed to free slots in blockIdx order, a CTA lasts as long as its longest ray (the warps march in
// lock-step), and the launch is over when the last one is -- list scheduling of the tile estimates on the resident slots,
// once per candidate order.  A band of a few tile rows is only 4-5 waves of CTAs: ending with full-length tiles leaves
// the last wave partly empty (estimated and measured: 12 % of the launch for the middle bands of 512^3 / 8 GPUs).
// An ESTIMATE that only orders the work; no result depends on it.
int best_tile_order(const FrameParams& p, const hp_plan_desc& d, uint32_t y0, uint32_t rows_px, uint32_t slots, double* out_ends4 = nullptr) {
    const uint32_t tw = kTileW * kWarpsX, th = kTileH * kWarpsY;
    const uint32_t tiles_x = (d.roi.width + tw - 1) / tw, rows = (rows_px + th - 1) / th;
    if (tiles_x == 0 || rows == 0 || slots == 0) return 0;
    std::vector<double> cost(static_cast<size_t>(tiles_x) * rows, 0.0);
    for (uint32_t r = 0; r < rows; ++r)
        for (uint32_t cx = 0; cx < tiles_x; ++cx) {
            double worst = 0.0;
            for (uint32_t sy = 0; sy < 3; ++sy)
                for (uint32_t sx = 0; sx < 3; ++sx) {
                    const uint32_t px = std::min(d.roi.width - 1, cx * tw + sx * (tw - 1) / 2);
                    const uint32_t py = std::min(rows_px - 1, r * th + sy * (th - 1) / 2);
                    worst = std::max(worst, pixel_work(p, d.roi.x + px, d.roi.y + y0 + py));
                }
            cost[static_cast<size_t>(r) * tiles_x + cx] = worst + 32.0;   // + prologue / epilogue of a CTA, in steps
        }
    std::vector<double> slot_free;
    const int candidates[4] = {0, 1, static_cast<int>(HPX_ORDER_COLUMNS), static_cast<int>(HPX_ORDER_COLUMNS) | 1};
    double ends[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < 4; ++k) {
        const int order = candidates[k];
        slot_free.assign(slots, 0.0);
        std::make_heap(slot_free.begin(), slot_free.end(), std::greater<double>());
        double end = 0.0;
        for (uint32_t b = 0; b < tiles_x * rows; ++b) {
            uint32_t cx = 0, r = 0;
            tile_of(b, tiles_x, rows, static_cast<uint32_t>(order), &cx, &r);
            std::pop_heap(slot_free.begin(), slot_free.end(), std::greater<double>());
            slot_free.back() += cost[static_cast<size_t>(r) * tiles_x + cx];
            end = std::max(end, slot_free.back());
            std::push_heap(slot_free.begin(), slot_free.end(), std::greater<double>());
        }
        ends[k] = end;
        if (out_ends4 != nullptr) out_ends4[k] = end;
    }
    // row by row unless a column order wins by more than the estimate's noise: where the two tie in the estimate (bands of
    // many tile rows) the row orders measured up to 2 % faster (512^3 / 2 GPUs: 28.4 vs 29.0 ms); hpx_shard_tune_order
    // settles it by measurement
    int best = candidates[0];
    double best_end = ends[0];
    for (int k : {1, 2, 3})
        if (ends[k] < best_end * 0.995) {
            best = candidates[k];
            best_end = ends[k];
        }
    return best;
}

// Contiguous bands of CTA tile rows with (nearly) equal work.
// Estimated marching work of every CTA tile row (8 image rows) of the plan's ROI.
std::vector<double> unit_costs(const FrameParams& p, const hp_plan_desc& d) {
    const uint32_t unit = kTileH * kWarpsY, h = d.roi.height;
    const uint32_t units = (h + unit - 1) / unit;
    std::vector<double> cost(units, 0.0);
    for (uint32_t u = 0; u < units; ++u)
        for (uint32_t r = u * unit; r < std::min(h, (u + 1) * unit); r += 2)   // every 2nd row of the unit
            cost[u] += row_work(p, d.roi.y + r, d.roi.x, d.roi.width);
    return cost;
}

std::vector<RowBand> partition_units(const std::vector<double>& cost, uint32_t h, uint32_t world) {
    const uint32_t unit = kTileH * kWarpsY;
    const uint32_t units = static_cast<uint32_t>(cost.size());
    std::vector<double> cum(units + 1, 0.0);
    for (uint32_t u = 0; u < units; ++u) cum[u + 1] = cum[u] + cost[u];
    // smallest cap such that `world` contiguous bands of at most `cap` work cover all units (binary search over the cap,
    // greedy packing as the feasibility test): minimises the work of the busiest rank
    auto bands_needed = [&](double cap, std::vector<uint32_t>* cuts) {
        uint32_t n = 0, start = 0;
        while (start < units) {
            uint32_t end = start + 1;   // at least one unit per band
            while (end < units && cum[end + 1] - cum[start] <= cap) ++end;
            if (cuts) cuts->push_back(end);
            start = end;
            ++n;
        }
        return n;
    };
    double lo = 0.0, hi = cum[units];
    for (uint32_t u = 0; u < units; ++u) lo = std::max(lo, cum[u + 1] - cum[u]);
    for (int it = 0; it < 50 && hi - lo > 1e-6 * hi; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (bands_needed(mid, nullptr) <= world) hi = mid; else lo = mid;
    }
    std::vector<uint32_t> cuts;
    bands_needed(hi, &cuts);
    std::vector<RowBand> out;
    uint32_t prev = 0;
    for (uint32_t b = 0; b < world; ++b) {
        const uint32_t cut = b < cuts.size() ? (b + 1 == world ? units : cuts[b]) : units;
        const uint32_t r0 = std::min(prev * unit, h), r1 = std::min(cut * unit, h);
        out.push_back(RowBand{r0, r1 - r0});   // y0 relative to the ROI
        prev = cut;
    }
    return out;
}

std::vector<RowBand> balanced_bands(const FrameParams& p, const hp_plan_desc& d, uint32_t world) {
    return partition_units(unit_costs(p, d), d.roi.height, world);
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


// Pure host logic of the owner search (see band_assign_owners): wedges [lo, hi) per rank, n slabs, touched hull.
static std::vector<int32_t> choose_owner_cuts(const std::vector<std::pair<int32_t, int32_t>>& wedges, int32_t n, int32_t hull_lo,
                                              int32_t hull_hi, bool replicated) {
    const int world = static_cast<int>(wedges.size());
    auto overlap_len = [&](int r, int32_t lo, int32_t hi) {
        return std::max(0, std::min(wedges[r].second, hi) - std::max(wedges[r].first, lo));
    };
    auto cost = [&](const std::vector<int32_t>& cuts) {
        double worst = 0.0, total = 0.0;
        for (int r = 0; r < world; ++r) {
            const int32_t own_lo = cuts[static_cast<size_t>(r)], own_hi = cuts[static_cast<size_t>(r) + 1];
            double out = (wedges[r].second - wedges[r].first) - overlap_len(r, own_lo, own_hi), in = 0.0;
            if (out < 0.0) out = 0.0;
            for (int q = 0; q < world; ++q)
                if (q != r) in += overlap_len(q, own_lo, own_hi);
            total += out;
            if (replicated) {
                const double own = std::max(0, std::min(own_hi, hull_hi) - std::max(own_lo, hull_lo));
                out += own * (world - 1);
                in += (hull_hi - hull_lo) - own;
            }
            worst = std::max(worst, std::max(out, in));
        }
        return worst + 1e-3 * total;
    };
    std::vector<int32_t> cuts(static_cast<size_t>(world) + 1, 0);
    cuts[static_cast<size_t>(world)] = n;
    {
        int32_t prev_hi = 0;
        for (int r = 0; r < world; ++r) {   // start: half way through the overlap (or gap) of neighbouring wedges
            const bool empty = wedges[r].first >= wedges[r].second;
            const int32_t lo = empty ? prev_hi : wedges[r].first, hi = empty ? prev_hi : wedges[r].second;
            if (r > 0) cuts[static_cast<size_t>(r)] = std::min(n, std::max(cuts[static_cast<size_t>(r) - 1], (lo + prev_hi) / 2));
            prev_hi = std::max(prev_hi, hi);
        }
    }
    double best = cost(cuts);
    for (int sweep = 0; sweep < 16; ++sweep) {
        bool moved = false;
        for (int b = 1; b < world; ++b) {
            const int32_t keep = cuts[static_cast<size_t>(b)];
            int32_t arg = keep;
            for (int32_t c = cuts[static_cast<size_t>(b) - 1]; c <= cuts[static_cast<size_t>(b) + 1]; ++c) {
                cuts[static_cast<size_t>(b)] = c;
                const double v = cost(cuts);
                if (v < best - 1e-9) { best = v; arg = c; }
            }
            cuts[static_cast<size_t>(b)] = arg;
            moved = moved || arg != keep;
        }
        if (!moved) break;
    }
    return cuts;
}

// Owners of the slabs.  Rank o adds, for every slab it owns, the partial sums of every OTHER rank whose wedge contains the
// slab; what a rank renders outside its own range it hands out.  The exchange is a set of concurrent point-to-point
// transfers over NVSwitch, so its duration follows the busiest port: the cuts minimise max over ranks of max(bytes out,
// bytes in) -- in the replicated mode including the second phase, in which every owner's sums go to all other ranks (that
// favours equal shares; the owned mode favours cuts through the middle of the wedge overlaps).  Coordinate descent over the
// world - 1 cuts from the mid-overlap start; every rank runs the same deterministic search on the same inputs.
static hp_status band_assign_owners(hpx_shard* s) {
    const int world = s->comm->world, me = s->comm->rank;
    const int32_t n = s->n_slabs;
    DV_CUDA(cudaStreamSynchronize(s->comm->ctx->stream));   // nothing in flight uses the old transfer lists / staging
    cudaFree(s->staging);
    s->staging = nullptr;
    s->sends.clear();
    s->recvs.clear();
    s->cuts = choose_owner_cuts(s->wedges, n, s->hull_lo, s->hull_hi, s->result == HPX_SHARD_RESULT_REPLICATED);
    auto overlap = [&](int r, int o) {   // slabs of rank r's wedge that rank o owns
        return std::pair<int32_t, int32_t>(std::max(s->wedges[r].first, s->cuts[o]), std::min(s->wedges[r].second, s->cuts[o + 1]));
    };
    size_t staging_floats = 0;
    for (int o = 0; o < world; ++o) {
        if (o == me) continue;
        const auto out = overlap(me, o);
        if (out.first < out.second) s->sends.push_back(hpx_shard::Xfer{o, out.first, out.second, 0});
        const auto in = overlap(o, me);
        if (in.first < in.second) {
            s->recvs.push_back(hpx_shard::Xfer{o, in.first, in.second, staging_floats});
            staging_floats += static_cast<size_t>(in.second - in.first) * s->slab_floats;
        }
    }
    if (!s->direct && staging_floats != 0 && cudaMalloc(&s->staging, staging_floats * sizeof(float)) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "cudaMalloc(shard staging)");
    return HP_STATUS_SUCCESS;
}

// (Re)builds everything that follows from the band cuts: this rank's frame, every rank's wedge, the owner cuts and the
// transfer lists.  Every rank derives all of it from the same inputs with the same code: the ranks agree without talking.
static hp_status band_configure(hpx_shard* s) {
    hpx_comm* c = s->comm;
    hpx_grid* g = s->grid;
    const hp_plan_desc& d = s->full_desc;
    const int world = c->world, me = c->rank;
    cudaStream_t main = c->ctx->stream;
    DV_CUDA(cudaStreamSynchronize(main));
    hpx_frame_release(s->frame);
    hp_plan_release(s->plan);
    s->frame = nullptr;
    s->plan = nullptr;
    s->band_row0.clear(); s->band_rows.clear(); s->wedges.clear();
    {
        const int32_t dims[3] = {g->nx, g->ny, g->nz};
        const int row_axis = s->slow_axis == 2 ? 1 : 2, fast_axis = s->slow_axis == 0 ? 1 : 0;
        s->rows_per_slab = dims[row_axis];
        s->row_floats = static_cast<size_t>(dims[fast_axis]) * 4;
        s->slab_rows.assign(static_cast<size_t>(world), std::vector<int2>(static_cast<size_t>(s->n_slabs), make_int2(0, 0)));
    }
    s->timed = false;
    const std::vector<RowBand> bands = partition_units(s->unit_cost, d.roi.height, static_cast<uint32_t>(world));
    hp_status st = HP_STATUS_SUCCESS;
    for (int r = 0; r < world; ++r) {
        s->band_row0.push_back(bands[r].y0);
        s->band_rows.push_back(bands[r].rows);
        std::pair<int32_t, int32_t> wedge(0, 0);
        if (bands[r].rows != 0) {
            hp_plan_desc bd = d;
            bd.roi.y = d.roi.y + bands[r].y0;
            bd.roi.height = bands[r].rows;
            bd.max_rays = 0;
            bd.max_samples = 0;
            hp_plan* band_plan = nullptr;
            hpx_frame* frame = nullptr;
            int32_t box[6] = {0, 0, 0, 0, 0, 0};
            st = hp_plan_create(c->ctx, &bd, &band_plan);
            if (st == HP_STATUS_SUCCESS) st = hpx_frame_create(band_plan, &frame);
            if (st == HP_STATUS_SUCCESS) st = hpx_frame_bounds(frame, g, box);
            if (st == HP_STATUS_SUCCESS) {
                std::vector<int32_t> lo(static_cast<size_t>(s->n_slabs)), hi(static_cast<size_t>(s->n_slabs));
                st = frame_slab_rows(frame, g, lo.data(), hi.data());
                for (int32_t y = 0; y < s->n_slabs && st == HP_STATUS_SUCCESS; ++y)
                    s->slab_rows[static_cast<size_t>(r)][static_cast<size_t>(y)] = make_int2(lo[static_cast<size_t>(y)], hi[static_cast<size_t>(y)]);
            }
            if (st == HP_STATUS_SUCCESS && r == me) {
                // stratified jitter hashes the ray's index in the WHOLE frame (reference samp_cpu.cpp:28-35)
                st = hpx_frame_set_view(frame, nullptr, d.seed, static_cast<uint64_t>(bands[r].y0) * d.roi.width);
                // end the launches with the band's CHEAP tiles (short tail): rows last-to-first when the rays get longer
                // downwards, columns centre-out when the rows cost the same and the edges of the image are cheap
                if (!s->order_tuned)   // (a measured choice survives the small band moves of hpx_shard_rebalance)
                    s->tile_order = best_tile_order(s->full_params, d, bands[r].y0, bands[r].rows, std::max(1u, c->ctx->usable_sms) * 5u);
                if (const char* env = std::getenv("DVREN_SHARD_TILE_ORDER")) {   // A/B timing: "rows" = round-2 behaviour before the column order
                    if (std::strcmp(env, "rows") == 0) {
                        const double first = row_work(s->full_params, d.roi.y + bands[r].y0, d.roi.x, d.roi.width);
                        const double last = row_work(s->full_params, d.roi.y + bands[r].y0 + bands[r].rows - 1, d.roi.x, d.roi.width);
                        s->tile_order = last > first ? 1 : 0;
                    } else if (std::strcmp(env, "columns") == 0) {
                        s->tile_order = static_cast<int>(HPX_ORDER_COLUMNS);
                    }
                }
                if (st == HP_STATUS_SUCCESS) st = hpx_frame_set_row_order(frame, s->tile_order);
                s->plan = band_plan;
                s->frame = frame;
                s->dl_offset_floats = static_cast<size_t>(bands[r].y0) * d.roi.width * 3;
            } else {
                hpx_frame_release(frame);
                hp_plan_release(band_plan);
            }
            if (st != HP_STATUS_SUCCESS) return st;
            if (box[3 + s->slow_axis] > 0) wedge = {box[s->slow_axis], box[s->slow_axis] + box[3 + s->slow_axis]};
        }
        s->wedges.push_back(wedge);
    }
    s->hull_lo = s->n_slabs;
    s->hull_hi = 0;
    for (int r = 0; r < world; ++r) {
        if (s->wedges[r].first >= s->wedges[r].second) continue;
        s->hull_lo = std::min(s->hull_lo, s->wedges[r].first);
        s->hull_hi = std::max(s->hull_hi, s->wedges[r].second);
    }
    if (s->hull_hi < s->hull_lo) s->hull_lo = s->hull_hi = 0;
    {
        const size_t per_rank = static_cast<size_t>(s->n_slabs);
        if (s->d_slab_rows == nullptr) DV_CUDA(cudaMalloc(&s->d_slab_rows, per_rank * world * sizeof(int2)));
        for (int r = 0; r < world; ++r)
            DV_CUDA(cudaMemcpyAsync(s->d_slab_rows + per_rank * r, s->slab_rows[static_cast<size_t>(r)].data(), per_rank * sizeof(int2),
                                    cudaMemcpyHostToDevice, main));
        DV_CUDA(cudaStreamSynchronize(main));
    }
    return band_assign_owners(s);
}


// ---- one frame over all ranks, contiguous bands -------------------------------------------------------------------
// Rank r renders a contiguous band of image rows; the bands are cut so that every rank has the same marching work
// (balanced_bands).  A band's rays stay inside a wedge of the volume, so with the gradient block laid out slab by slab
// along the world axis the image rows advance along, rank r's backward touches ONE contiguous slab range (its wedge) --
// about 2 / world of the grid -- and the exchange is sparse: every slab has an owner (cuts between neighbouring
// wedges); a rank sends the parts of its wedge it does not own to their owners (point-to-point, a few hundred MB at
// 512^3 / 8 GPUs instead of a 2.1 GB all-reduce), the owners add what they receive in rank order (deterministic).
// HPX_SHARD_RESULT_OWNED stops there: rank r holds the finished sum of slabs [cuts[r], cuts[r+1]) -- the hand-over for a
// slab-sharded optimiser.  HPX_SHARD_RESULT_REPLICATED then broadcasts every owned range, so that all ranks hold the
// whole summed gradient (what an all-reduce leaves behind).
HP_API hp_status hpx_shard_create_bands(hpx_comm* c, const hp_plan* full_plan, hpx_grid* g, uint32_t result,
                                        hpx_shard** out_shard) {
    DV_RANGE("hpx_shard_create_bands");
    if (c == nullptr || full_plan == nullptr || g == nullptr || out_shard == nullptr ||
        (result != HPX_SHARD_RESULT_OWNED && result != HPX_SHARD_RESULT_REPLICATED))
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
    s->bands = true;
    s->result = static_cast<int>(result);
    auto fail = [&](hp_status st) {
        hpx_shard_release(s);
        return st;
    };
    const hp_plan_desc& d = full_plan->desc;
    const float down[3] = {std::fabs(d.camera.c2w[1]), std::fabs(d.camera.c2w[5]), std::fabs(d.camera.c2w[9])};
    s->slow_axis = down[0] > down[1] ? (down[0] > down[2] ? 0 : 2) : (down[1] >= down[2] ? 1 : 2);
    hp_status st = hpx_grid_set_grad_layout(g, s->slow_axis, &s->slab_floats, &s->n_slabs);
    if (st != HP_STATUS_SUCCESS) return fail(st);
    if (cudaEventCreateWithFlags(&s->ev_zero, cudaEventDisableTiming) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "event"));

    s->full_desc = d;
    s->full_params = frame_params_from_plan(*full_plan);
    s->unit_cost = unit_costs(s->full_params, d);
    if (cudaEventCreate(&s->ev_t0) != cudaSuccess || cudaEventCreate(&s->ev_t1) != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "event"));
    st = map_peer_blocks(s);
    if (st == HP_STATUS_SUCCESS) st = band_configure(s);
    if (st != HP_STATUS_SUCCESS) return fail(st);
    *out_shard = s;
    return HP_STATUS_SUCCESS;
}

// Collective.  Re-cuts the bands from MEASURED time: every rank contributes the GPU time of its last step's forward +
// backward; a band that took longer than its share has the estimated work of its rows scaled up by that ratio (and vice
// versa), and the rows are partitioned again.  The estimate (in-cube steps per ray) does not know that tile rows differ
// in cache behaviour and idle lanes; two or three rounds during warm-up bring the ranks within a few percent.
// out_changed: 1 when the bands moved (frames, wedges and owner cuts were rebuilt; hpx_shard_frame / hpx_shard_owned /
// hpx_shard_bands must be queried again).
HP_API hp_status hpx_shard_rebalance(hpx_shard* s, int32_t* out_changed) {
    DV_RANGE("hpx_shard_rebalance");
    if (s == nullptr || !s->bands) return HP_STATUS_INVALID_ARGUMENT;
    if (out_changed) *out_changed = 0;
    hpx_comm* c = s->comm;
    if (c->world == 1 || !s->timed) return HP_STATUS_SUCCESS;
    DV_ENTER(c->ctx);
    cudaStream_t main = c->ctx->stream;
    DV_CUDA(cudaStreamSynchronize(main));
    float ms = 0.0f;
    if (s->frame != nullptr) DV_CUDA(cudaEventElapsedTime(&ms, s->ev_t0, s->ev_t1));
    const int world = c->world;
    float* d_ms = nullptr;
    DV_CUDA(cudaMalloc(&d_ms, sizeof(float) * world));
    std::vector<float> all(static_cast<size_t>(world), 0.0f);
    hp_status st = HP_STATUS_SUCCESS;
    if (cudaMemcpyAsync(d_ms + c->rank, &ms, sizeof(float), cudaMemcpyHostToDevice, main) != cudaSuccess) st = cuda_fail(cudaGetLastError(), "rebalance");
    if (st == HP_STATUS_SUCCESS) {
        const ncclResult_t r = nccl().AllGather(d_ms + c->rank, d_ms, 1, ncclFloat32, c->comm, main);
        if (r != ncclSuccess) st = nccl_fail(r, "ncclAllGather(step times)");
    }
    if (st == HP_STATUS_SUCCESS && (cudaMemcpyAsync(all.data(), d_ms, sizeof(float) * world, cudaMemcpyDeviceToHost, main) != cudaSuccess ||
                                    cudaStreamSynchronize(main) != cudaSuccess))
        st = cuda_fail(cudaGetLastError(), "rebalance");
    cudaFree(d_ms);
    if (st != HP_STATUS_SUCCESS) return st;
    const uint32_t unit = kTileH * kWarpsY;
    double t_sum = 0.0, w_sum = 0.0;
    std::vector<double> w(static_cast<size_t>(world), 0.0);
    for (int r = 0; r < world; ++r) {
        for (uint32_t u = s->band_row0[r] / unit; u < (s->band_row0[r] + s->band_rows[r] + unit - 1) / unit && u < s->unit_cost.size(); ++u) w[r] += s->unit_cost[u];
        if (s->band_rows[r] != 0) { t_sum += all[r]; w_sum += w[r]; }
    }
    if (!(t_sum > 0.0) || !(w_sum > 0.0)) return HP_STATUS_SUCCESS;
    for (int r = 0; r < world; ++r) {
        if (s->band_rows[r] == 0 || !(w[r] > 0.0) || !(all[r] > 0.0f)) continue;
        const double ratio = std::min(1.25, std::max(0.8, (all[r] / t_sum) / (w[r] / w_sum)));   // measured share / estimated share
        for (uint32_t u = s->band_row0[r] / unit; u < (s->band_row0[r] + s->band_rows[r] + unit - 1) / unit && u < s->unit_cost.size(); ++u) s->unit_cost[u] *= ratio;
    }
    const std::vector<RowBand> bands = partition_units(s->unit_cost, s->full_desc.roi.height, static_cast<uint32_t>(world));
    bool same = true;
    for (int r = 0; r < world; ++r) same = same && bands[r].y0 == s->band_row0[r] && bands[r].rows == s->band_rows[r];
    if (same) return HP_STATUS_SUCCESS;
    DV_TRY(band_configure(s));
    if (out_changed) *out_changed = 1;
    return HP_STATUS_SUCCESS;
}

// Host-only (works without a GPU): the owner cuts hpx_shard_* chooses for the given per-rank wedges.
HP_API hp_status hpx_plan_owner_cuts(uint32_t world, int32_t n_slabs, const int32_t* wedges, uint32_t result, int32_t* out_cuts) {
    DV_RANGE("hpx_plan_owner_cuts");
    if (world == 0 || world > 4096 || n_slabs <= 0 || wedges == nullptr || out_cuts == nullptr ||
        (result != HPX_SHARD_RESULT_OWNED && result != HPX_SHARD_RESULT_REPLICATED))
        return HP_STATUS_INVALID_ARGUMENT;
    std::vector<std::pair<int32_t, int32_t>> w;
    int32_t hull_lo = n_slabs, hull_hi = 0;
    for (uint32_t r = 0; r < world; ++r) {
        const int32_t lo = wedges[2 * r], hi = wedges[2 * r + 1];
        if (lo < 0 || hi > n_slabs) return HP_STATUS_INVALID_ARGUMENT;
        w.emplace_back(lo, hi);
        if (lo < hi) { hull_lo = std::min(hull_lo, lo); hull_hi = std::max(hull_hi, hi); }
    }
    if (hull_hi < hull_lo) hull_lo = hull_hi = 0;
    const std::vector<int32_t> cuts = choose_owner_cuts(w, n_slabs, hull_lo, hull_hi, result == HPX_SHARD_RESULT_REPLICATED);
    std::copy(cuts.begin(), cuts.end(), out_cuts);
    return HP_STATUS_SUCCESS;
}

// Host-only (works without a GPU): the bands hpx_shard_create_bands cuts for `world` ranks.
HP_API hp_status hpx_plan_balanced_bands(const hp_plan* plan, uint32_t world, uint32_t* out_row0, uint32_t* out_rows, double* out_work) {
    DV_RANGE("hpx_plan_balanced_bands");
    if (plan == nullptr || world == 0 || world > 4096 || out_row0 == nullptr || out_rows == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    const FrameParams p = frame_params_from_plan(*plan);
    const std::vector<RowBand> bands = balanced_bands(p, plan->desc, world);
    for (uint32_t r = 0; r < world; ++r) {
        out_row0[r] = bands[r].y0;
        out_rows[r] = bands[r].rows;
        if (out_work != nullptr) {
            double w = 0.0;
            for (uint32_t y = bands[r].y0; y < bands[r].y0 + bands[r].rows; ++y) w += row_work(p, plan->desc.roi.y + y, plan->desc.roi.x, plan->desc.roi.width);
            out_work[r] = w;
        }
    }
    return HP_STATUS_SUCCESS;
}

// Host-only (works without a GPU): best_tile_order for a band of the plan's ROI; out_ends4 (optional): the estimated end of
// a launch, in marching steps, under the orders {0, 1, HPX_ORDER_COLUMNS, HPX_ORDER_COLUMNS | 1}.
HP_API hp_status hpx_plan_best_tile_order(const hp_plan* plan, uint32_t row0, uint32_t rows, uint32_t slots, int32_t* out_order,
                                          double* out_ends4) {
    DV_RANGE("hpx_plan_best_tile_order");
    if (plan == nullptr || out_order == nullptr || rows == 0 || slots == 0 || slots > (1u << 20) ||
        static_cast<uint64_t>(row0) + rows > plan->desc.roi.height)
        return HP_STATUS_INVALID_ARGUMENT;
    *out_order = best_tile_order(frame_params_from_plan(*plan), plan->desc, row0, rows, slots, out_ends4);
    return HP_STATUS_SUCCESS;
}

// Rank-local (no collective): renders this rank's band -- forward + backward, no exchange -- under every candidate dispatch
// order and keeps the fastest.  The estimate behind hpx_shard_create_bands models a CTA as lasting as long as its longest
// ray; it knows nothing about the memory system (column order walks the slab axis with a stride; measured 0-2 % slower
// than its estimate on bands of 50-130 tile rows), so two orders that tie in the estimate are told apart here.  The
// backward passes ACCUMULATE into the gradient block: call it during warm-up and clear the block (HPX_BACKWARD_ZERO) on
// the next step.  The choice survives hpx_shard_rebalance.
HP_API hp_status hpx_shard_tune_order(hpx_shard* s, const float* dL_dI_device, uint32_t flags, int32_t* out_order) {
    DV_RANGE("hpx_shard_tune_order");
    if (s == nullptr || !s->bands || dL_dI_device == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (out_order != nullptr) *out_order = s->tile_order;
    if (s->frame == nullptr || std::getenv("DVREN_SHARD_TILE_ORDER") != nullptr) return HP_STATUS_SUCCESS;
    hpx_comm* c = s->comm;
    DV_ENTER(c->ctx);
    cudaStream_t main = c->ctx->stream;
    const int current = s->tile_order;
    int best = current;
    float best_ms = 0.0f;
    hp_status st = HP_STATUS_SUCCESS;
    for (int order : {current, 0, 1, static_cast<int>(HPX_ORDER_COLUMNS), static_cast<int>(HPX_ORDER_COLUMNS) | 1}) {
        if (order == current && best_ms > 0.0f) continue;   // (measured first)
        st = hpx_frame_set_row_order(s->frame, order);
        for (int rep = 0; rep < 3 && st == HP_STATUS_SUCCESS; ++rep) {   // one untimed, two timed
            if (rep == 1 && cudaEventRecord(s->ev_t0, main) != cudaSuccess) st = cuda_fail(cudaGetLastError(), "tune");
            if (st == HP_STATUS_SUCCESS) st = hpx_forward(s->frame, s->grid);
            if (st == HP_STATUS_SUCCESS)
                st = hpx_backward(s->frame, s->grid, dL_dI_device + s->dl_offset_floats, HP_MEMSPACE_DEVICE, flags & ~HPX_BACKWARD_ZERO);
        }
        float ms = 0.0f;
        if (st == HP_STATUS_SUCCESS && (cudaEventRecord(s->ev_t1, main) != cudaSuccess || cudaStreamSynchronize(main) != cudaSuccess ||
                                        cudaEventElapsedTime(&ms, s->ev_t0, s->ev_t1) != cudaSuccess))
            st = cuda_fail(cudaGetLastError(), "tune");
        if (st != HP_STATUS_SUCCESS) break;
        if (best_ms == 0.0f || ms < best_ms * 0.995f) {   // the order in place stays unless another is measurably faster
            best = order;
            best_ms = ms;
        }
    }
    const hp_status restored = hpx_frame_set_row_order(s->frame, st == HP_STATUS_SUCCESS ? best : current);
    if (st != HP_STATUS_SUCCESS) return st;
    DV_TRY(restored);
    s->tile_order = best;
    s->order_tuned = true;
    s->timed = false;   // ev_t0 / ev_t1 no longer bracket a step
    if (out_order != nullptr) *out_order = best;
    return HP_STATUS_SUCCESS;
}

// The dispatch order (hpx_frame_set_row_order) the library chose for this rank's band.
HP_API hp_status hpx_shard_tile_order(const hpx_shard* s, int32_t* out_order) {
    DV_RANGE("hpx_shard_tile_order");
    if (s == nullptr || out_order == nullptr || !s->bands) return HP_STATUS_INVALID_ARGUMENT;
    *out_order = s->tile_order;
    return HP_STATUS_SUCCESS;
}

// 1: the exchange runs as this library's own kernels over mapped peer memory; 0: NCCL point-to-point + broadcast.
HP_API hp_status hpx_shard_exchange_is_direct(const hpx_shard* s, int32_t* out_direct) {
    DV_RANGE("hpx_shard_exchange_is_direct");
    if (s == nullptr || out_direct == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    *out_direct = s->direct ? 1 : 0;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_shard_set_result(hpx_shard* s, uint32_t result) {
    DV_RANGE("hpx_shard_set_result");
    if (s == nullptr || !s->bands || (result != HPX_SHARD_RESULT_OWNED && result != HPX_SHARD_RESULT_REPLICATED))
        return HP_STATUS_INVALID_ARGUMENT;
    if (s->result == static_cast<int>(result)) return HP_STATUS_SUCCESS;
    s->result = static_cast<int>(result);
    DV_ENTER(s->comm->ctx);
    return band_assign_owners(s);   // the best owner cuts differ between the two result modes
}

HP_API hp_status hpx_shard_bands(const hpx_shard* s, uint32_t* out_row0, uint32_t* out_rows, int32_t* out_wedges, int32_t* out_cuts,
                                 size_t* out_send_floats, size_t* out_recv_floats) {
    DV_RANGE("hpx_shard_bands");
    if (s == nullptr || !s->bands) return HP_STATUS_INVALID_ARGUMENT;
    for (size_t r = 0; r < s->band_rows.size(); ++r) {
        if (out_row0) out_row0[r] = s->band_row0[r];
        if (out_rows) out_rows[r] = s->band_rows[r];
        if (out_wedges) {
            out_wedges[2 * r] = s->wedges[r].first;
            out_wedges[2 * r + 1] = s->wedges[r].second;
        }
    }
    if (out_cuts) std::copy(s->cuts.begin(), s->cuts.end(), out_cuts);
    size_t out = 0, in = 0;
    auto volume = [&](int rank, int32_t lo, int32_t hi) {   // floats of rank's partial sums inside slabs [lo, hi) that travel
        if (!s->direct) return static_cast<size_t>(hi - lo) * s->slab_floats;
        size_t rows = 0;
        for (int32_t y = lo; y < hi; ++y) {
            const int2 r = s->slab_rows[static_cast<size_t>(rank)][static_cast<size_t>(y)];
            rows += r.y > r.x ? static_cast<size_t>(r.y - r.x) : 0;
        }
        return rows * s->row_floats;
    };
    for (const auto& x : s->sends) out += volume(s->comm->rank, x.lo, x.hi);
    for (const auto& x : s->recvs) in += volume(x.peer, x.lo, x.hi);
    if (out_send_floats) *out_send_floats = out;
    if (out_recv_floats) *out_recv_floats = in;
    return HP_STATUS_SUCCESS;
}

// The slabs this rank owns, as they lie in the gradient block (hpx_grid_set_grad_layout order, {dr,dg,db,dsigma} per
// voxel): [first_slab, first_slab + slabs) x slab_floats floats at *out_device_ptr.
HP_API hp_status hpx_shard_owned(const hpx_shard* s, float** out_device_ptr, int32_t* out_first_slab, int32_t* out_slabs,
                                 size_t* out_slab_floats, int32_t* out_slow_axis) {
    DV_RANGE("hpx_shard_owned");
    if (s == nullptr || !s->bands) return HP_STATUS_INVALID_ARGUMENT;
    float* block = nullptr;
    size_t floats = 0;
    DV_TRY(hpx_grid_grad_buffer(s->grid, &block, &floats));
    const int32_t lo = s->cuts[static_cast<size_t>(s->comm->rank)], hi = s->cuts[static_cast<size_t>(s->comm->rank) + 1];
    if (out_device_ptr) *out_device_ptr = block + static_cast<size_t>(lo) * s->slab_floats;
    if (out_first_slab) *out_first_slab = lo;
    if (out_slabs) *out_slabs = hi - lo;
    if (out_slab_floats) *out_slab_floats = s->slab_floats;
    if (out_slow_axis) *out_slow_axis = s->slow_axis;
    return HP_STATUS_SUCCESS;
}

static hp_status band_step(hpx_shard* s, const float* dL_dI_device, uint32_t flags) {
    hpx_comm* c = s->comm;
    cudaStream_t main = c->ctx->stream, side = c->side;
    const int me = c->rank;
    float* block = nullptr;
    size_t floats = 0;
    DV_TRY(hpx_grid_grad_buffer(s->grid, &block, &floats));
    const size_t S = s->slab_floats;
    if (flags & HPX_BACKWARD_ZERO) {
        // only what this rank can write or must hand out as a finished sum: its wedge and the slabs it owns; cleared on the
        // side stream while the forward kernel runs
        const int32_t lo = std::min(s->wedges[me].first < s->wedges[me].second ? s->wedges[me].first : s->cuts[me], s->cuts[me]);
        const int32_t hi = std::max(s->wedges[me].first < s->wedges[me].second ? s->wedges[me].second : s->cuts[me + 1], s->cuts[me + 1]);
        DV_CUDA(cudaEventRecord(c->ev_main, main));
        DV_CUDA(cudaStreamWaitEvent(side, c->ev_main, 0));
        if (hi > lo) DV_CUDA(cudaMemsetAsync(block + static_cast<size_t>(lo) * S, 0, static_cast<size_t>(hi - lo) * S * sizeof(float), side));
        DV_CUDA(cudaMemsetAsync(block + floats - 16, 0, 16 * sizeof(float), side));
        DV_CUDA(cudaEventRecord(s->ev_zero, side));
    }
    DV_CUDA(cudaEventRecord(s->ev_t0, main));
    if (s->frame != nullptr) DV_TRY(hpx_forward(s->frame, s->grid));
    if (flags & HPX_BACKWARD_ZERO) DV_CUDA(cudaStreamWaitEvent(main, s->ev_zero, 0));
    if (s->frame != nullptr)
        DV_TRY(hpx_backward(s->frame, s->grid, dL_dI_device + s->dl_offset_floats, HP_MEMSPACE_DEVICE, flags & ~HPX_BACKWARD_ZERO));
    DV_CUDA(cudaEventRecord(s->ev_t1, main));
    s->timed = true;
    if (c->world == 1 || !s->reduce) return HP_STATUS_SUCCESS;
    if (s->direct) {
        // ---- own kernels over peer memory (NVLink): every owner PULLS the wedge parts of its slabs out of its neighbours'
        // gradient blocks and adds them in rank order -- no staging copy, no separate add pass.
        DV_TRY(stream_barrier(s));   // every rank's backward has finished: the blocks hold the final partial sums
        const size_t slab4 = S / 4;
        const unsigned bx = static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>((slab4 + 1023) / 1024, 64)));
        if (!s->recvs.empty() && s->cuts[me + 1] > s->cuts[me]) {
            PullTable t{};
            t.n = 0;
            for (const auto& x : s->recvs) {   // ascending peer order = order of the additions
                if (t.n >= kMaxPeers) break;
                t.src[t.n] = reinterpret_cast<const float4*>(s->peer_block[static_cast<size_t>(x.peer)]);
                t.rows[t.n] = s->d_slab_rows + static_cast<size_t>(s->n_slabs) * static_cast<size_t>(x.peer);
                t.lo[t.n] = x.lo;
                t.hi[t.n] = x.hi;
                ++t.n;
            }
            t.first_slab = s->cuts[me];
            t.row4 = static_cast<uint32_t>(s->row_floats / 4);
            peer_reduce_kernel<<<dim3(bx, static_cast<unsigned>(s->cuts[me + 1] - s->cuts[me])), 256, 0, main>>>(
                reinterpret_cast<float4*>(block), t, slab4);
            DV_CUDA(cudaGetLastError());
        }
        if (flags & HPX_BACKWARD_CAMERA)
            DV_NCCL(nccl().AllReduce(block + floats - 16, block + floats - 16, 16, ncclFloat32, ncclSum, c->comm, main));
        if (s->result == HPX_SHARD_RESULT_REPLICATED && s->hull_hi > s->hull_lo) {
            DV_TRY(stream_barrier(s));   // every owner has finished its sums
            OwnerTable t{};
            for (int r = 0; r < c->world; ++r) t.src[r] = reinterpret_cast<const float4*>(s->peer_block[static_cast<size_t>(r)]);
            for (int r = 0; r <= c->world; ++r) t.cuts[r] = s->cuts[static_cast<size_t>(r)];
            t.world = c->world;
            t.me = me;
            t.first_slab = s->hull_lo;
            t.rotate = std::max(0, std::min(s->cuts[me + 1], s->hull_hi) - s->hull_lo);
            peer_gather_kernel<<<dim3(bx, static_cast<unsigned>(s->hull_hi - s->hull_lo)), 256, 0, main>>>(reinterpret_cast<float4*>(block), t, slab4);
            DV_CUDA(cudaGetLastError());
        }
        return stream_barrier(s);        // nobody still reads this rank's block when it is cleared for the next step
    }
    // ---- NCCL fallback (DVREN_SHARD_EXCHANGE=nccl, or a peer's block cannot be mapped): sparse reduce-scatter with
    // point-to-point sends into a staging buffer
    if (!s->sends.empty() || !s->recvs.empty()) {
        DV_NCCL(nccl().GroupStart());
        ncclResult_t r = ncclSuccess;
        for (const auto& x : s->sends)
            if (r == ncclSuccess)
                r = nccl().Send(block + static_cast<size_t>(x.lo) * S, static_cast<size_t>(x.hi - x.lo) * S, ncclFloat32, x.peer, c->comm, main);
        for (const auto& x : s->recvs)
            if (r == ncclSuccess)
                r = nccl().Recv(s->staging + x.staging_off, static_cast<size_t>(x.hi - x.lo) * S, ncclFloat32, x.peer, c->comm, main);
        const ncclResult_t e = nccl().GroupEnd();
        if (r != ncclSuccess) return nccl_fail(r, "ncclSend/ncclRecv(slabs)");
        if (e != ncclSuccess) return nccl_fail(e, "ncclGroupEnd(slabs)");
        for (const auto& x : s->recvs) {   // ascending peer order: the sum does not depend on arrival order
            const size_t n4 = static_cast<size_t>(x.hi - x.lo) * S / 4;
            add_slabs_kernel<<<static_cast<unsigned>(std::min<size_t>((n4 + 255) / 256, 148 * 8)), 256, 0, main>>>(
                reinterpret_cast<float4*>(block + static_cast<size_t>(x.lo) * S), reinterpret_cast<const float4*>(s->staging + x.staging_off), n4);
        }
        DV_CUDA(cudaGetLastError());
    }
    if (flags & HPX_BACKWARD_CAMERA) DV_NCCL(nccl().AllReduce(block + floats - 16, block + floats - 16, 16, ncclFloat32, ncclSum, c->comm, main));
    if (s->result != HPX_SHARD_RESULT_REPLICATED) return HP_STATUS_SUCCESS;
    // ---- all-gather of the finished sums: every owner broadcasts its slabs inside the touched hull
    DV_NCCL(nccl().GroupStart());
    ncclResult_t r = ncclSuccess;
    for (int o = 0; o < c->world && r == ncclSuccess; ++o) {
        const int32_t lo = std::max(s->cuts[o], s->hull_lo), hi = std::min(s->cuts[o + 1], s->hull_hi);
        if (lo >= hi) continue;
        float* p = block + static_cast<size_t>(lo) * S;
        r = nccl().Broadcast(p, p, static_cast<size_t>(hi - lo) * S, ncclFloat32, o, c->comm, main);
    }
    const ncclResult_t e = nccl().GroupEnd();
    if (r != ncclSuccess) return nccl_fail(r, "ncclBroadcast(slabs)");
    if (e != ncclSuccess) return nccl_fail(e, "ncclGroupEnd(broadcast)");
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
        cudaFree(s->staging);
        cudaFree(s->d_flag);
        cudaFree(s->d_slab_rows);
        if (s->ev_t0 != nullptr) cudaEventDestroy(s->ev_t0);
        if (s->ev_t1 != nullptr) cudaEventDestroy(s->ev_t1);
        for (void* p : s->ipc_opened) cudaIpcCloseMemHandle(p);
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
    if (s->bands) return band_step(s, dL_dI_device, flags);
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

// ---- hpx_backward_streamed: the gradient read-back runs UNDER the backward kernel ----------------------------------
// A training loop on the host (the dvren::Renderer contract, reference renderer.cpp:441-442) reads the whole gradient
// every step: 2.1 GB at 512^3, as long over PCIe as the backward kernel itself takes.  The backward is therefore launched
// with per-row-group completion signals (hpx_backward_signalled) on a gradient block laid out slab by slab along the
// world axis the image rows advance along; a high-priority copy stream waits for each group (cuStreamWaitValue32: no SM
// is occupied by the wait), un-interleaves the slabs that group finished and no later group will touch, and copies them
// into the caller's arrays in the reference layout while the later rows still render.
namespace {
struct StreamPlan {
    const hpx_grid* grid = nullptr;
    CameraParams cam{};
    RoiParams roi{};
    int slow_axis = 2;
    std::vector<uint32_t> group_end_rows;
    std::vector<std::vector<std::pair<int32_t, int32_t>>> runs;
    std::vector<std::pair<int32_t, int32_t>> untouched;
    cudaStream_t side = nullptr, side2 = nullptr;   // side2: second copy engine for the colour copies
    cudaEvent_t ev_main = nullptr, ev_side = nullptr, ev_side2 = nullptr;
    std::vector<cudaEvent_t> ev_unpack;              // one per slab run in flight (staging regions differ per run)
};

void stream_plan_free(void* p) {
    StreamPlan* sp = static_cast<StreamPlan*>(p);
    if (sp == nullptr) return;
    if (sp->side != nullptr) { cudaStreamSynchronize(sp->side); cudaStreamDestroy(sp->side); }
    if (sp->side2 != nullptr) { cudaStreamSynchronize(sp->side2); cudaStreamDestroy(sp->side2); }
    if (sp->ev_main != nullptr) cudaEventDestroy(sp->ev_main);
    if (sp->ev_side != nullptr) cudaEventDestroy(sp->ev_side);
    if (sp->ev_side2 != nullptr) cudaEventDestroy(sp->ev_side2);
    for (cudaEvent_t e : sp->ev_unpack) cudaEventDestroy(e);
    delete sp;
}

constexpr uint32_t kStreamGroups = 16;   // LeanBuffers::group_end

hp_status stream_plan_build(hpx_frame* f, hpx_grid* g, StreamPlan** out) {
    StreamPlan* sp = static_cast<StreamPlan*>(f->stream_plan);
    const CameraParams& cam = f->h_params.cam;
    const RoiParams& roi = f->h_params.roi;
    const float down[3] = {std::fabs(cam.r01), std::fabs(cam.r11), std::fabs(cam.r21)};
    const int axis = down[0] > down[1] ? (down[0] > down[2] ? 0 : 2) : (down[1] >= down[2] ? 1 : 2);
    if (sp != nullptr && sp->grid == g && std::memcmp(&sp->cam, &cam, sizeof(cam)) == 0 && std::memcmp(&sp->roi, &roi, sizeof(roi)) == 0 &&
        g->grad_slow_axis == sp->slow_axis) {
        *out = sp;
        return HP_STATUS_SUCCESS;
    }
    if (sp == nullptr) {
        sp = new (std::nothrow) StreamPlan();
        if (sp == nullptr) return HP_STATUS_OUT_OF_MEMORY;
        f->stream_plan = sp;
        f->stream_plan_free = stream_plan_free;
        int lo = 0, hi = 0;
        cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&sp->side, cudaStreamNonBlocking, hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&sp->side2, cudaStreamNonBlocking, hi);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sp->ev_main, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sp->ev_side, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sp->ev_side2, cudaEventDisableTiming);
        if (e != cudaSuccess) return cuda_fail(e, "stream plan");
    }
    sp->grid = nullptr;   // invalid until rebuilt
    sp->group_end_rows.clear();
    sp->runs.clear();
    sp->untouched.clear();
    sp->slow_axis = axis;
    if (g->grad_slow_axis != axis) DV_TRY(hpx_grid_set_grad_layout(g, axis, nullptr, nullptr));   // (clears the block: the caller asked for ZERO)
    const int32_t n_slabs = axis == 0 ? g->nx : axis == 1 ? g->ny : g->nz;
    // Row groups in DISPATCH order of the centre-out row order (RoiParams::tile_row_reverse = 2): group k is the k-th ring of
    // tile rows around the middle of the image -- one band below the centre and one above it.  With a perspective camera the
    // rows further out can only touch slabs further out (or the far ends of the inner ones), so the slabs around the centre
    // are final after the first group and the copy stream has work from the start.
    const uint32_t tile_rows_px = kTileH * kWarpsY;
    const uint32_t tile_rows = (roi.h + tile_rows_px - 1) / tile_rows_px;
    uint32_t want = kStreamGroups;   // (DVREN_STREAM_GROUPS = 1..16 overrides it: profiles/README.md has the sweep)
    if (const char* env = std::getenv("DVREN_STREAM_GROUPS")) want = static_cast<uint32_t>(std::max(1, std::min(16, std::atoi(env))));
    const uint32_t groups = std::min<uint32_t>(want, std::max<uint32_t>(1u, tile_rows));
    std::vector<SlabIntervals> ranges;
    for (uint32_t k = 0; k < groups; ++k) {
        const uint32_t i0 = static_cast<uint32_t>(static_cast<uint64_t>(tile_rows) * k / groups);
        const uint32_t i1 = static_cast<uint32_t>(static_cast<uint64_t>(tile_rows) * (k + 1) / groups);
        if (i1 <= i0) continue;
        // the dispatched rows [i0, i1) as (at most) two runs of tile rows: those below and those above the centre
        uint32_t lo[2] = {UINT32_MAX, UINT32_MAX}, hi[2] = {0, 0};
        const uint32_t c = tile_rows / 2u;
        for (uint32_t i = i0; i < i1; ++i) {
            const uint32_t r = tile_row_of(i, tile_rows, 2u);
            const int side_of = r >= c ? 0 : 1;
            lo[side_of] = std::min(lo[side_of], r);
            hi[side_of] = std::max(hi[side_of], r + 1u);
        }
        SlabIntervals set;
        for (int h = 0; h < 2; ++h) {
            if (lo[h] == UINT32_MAX) continue;
            int32_t box[6] = {0, 0, 0, 0, 0, 0};
            const uint32_t row0 = lo[h] * tile_rows_px, rows = std::min(roi.h, hi[h] * tile_rows_px) - row0;
            DV_TRY(frame_rows_bounds(f, g, row0, rows, box));
            if (box[3 + axis] > 0) set.emplace_back(box[axis], box[axis] + box[3 + axis]);
        }
        ranges.push_back(std::move(set));
        sp->group_end_rows.push_back(i1);
    }
    if (sp->group_end_rows.empty()) return HP_STATUS_INVALID_ARGUMENT;
    sp->runs = final_slab_runs(ranges);
    std::vector<char> touched(static_cast<size_t>(n_slabs), 0);
    for (const auto& set : ranges)
        for (const auto& r : set)
            for (int32_t y = std::max(r.first, 0); y < std::min(r.second, n_slabs); ++y) touched[static_cast<size_t>(y)] = 1;
    for (int32_t y = 0; y < n_slabs; ++y) {
        if (touched[static_cast<size_t>(y)]) continue;
        if (!sp->untouched.empty() && sp->untouched.back().second == y) sp->untouched.back().second = y + 1;
        else sp->untouched.emplace_back(y, y + 1);
    }
    sp->cam = cam;
    sp->roi = roi;
    sp->grid = g;
    *out = sp;
    return HP_STATUS_SUCCESS;
}
}  // namespace

extern "C" {

HP_API hp_status hpx_backward_streamed(hpx_frame* f, hpx_grid* g, const float* dL_dI, hp_memspace memspace, uint32_t flags,
                                       float* sigma_grad_host, float* color_grad_host, float* camera16_host) {
    DV_RANGE("hpx_backward_streamed");
    if (f == nullptr || g == nullptr || dL_dI == nullptr || f->ctx != g->ctx) return HP_STATUS_INVALID_ARGUMENT;
    if ((flags & HPX_BACKWARD_GRID) == 0u) return HP_STATUS_INVALID_ARGUMENT;
    const RoiParams& roi = f->h_params.roi;
    const float down0 = std::fabs(f->h_params.cam.r01), down1 = std::fabs(f->h_params.cam.r11), down2 = std::fabs(f->h_params.cam.r21);
    const bool x_slow = down0 > down1 && down0 > down2;
    const bool can_stream = g->linear && !g->clamp && scatter_params(*g).unit_bbox != 0u && (flags & HPX_BACKWARD_ZERO) != 0u &&
                            (flags & HPX_BACKWARD_DETERMINISTIC) == 0u && wait_value() != nullptr && !x_slow && roi.tile_row_stride <= 1 &&
                            roi.tile_row_reverse == 0u && (sigma_grad_host != nullptr || color_grad_host != nullptr);
    if (!can_stream) {   // same result, read back after the kernel
        DV_TRY(hpx_backward(f, g, dL_dI, memspace, flags));
        return hpx_grid_read_grad(g, sigma_grad_host, color_grad_host, camera16_host, HP_MEMSPACE_HOST);
    }
    DV_ENTER(f->ctx);
    StreamPlan* sp = nullptr;
    DV_TRY(stream_plan_build(f, g, &sp));
    cudaStream_t main = f->ctx->stream, side = sp->side;
    float* block = nullptr;
    size_t floats = 0;
    DV_TRY(hpx_grid_grad_buffer(g, &block, &floats));
    DV_CUDA(cudaMemsetAsync(block, 0, floats * sizeof(float), main));
    uint32_t* counters = nullptr;
    DV_TRY(hpx_frame_reset_group_counters(f, &counters));
    DV_CUDA(cudaEventRecord(sp->ev_main, main));   // counters cleared: the previous call's counts cannot satisfy the waits
    uint32_t expected[16] = {};
    const uint32_t n_groups = static_cast<uint32_t>(sp->group_end_rows.size());
    f->h_params.roi.tile_row_reverse = 2u;   // centre-out for this launch (stream_plan_build cut the groups for it)
    f->params_dirty = true;
    const hp_status launched = hpx_backward_signalled(f, g, dL_dI, memspace, flags & ~HPX_BACKWARD_ZERO, sp->group_end_rows.data(),
                                                      n_groups, &counters, expected);
    f->h_params.roi.tile_row_reverse = 0u;
    f->params_dirty = true;
    DV_TRY(launched);
    DV_CUDA(cudaStreamWaitEvent(side, sp->ev_main, 0));
    size_t run_index = 0;
    for (uint32_t i = 0; i < n_groups; ++i) {
        const int rc = wait_value()(side, reinterpret_cast<unsigned long long>(counters + i), expected[i], 0u /* GEQ */);
        if (rc != 0) {
            set_last_error("cuStreamWaitValue32 failed with driver error " + std::to_string(rc));
            return HP_STATUS_INTERNAL_ERROR;
        }
        auto emit = [&](const std::pair<int32_t, int32_t>& run) -> hp_status {
            if (run_index >= sp->ev_unpack.size()) {
                cudaEvent_t e = nullptr;
                DV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                sp->ev_unpack.push_back(e);
            }
            return grid_slabs_to_host(g, side, run.first, run.second, sigma_grad_host, color_grad_host, sp->side2, sp->ev_unpack[run_index++]);
        };
        if (i == 0)
            for (const auto& run : sp->untouched) DV_TRY(emit(run));
        for (const auto& run : sp->runs[i]) DV_TRY(emit(run));
    }
    if (camera16_host != nullptr)   // (the camera reduction kernel runs after the backward kernel on the main stream)
        DV_CUDA(cudaMemcpyAsync(camera16_host, block + floats - 16, 16 * sizeof(float), cudaMemcpyDeviceToHost, main));
    DV_CUDA(cudaEventRecord(sp->ev_side, side));
    DV_CUDA(cudaStreamWaitEvent(main, sp->ev_side, 0));
    DV_CUDA(cudaEventRecord(sp->ev_side2, sp->side2));
    DV_CUDA(cudaStreamWaitEvent(main, sp->ev_side2, 0));
    return HP_STATUS_SUCCESS;
}

}  // extern "C"
