// dv_lean.cu -- the non-materialising hot path.
//
//   lean_forward_kernel   ray generation + marching + emission-absorption
//                         integration + image composition, one warp per 8x4
//                         pixel tile, transmittance in registers, warp-vote
//                         early exit.  Replaces hp_ray -> hp_samp_int_fused ->
//                         hp_img (reference src/render/renderer.cpp:259-365).
//   lean_backward_kernel  reverse march over 8-sample segments: each segment is
//                         recomputed from a transmittance checkpoint written by
//                         the forward kernel, then swept in reverse with the
//                         reference's adjoint recurrence and scattered to the
//                         packed gradient grid.  Replaces hp_diff +
//                         DenseGridField::AccumulateSampleGradients (reference
//                         hotpath/src/cpu/diff_cpu.cpp:156-195,
//                         src/fields/dense_grid.cpp:198-306).
//   camera_adjoint_kernel forward-order evaluation of d/d c2w, d/d intrinsics
//                         (no reference counterpart: the reference returns
//                         zeros, diff_cpu.cpp:25,73-74).
//
// Compiled with -fmad=false: see dv_device.cuh.
#include "dv_lean.h"

#include <algorithm>
#include <climits>

#include "dv_device.cuh"

namespace dv {

namespace {

struct TilePixel {
    uint32_t lx, ly;   // pixel inside the ROI
    uint32_t ray;      // plan ray index = ly * roi.w + lx
    bool inside;
};

__device__ __forceinline__ TilePixel tile_pixel(const RoiParams& roi) {
    const uint32_t tiles_x = (roi.w + kTileW * kWarpsX - 1) / (kTileW * kWarpsX);
    uint32_t tile_x, owned_row;
    tile_of(blockIdx.x, tiles_x, gridDim.x / tiles_x, roi.tile_row_reverse, &tile_x, &owned_row);
    const uint32_t tile_y = owned_row * roi.tile_row_stride + roi.tile_row_phase;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TilePixel p;
    p.lx = tile_x * (kTileW * kWarpsX) + (warp % kWarpsX) * kTileW + (lane % kTileW);
    p.ly = tile_y * (kTileH * kWarpsY) + (warp / kWarpsX) * kTileH + (lane / kTileW);
    p.inside = p.lx < roi.w && p.ly < roi.h;
    p.ray = p.ly * roi.w + p.lx;
    return p;
}

template <bool kLinear, bool kClamp, class V>
__device__ __forceinline__ float4 lean_sample(const V* __restrict__ g, int32_t nx, int32_t ny, int32_t nz, float px,
                                              float py, float pz) {
    if (kLinear) return sample_packed_lean<kClamp>(g, nx, ny, nz, px, py, pz);
    return sample_packed<false, kClamp, false, V>(g, nx, ny, nz, px, py, pz);
}

// d/d(position) of sigma and of h = g . rgb inside one trilinear cell: differences of the lerped
// faces, scaled by (n - 1) per axis; an axis on which the position was clamped (OOB clamp) has zero derivative.
__device__ __forceinline__ void corner_gradients(const Corners& k, const Cell& c, float g0, float g1, float g2, float sx,
                                                 float sy, float sz, float grad_sigma[3], float grad_h[3]) {
    const float4 v[8] = {k.v000, k.v100, k.v010, k.v110, k.v001, k.v101, k.v011, k.v111};
    float s[8], h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s[i] = v[i].w;
        h[i] = g0 * v[i].x + g1 * v[i].y + g2 * v[i].z;
    }
    const float ux = 1.0f - c.tx, uy = 1.0f - c.ty, uz = 1.0f - c.tz;
    auto partials = [&](const float* f, float* out) {
        out[0] = ((f[1] - f[0]) * uy + (f[3] - f[2]) * c.ty) * uz + ((f[5] - f[4]) * uy + (f[7] - f[6]) * c.ty) * c.tz;
        out[1] = ((f[2] - f[0]) * ux + (f[3] - f[1]) * c.tx) * uz + ((f[6] - f[4]) * ux + (f[7] - f[5]) * c.tx) * c.tz;
        out[2] = ((f[4] - f[0]) * ux + (f[5] - f[1]) * c.tx) * uy + ((f[6] - f[2]) * ux + (f[7] - f[3]) * c.tx) * c.ty;
    };
    partials(s, grad_sigma);
    partials(h, grad_h);
    grad_sigma[0] *= sx; grad_sigma[1] *= sy; grad_sigma[2] *= sz;
    grad_h[0] *= sx; grad_h[1] *= sy; grad_h[2] *= sz;
}

// Sample + scatter cell + (optionally) the field gradients the camera adjoint needs, from ONE set of corner loads.
template <bool kClamp, bool kGradients, bool kOcc = false, class V = float4>
__device__ __forceinline__ float4 sample_cell_gradients(const V* __restrict__ g, int32_t nx, int32_t ny, int32_t nz,
                                                        float px, float py, float pz, float g0, float g1, float g2,
                                                        float4& cell, float grad_sigma[3], float grad_h[3],
                                                        const uint32_t* __restrict__ occ = nullptr, int32_t obx = 0, int32_t oby = 0) {
    if (kGradients) {
        grad_sigma[0] = grad_sigma[1] = grad_sigma[2] = 0.f;
        grad_h[0] = grad_h[1] = grad_h[2] = 0.f;
    }
    const bool ox = px < 0.0f || px > 1.0f, oy = py < 0.0f || py > 1.0f, oz = pz < 0.0f || pz > 1.0f;
    float fx, fy, fz;
    if (!grid_coords(px, py, pz, kClamp, nx, ny, nz, fx, fy, fz)) {
        cell = make_float4(__uint_as_float(0xffffffffu), 0.f, 0.f, 0.f);
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const Cell c = make_cell(fx, fy, fz, nx, ny, nz);
    cell = make_float4(__uint_as_float(static_cast<uint32_t>(c.x0) | (static_cast<uint32_t>(c.y0) << 10) |
                                       (static_cast<uint32_t>(c.z0) << 20)),
                       c.tx, c.ty, c.tz);
    // empty brick (every channel zero at all eight corners): value and field gradients are exactly zero, no gather
    if (kOcc && occupancy_bits(occ, c.x0, c.y0, c.z0, obx, oby) == 0u) return make_float4(0.f, 0.f, 0.f, 0.f);
    const Corners k = load_corners(g, c, nx, ny);
    if (kGradients)
        corner_gradients(k, c, g0, g1, g2, (kClamp && ox) ? 0.f : static_cast<float>(nx - 1),
                         (kClamp && oy) ? 0.f : static_cast<float>(ny - 1), (kClamp && oz) ? 0.f : static_cast<float>(nz - 1),
                         grad_sigma, grad_h);
    return trilerp_pairs(k, c.tx, c.ty, c.tz);
}

// Warp-wide range of steps that can touch the unit cube (OOB-zero fields only): [lo, hi), lo aligned down to a
// segment start.  Forward and backward kernels evaluate exactly this function, so they agree on the range.
struct WarpRange { uint32_t lo, hi; };

template <bool kSkipOutside>
__device__ __forceinline__ WarpRange warp_step_range(const MarchParams& mp, const Ray& ray, bool inside, float& t_in,
                                                     float& t_out) {
    const uint32_t count = mp.uniform_count;
    t_in = -CUDART_INF_F;
    t_out = CUDART_INF_F;
    uint32_t k_lo = 0, k_hi = count;
    if (kSkipOutside) {
        cube_interval(ray, t_in, t_out);
        step_range(mp.t_near, mp.dt, count, t_in, t_out, k_lo, k_hi);
    }
    if (!inside) { k_lo = count; k_hi = 0; }
    WarpRange r;
    r.lo = __reduce_min_sync(0xffffffffu, k_lo) & ~static_cast<uint32_t>(kSegment - 1);
    r.hi = __reduce_max_sync(0xffffffffu, k_hi);
    if (r.hi < r.lo) r.hi = r.lo;
    return r;
}

template <bool kLinear, bool kClamp, bool kStratified, bool kOcc = false, class V = float4>
__global__ void __launch_bounds__(kLeanThreads)
lean_forward_kernel(const FrameParams* __restrict__ P, const V* __restrict__ grid, int32_t nx, int32_t ny,
                    int32_t nz, LeanBuffers out, const uint32_t* __restrict__ occ = nullptr, int32_t obx = 0, int32_t oby = 0) {
    const CameraParams cam = P->cam;
    const MarchParams mp = P->march;
    const RoiParams roi = P->roi;
    const TilePixel px = tile_pixel(roi);

    const uint32_t gx = roi.x + px.lx, gy = roi.y + px.ly;
    const Ray ray = make_ray(cam, gx, gy);
    const uint64_t ray_index = mp.ray_index_base + px.ray;

    RayAccum acc;
    bool alive = px.inside;
    const uint32_t count = mp.uniform_count;
    uint32_t live = count;                       // samples integrated before the stop (all of them if the ray never stops)
    float t_in, t_out;
    const WarpRange wr = warp_step_range<!kClamp>(mp, ray, px.inside, t_in, t_out);
    float* const ckpt = out.ckpt != nullptr && px.inside ? out.ckpt + px.ray : nullptr;

    // segments before the range: nothing has been absorbed yet
    if (ckpt != nullptr)
        for (uint32_t s = 0; s < wr.lo; s += kSegment) ckpt[static_cast<size_t>(s / kSegment) * out.ckpt_stride] = 1.0f;

    uint32_t step = wr.lo;
    for (; step < wr.hi; ++step) {
        if (!__any_sync(0xffffffffu, alive)) break;
        if ((step % kSegment) == 0 && alive && ckpt != nullptr) ckpt[static_cast<size_t>(step / kSegment) * out.ckpt_stride] = acc.T;
        if (alive) {
            const float4 tab = __ldg(out.steps + step);
            // whole step outside the cube: sigma = 0, nothing changes (the depth cursor comes from the table)
            if (!kClamp && (tab.x > t_out || tab.x + mp.dt < t_in)) continue;
            const float t = step_time<kStratified>(tab, mp.t_near, mp.t_far, mp.dt, mp.seed, ray_index, step);
            const float pxw = ray.ox + ray.dx * t;
            const float pyw = ray.oy + ray.dy * t;
            const float pzw = ray.oz + ray.dz * t;
            float4 v;
            if (kOcc) {   // (linear fields only) a sample in an empty brick changes nothing: skip its gather and its arithmetic
                if (!sample_packed_lean_occ<kClamp>(grid, occ, obx, oby, nx, ny, nz, pxw, pyw, pzw, v)) continue;
            } else {
                v = lean_sample<kLinear, kClamp>(grid, nx, ny, nz, pxw, pyw, pzw);
            }
            float a, w, tb;
            acc.t_cursor = tab.w;
            if (integrate_sample<false>(acc, tab.z, v, a, w, tb)) {
                alive = false;
                live = step + 1;
            }
        }
    }
    // segments after the range (or after the whole warp has stopped): transmittance no longer changes
    if (ckpt != nullptr && alive)
        for (uint32_t s = (step + kSegment - 1) & ~static_cast<uint32_t>(kSegment - 1); s < count; s += kSegment)
            ckpt[static_cast<size_t>(s / kSegment) * out.ckpt_stride] = acc.T;

    if (px.inside) {
        float opacity, depth;
        finish_ray(acc, mp.t_far, opacity, depth);
        const size_t pid = static_cast<size_t>(gy) * roi.img_w + gx;
        out.image[pid * 3 + 0] = acc.cr;
        out.image[pid * 3 + 1] = acc.cg;
        out.image[pid * 3 + 2] = acc.cb;
        out.trans[pid] = acc.T;
        out.opacity[pid] = opacity;
        out.depth[pid] = depth;
        out.hitmask[pid] = 1u;
        if (out.live != nullptr) out.live[px.ray] = live;
    }
    if (out.live_total != nullptr) {
        // one 64-bit atomic per warp
        uint32_t s = px.inside ? live : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0 && s != 0) atomicAdd(out.live_total, static_cast<unsigned long long>(s));
    }
}

// End-of-CTA completion signal, see LeanBuffers::group_done.
__device__ __forceinline__ void signal_group_done(const LeanBuffers& st, const RoiParams& roi) {
    if (st.group_done == nullptr) return;
    __syncthreads();                       // every warp of the CTA has issued its reds
    if (threadIdx.x == 0) {
        __threadfence();                   // ... and they are visible device-wide before the counter moves
        const uint32_t tiles_x = (roi.w + kTileW * kWarpsX - 1) / (kTileW * kWarpsX);
        // the i-th DISPATCHED row (column order: a row is complete only with the launch's last column)
        const uint32_t row = (roi.tile_row_reverse & kTileOrderColumns) ? blockIdx.x % (gridDim.x / tiles_x) : blockIdx.x / tiles_x;
        uint32_t g = 0;
        while (g + 1 < st.group_count && row >= st.group_end[g]) ++g;
        atomicAdd(st.group_done + g, 1u);
    }
}

// Per-sample state of one segment, stashed in shared memory between the forward recompute and the
// reverse sweep: [sample][field][thread] so that a warp touches 32 consecutive banks.  Keeping the
// two loops rolled (dynamic index into shared memory instead of unrolled register arrays) keeps the
// kernel body small: the fully unrolled form stalled on instruction fetch (ncu: no_instruction).
struct SegmentStash {
    float alpha[kSegment][kLeanThreads];
    float T_prev[kSegment][kLeanThreads];
    float dot[kSegment][kLeanThreads];
    float t[kSegment][kLeanThreads];
    float dt[kSegment][kLeanThreads];
};

template <bool kLinear, bool kClamp, bool kStratified, bool kOcc = false, class V = float4>
__global__ void __launch_bounds__(kLeanThreads)
lean_backward_kernel(const FrameParams* __restrict__ P, const V* __restrict__ grid, int32_t nx, int32_t ny,
                     int32_t nz, ScatterParams sp, const float* __restrict__ dL_dI, LeanBuffers st,
                     const uint32_t* __restrict__ occ = nullptr, int32_t obx = 0, int32_t oby = 0) {
    __shared__ SegmentStash stash;
    const float inv_q = scatter_inv_quantum(sp);
    const CameraParams cam = P->cam;
    const MarchParams mp = P->march;
    const RoiParams roi = P->roi;
    const TilePixel px = tile_pixel(roi);
    const uint32_t tid = threadIdx.x;

    const Ray ray = make_ray(cam, roi.x + px.lx, roi.y + px.ly);
    const uint64_t ray_index = mp.ray_index_base + px.ray;

    uint32_t live = 0;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (px.inside) {
        live = st.live[px.ray];
        g0 = dL_dI[static_cast<size_t>(px.ray) * 3 + 0];
        g1 = dL_dI[static_cast<size_t>(px.ray) * 3 + 1];
        g2 = dL_dI[static_cast<size_t>(px.ray) * 3 + 2];
    }
    const uint32_t nseg = (live + kSegment - 1) / kSegment;
    uint32_t warp_nseg = nseg;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_nseg = max(warp_nseg, __shfl_xor_sync(0xffffffffu, warp_nseg, o));

    const bool skippable = !kClamp && sp.unit_bbox != 0u;   // outside samples touch neither T nor the grid
    float t_in, t_out;
    const WarpRange wr = skippable ? warp_step_range<true>(mp, ray, px.inside, t_in, t_out)
                                   : warp_step_range<false>(mp, ray, px.inside, t_in, t_out);
    const uint32_t seg_lo = wr.lo / kSegment, seg_hi = min(warp_nseg, (wr.hi + kSegment - 1) / kSegment);

    float adj_T = 0.0f;
    // all lanes of a warp walk the same segment index in the same iteration: the warp's gathers and
    // scatters of one iteration stay inside one thin slab of the grid
    for (uint32_t seg = seg_hi; seg-- > seg_lo;) {
        // Re-converge the warp every segment.  Lanes leave the two inner loops at different trip counts on early-
        // terminating volumes; without this barrier the compiler's reconvergence point moved behind the whole segment
        // loop after a harmless-looking change to the scatter code and the kernel ran with 13 of 32 lanes active
        // (ncu: 8.2 G instead of 2.3 G warp instructions, 10.2 ms instead of 5.6 ms on the dense config-2 volume).
        __syncwarp();
        if (seg >= nseg) continue;
        const uint32_t first = seg * kSegment;
        const uint32_t count = min(static_cast<uint32_t>(kSegment), live - first);
        float T = st.ckpt[static_cast<size_t>(seg) * st.ckpt_stride + px.ray];
        // forward part: recompute the segment from its transmittance checkpoint
#pragma unroll 1
        for (uint32_t j = 0; j < count; ++j) {
            const float4 tab = __ldg(st.steps + first + j);
            if (skippable && (tab.x > t_out || tab.x + mp.dt < t_in)) {
                stash.alpha[j][tid] = -1.0f;   // marker: no contribution, adj_T unchanged
                continue;
            }
            const float t = step_time<kStratified>(tab, mp.t_near, mp.t_far, mp.dt, mp.seed, ray_index, first + j);
            float4 v;
            if (kOcc) {   // empty brick (all channels zero): no gather; the sample still receives d sigma = -adj_T T dt below
                float4 cell;
                v = sample_cell_gradients<kClamp, false, true>(grid, nx, ny, nz, ray.ox + ray.dx * t, ray.oy + ray.dy * t,
                                                               ray.oz + ray.dz * t, g0, g1, g2, cell, nullptr, nullptr, occ, obx, oby);
            } else {
                v = lean_sample<kLinear, kClamp>(grid, nx, ny, nz, ray.ox + ray.dx * t, ray.oy + ray.dy * t, ray.oz + ray.dz * t);
            }
            const float a = alpha_of(v.w, tab.z);
            stash.alpha[j][tid] = a;
            stash.T_prev[j][tid] = T;
            stash.dot[j][tid] = g0 * v.x + g1 * v.y + g2 * v.z;
            stash.t[j][tid] = t;
            stash.dt[j][tid] = tab.z;
            T = T * fmaxf(1.0f - a, 0.0f);
        }
        // reverse part: the reference's adjoint recurrence, then the grid scatter
#pragma unroll 1
        for (uint32_t j = count; j-- > 0;) {
            const float a = stash.alpha[j][tid], Tp = stash.T_prev[j][tid];
            if (a < 0.0f) continue;
            const float w = Tp * a;
            float dsigma;
            adjoint_sample(stash.dot[j][tid], a, Tp, stash.dt[j][tid], adj_T, dsigma);
            const float t = stash.t[j][tid];
            scatter_sample(sp, ray.ox + ray.dx * t, ray.oy + ray.dy * t, ray.oz + ray.dz * t,
                           make_float4(g0 * w, g1 * w, g2 * w, dsigma), inv_q);
        }
    }
    signal_group_done(st, roi);
}

// ---- backward, merged scatter ------------------------------------------------
// The scatter primitive (red.global.add.v4.f32) costs per ACTIVE LANE: tools/mem_probe measures
// ~300 G lane-reds/s on a B200 and ncu shows no merging of equal addresses inside a request
// (profiles/r01_c2_lean_v1.md).  When pixels are denser than voxels (config 2: 0.3 voxel between
// neighbouring rays, 0.75 voxel between steps) most of a warp's 256 reds per step hit the same few
// voxels.  This kernel cuts the lane-reds by merging in REGISTERS before issuing them:
//
//   phase A  (lane = ray)      recompute the segment forward from its checkpoint -> stash {alpha, T_prev, g.c, t}
//   phase A' (lane = ray)      reverse sweep with the reference's adjoint recurrence; per sample the gradient
//                              {g w, d sigma} and the scatter cell {x0|y0|z0, tx, ty, tz} go to shared memory
//   phase B  (lane = 2x2 pixel quad x 2 steps)   walks its 8 samples, accumulates the 8 corner contributions while
//                              the cell stays the same and issues 8 reds only when it changes.
//
// Only __syncwarp() separates the phases: a quad's rays and all steps of a segment live in one warp.
constexpr uint32_t kNoCell = 0xffffffffu;
constexpr int kStashRow = kLeanThreads + kLeanThreads / 16;   // skewed by one float4 per 16 threads: phase B reads conflict-free

// Chain rule of one ray (d = v / |v|, v = R q) followed by a deterministic block reduction:
// warp shuffles, then the warps summed in order; one 16-double partial per CTA.
__device__ __forceinline__ void camera_block_reduce(const CameraParams& cam, const Ray& ray, const RayAux& ra, bool inside,
                                                    const float acc_o[3], const float acc_d[3],
                                                    double* __restrict__ partials) {
    double out[16];
    {
        const double d[3] = {ray.dx, ray.dy, ray.dz};
        const double dd = d[0] * acc_d[0] + d[1] * acc_d[1] + d[2] * acc_d[2];
        double dv[3];
        for (int i = 0; i < 3; ++i) dv[i] = (static_cast<double>(acc_d[i]) - d[i] * dd) * static_cast<double>(ra.inv_len);
        const double q[3] = {ra.qx, ra.qy, ra.qz};
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) out[4 * i + j] = dv[i] * q[j];
            out[4 * i + 3] = acc_o[i];
        }
        if (cam.ortho) {
            out[12] = out[13] = out[14] = out[15] = 0.0;
        } else {
            const double dq0 = cam.r00 * dv[0] + cam.r10 * dv[1] + cam.r20 * dv[2];
            const double dq1 = cam.r01 * dv[0] + cam.r11 * dv[1] + cam.r21 * dv[2];
            out[12] = -q[0] / cam.fx * dq0;
            out[13] = -q[1] / cam.fy * dq1;
            out[14] = -dq0 / cam.fx;
            out[15] = -dq1 / cam.fy;
        }
        if (!inside) {
            for (int i = 0; i < 16; ++i) out[i] = 0.0;
        }
    }
    __shared__ double smem[kLeanThreads / 32][16];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        double x = out[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) smem[warp][i] = x;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double x = 0.0;
        for (int w = 0; w < kLeanThreads / 32; ++w) x += smem[w][threadIdx.x];
        partials[static_cast<size_t>(blockIdx.x) * 16 + threadIdx.x] = x;
    }
}

template <bool kCamera>
struct MergeStash {
    float4 a[kSegment][kStashRow];   // phase A: {alpha, T_prev, g.c, t};  phase A': {g0 w, g1 w, g2 w, d sigma}
    float4 c[kSegment][kStashRow];   // scatter cell {x0 | y0 << 10 | z0 << 20, tx, ty, tz}; key 0xffffffff = nothing to scatter
    float s[kCamera ? 3 : 1][kCamera ? kSegment : 1][kCamera ? kStashRow : 1];   // camera adjoint only: d sigma / d position
                                                                                  // (3 planes: 48 KB of static shared memory in all)
};

__device__ __forceinline__ uint32_t stash_slot(uint32_t tid) { return tid + (tid >> 4); }

// scatter cell of a position: src/fields/dense_grid.cpp:206-246 (linear, every axis >= 2 voxels)
__device__ __forceinline__ float4 scatter_cell(const ScatterParams& sp, float px, float py, float pz) {
    const float ex = sp.bmax[0] - sp.bmin[0], ey = sp.bmax[1] - sp.bmin[1], ez = sp.bmax[2] - sp.bmin[2];
    float lx = ex != 0.0f ? (px - sp.bmin[0]) / ex : 0.0f;
    float ly = ey != 0.0f ? (py - sp.bmin[1]) / ey : 0.0f;
    float lz = ez != 0.0f ? (pz - sp.bmin[2]) / ez : 0.0f;
    const bool outside = lx < 0.0f || lx > 1.0f || ly < 0.0f || ly > 1.0f || lz < 0.0f || lz > 1.0f;
    if (outside) {
        if (!sp.clamp) return make_float4(__uint_as_float(kNoCell), 0.f, 0.f, 0.f);
        lx = fmaxf(0.0f, fminf(1.0f, lx));
        ly = fmaxf(0.0f, fminf(1.0f, ly));
        lz = fmaxf(0.0f, fminf(1.0f, lz));
    }
    const Cell c = make_cell(lx * static_cast<float>(sp.nx - 1), ly * static_cast<float>(sp.ny - 1),
                             lz * static_cast<float>(sp.nz - 1), sp.nx, sp.ny, sp.nz);
    const uint32_t key = static_cast<uint32_t>(c.x0) | (static_cast<uint32_t>(c.y0) << 10) | (static_cast<uint32_t>(c.z0) << 20);
    return make_float4(__uint_as_float(key), c.tx, c.ty, c.tz);
}

// 8 corner accumulators {d r, d g, d b, d sigma}: corner k = dx + 2 dy + 4 dz.  Kept as float4 so that the register
// allocator can place each one in an aligned quad: FFMA2 updates its (x,y) / (z,w) halves in place and
// red.global.add.v4.f32 consumes the quad without moves.
// kPlain: float reds into the grid's own gradient block in its default order (the common case) -- no box test, no
// stride multiplies, no fixed-point branch.
template <bool kPlain>
__device__ __forceinline__ void flush_cell(const ScatterParams& sp, uint32_t key, const float4 (&acc)[8], float inv_q) {
    const int32_t x0 = key & 1023u, y0 = (key >> 10) & 1023u, z0 = key >> 20;
    const int32_t x1 = min(x0 + 1, sp.nx - 1), y1 = min(y0 + 1, sp.ny - 1), z1 = min(z0 + 1, sp.nz - 1);
    if (kPlain) {
        const uint32_t p00 = voxel_index32(0, y0, z0, sp.nx, sp.ny), p10 = voxel_index32(0, y1, z0, sp.nx, sp.ny);
        const uint32_t p01 = voxel_index32(0, y0, z1, sp.nx, sp.ny), p11 = voxel_index32(0, y1, z1, sp.nx, sp.ny);
        red_add4(sp.grad + (p00 + x0), acc[0]); red_add4(sp.grad + (p00 + x1), acc[1]);
        red_add4(sp.grad + (p10 + x0), acc[2]); red_add4(sp.grad + (p10 + x1), acc[3]);
        red_add4(sp.grad + (p01 + x0), acc[4]); red_add4(sp.grad + (p01 + x1), acc[5]);
        red_add4(sp.grad + (p11 + x0), acc[6]); red_add4(sp.grad + (p11 + x1), acc[7]);
        return;
    }
    if (sp.boxed) {
        const int32_t lx0 = x0 - sp.box_ox, ly0 = y0 - sp.box_oy, lz0 = z0 - sp.box_oz;
        const int32_t lx1 = x1 - sp.box_ox, ly1 = y1 - sp.box_oy, lz1 = z1 - sp.box_oz;
        if (lx0 < 0 || lx1 >= sp.box_nx || ly0 < 0 || ly1 >= sp.box_ny || lz0 < 0 || lz1 >= sp.box_nz) {
            atomicAdd(sp.box_miss, 1u);   // cannot happen with the bounds of hpx_frame_bounds; never write outside the box
            return;
        }
    }
    // sp.grad is biased by the box origin on the host: grid coordinates index it directly
    const uint32_t bx0 = static_cast<uint32_t>(x0) * sp.box_sx, bx1 = static_cast<uint32_t>(x1) * sp.box_sx;
    const uint32_t ry0 = static_cast<uint32_t>(y0) * sp.box_sy, ry1 = static_cast<uint32_t>(y1) * sp.box_sy;
    const uint32_t rz0 = static_cast<uint32_t>(z0) * sp.box_sz, rz1 = static_cast<uint32_t>(z1) * sp.box_sz;
    const uint32_t r00 = rz0 + ry0, r10 = rz0 + ry1, r01 = rz1 + ry0, r11 = rz1 + ry1;
    // (a rolled loop over the corners would index acc[] dynamically: the accumulators would move to local memory and
    // the kernel doubles in time -- measured)
    scatter_add(sp, r00 + bx0, acc[0], inv_q); scatter_add(sp, r00 + bx1, acc[1], inv_q);
    scatter_add(sp, r10 + bx0, acc[2], inv_q); scatter_add(sp, r10 + bx1, acc[3], inv_q);
    scatter_add(sp, r01 + bx0, acc[4], inv_q); scatter_add(sp, r01 + bx1, acc[5], inv_q);
    scatter_add(sp, r11 + bx0, acc[6], inv_q); scatter_add(sp, r11 + bx1, acc[7], inv_q);
}

// compare-exchange of the 19-comparator sorting network for 8 keys
#define DV_CE(a, b) { const uint32_t lo_ = min(k[a], k[b]), hi_ = max(k[a], k[b]); k[a] = lo_; k[b] = hi_; }

#ifndef DV_MERGE_MIN_BLOCKS
#define DV_MERGE_MIN_BLOCKS 5
#endif
// kCamera: also accumulate d L / d (ray origin, ray direction) = sum_s (d sigma_s grad sigma(x_s) + w_s grad (g . rgb)(x_s)) {1, t_s}
// from the corners phase A has in registers anyway, and reduce it to the camera parameters at the end
// (replaces a separate camera_adjoint_kernel pass over every sample).
template <bool kClamp, bool kStratified, bool kUnitBox, bool kCamera, bool kPlain, bool kOcc = false, class V = float4>
__global__ void __launch_bounds__(kLeanThreads, kCamera ? 4 : DV_MERGE_MIN_BLOCKS)
lean_backward_merge_kernel(const FrameParams* __restrict__ P, const V* __restrict__ grid, int32_t nx, int32_t ny,
                           int32_t nz, ScatterParams sp, const float* __restrict__ dL_dI, LeanBuffers st,
                           double* __restrict__ cam_partials, const uint32_t* __restrict__ occ = nullptr, int32_t obx = 0,
                           int32_t oby = 0) {
    __shared__ MergeStash<kCamera> stash;
    const CameraParams cam = P->cam;
    const MarchParams mp = P->march;
    const RoiParams roi = P->roi;
    const TilePixel px = tile_pixel(roi);
    const uint32_t tid = threadIdx.x, slot = stash_slot(tid);
    const float inv_q = scatter_inv_quantum(sp);

    const Ray ray = make_ray(cam, roi.x + px.lx, roi.y + px.ly);
    const uint64_t ray_index = mp.ray_index_base + px.ray;
    float cam_o[3] = {0.f, 0.f, 0.f}, cam_d[3] = {0.f, 0.f, 0.f};

    uint32_t live = 0;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (px.inside) {
        live = st.live[px.ray];
        g0 = dL_dI[static_cast<size_t>(px.ray) * 3 + 0];
        g1 = dL_dI[static_cast<size_t>(px.ray) * 3 + 1];
        g2 = dL_dI[static_cast<size_t>(px.ray) * 3 + 2];
    }
    const uint32_t nseg = (live + kSegment - 1) / kSegment;
    const uint32_t warp_nseg = __reduce_max_sync(0xffffffffu, nseg);

    constexpr bool skippable = !kClamp && kUnitBox;   // outside samples touch neither T nor the grid
    float t_in, t_out;
    const WarpRange wr = warp_step_range<skippable>(mp, ray, px.inside, t_in, t_out);
    const uint32_t seg_lo = wr.lo / kSegment, seg_hi = min(warp_nseg, (wr.hi + kSegment - 1) / kSegment);

    // phase B role of this lane: quad q (4 x 2 quads in the warp's 8 x 4 pixel tile), step pair sb
    const uint32_t lane = tid & 31u, warp_base = tid & ~31u;
    const uint32_t sb = lane >> 3, qx = lane & 3u, qy = (lane >> 2) & 1u;
    const uint32_t quad_slot = stash_slot(warp_base + (2u * qy) * kTileW + 2u * qx);   // slot of the quad's top-left ray
    // the quad's rays sit at lanes +0, +1, +kTileW, +kTileW+1: a quad never straddles a 16-thread skew boundary

    float adj_T = 0.0f;
    for (uint32_t seg = seg_hi; seg-- > seg_lo;) {
        const uint32_t first = seg * kSegment;
        const uint32_t count = seg < nseg ? min(static_cast<uint32_t>(kSegment), live - first) : 0u;
        // ---- phase A: forward recompute from the checkpoint
        if (count != 0) {
            float T = st.ckpt[static_cast<size_t>(seg) * st.ckpt_stride + px.ray];
#pragma unroll 1
            for (uint32_t j = 0; j < count; ++j) {
                const float4 tab = __ldg(st.steps + first + j);
                if (skippable && (tab.x > t_out || tab.x + mp.dt < t_in)) {
                    stash.a[j][slot] = make_float4(-1.0f, 0.f, 0.f, 0.f);   // marker: no contribution, adj_T unchanged
                    continue;
                }
                const float t = step_time<kStratified>(tab, mp.t_near, mp.t_far, mp.dt, mp.seed, ray_index, first + j);
                float4 cell;
                float gs[3], gh[3];
                const float4 v = sample_cell_gradients<kClamp, kCamera, kOcc>(grid, nx, ny, nz, ray.ox + ray.dx * t, ray.oy + ray.dy * t,
                                                                              ray.oz + ray.dz * t, g0, g1, g2, cell, gs, gh, occ, obx, oby);
                if (kUnitBox) {
                    stash.c[j][slot] = cell;
                    if (!kClamp && __float_as_uint(cell.x) == kNoCell) {   // outside the cube: sigma = rgb = 0, adj_T unchanged, no scatter
                        stash.a[j][slot] = make_float4(-1.0f, 0.f, 0.f, 0.f);
                        continue;
                    }
                }
                const float a = alpha_of(v.w, tab.z);
                stash.a[j][slot] = make_float4(a, T, g0 * v.x + g1 * v.y + g2 * v.z, t);
                if (kCamera) {
                    const float w = T * a;   // the colour part needs nothing from later samples
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        cam_o[i] += w * gh[i];
                        cam_d[i] += t * (w * gh[i]);
                    }
                    stash.s[0][j][slot] = gs[0];
                    stash.s[1][j][slot] = gs[1];
                    stash.s[2][j][slot] = gs[2];
                }
                T = T * fmaxf(1.0f - a, 0.0f);
            }
        }
        // ---- phase A': reverse sweep (diff_cpu.cpp:170-194): per-sample gradients to shared memory
#pragma unroll 1
        for (uint32_t j = kSegment; j-- > 0;) {
            bool valid = false;
            if (j < count) {
                const float4 s = stash.a[j][slot];
                if (s.x >= 0.0f) {
                    const float a = s.x, Tp = s.y;
                    const float dtv = __ldg(st.steps + first + j).z;
                    const float w = Tp * a;
                    float dsigma;
                    adjoint_sample(s.z, a, Tp, dtv, adj_T, dsigma);
                    stash.a[j][slot] = make_float4(g0 * w, g1 * w, g2 * w, dsigma);
                    if (kCamera) {
                        const float t = s.w;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const float x = dsigma * stash.s[i][j][slot];
                            cam_o[i] += x;
                            cam_d[i] += t * x;
                        }
                    }
                    if (!kUnitBox) {
                        const float t = s.w;
                        stash.c[j][slot] = scatter_cell(sp, ray.ox + ray.dx * t, ray.oy + ray.dy * t, ray.oz + ray.dz * t);
                    }
                    valid = true;
                }
            }
            if (!valid) stash.c[j][slot].x = __uint_as_float(kNoCell);
        }
        __syncwarp();
        // ---- phase B: this lane merges steps {2 sb, 2 sb + 1} of its quad's four rays
        {
            // sort the 8 samples by cell so that equal cells are adjacent.  Sort key = low 3 bits of each cell
            // coordinate (neighbouring samples differ by far less than 8 cells; a collision would only cost a missed
            // merge, equality below is decided on the full key) with the sample number in the low 3 bits.
            uint32_t k[8];
#pragma unroll
            for (uint32_t i = 0; i < 8; ++i) {
                const uint32_t key = __float_as_uint(stash.c[2u * sb + (i >> 2)][quad_slot + (i & 1u) + ((i >> 1) & 1u) * kTileW].x);
                const uint32_t rel = (key & 7u) | ((key >> 7) & 0x38u) | ((key >> 14) & 0x1c0u);
                k[i] = ((key == kNoCell ? 0xfffu : rel) << 3) | i;
            }
            DV_CE(0, 1) DV_CE(2, 3) DV_CE(4, 5) DV_CE(6, 7) DV_CE(0, 2) DV_CE(1, 3) DV_CE(4, 6) DV_CE(5, 7) DV_CE(1, 2) DV_CE(5, 6)
            DV_CE(0, 4) DV_CE(3, 7) DV_CE(1, 5) DV_CE(2, 6) DV_CE(1, 4) DV_CE(3, 6) DV_CE(2, 4) DV_CE(3, 5) DV_CE(3, 4)
            uint32_t order = 0;
#pragma unroll
            for (uint32_t i = 0; i < 8; ++i) order |= (k[i] & 7u) << (3u * i);

            float4 acc[8];
            uint32_t cur = kNoCell;
            const uint64_t one2 = pack2(1.0f, 1.0f);
#pragma unroll 1
            for (uint32_t n = 0; n <= 8; ++n) {
                uint32_t key = kNoCell;
                float4 c = make_float4(0.f, 0.f, 0.f, 0.f), gv = c;
                if (n < 8) {
                    const uint32_t i = (order >> (3u * n)) & 7u;
                    const uint32_t j = 2u * sb + (i >> 2);
                    const uint32_t rs = quad_slot + (i & 1u) + ((i >> 1) & 1u) * kTileW;
                    c = stash.c[j][rs];
                    key = __float_as_uint(c.x);
                    if (key != kNoCell) gv = stash.a[j][rs];
                }
                const bool fresh = key != cur;
                if (fresh) {
                    if (cur != kNoCell) flush_cell<kPlain>(sp, cur, acc, inv_q);
                    cur = key;
                }
                if (key != kNoCell) {
                    // (1-tx)(1-ty)(1-tz) ... as duplicated pairs so that the products feed FFMA2 directly
                    const uint64_t tx2 = pack2(c.y, c.y), ty2 = pack2(c.z, c.z), tz2 = pack2(c.w, c.w);
                    const uint64_t ux2 = sub2(one2, tx2), uy2 = sub2(one2, ty2), uz2 = sub2(one2, tz2);
                    const uint64_t wxy[4] = {mul2(ux2, uy2), mul2(tx2, uy2), mul2(ux2, ty2), mul2(tx2, ty2)};
                    const uint64_t glo = pack2(gv.x, gv.y), ghi = pack2(gv.z, gv.w);
                    if (fresh) {   // first sample of a run: plain products, no zero fill of the accumulators
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const uint64_t w2 = mul2(wxy[q & 3], (q & 4) ? tz2 : uz2);
                            unpack2(mul2(glo, w2), acc[q].x, acc[q].y);
                            unpack2(mul2(ghi, w2), acc[q].z, acc[q].w);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const uint64_t w2 = mul2(wxy[q & 3], (q & 4) ? tz2 : uz2);
                            unpack2(fma2(glo, w2, pack2(acc[q].x, acc[q].y)), acc[q].x, acc[q].y);
                            unpack2(fma2(ghi, w2, pack2(acc[q].z, acc[q].w)), acc[q].z, acc[q].w);
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    if (kCamera) {
        RayAux ra;
        make_ray(cam, roi.x + px.lx, roi.y + px.ly, &ra);
        camera_block_reduce(cam, ray, ra, px.inside, cam_o, cam_d, cam_partials);
    }
    signal_group_done(st, roi);
}
#undef DV_CE

// ---- camera adjoint ---------------------------------------------------------
// value is not needed, only d/d(position) of sigma and of h = g . rgb
template <bool kClamp, class V>
__device__ __forceinline__ void field_gradients(const V* __restrict__ g, int32_t nx, int32_t ny, int32_t nz,
                                                float px, float py, float pz, float g0, float g1, float g2,
                                                float4& value, float grad_sigma[3], float grad_h[3]) {
    grad_sigma[0] = grad_sigma[1] = grad_sigma[2] = 0.f;
    grad_h[0] = grad_h[1] = grad_h[2] = 0.f;
    value = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool ox = px < 0.0f || px > 1.0f, oy = py < 0.0f || py > 1.0f, oz = pz < 0.0f || pz > 1.0f;
    float fx, fy, fz;
    if (!grid_coords(px, py, pz, kClamp, nx, ny, nz, fx, fy, fz)) return;
    const Cell c = make_cell(fx, fy, fz, nx, ny, nz);
    const Corners k = load_corners(g, c, nx, ny);
    value = trilerp_pairs(k, c.tx, c.ty, c.tz);   // sigma in the reference's operation order: alpha and T match the forward pass
    corner_gradients(k, c, g0, g1, g2, (kClamp && ox) ? 0.f : static_cast<float>(nx - 1),
                     (kClamp && oy) ? 0.f : static_cast<float>(ny - 1), (kClamp && oz) ? 0.f : static_cast<float>(nz - 1),
                     grad_sigma, grad_h);
}

template <bool kClamp, bool kStratified, class V = float4>
__global__ void __launch_bounds__(kLeanThreads)
camera_adjoint_kernel(const FrameParams* __restrict__ P, const V* __restrict__ grid, int32_t nx, int32_t ny,
                      int32_t nz, const float* __restrict__ dL_dI, const uint32_t* __restrict__ live_counts,
                      const float4* __restrict__ steps, double* __restrict__ partials) {
    const CameraParams cam = P->cam;
    const MarchParams mp = P->march;
    const RoiParams roi = P->roi;
    const TilePixel px = tile_pixel(roi);
    RayAux ra;
    const Ray ray = make_ray(cam, roi.x + px.lx, roi.y + px.ly, &ra);
    const uint64_t ray_index = mp.ray_index_base + px.ray;

    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    uint32_t live = 0;
    if (px.inside) {
        live = live_counts[px.ray];
        g0 = dL_dI[static_cast<size_t>(px.ray) * 3 + 0];
        g1 = dL_dI[static_cast<size_t>(px.ray) * 3 + 1];
        g2 = dL_dI[static_cast<size_t>(px.ray) * 3 + 2];
    }
    // forward-order form of the reverse recurrence:
    //   d sigma_s = B_s (dot_s - adjT_{s+1}),  B_s = dt_s (1 - alpha_s) T_s
    //   sum_s X_s adjT_{s+1} = sum_j dot_j alpha_j Y_j,  Y_{j+1} = Y_j (1 - alpha_j) + X_j
    float T = 1.0f;
    float acc_o[3] = {0.f, 0.f, 0.f}, acc_d[3] = {0.f, 0.f, 0.f};
    float Yo[3] = {0.f, 0.f, 0.f}, Yd[3] = {0.f, 0.f, 0.f};
    // samples outside the cube (OOB zero) have sigma = 0 and zero field gradients: T, Y and the sums do not change
    float t_in, t_out;
    const WarpRange wr = warp_step_range<!kClamp>(mp, ray, px.inside, t_in, t_out);
    const uint32_t k_end = min(wr.hi, __reduce_max_sync(0xffffffffu, live));
    for (uint32_t k = wr.lo; k < k_end; ++k) {
        __syncwarp();   // keep the warp converged: lanes skip different steps
        if (k >= live) continue;
        const float4 tab = __ldg(steps + k);
        if (!kClamp && (tab.x > t_out || tab.x + mp.dt < t_in)) continue;
        const float t = step_time<kStratified>(tab, mp.t_near, mp.t_far, mp.dt, mp.seed, ray_index, k);
        const float dtv = tab.z;
        float4 v;
        float gs[3], gh[3];
        field_gradients<kClamp>(grid, nx, ny, nz, ray.ox + ray.dx * t, ray.oy + ray.dy * t, ray.oz + ray.dz * t, g0,
                                g1, g2, v, gs, gh);
        const float a = alpha_of(v.w, dtv);
        const float dot = g0 * v.x + g1 * v.y + g2 * v.z;
        const float w = T * a;
        const float B = dtv * (1.0f - a) * T;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float X = B * gs[i];
            const float term = X * dot + w * gh[i];
            acc_o[i] += term - dot * a * Yo[i];
            acc_d[i] += t * term - dot * a * Yd[i];
            Yo[i] = Yo[i] * (1.0f - a) + X;
            Yd[i] = Yd[i] * (1.0f - a) + t * X;
        }
        T = T * fmaxf(1.0f - a, 0.0f);
    }
    camera_block_reduce(cam, ray, ra, px.inside, acc_o, acc_d, partials);
}

__global__ void camera_reduce_kernel(const double* __restrict__ partials, uint32_t blocks, float* __restrict__ cam16) {
    // 16 warps, one output each; lanes stride over blocks, fixed-order tree at the end
    const uint32_t lane = threadIdx.x & 31, o = threadIdx.x >> 5;
    double x = 0.0;
    for (uint32_t b = lane; b < blocks; b += 32) x += partials[static_cast<size_t>(b) * 16 + o];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
    if (lane == 0) cam16[o] += static_cast<float>(x);
}

__global__ void background_kernel(LeanBuffers out, RoiParams roi, float t_far) {
    const size_t n = static_cast<size_t>(roi.img_w) * roi.img_h;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        out.image[3 * i] = 0.f; out.image[3 * i + 1] = 0.f; out.image[3 * i + 2] = 0.f;
        out.trans[i] = 1.0f;
        out.opacity[i] = 0.0f;
        out.depth[i] = t_far;
        out.hitmask[i] = 0u;
    }
}

uint32_t tile_blocks(const RoiParams& roi) {
    const uint32_t tx = (roi.w + kTileW * kWarpsX - 1) / (kTileW * kWarpsX);
    const uint32_t ty = (roi.h + kTileH * kWarpsY - 1) / (kTileH * kWarpsY);
    const uint32_t stride = roi.tile_row_stride ? roi.tile_row_stride : 1u;
    const uint32_t owned = ty > roi.tile_row_phase ? (ty - roi.tile_row_phase + stride - 1) / stride : 0u;
    return tx * owned;
}

}  // namespace

// ---- deterministic (fixed-point) gradient accumulation -----------------------------------------------------
namespace {
// skip_w: v is a packed {r,g,b,sigma} array and only the colours count (sigma never scales a gradient contribution)
__global__ void abs_max_kernel(const float* __restrict__ v, size_t n, uint32_t* __restrict__ out_bits, bool skip_w) {
    float m = 0.0f;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        if (skip_w && (i & 3u) == 3u) continue;
        const float a = fabsf(v[i]);
        if (a < CUDART_INF_F) m = fmaxf(m, a);   // ignore inf / nan: they poison the gradient either way
    }
    const uint32_t bits = __reduce_max_sync(0xffffffffu, __float_as_uint(m));   // non-negative floats order like their bit patterns
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, bits);
}

// quantum = 2^e with 2^44 quanta per bound B on one contribution: |d colour| <= max|dL/dI|, |d sigma| = |adj_alpha| dt (1 - alpha)
// <= 6 max|dL/dI| max|rgb| dt (adj_alpha = (g.c - adj_T) T_prev with |g.c|, |adj_T| <= 3 max|dL/dI| max|rgb|), so
// B = 8 max|dL/dI| max(1, max|rgb|) max(1, dt); int64 then holds 2^19 contributions of the largest possible size per voxel.
// Rounding error per contribution <= quantum / 2 ~ 3e-14 B: far inside the float32 rounding of the plain path.
__global__ void fixed_scale_kernel(float* __restrict__ meta, float dt) {
    const float gmax = __uint_as_float(reinterpret_cast<const uint32_t*>(meta)[1]);
    const float cmax = __uint_as_float(reinterpret_cast<const uint32_t*>(meta)[0]);
    const float bound = 8.0f * gmax * fmaxf(1.0f, cmax) * fmaxf(1.0f, dt);
    int e = 0;
    if (bound > 0.0f && bound < CUDART_INF_F) e = ilogbf(bound) + 1 - 44;
    e = max(-100, min(100, e));
    meta[2] = ldexpf(1.0f, -e);
    meta[3] = ldexpf(1.0f, e);
}

__global__ void fixed_to_float_kernel(unsigned long long* __restrict__ fixed, float4* __restrict__ grad, size_t voxels,
                                      const float* __restrict__ meta) {
    const float q = meta[3];
    for (size_t v = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; v < voxels;
         v += static_cast<size_t>(gridDim.x) * blockDim.x) {
        ulonglong2* p = reinterpret_cast<ulonglong2*>(fixed + 4 * v);
        const ulonglong2 a = p[0], b = p[1];
        if ((a.x | a.y | b.x | b.y) == 0ull) continue;
        float4 g = grad[v];
        g.x += static_cast<float>(static_cast<long long>(a.x)) * q;
        g.y += static_cast<float>(static_cast<long long>(a.y)) * q;
        g.z += static_cast<float>(static_cast<long long>(b.x)) * q;
        g.w += static_cast<float>(static_cast<long long>(b.y)) * q;
        grad[v] = g;
        p[0] = make_ulonglong2(0ull, 0ull);   // leave the integer grid clean for the next backward
        p[1] = make_ulonglong2(0ull, 0ull);
    }
}
}  // namespace

cudaError_t launch_abs_max(cudaStream_t stream, const float* d_values, size_t n, uint32_t* d_out_bits, bool packed_rgb_only) {
    if (n == 0) return cudaSuccess;
    abs_max_kernel<<<static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 148 * 8)), 256, 0, stream>>>(d_values, n, d_out_bits,
                                                                                                      packed_rgb_only);
    return cudaGetLastError();
}

cudaError_t launch_fixed_scale(cudaStream_t stream, float* d_meta, float dt) {
    fixed_scale_kernel<<<1, 1, 0, stream>>>(d_meta, dt);
    return cudaGetLastError();
}

cudaError_t launch_fixed_to_float(cudaStream_t stream, unsigned long long* d_fixed, float4* d_grad, size_t voxels,
                                  const float* d_meta) {
    if (voxels == 0) return cudaSuccess;
    fixed_to_float_kernel<<<148 * 8, 256, 0, stream>>>(d_fixed, d_grad, voxels, d_meta);
    return cudaGetLastError();
}

// ---- voxel bounds of a frame's rays ------------------------------------------------------------------------------
namespace {
// Box of grid voxels the backward of this launch can touch: along a ray the position is linear in t, so the cells of
// all in-cube samples lie between the cells of the two ends of the (padded) in-cube interval.  bounds = {min x,y,z,
// max x,y,z} as ints, pre-set to {INT_MAX.., INT_MIN..}; the upper corner of a cell (+1) and one voxel of slack for
// rounding are added on the host.
__global__ void __launch_bounds__(kLeanThreads)
ray_bounds_kernel(const FrameParams* __restrict__ P, int32_t nx, int32_t ny, int32_t nz, int* __restrict__ bounds) {
    const CameraParams cam = P->cam;
    const MarchParams mp = P->march;
    const RoiParams roi = P->roi;
    const TilePixel px = tile_pixel(roi);
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    if (px.inside && mp.uniform_count != 0) {
        const Ray ray = make_ray(cam, roi.x + px.lx, roi.y + px.ly);
        float t_in, t_out;
        cube_interval(ray, t_in, t_out);
        const float t_last = mp.t_near + static_cast<float>(mp.uniform_count) * mp.dt;
        const float ta = fmaxf(t_in, mp.t_near), tb = fminf(t_out, fminf(t_last, mp.t_far));
        if (ta <= tb) {
            const float n[3] = {static_cast<float>(nx - 1), static_cast<float>(ny - 1), static_cast<float>(nz - 1)};
            const float o[3] = {ray.ox, ray.oy, ray.oz}, d[3] = {ray.dx, ray.dy, ray.dz};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float pa = fminf(fmaxf(o[i] + d[i] * ta, 0.0f), 1.0f) * n[i];
                const float pb = fminf(fmaxf(o[i] + d[i] * tb, 0.0f), 1.0f) * n[i];
                lo[i] = static_cast<int>(floorf(fminf(pa, pb)));
                hi[i] = static_cast<int>(floorf(fmaxf(pa, pb)));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int l = __reduce_min_sync(0xffffffffu, lo[i]), h = __reduce_max_sync(0xffffffffu, hi[i]);
        if ((threadIdx.x & 31) == 0) {
            if (l != INT_MAX) atomicMin(bounds + i, l);
            if (h != INT_MIN) atomicMax(bounds + 3 + i, h);
        }
    }
}
}  // namespace

cudaError_t launch_ray_bounds(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params, int32_t nx,
                              int32_t ny, int32_t nz, int* d_bounds) {
    const uint32_t blocks = tile_blocks(h_params.roi);
    if (blocks == 0) return cudaSuccess;
    ray_bounds_kernel<<<blocks, kLeanThreads, 0, stream>>>(d_params, nx, ny, nz, d_bounds);
    return cudaGetLastError();
}

// ---- rows of every slab the backward of this launch can touch ------------------------------------------------------
namespace {
// A gradient block laid out slab by slab along axis `a` is, inside a slab, a stack of ROWS along axis `b` (the second
// slowest one).  Per slab: the range of rows that samples of this launch's rays can reach -- a sample at grid coordinate
// p touches the cells floor(p), floor(p) + 1 on every axis, and along a ray p is linear in t, so the rows a ray reaches in
// slab s are those between its positions at the two ends of the t-interval in which p_a lies within [s - 1, s + 1] (the
// bounds below widen both intervals by one more voxel for rounding).  lo / hi: [slabs] ints pre-set to INT_MAX / INT_MIN;
// hi is inclusive.  A sharded frame exchanges only these rows of a slab instead of the whole slab (dv_comm.cu).
__global__ void __launch_bounds__(kLeanThreads)
ray_slab_rows_kernel(const FrameParams* __restrict__ P, int32_t nx, int32_t ny, int32_t nz, int a, int b, int* __restrict__ lo,
                     int* __restrict__ hi) {
    const CameraParams cam = P->cam;
    const MarchParams mp = P->march;
    const RoiParams roi = P->roi;
    const TilePixel px = tile_pixel(roi);
    const int32_t dims[3] = {nx, ny, nz};
    const float na = static_cast<float>(dims[a] - 1), nb = static_cast<float>(dims[b] - 1);
    int s_lo = INT_MAX, s_hi = INT_MIN;
    float oa = 0.f, da = 0.f, ob = 0.f, db = 0.f, ta = 0.f, tb = -1.f;
    if (px.inside && mp.uniform_count != 0) {
        const Ray ray = make_ray(cam, roi.x + px.lx, roi.y + px.ly);
        float t_in, t_out;
        cube_interval(ray, t_in, t_out);
        const float t_last = mp.t_near + static_cast<float>(mp.uniform_count) * mp.dt;
        ta = fmaxf(t_in, mp.t_near);
        tb = fminf(t_out, fminf(t_last, mp.t_far));
        if (ta <= tb) {
            const float o[3] = {ray.ox, ray.oy, ray.oz}, d[3] = {ray.dx, ray.dy, ray.dz};
            oa = o[a] * na; da = d[a] * na; ob = o[b] * nb; db = d[b] * nb;   // grid coordinates, linear in t
            const float pa0 = fminf(fmaxf(oa + da * ta, 0.0f), na), pa1 = fminf(fmaxf(oa + da * tb, 0.0f), na);
            s_lo = max(0, static_cast<int>(floorf(fminf(pa0, pa1))) - 1);
            s_hi = min(dims[a] - 1, static_cast<int>(floorf(fmaxf(pa0, pa1))) + 2);
        }
    }
    const int w_lo = __reduce_min_sync(0xffffffffu, s_lo), w_hi = __reduce_max_sync(0xffffffffu, s_hi);
    for (int s = w_lo; s <= w_hi; ++s) {
        int r_lo = INT_MAX, r_hi = INT_MIN;
        if (s >= s_lo && s <= s_hi) {
            float t0 = ta, t1 = tb;
            if (fabsf(da) > 1e-6f) {   // t-interval in which p_a is within [s - 2, s + 2]
                const float u = (static_cast<float>(s) - 2.0f - oa) / da, v = (static_cast<float>(s) + 2.0f - oa) / da;
                t0 = fmaxf(ta, fminf(u, v));
                t1 = fminf(tb, fmaxf(u, v));
            }
            if (t0 <= t1) {
                const float pb0 = fminf(fmaxf(ob + db * t0, 0.0f), nb), pb1 = fminf(fmaxf(ob + db * t1, 0.0f), nb);
                r_lo = max(0, static_cast<int>(floorf(fminf(pb0, pb1))) - 1);
                r_hi = min(dims[b] - 1, static_cast<int>(floorf(fmaxf(pb0, pb1))) + 2);
            }
        }
        r_lo = __reduce_min_sync(0xffffffffu, r_lo);
        r_hi = __reduce_max_sync(0xffffffffu, r_hi);
        if ((threadIdx.x & 31) == 0 && r_lo != INT_MAX) {
            atomicMin(lo + s, r_lo);
            atomicMax(hi + s, r_hi);
        }
    }
}
}  // namespace

cudaError_t launch_ray_slab_rows(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params, int32_t nx, int32_t ny,
                                 int32_t nz, int slab_axis, int row_axis, int* d_lo, int* d_hi) {
    const uint32_t blocks = tile_blocks(h_params.roi);
    if (blocks == 0) return cudaSuccess;
    ray_slab_rows_kernel<<<blocks, kLeanThreads, 0, stream>>>(d_params, nx, ny, nz, slab_axis, row_axis, d_lo, d_hi);
    return cudaGetLastError();
}

uint32_t lean_block_count(const RoiParams& roi) { return tile_blocks(roi); }

// ---- measurement helpers (bench.py roofline) ----------------------------------------------------------------------
namespace {
// Live samples inside the unit cube: the samples that actually gather 8 corners (forward) and issue 8 reds (backward).
// Re-marches the rays with the forward kernel's own step logic, without touching the grid.
template <bool kStratified>
__global__ void __launch_bounds__(kLeanThreads)
cube_count_kernel(const FrameParams* __restrict__ P, const uint32_t* __restrict__ live_counts, const float4* __restrict__ steps,
                  unsigned long long* __restrict__ total) {
    const CameraParams cam = P->cam;
    const MarchParams mp = P->march;
    const RoiParams roi = P->roi;
    const TilePixel px = tile_pixel(roi);
    const Ray ray = make_ray(cam, roi.x + px.lx, roi.y + px.ly);
    const uint64_t ray_index = mp.ray_index_base + px.ray;
    const uint32_t live = px.inside ? live_counts[px.ray] : 0u;
    float t_in, t_out;
    const WarpRange wr = warp_step_range<true>(mp, ray, px.inside, t_in, t_out);
    uint32_t n = 0;
    for (uint32_t k = wr.lo; k < min(wr.hi, live); ++k) {
        const float4 tab = __ldg(steps + k);
        if (tab.x > t_out || tab.x + mp.dt < t_in) continue;
        const float t = step_time<kStratified>(tab, mp.t_near, mp.t_far, mp.dt, mp.seed, ray_index, k);
        const float x = ray.ox + ray.dx * t, y = ray.oy + ray.dy * t, z = ray.oz + ray.dz * t;
        n += (x < 0.0f || x > 1.0f || y < 0.0f || y > 1.0f || z < 0.0f || z > 1.0f) ? 0u : 1u;
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0 && n != 0) atomicAdd(total, static_cast<unsigned long long>(n));
}

__global__ void touched_voxels_kernel(const float4* __restrict__ grad, size_t voxels, unsigned long long* __restrict__ total) {
    uint32_t n = 0;
    for (size_t v = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; v < voxels;
         v += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float4 g = grad[v];
        n += (g.x != 0.0f || g.y != 0.0f || g.z != 0.0f || g.w != 0.0f) ? 1u : 0u;
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0 && n != 0) atomicAdd(total, static_cast<unsigned long long>(n));
}
}  // namespace

cudaError_t launch_cube_count(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params,
                              const LeanBuffers& buf, unsigned long long* d_total) {
    const uint32_t blocks = tile_blocks(h_params.roi);
    if (blocks == 0) return cudaSuccess;
    if (h_params.march.stratified != 0)
        cube_count_kernel<true><<<blocks, kLeanThreads, 0, stream>>>(d_params, buf.live, buf.steps, d_total);
    else
        cube_count_kernel<false><<<blocks, kLeanThreads, 0, stream>>>(d_params, buf.live, buf.steps, d_total);
    return cudaGetLastError();
}

cudaError_t launch_touched_voxels(cudaStream_t stream, const float4* d_grad, size_t voxels, unsigned long long* d_total) {
    if (voxels == 0) return cudaSuccess;
    touched_voxels_kernel<<<148 * 8, 256, 0, stream>>>(d_grad, voxels, d_total);
    return cudaGetLastError();
}

// ---- empty-space skipping: occupancy bits ---------------------------------------------------------------------------
namespace {
// One warp per brick of 8^3 cells: ORs over the 9^3 voxels its cells can read (x0 .. x0 + 1 for x0 in the brick).
template <class V>
__global__ void __launch_bounds__(256)
occupancy_build_kernel(const V* __restrict__ values, int32_t nx, int32_t ny, int32_t nz, uint32_t* __restrict__ occ,
                       int32_t obx, int32_t oby, int32_t obz, unsigned int* __restrict__ counts) {
    const uint32_t lane = threadIdx.x & 31u;
    const size_t bricks = static_cast<size_t>(obx) * oby * obz;
    for (size_t b = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5; b < bricks;
         b += (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5) {
        const int32_t bx = static_cast<int32_t>(b % obx), by = static_cast<int32_t>((b / obx) % oby), bz = static_cast<int32_t>(b / (static_cast<size_t>(obx) * oby));
        const int32_t x0 = bx * 8, y0 = by * 8, z0 = bz * 8;
        const int32_t ex = min(9, nx - x0), ey = min(9, ny - y0), ez = min(9, nz - z0);
        uint32_t bits = 0;
        for (int32_t i = lane; i < ex * ey * ez; i += 32) {
            const int32_t x = i % ex, y = (i / ex) % ey, z = i / (ex * ey);
            const float4 v = load_voxel(values, voxel_index32(x0 + x, y0 + y, z0 + z, nx, ny));
            if (v.w != 0.0f) bits |= 3u;
            else if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f) bits |= 2u;
        }
        bits = __reduce_or_sync(0xffffffffu, bits);
        if (lane == 0) {
            if (bits != 0u) atomicOr(occ + (b >> 4), bits << ((b & 15u) * 2u));
            if ((bits & 1u) == 0u) atomicAdd(counts + 0, 1u);
            if ((bits & 2u) == 0u) atomicAdd(counts + 1, 1u);
        }
    }
}
}  // namespace

cudaError_t launch_build_occupancy(cudaStream_t stream, const float4* values, int32_t nx, int32_t ny, int32_t nz, uint32_t* d_occ,
                                   size_t occ_words, unsigned int* d_counts, bool values_are_half) {
    const int32_t obx = (nx + 7) / 8, oby = (ny + 7) / 8, obz = (nz + 7) / 8;
    cudaError_t e = cudaMemsetAsync(d_occ, 0, occ_words * sizeof(uint32_t), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    const size_t bricks = static_cast<size_t>(obx) * oby * obz;
    const unsigned blocks = static_cast<unsigned>(std::min<size_t>((bricks + 7) / 8, 148 * 16));
    if (values_are_half)
        occupancy_build_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const HalfVoxel*>(values), nx, ny, nz, d_occ, obx, oby, obz, d_counts);
    else
        occupancy_build_kernel<<<blocks, 256, 0, stream>>>(values, nx, ny, nz, d_occ, obx, oby, obz, d_counts);
    return cudaGetLastError();
}

// ---- half storage ------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ HalfVoxel to_half_voxel(float4 v) {
    HalfVoxel h;
    const __half2 rg = __floats2half2_rn(v.x, v.y), bs = __floats2half2_rn(v.z, v.w);
    h.bits.x = *reinterpret_cast<const uint32_t*>(&rg);
    h.bits.y = *reinterpret_cast<const uint32_t*>(&bs);
    return h;
}

__global__ void convert_storage_kernel(float4* __restrict__ f32, HalfVoxel* __restrict__ half, size_t voxels, bool to_half) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < voxels; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        if (to_half) half[i] = to_half_voxel(f32[i]);
        else f32[i] = load_voxel(half, static_cast<uint32_t>(i));
    }
}

// hpx_grid_update on a half grid: channels the caller does not pass keep their stored values
__global__ void pack_grid_half_kernel(const float* __restrict__ sigma, const float* __restrict__ color, HalfVoxel* __restrict__ half, size_t voxels) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < voxels; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float4 v = (sigma == nullptr || color == nullptr) ? load_voxel(half, static_cast<uint32_t>(i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (color != nullptr) { v.x = color[3 * i]; v.y = color[3 * i + 1]; v.z = color[3 * i + 2]; }
        if (sigma != nullptr) v.w = sigma[i];
        half[i] = to_half_voxel(v);
    }
}

__global__ void abs_max_half_kernel(const HalfVoxel* __restrict__ half, size_t voxels, uint32_t* __restrict__ out_bits) {
    float m = 0.0f;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < voxels; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float4 v = load_voxel(half, static_cast<uint32_t>(i));
        const float a = fmaxf(fabsf(v.x), fmaxf(fabsf(v.y), fabsf(v.z)));
        if (a < CUDART_INF_F) m = fmaxf(m, a);
    }
    const uint32_t bits = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, bits);
}
}  // namespace

cudaError_t launch_convert_storage(cudaStream_t stream, const float4* f32, void* half, size_t voxels, bool to_half) {
    if (voxels == 0) return cudaSuccess;
    convert_storage_kernel<<<148 * 8, 256, 0, stream>>>(const_cast<float4*>(f32), static_cast<HalfVoxel*>(half), voxels, to_half);
    return cudaGetLastError();
}

cudaError_t launch_pack_grid_half(cudaStream_t stream, const float* sigma, const float* color, void* half, size_t voxels) {
    if (voxels == 0) return cudaSuccess;
    pack_grid_half_kernel<<<148 * 8, 256, 0, stream>>>(sigma, color, static_cast<HalfVoxel*>(half), voxels);
    return cudaGetLastError();
}

cudaError_t launch_abs_max_half(cudaStream_t stream, const void* half, size_t voxels, uint32_t* d_out_bits) {
    if (voxels == 0) return cudaSuccess;
    abs_max_half_kernel<<<148 * 8, 256, 0, stream>>>(static_cast<const HalfVoxel*>(half), voxels, d_out_bits);
    return cudaGetLastError();
}

namespace {
__global__ void upload_params_kernel(FrameParams* dst, const FrameParams src) { *dst = src; }
}  // namespace

cudaError_t launch_upload_params(cudaStream_t stream, FrameParams* d_params, const FrameParams& h_params) {
    upload_params_kernel<<<1, 1, 0, stream>>>(d_params, h_params);
    return cudaGetLastError();
}

// Host restatement of the ray-independent part of the marching loop (samp_cpu.cpp:227-241) and of the depth
// cursor (int_cpu.cpp:170,211).  volatile forces every intermediate to be rounded to float; the host objects are
// built with -ffp-contract=off, so the values equal what the kernels used to compute per step.
void build_step_table(const MarchParams& mp, float4* table) {
    const float tn = mp.t_near, tf = mp.t_far, dts = mp.dt;
    volatile float cursor = tn;
    for (uint32_t k = 0; k < mp.uniform_count; ++k) {
        volatile float prod = static_cast<float>(k) * dts;
        volatile float base = tn + prod;
        volatile float half = 0.5f * dts;
        volatile float mid = base + half;
        if (mid >= tf) mid = nextafterf(tf, tn);
        volatile float end = base + dts;
        if (tf < end) end = tf;
        volatile float dta = end - base;
        table[k] = make_float4(base, mid, dta, cursor);
        cursor = cursor + dta;
    }
}

#define DV_DISPATCH3(FN, lin, clampo, strat, ...)                                             \
    do {                                                                                      \
        if (lin) {                                                                            \
            if (clampo) { if (strat) FN<true, true, true> __VA_ARGS__; else FN<true, true, false> __VA_ARGS__; }       \
            else        { if (strat) FN<true, false, true> __VA_ARGS__; else FN<true, false, false> __VA_ARGS__; }     \
        } else {                                                                              \
            if (clampo) { if (strat) FN<false, true, true> __VA_ARGS__; else FN<false, true, false> __VA_ARGS__; }     \
            else        { if (strat) FN<false, false, true> __VA_ARGS__; else FN<false, false, false> __VA_ARGS__; }   \
        }                                                                                     \
    } while (0)

cudaError_t launch_lean_forward(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params,
                                const PackedGrid& grid, const LeanBuffers& out, bool fill_background) {
    const RoiParams& roi = h_params.roi;
    if (fill_background) {
        background_kernel<<<148 * 4, 256, 0, stream>>>(out, roi, h_params.march.t_far);
    }
    const uint32_t blocks = tile_blocks(roi);
    if (blocks == 0) return cudaGetLastError();
    const bool strat = h_params.march.stratified != 0;
    if (grid.half_values != nullptr) {   // half storage: linear OOB-zero grids (hpx_grid_set_storage checks)
        const HalfVoxel* hv = static_cast<const HalfVoxel*>(grid.half_values);
        const bool occ = grid.occ != nullptr;
#define DV_FWD_HALF(S, O) lean_forward_kernel<true, false, S, O, HalfVoxel><<<blocks, kLeanThreads, 0, stream>>>(      \
        d_params, hv, grid.nx, grid.ny, grid.nz, out, grid.occ, grid.obx, grid.oby)
        if (strat) { if (occ) DV_FWD_HALF(true, true); else DV_FWD_HALF(true, false); }
        else       { if (occ) DV_FWD_HALF(false, true); else DV_FWD_HALF(false, false); }
#undef DV_FWD_HALF
        return cudaGetLastError();
    }
    if (grid.occ != nullptr && grid.linear && !grid.clamp) {   // empty-space skipping (linear, OOB-zero fields)
        if (strat) lean_forward_kernel<true, false, true, true><<<blocks, kLeanThreads, 0, stream>>>(
                       d_params, grid.values, grid.nx, grid.ny, grid.nz, out, grid.occ, grid.obx, grid.oby);
        else       lean_forward_kernel<true, false, false, true><<<blocks, kLeanThreads, 0, stream>>>(
                       d_params, grid.values, grid.nx, grid.ny, grid.nz, out, grid.occ, grid.obx, grid.oby);
        return cudaGetLastError();
    }
    DV_DISPATCH3(lean_forward_kernel, grid.linear, grid.clamp, strat,
                 <<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, out));
    return cudaGetLastError();
}

// Pixel spacing, in voxels, between neighbouring rays where they cross the cube centre.  Measured crossover
// (tools/heuristic_time.py, profiles/README.md): merged wins by 25-35 % at 0.3-0.5 voxel, 7-19 % at 0.6-0.83,
// loses 3-8 % at 0.93 and 12-27 % at 1.25.
static bool merge_scatter_pays(const FrameParams& h, const PackedGrid& grid, const ScatterParams& sp) {
    if (!grid.linear || sp.nearest) return false;
    if (sp.nx < 2 || sp.ny < 2 || sp.nz < 2 || sp.nx > 1024 || sp.ny > 1024 || sp.nz > 1024) return false;   // 10-bit cell keys
    if (h.cam.ortho) return true;   // the reference's orthographic rays all share one origin and direction (ray_cpu.cpp:189-199)
    const float dx = 0.5f - h.cam.ox, dy = 0.5f - h.cam.oy, dz = 0.5f - h.cam.oz;
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    const float f = fminf(fabsf(h.cam.fx), fabsf(h.cam.fy));
    if (!(f > 0.0f)) return false;
    const float n = static_cast<float>(max(sp.nx, max(sp.ny, sp.nz)) - 1);
    return dist / f * n < 0.9f;
}

int resolve_scatter_mode(const FrameParams& h_params, const PackedGrid& grid, const ScatterParams& sp, int scatter_mode) {
    const bool can_merge = grid.linear && !sp.nearest && sp.nx >= 2 && sp.ny >= 2 && sp.nz >= 2 && sp.nx <= 1024 &&
                           sp.ny <= 1024 && sp.nz <= 1024;
    const bool merge = scatter_mode == kScatterMerge ? can_merge
                     : scatter_mode == kScatterPerRay ? false : merge_scatter_pays(h_params, grid, sp);
    return merge ? kScatterMerge : kScatterPerRay;
}

cudaError_t launch_lean_backward(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params,
                                 const PackedGrid& grid, const ScatterParams& sp, const float* d_dL_dI,
                                 const LeanBuffers& state, int scatter_mode, double* cam_partials, float* d_cam16) {
    const uint32_t blocks = tile_blocks(h_params.roi);
    if (blocks == 0) return cudaSuccess;
    const bool strat = h_params.march.stratified != 0;
    const bool merge = resolve_scatter_mode(h_params, grid, sp, scatter_mode) == kScatterMerge;
    if (merge) {
        // plain = float reds into the grid's own block in default order: the specialised flush
        const bool plain = sp.fixed == nullptr && sp.boxed == 0u && sp.box_sx == 1u && sp.box_sy == static_cast<uint32_t>(sp.nx) &&
                           sp.box_sz == static_cast<uint32_t>(sp.nx) * static_cast<uint32_t>(sp.ny);
#define DV_MERGE2(C, S, U, K)                                                                                          \
    do {                                                                                                               \
        if (plain)                                                                                                     \
            lean_backward_merge_kernel<C, S, U, K, true><<<blocks, kLeanThreads, 0, stream>>>(                         \
                d_params, grid.values, grid.nx, grid.ny, grid.nz, sp, d_dL_dI, state, cam_partials);                   \
        else                                                                                                           \
            lean_backward_merge_kernel<C, S, U, K, false><<<blocks, kLeanThreads, 0, stream>>>(                        \
                d_params, grid.values, grid.nx, grid.ny, grid.nz, sp, d_dL_dI, state, cam_partials);                   \
    } while (0)
#define DV_MERGE(C, S, U)                                                                                              \
    do {                                                                                                               \
        if (cam_partials != nullptr) DV_MERGE2(C, S, U, true);                                                         \
        else DV_MERGE2(C, S, U, false);                                                                                \
    } while (0)
        const bool unit = sp.unit_bbox != 0;
        if (grid.half_values != nullptr) {
            if (grid.clamp || !unit) return cudaErrorInvalidValue;   // hpx_grid_set_storage refuses such grids
            const HalfVoxel* hv = static_cast<const HalfVoxel*>(grid.half_values);
            const bool occ = grid.occ != nullptr, cam = cam_partials != nullptr;
#define DV_MERGE_HALF(S, K, PL, O)                                                                                     \
    lean_backward_merge_kernel<false, S, true, K, PL, O, HalfVoxel><<<blocks, kLeanThreads, 0, stream>>>(             \
        d_params, hv, grid.nx, grid.ny, grid.nz, sp, d_dL_dI, state, cam_partials, grid.occ, grid.obx, grid.oby)
#define DV_MERGE_HALF3(S, K, PL) do { if (occ) DV_MERGE_HALF(S, K, PL, true); else DV_MERGE_HALF(S, K, PL, false); } while (0)
#define DV_MERGE_HALF2(S, K) do { if (plain) DV_MERGE_HALF3(S, K, true); else DV_MERGE_HALF3(S, K, false); } while (0)
            if (strat) { if (cam) DV_MERGE_HALF2(true, true); else DV_MERGE_HALF2(true, false); }
            else       { if (cam) DV_MERGE_HALF2(false, true); else DV_MERGE_HALF2(false, false); }
#undef DV_MERGE_HALF2
#undef DV_MERGE_HALF3
#undef DV_MERGE_HALF
        } else if (grid.occ != nullptr && !grid.clamp && unit) {   // empty-space skipping
#define DV_MERGE_OCC(S, K)                                                                                             \
    do {                                                                                                               \
        if (plain)                                                                                                     \
            lean_backward_merge_kernel<false, S, true, K, true, true><<<blocks, kLeanThreads, 0, stream>>>(            \
                d_params, grid.values, grid.nx, grid.ny, grid.nz, sp, d_dL_dI, state, cam_partials, grid.occ, grid.obx, grid.oby); \
        else                                                                                                           \
            lean_backward_merge_kernel<false, S, true, K, false, true><<<blocks, kLeanThreads, 0, stream>>>(           \
                d_params, grid.values, grid.nx, grid.ny, grid.nz, sp, d_dL_dI, state, cam_partials, grid.occ, grid.obx, grid.oby); \
    } while (0)
            if (strat) { if (cam_partials != nullptr) DV_MERGE_OCC(true, true); else DV_MERGE_OCC(true, false); }
            else       { if (cam_partials != nullptr) DV_MERGE_OCC(false, true); else DV_MERGE_OCC(false, false); }
#undef DV_MERGE_OCC
        } else if (grid.clamp) {
            if (strat) { if (unit) DV_MERGE(true, true, true); else DV_MERGE(true, true, false); }
            else       { if (unit) DV_MERGE(true, false, true); else DV_MERGE(true, false, false); }
        } else {
            if (strat) { if (unit) DV_MERGE(false, true, true); else DV_MERGE(false, true, false); }
            else       { if (unit) DV_MERGE(false, false, true); else DV_MERGE(false, false, false); }
        }
#undef DV_MERGE
#undef DV_MERGE2
        if (cam_partials != nullptr) camera_reduce_kernel<<<1, 512, 0, stream>>>(cam_partials, blocks, d_cam16);
        return cudaGetLastError();
    }
    if (cam_partials != nullptr) return cudaErrorInvalidValue;   // the fused camera adjoint exists in the merged kernel only
    if (grid.half_values != nullptr) {
        const HalfVoxel* hv = static_cast<const HalfVoxel*>(grid.half_values);
        const bool occ = grid.occ != nullptr;
#define DV_BWD_HALF(S, O) lean_backward_kernel<true, false, S, O, HalfVoxel><<<blocks, kLeanThreads, 0, stream>>>(     \
        d_params, hv, grid.nx, grid.ny, grid.nz, sp, d_dL_dI, state, grid.occ, grid.obx, grid.oby)
        if (strat) { if (occ) DV_BWD_HALF(true, true); else DV_BWD_HALF(true, false); }
        else       { if (occ) DV_BWD_HALF(false, true); else DV_BWD_HALF(false, false); }
#undef DV_BWD_HALF
        return cudaGetLastError();
    }
    if (grid.occ != nullptr && grid.linear && !grid.clamp) {
        if (strat) lean_backward_kernel<true, false, true, true><<<blocks, kLeanThreads, 0, stream>>>(
                       d_params, grid.values, grid.nx, grid.ny, grid.nz, sp, d_dL_dI, state, grid.occ, grid.obx, grid.oby);
        else       lean_backward_kernel<true, false, false, true><<<blocks, kLeanThreads, 0, stream>>>(
                       d_params, grid.values, grid.nx, grid.ny, grid.nz, sp, d_dL_dI, state, grid.occ, grid.obx, grid.oby);
        return cudaGetLastError();
    }
    DV_DISPATCH3(lean_backward_kernel, grid.linear, grid.clamp, strat,
                 <<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, sp, d_dL_dI,
                                                      state));
    return cudaGetLastError();
}

cudaError_t launch_camera_adjoint(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params,
                                  const PackedGrid& grid, const float* d_dL_dI, const uint32_t* d_live,
                                  const float4* d_steps, double* d_partials, float* d_cam16) {
    const uint32_t blocks = tile_blocks(h_params.roi);
    if (blocks == 0 || !grid.linear) return cudaSuccess;  // nearest-neighbour fields have zero spatial gradient
    const bool strat = h_params.march.stratified != 0;
    if (grid.half_values != nullptr) {
        const HalfVoxel* hv = static_cast<const HalfVoxel*>(grid.half_values);
        if (strat) camera_adjoint_kernel<false, true, HalfVoxel><<<blocks, kLeanThreads, 0, stream>>>(d_params, hv, grid.nx, grid.ny, grid.nz, d_dL_dI, d_live, d_steps, d_partials);
        else       camera_adjoint_kernel<false, false, HalfVoxel><<<blocks, kLeanThreads, 0, stream>>>(d_params, hv, grid.nx, grid.ny, grid.nz, d_dL_dI, d_live, d_steps, d_partials);
        camera_reduce_kernel<<<1, 512, 0, stream>>>(d_partials, blocks, d_cam16);
        return cudaGetLastError();
    }
    if (grid.clamp) {
        if (strat) camera_adjoint_kernel<true, true><<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, d_dL_dI, d_live, d_steps, d_partials);
        else       camera_adjoint_kernel<true, false><<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, d_dL_dI, d_live, d_steps, d_partials);
    } else {
        if (strat) camera_adjoint_kernel<false, true><<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, d_dL_dI, d_live, d_steps, d_partials);
        else       camera_adjoint_kernel<false, false><<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, d_dL_dI, d_live, d_steps, d_partials);
    }
    camera_reduce_kernel<<<1, 512, 0, stream>>>(d_partials, blocks, d_cam16);
    return cudaGetLastError();
}

}  // namespace dv
