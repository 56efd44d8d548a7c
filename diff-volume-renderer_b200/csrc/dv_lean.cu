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
    const uint32_t tile_x = blockIdx.x % tiles_x;
    const uint32_t tile_y = blockIdx.x / tiles_x;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TilePixel p;
    p.lx = tile_x * (kTileW * kWarpsX) + (warp % kWarpsX) * kTileW + (lane % kTileW);
    p.ly = tile_y * (kTileH * kWarpsY) + (warp / kWarpsX) * kTileH + (lane / kTileW);
    p.inside = p.lx < roi.w && p.ly < roi.h;
    p.ray = p.ly * roi.w + p.lx;
    return p;
}

template <bool kLinear, bool kClamp, bool kStratified>
__global__ void __launch_bounds__(kLeanThreads)
lean_forward_kernel(const FrameParams* __restrict__ P, const float4* __restrict__ grid, int32_t nx, int32_t ny,
                    int32_t nz, LeanBuffers out) {
    const CameraParams cam = P->cam;
    const MarchParams mp = P->march;
    const RoiParams roi = P->roi;
    const TilePixel px = tile_pixel(roi);

    const uint32_t gx = roi.x + px.lx, gy = roi.y + px.ly;
    const Ray ray = make_ray(cam, gx, gy);
    const uint64_t ray_index = mp.ray_index_base + px.ray;

    RayAccum acc;
    acc.t_cursor = mp.t_near;
    bool alive = px.inside;
    uint32_t live = 0;
    const uint32_t count = mp.uniform_count;
    float t_in = -CUDART_INF_F, t_out = CUDART_INF_F;
    if (!kClamp) cube_interval(ray, t_in, t_out);

    for (uint32_t step = 0; step < count; ++step) {
        if (!__any_sync(0xffffffffu, alive)) break;
        if ((step % kSegment) == 0 && alive && out.ckpt != nullptr) {
            out.ckpt[static_cast<size_t>(step / kSegment) * out.ckpt_stride + px.ray] = acc.T;
        }
        if (alive) {
            ++live;
            const float base = mp.t_near + static_cast<float>(step) * mp.dt;
            if (!kClamp && (base > t_out || base + mp.dt < t_in)) {
                // whole step outside the cube: sigma = 0, only the depth cursor moves (same dt arithmetic)
                acc.t_cursor += fminf(base + mp.dt, mp.t_far) - base;
                continue;
            }
            float t, dtv;
            march_step<kStratified>(mp.t_near, mp.t_far, mp.dt, mp.seed, ray_index, step, t, dtv);
            const float pxw = ray.ox + ray.dx * t;
            const float pyw = ray.oy + ray.dy * t;
            const float pzw = ray.oz + ray.dz * t;
            const float4 v = sample_packed<kLinear, kClamp, false>(grid, nx, ny, nz, pxw, pyw, pzw);
            float a, w, tb;
            if (integrate_sample<false>(acc, dtv, v, a, w, tb)) alive = false;
        }
    }

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
        uint32_t s = live;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0 && s != 0) atomicAdd(out.live_total, static_cast<unsigned long long>(s));
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

template <bool kLinear, bool kClamp, bool kStratified>
__global__ void __launch_bounds__(kLeanThreads)
lean_backward_kernel(const FrameParams* __restrict__ P, const float4* __restrict__ grid, int32_t nx, int32_t ny,
                     int32_t nz, ScatterParams sp, const float* __restrict__ dL_dI, LeanBuffers st) {
    __shared__ SegmentStash stash;
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

    float t_in = -CUDART_INF_F, t_out = CUDART_INF_F;
    const bool skippable = !kClamp && sp.unit_bbox != 0u;   // outside samples touch neither T nor the grid
    if (skippable) cube_interval(ray, t_in, t_out);

    float adj_T = 0.0f;
    // all lanes of a warp walk the same segment index in the same iteration: the warp's gathers and
    // scatters of one iteration stay inside one thin slab of the grid
    for (uint32_t seg = warp_nseg; seg-- > 0;) {
        if (seg >= nseg) continue;
        const uint32_t first = seg * kSegment;
        const uint32_t count = min(static_cast<uint32_t>(kSegment), live - first);
        float T = st.ckpt[static_cast<size_t>(seg) * st.ckpt_stride + px.ray];
        // forward part: recompute the segment from its transmittance checkpoint
#pragma unroll 1
        for (uint32_t j = 0; j < count; ++j) {
            const float base = mp.t_near + static_cast<float>(first + j) * mp.dt;
            if (skippable && (base > t_out || base + mp.dt < t_in)) {
                stash.alpha[j][tid] = -1.0f;   // marker: no contribution, adj_T unchanged
                continue;
            }
            float t, dtv;
            march_step<kStratified>(mp.t_near, mp.t_far, mp.dt, mp.seed, ray_index, first + j, t, dtv);
            const float4 v = sample_packed<kLinear, kClamp, false>(grid, nx, ny, nz, ray.ox + ray.dx * t,
                                                                   ray.oy + ray.dy * t, ray.oz + ray.dz * t);
            const float a = alpha_of(v.w, dtv);
            stash.alpha[j][tid] = a;
            stash.T_prev[j][tid] = T;
            stash.dot[j][tid] = g0 * v.x + g1 * v.y + g2 * v.z;
            stash.t[j][tid] = t;
            stash.dt[j][tid] = dtv;
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
                           make_float4(g0 * w, g1 * w, g2 * w, dsigma));
        }
    }
}

// ---- camera adjoint ---------------------------------------------------------
// value is not needed, only d/d(position) of sigma and of h = g . rgb
template <bool kClamp>
__device__ __forceinline__ void field_gradients(const float4* __restrict__ g, int32_t nx, int32_t ny, int32_t nz,
                                                float px, float py, float pz, float g0, float g1, float g2,
                                                float4& value, float grad_sigma[3], float grad_h[3]) {
    grad_sigma[0] = grad_sigma[1] = grad_sigma[2] = 0.f;
    grad_h[0] = grad_h[1] = grad_h[2] = 0.f;
    value = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool ox = px < 0.0f || px > 1.0f, oy = py < 0.0f || py > 1.0f, oz = pz < 0.0f || pz > 1.0f;
    float fx, fy, fz;
    if (!grid_coords(px, py, pz, kClamp, nx, ny, nz, fx, fy, fz)) return;
    const Cell c = make_cell(fx, fy, fz, nx, ny, nz);
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int32_t x = (i & 1) ? c.x1 : c.x0, y = (i & 2) ? c.y1 : c.y0, z = (i & 4) ? c.z1 : c.z0;
        v[i] = __ldg(g + voxel_index(x, y, z, nx, ny));
    }
    value.x = trilerp(v[0].x, v[1].x, v[2].x, v[3].x, v[4].x, v[5].x, v[6].x, v[7].x, c.tx, c.ty, c.tz);
    value.y = trilerp(v[0].y, v[1].y, v[2].y, v[3].y, v[4].y, v[5].y, v[6].y, v[7].y, c.tx, c.ty, c.tz);
    value.z = trilerp(v[0].z, v[1].z, v[2].z, v[3].z, v[4].z, v[5].z, v[6].z, v[7].z, c.tx, c.ty, c.tz);
    value.w = trilerp(v[0].w, v[1].w, v[2].w, v[3].w, v[4].w, v[5].w, v[6].w, v[7].w, c.tx, c.ty, c.tz);
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
    const float sx = (kClamp && ox) ? 0.f : static_cast<float>(nx - 1);
    const float sy = (kClamp && oy) ? 0.f : static_cast<float>(ny - 1);
    const float sz = (kClamp && oz) ? 0.f : static_cast<float>(nz - 1);
    grad_sigma[0] *= sx; grad_sigma[1] *= sy; grad_sigma[2] *= sz;
    grad_h[0] *= sx; grad_h[1] *= sy; grad_h[2] *= sz;
}

template <bool kClamp, bool kStratified>
__global__ void __launch_bounds__(kLeanThreads)
camera_adjoint_kernel(const FrameParams* __restrict__ P, const float4* __restrict__ grid, int32_t nx, int32_t ny,
                      int32_t nz, const float* __restrict__ dL_dI, const uint32_t* __restrict__ live_counts,
                      double* __restrict__ partials) {
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
    for (uint32_t k = 0; k < live; ++k) {
        float t, dtv;
        march_step<kStratified>(mp.t_near, mp.t_far, mp.dt, mp.seed, ray_index, k, t, dtv);
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
    // chain rule: d = v / |v|, v = R q   
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
        if (!px.inside) {
            for (int i = 0; i < 16; ++i) out[i] = 0.0;
        }
    }
    // deterministic block reduction: warp shuffles, then warp 0 sums the warps in order
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
    return tx * ty;
}

}  // namespace

uint32_t lean_block_count(const RoiParams& roi) { return tile_blocks(roi); }

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
    DV_DISPATCH3(lean_forward_kernel, grid.linear, grid.clamp, strat,
                 <<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, out));
    return cudaGetLastError();
}

cudaError_t launch_lean_backward(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params,
                                 const PackedGrid& grid, const ScatterParams& sp, const float* d_dL_dI,
                                 const LeanBuffers& state) {
    const uint32_t blocks = tile_blocks(h_params.roi);
    if (blocks == 0) return cudaSuccess;
    const bool strat = h_params.march.stratified != 0;
    DV_DISPATCH3(lean_backward_kernel, grid.linear, grid.clamp, strat,
                 <<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, sp, d_dL_dI,
                                                      state));
    return cudaGetLastError();
}

cudaError_t launch_camera_adjoint(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params,
                                  const PackedGrid& grid, const float* d_dL_dI, const uint32_t* d_live,
                                  double* d_partials, float* d_cam16) {
    const uint32_t blocks = tile_blocks(h_params.roi);
    if (blocks == 0 || !grid.linear) return cudaSuccess;  // nearest-neighbour fields have zero spatial gradient
    const bool strat = h_params.march.stratified != 0;
    if (grid.clamp) {
        if (strat) camera_adjoint_kernel<true, true><<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, d_dL_dI, d_live, d_partials);
        else       camera_adjoint_kernel<true, false><<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, d_dL_dI, d_live, d_partials);
    } else {
        if (strat) camera_adjoint_kernel<false, true><<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, d_dL_dI, d_live, d_partials);
        else       camera_adjoint_kernel<false, false><<<blocks, kLeanThreads, 0, stream>>>(d_params, grid.values, grid.nx, grid.ny, grid.nz, d_dL_dI, d_live, d_partials);
    }
    camera_reduce_kernel<<<1, 512, 0, stream>>>(d_partials, blocks, d_cam16);
    return cudaGetLastError();
}

}  // namespace dv
