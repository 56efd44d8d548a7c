// dv_staged.cu -- materialising kernels behind the hp.h entry points.
//
// These honour the reference's buffer contracts (per-sample positions / dt /
// sigma / colour, prefix offsets, aux rows) so existing callers of hp_ray,
// hp_samp, hp_int, hp_samp_int_fused, hp_diff and hp_img keep working on
// DEVICE tensors.  They share every arithmetic routine with the lean kernels
// (dv_device.cuh), so staged == fused bit for bit (reference test
// hotpath/tests/runner/hp_runner.cpp:1737-1760).
//
// Compiled with -fmad=false: see dv_device.cuh.
#include "dv_staged.h"

#include <cstdlib>

#include "dv_device.cuh"

namespace dv {

namespace {

constexpr int kThreads = 128;

inline uint32_t blocks_for(size_t n, int threads) { return static_cast<uint32_t>((n + threads - 1) / threads); }

// ---- K1: rays (reference hotpath/src/cpu/ray_cpu.cpp:183-226) ---------------
__global__ void rays_kernel(FrameParams p, RayArrays out, uint32_t n_rays) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const uint32_t lx = i % p.roi.w, ly = i / p.roi.w;
    const uint32_t px = p.roi.x + lx, py = p.roi.y + ly;
    const Ray r = make_ray(p.cam, px, py);
    out.origins[3 * i + 0] = r.ox; out.origins[3 * i + 1] = r.oy; out.origins[3 * i + 2] = r.oz;
    out.directions[3 * i + 0] = r.dx; out.directions[3 * i + 1] = r.dy; out.directions[3 * i + 2] = r.dz;
    out.t_near[i] = p.march.t_near;
    out.t_far[i] = p.march.t_far;
    out.pixel_ids[i] = py * p.roi.img_w + px;
}

// ---- sample counts + exclusive scan ----------------------------------------
__device__ __forceinline__ uint32_t ray_sample_count(const MarchParams& mp, float tn, float tf) {
    if (!(tf > tn)) return 0;
    uint32_t n = 0;
    for (uint32_t step = 0; step < mp.max_steps; ++step) {
        float t, dtv;
        const int r = march_step<false>(tn, tf, mp.dt, 0, 0, step, t, dtv);
        if (r == 2) break;
        if (r == 0) ++n;
    }
    return n;
}

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void count_kernel(MarchParams mp, RayArrays rays, uint32_t n_rays, uint32_t* counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    counts[i] = ray_sample_count(mp, rays.t_near[i], rays.t_far[i]);
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t* counts, uint32_t n,
                                                                      unsigned long long* tile_sums) {
    __shared__ unsigned long long warp_sums[kScanThreads / 32];
    const size_t base = static_cast<size_t>(blockIdx.x) * kScanTile;
    unsigned long long s = 0;
    for (int k = 0; k < kScanItems; ++k) {
        const size_t i = base + static_cast<size_t>(k) * kScanThreads + threadIdx.x;
        if (i < n) s += counts[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += warp_sums[w];
        tile_sums[blockIdx.x] = t;
    }
}

// one block: exclusive scan of the tile sums in place, grand total to *total
__global__ void __launch_bounds__(1024) scan_spine_kernel(unsigned long long* tile_sums, uint32_t tiles,
                                                          unsigned long long* total) {
    __shared__ unsigned long long warp_part[32];
    __shared__ unsigned long long carry, chunk_total;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < tiles ? tile_sums[i] : 0ULL;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) warp_part[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = warp_part[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += up;
            }
            warp_part[lane] = wi - w;  // exclusive offset of each warp inside the chunk
            if (lane == 31) chunk_total = wi;
        }
        __syncthreads();
        if (i < tiles) tile_sums[i] = carry + warp_part[warp] + (incl - v);
        __syncthreads();
        if (threadIdx.x == 0) carry += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* counts, uint32_t n,
                                                                  const unsigned long long* tile_offsets,
                                                                  const unsigned long long* total,
                                                                  uint32_t* ray_offset) {
    // blocked arrangement: thread t owns items [t*4, t*4+4) of the tile
    __shared__ uint32_t warp_tot[kScanThreads / 32];
    const size_t base = static_cast<size_t>(blockIdx.x) * kScanTile + static_cast<size_t>(threadIdx.x) * kScanItems;
    uint32_t v[kScanItems];
    uint32_t local = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? counts[base + k] : 0u;
        local += v[k];
    }
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += up;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t warp_base = 0;
    for (int w = 0; w < static_cast<int>(threadIdx.x >> 5); ++w) warp_base += warp_tot[w];
    // the reference keeps offsets in u32 (static_cast<uint32_t>(total_samples), samp_cpu.cpp:208)
    uint32_t run = static_cast<uint32_t>(tile_offsets[blockIdx.x]) + warp_base + (incl - local);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) ray_offset[base + k] = run;
        run += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ray_offset[n] = static_cast<uint32_t>(*total);
}

// offsets of a plan whose rays all emit the same number of samples (graph path: no host read-back)
__global__ void uniform_offsets_kernel(uint32_t* ray_offset, uint32_t n_rays, uint32_t per_ray) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i <= n_rays) ray_offset[i] = static_cast<uint32_t>(i * per_ray);
}

// ---- K2: sampler, optionally with the integrator inline --------------------
template <bool kStratified, bool kIntegrate>
__global__ void __launch_bounds__(kThreads)
sample_kernel(MarchParams mp, float plan_t_near, float plan_t_far, FieldPair fields, RayArrays rays, uint32_t n_rays,
              SampleArrays samp, IntegralArrays intl, bool aux_aligned) {
    const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const float ox = rays.origins[3 * ray], oy = rays.origins[3 * ray + 1], oz = rays.origins[3 * ray + 2];
    const float dx = rays.directions[3 * ray], dy = rays.directions[3 * ray + 1], dz = rays.directions[3 * ray + 2];
    const float tn = rays.t_near[ray], tf = rays.t_far[ray];
    const uint64_t ray_index = mp.ray_index_base + ray;
    size_t idx = samp.ray_offset[ray];
    RayAccum acc;
    acc.t_cursor = plan_t_near;
    bool stopped = false;
    if (tf > tn) {
        for (uint32_t step = 0; step < mp.max_steps; ++step) {
            float t, dtv;
            const int r = march_step<kStratified>(tn, tf, mp.dt, mp.seed, ray_index, step, t, dtv);
            if (r == 2) break;
            if (r == 1) continue;
            const float px = ox + dx * t, py = oy + dy * t, pz = oz + dz * t;
            const float4 v = sample_fields(fields, px, py, pz);
            samp.positions[3 * idx] = px; samp.positions[3 * idx + 1] = py; samp.positions[3 * idx + 2] = pz;
            samp.dt[idx] = dtv;
            samp.sigma[idx] = v.w;
            samp.color[3 * idx] = v.x; samp.color[3 * idx + 1] = v.y; samp.color[3 * idx + 2] = v.z;
            if (kIntegrate) {
                float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!stopped) {
                    float a, w, tb;
                    stopped = integrate_sample(acc, dtv, v, a, w, tb);
                    row = make_float4(a, w, tb, logf(fmaxf(tb, 1e-30f)));
                }
                if (aux_aligned) {
                    reinterpret_cast<float4*>(intl.aux)[idx] = row;
                } else {
                    float* a4 = intl.aux + 4 * idx;
                    a4[0] = row.x; a4[1] = row.y; a4[2] = row.z; a4[3] = row.w;
                }
            }
            ++idx;
        }
    }
    if (kIntegrate) {
        float opacity, depth;
        finish_ray(acc, plan_t_far, opacity, depth);
        intl.radiance[3 * ray] = acc.cr; intl.radiance[3 * ray + 1] = acc.cg; intl.radiance[3 * ray + 2] = acc.cb;
        intl.transmittance[ray] = acc.T;
        intl.opacity[ray] = opacity;
        intl.depth[ray] = depth;
    }
}

// ---- K2, warp-staged: the same sampler, but a warp's 32 rays buffer 16 samples each in shared memory and the warp writes
// them out ray by ray in whole lines (positions 192 B, dt / sigma 64 B, colour 192 B, aux 256 B per ray) instead of every
// lane poking 4-16 bytes into its own far-away region after every sample.  Arithmetic, order and results are those of
// sample_kernel (the fused == staged bit-for-bit contract of hp_runner.cpp:1737-1760 still holds).
constexpr int kSampTile = 16;

template <bool kStratified, bool kIntegrate>
__global__ void __launch_bounds__(32)
sample_tile_kernel(MarchParams mp, float plan_t_near, float plan_t_far, FieldPair fields, RayArrays rays, uint32_t n_rays,
                   SampleArrays samp, IntegralArrays intl, bool aux_aligned) {
    __shared__ float sm[kIntegrate ? 12 : 8][32][kSampTile + 1];   // px py pz | dt | sigma | r g b | aux x4
    const uint32_t lane = threadIdx.x, ray = blockIdx.x * 32u + lane;
    const bool valid = ray < n_rays;
    float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, tn = 0.f, tf = 0.f;
    size_t idx = 0;
    if (valid) {
        ox = rays.origins[3 * ray]; oy = rays.origins[3 * ray + 1]; oz = rays.origins[3 * ray + 2];
        dx = rays.directions[3 * ray]; dy = rays.directions[3 * ray + 1]; dz = rays.directions[3 * ray + 2];
        tn = rays.t_near[ray]; tf = rays.t_far[ray];
        idx = samp.ray_offset[ray];
    }
    const uint64_t ray_index = mp.ray_index_base + ray;
    RayAccum acc;
    acc.t_cursor = plan_t_near;
    bool stopped = false, active = valid && tf > tn;
    uint32_t cnt = 0;
    const uint32_t half = lane >> 4, l = lane & 15u;
    for (uint32_t step = 0;; ++step) {
        if (active && step >= mp.max_steps) active = false;
        if (active) {
            float t, dtv;
            const int r = march_step<kStratified>(tn, tf, mp.dt, mp.seed, ray_index, step, t, dtv);
            if (r == 2) {
                active = false;
            } else if (r == 0) {
                const float px = ox + dx * t, py = oy + dy * t, pz = oz + dz * t;
                const float4 v = sample_fields(fields, px, py, pz);
                sm[0][lane][cnt] = px; sm[1][lane][cnt] = py; sm[2][lane][cnt] = pz;
                sm[3][lane][cnt] = dtv;
                sm[4][lane][cnt] = v.w;
                sm[5][lane][cnt] = v.x; sm[6][lane][cnt] = v.y; sm[7][lane][cnt] = v.z;
                if (kIntegrate) {
                    float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!stopped) {
                        float a, w, tb;
                        stopped = integrate_sample(acc, dtv, v, a, w, tb);
                        row = make_float4(a, w, tb, logf(fmaxf(tb, 1e-30f)));
                    }
                    sm[8][lane][cnt] = row.x; sm[9][lane][cnt] = row.y; sm[10][lane][cnt] = row.z; sm[11][lane][cnt] = row.w;
                }
                ++cnt;
            }
        }
        const bool any_active = __any_sync(0xffffffffu, active);
        if (__any_sync(0xffffffffu, cnt == kSampTile) || !any_active) {
            __syncwarp();
            // ---- flush: two rays per pass (lanes 0-15 / 16-31), every ray's buffered samples as contiguous runs
#pragma unroll 4
            for (uint32_t rr = 0; rr < 32u; rr += 2u) {
                const uint32_t r = rr + half;
                const uint32_t n = __shfl_sync(0xffffffffu, cnt, r);
                const unsigned long long base = __shfl_sync(0xffffffffu, static_cast<unsigned long long>(idx), r);
                if (n == 0u) continue;
                float* pos = samp.positions + 3 * base;
                float* col = samp.color + 3 * base;
                for (uint32_t k = l; k < 3u * n; k += 16u) {
                    pos[k] = sm[k % 3u][r][k / 3u];
                    col[k] = sm[5 + k % 3u][r][k / 3u];
                }
                if (l < n) {
                    samp.dt[base + l] = sm[3][r][l];
                    samp.sigma[base + l] = sm[4][r][l];
                    if (kIntegrate) {
                        if (aux_aligned) {
                            reinterpret_cast<float4*>(intl.aux)[base + l] = make_float4(sm[8][r][l], sm[9][r][l], sm[10][r][l], sm[11][r][l]);
                        } else {
                            float* a4 = intl.aux + 4 * (base + l);
                            a4[0] = sm[8][r][l]; a4[1] = sm[9][r][l]; a4[2] = sm[10][r][l]; a4[3] = sm[11][r][l];
                        }
                    }
                }
            }
            __syncwarp();
            idx += cnt;
            cnt = 0;
        }
        if (!any_active) break;
    }
    if (kIntegrate && valid) {
        float opacity, depth;
        finish_ray(acc, plan_t_far, opacity, depth);
        intl.radiance[3 * ray] = acc.cr; intl.radiance[3 * ray + 1] = acc.cg; intl.radiance[3 * ray + 2] = acc.cb;
        intl.transmittance[ray] = acc.T;
        intl.opacity[ray] = opacity;
        intl.depth[ray] = depth;
    }
}

// ---- K4: integrator over materialised samples (int_cpu.cpp:160-226) ---------
// kVec: aligned blocks of four samples, like diff_kernel below (16-byte aligned arrays, checked on the host).
template <bool kVec>
__global__ void __launch_bounds__(kThreads)
integrate_kernel(float plan_t_near, float plan_t_far, SampleArrays samp, uint32_t n_rays, uint32_t n_samples,
                 IntegralArrays intl, bool aux_aligned, uint32_t* status) {
    const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const uint32_t b = samp.ray_offset[ray], e = samp.ray_offset[ray + 1];
    RayAccum acc;
    acc.t_cursor = plan_t_near;
    if (e < b || e > n_samples) {
        atomicOr(status, kErrBadOffsets);
    } else {
        bool stopped = false;
        auto one = [&](float r, float g, float bl, float sigma, float dtv) {
            float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!stopped) {
                float a, w, tb;
                stopped = integrate_sample(acc, dtv, make_float4(r, g, bl, sigma), a, w, tb);
                row = make_float4(a, w, tb, logf(fmaxf(tb, 1e-30f)));
            }
            return row;
        };
        auto scalar = [&](uint32_t i) {
            const float4 row = one(samp.color[3 * size_t(i)], samp.color[3 * size_t(i) + 1], samp.color[3 * size_t(i) + 2], samp.sigma[i],
                                   samp.dt[i]);
            if (intl.aux != nullptr) {
                if (aux_aligned) {
                    reinterpret_cast<float4*>(intl.aux)[i] = row;
                } else {
                    float* a4 = intl.aux + 4 * size_t(i);
                    a4[0] = row.x; a4[1] = row.y; a4[2] = row.z; a4[3] = row.w;
                }
            }
        };
        uint32_t i = b;
        if (kVec)
            while (i < e && (i & 3u) != 0u) scalar(i++);
        while (i + 4u <= e) {
            float4 C0, C1, C2, SG, DT;   // colour: r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
            if (kVec) {
                const float4* c4 = reinterpret_cast<const float4*>(samp.color + 3 * size_t(i));
                C0 = __ldcs(c4); C1 = __ldcs(c4 + 1); C2 = __ldcs(c4 + 2);
                SG = __ldcs(reinterpret_cast<const float4*>(samp.sigma + i));
                DT = __ldcs(reinterpret_cast<const float4*>(samp.dt + i));
            } else {   // 4-byte aligned arrays (the reference's workspace rule): same block of four, scalar accesses
                const float* c = samp.color + 3 * size_t(i);
                C0 = make_float4(c[0], c[1], c[2], c[3]); C1 = make_float4(c[4], c[5], c[6], c[7]); C2 = make_float4(c[8], c[9], c[10], c[11]);
                SG = make_float4(samp.sigma[i], samp.sigma[i + 1], samp.sigma[i + 2], samp.sigma[i + 3]);
                DT = make_float4(samp.dt[i], samp.dt[i + 1], samp.dt[i + 2], samp.dt[i + 3]);
            }
            const float4 R0 = one(C0.x, C0.y, C0.z, SG.x, DT.x);
            const float4 R1 = one(C0.w, C1.x, C1.y, SG.y, DT.y);
            const float4 R2 = one(C1.z, C1.w, C2.x, SG.z, DT.z);
            const float4 R3 = one(C2.y, C2.z, C2.w, SG.w, DT.w);
            if (intl.aux != nullptr) {
                if (aux_aligned) {
                    float4* a4 = reinterpret_cast<float4*>(intl.aux) + i;
                    __stcs(a4, R0); __stcs(a4 + 1, R1); __stcs(a4 + 2, R2); __stcs(a4 + 3, R3);
                } else {
                    float* a = intl.aux + 4 * size_t(i);
                    a[0] = R0.x; a[1] = R0.y; a[2] = R0.z; a[3] = R0.w; a[4] = R1.x; a[5] = R1.y; a[6] = R1.z; a[7] = R1.w;
                    a[8] = R2.x; a[9] = R2.y; a[10] = R2.z; a[11] = R2.w; a[12] = R3.x; a[13] = R3.y; a[14] = R3.z; a[15] = R3.w;
                }
            }
            i += 4u;
        }
        while (i < e) scalar(i++);
    }
    float opacity, depth;
    finish_ray(acc, plan_t_far, opacity, depth);
    intl.radiance[3 * ray] = acc.cr; intl.radiance[3 * ray + 1] = acc.cg; intl.radiance[3 * ray + 2] = acc.cb;
    intl.transmittance[ray] = acc.T;
    intl.opacity[ray] = opacity;
    intl.depth[ray] = depth;
}

// ---- K4, warp-transposed (see diff_tile_kernel below for the access pattern): tiles of 16 samples, first to last -----
constexpr int kIntTile = 16;

__global__ void __launch_bounds__(32)
integrate_tile_kernel(float plan_t_near, float plan_t_far, SampleArrays samp, uint32_t n_rays, uint32_t n_samples, IntegralArrays intl,
                      bool aux_aligned, uint32_t* status) {
    __shared__ float sm[9][32][kIntTile + 1];   // r g b | sigma | dt | aux x4
    const uint32_t lane = threadIdx.x, ray = blockIdx.x * 32u + lane;
    const bool valid = ray < n_rays;
    uint32_t b = 0, e = 0;
    if (valid) {
        b = samp.ray_offset[ray];
        e = samp.ray_offset[ray + 1];
        if (e < b || e > n_samples) {
            atomicOr(status, kErrBadOffsets);
            e = b;
        }
    }
    const uint32_t longest = __reduce_max_sync(0xffffffffu, e - b);
    RayAccum acc;
    acc.t_cursor = plan_t_near;
    bool stopped = false;
    const uint32_t half = lane >> 4, l = lane & 15u;
    for (uint32_t done = 0; done < longest; done += kIntTile) {
#pragma unroll 8
        for (uint32_t rr = 0; rr < 32u; rr += 2u) {
            const uint32_t r = rr + half;
            const uint32_t rb = __shfl_sync(0xffffffffu, b, r), re = __shfl_sync(0xffffffffu, e, r);
            if (re - rb <= done) continue;
            const uint32_t lo = rb + done, count = min(static_cast<uint32_t>(kIntTile), re - lo);
            if (l < count) {
                sm[3][r][l] = __ldcs(samp.sigma + lo + l);
                sm[4][r][l] = __ldcs(samp.dt + lo + l);
            }
            const float* c = samp.color + 3 * static_cast<size_t>(lo);
            for (uint32_t k = l; k < 3u * count; k += 16u) sm[k % 3u][r][k / 3u] = __ldcs(c + k);
        }
        __syncwarp();
        if (e - b > done) {
            const uint32_t count = min(static_cast<uint32_t>(kIntTile), e - b - done);
            for (uint32_t j = 0; j < count; ++j) {
                float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!stopped) {
                    float a, w, tb;
                    stopped = integrate_sample(acc, sm[4][lane][j], make_float4(sm[0][lane][j], sm[1][lane][j], sm[2][lane][j], sm[3][lane][j]), a, w, tb);
                    row = make_float4(a, w, tb, logf(fmaxf(tb, 1e-30f)));
                }
                sm[5][lane][j] = row.x; sm[6][lane][j] = row.y; sm[7][lane][j] = row.z; sm[8][lane][j] = row.w;
            }
        }
        __syncwarp();
        if (intl.aux != nullptr) {
#pragma unroll 8
            for (uint32_t rr = 0; rr < 32u; rr += 2u) {
                const uint32_t r = rr + half;
                const uint32_t rb = __shfl_sync(0xffffffffu, b, r), re = __shfl_sync(0xffffffffu, e, r);
                if (re - rb <= done) continue;
                const uint32_t lo = rb + done, count = min(static_cast<uint32_t>(kIntTile), re - lo);
                if (l < count) {
                    if (aux_aligned) {
                        __stcs(reinterpret_cast<float4*>(intl.aux) + lo + l, make_float4(sm[5][r][l], sm[6][r][l], sm[7][r][l], sm[8][r][l]));
                    } else {
                        float* a4 = intl.aux + 4 * static_cast<size_t>(lo + l);
                        a4[0] = sm[5][r][l]; a4[1] = sm[6][r][l]; a4[2] = sm[7][r][l]; a4[3] = sm[8][r][l];
                    }
                }
            }
        }
        __syncwarp();
    }
    if (valid) {
        float opacity, depth;
        finish_ray(acc, plan_t_far, opacity, depth);
        intl.radiance[3 * ray] = acc.cr; intl.radiance[3 * ray + 1] = acc.cg; intl.radiance[3 * ray + 2] = acc.cb;
        intl.transmittance[ray] = acc.T;
        intl.opacity[ray] = opacity;
        intl.depth[ray] = depth;
    }
}

// ---- K5: per-sample backward (diff_cpu.cpp:156-195) -------------------------
// One thread walks one ray's samples last-to-first with the reference's recurrence (the carry adj_T makes a ray
// sequential).  A ray's samples are contiguous in every array, so the thread works in ALIGNED BLOCKS OF FOUR samples:
// aux as 4 x 16 B, colour as 3 x 16 B (48 B), dt as 16 B in, colour gradient as 3 x 16 B and sigma gradient as 16 B out,
// issued back to back -- every 32-byte sector that is fetched or written is used completely, where sample-at-a-time
// scalar accesses (the reference's backward_kernel, diff_cuda.cu:11-63, and this kernel's first version) move a sector
// per 4-byte access.  kVec needs 16-byte aligned arrays (checked on the host); ragged heads / tails are scalar.
__device__ __forceinline__ void diff_one(float g0, float g1, float g2, float alpha, float w, float T_prev, float c0, float c1,
                                         float c2, float dtv, float& adj_T, float& dsigma, float& o0, float& o1, float& o2) {
    const float dot = g0 * c0 + g1 * c1 + g2 * c2;
    adjoint_sample(dot, alpha, T_prev, dtv, adj_T, dsigma);
    o0 = g0 * w; o1 = g1 * w; o2 = g2 * w;
}

template <bool kVec>
__global__ void __launch_bounds__(kThreads)
diff_kernel(const float* __restrict__ dL_dI, int64_t stride_ray, int64_t stride_c, SampleArrays samp,
            const float* __restrict__ aux, uint32_t n_rays, uint32_t n_samples, float* __restrict__ grad_sigma,
            float* __restrict__ grad_color, uint32_t* status) {
    const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const uint32_t b = samp.ray_offset[ray], e = samp.ray_offset[ray + 1];
    if (e < b || e > n_samples) {
        atomicOr(status, kErrBadOffsets);
        return;
    }
    const float* gp = dL_dI + static_cast<int64_t>(ray) * stride_ray;
    const float g0 = gp[0], g1 = gp[stride_c], g2 = gp[2 * stride_c];
    float adj_T = 0.0f;
    auto scalar = [&](uint32_t i) {
        float ds, o0, o1, o2;
        diff_one(g0, g1, g2, aux[4 * size_t(i)], aux[4 * size_t(i) + 1], aux[4 * size_t(i) + 2], samp.color[3 * size_t(i)],
                 samp.color[3 * size_t(i) + 1], samp.color[3 * size_t(i) + 2], samp.dt[i], adj_T, ds, o0, o1, o2);
        grad_color[3 * size_t(i)] = o0;
        grad_color[3 * size_t(i) + 1] = o1;
        grad_color[3 * size_t(i) + 2] = o2;
        grad_sigma[i] = ds;
    };
    uint32_t i = e;
    if (kVec)
        while (i > b && (i & 3u) != 0u) scalar(--i);
    while (i >= b + 4u) {
        i -= 4u;
        float4 A0, A1, A2, A3, C0, C1, C2, DT;   // colour: r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
        if (kVec) {
            const float4* a4 = reinterpret_cast<const float4*>(aux) + i;
            const float4* c4 = reinterpret_cast<const float4*>(samp.color + 3 * size_t(i));
            A0 = __ldcs(a4); A1 = __ldcs(a4 + 1); A2 = __ldcs(a4 + 2); A3 = __ldcs(a4 + 3);
            C0 = __ldcs(c4); C1 = __ldcs(c4 + 1); C2 = __ldcs(c4 + 2);
            DT = __ldcs(reinterpret_cast<const float4*>(samp.dt + i));
        } else {   // arrays that are only 4-byte aligned (the reference's workspace rule): same block, scalar accesses
            const float* a = aux + 4 * size_t(i);
            const float* c = samp.color + 3 * size_t(i);
            A0 = make_float4(a[0], a[1], a[2], 0.f); A1 = make_float4(a[4], a[5], a[6], 0.f);
            A2 = make_float4(a[8], a[9], a[10], 0.f); A3 = make_float4(a[12], a[13], a[14], 0.f);
            C0 = make_float4(c[0], c[1], c[2], c[3]); C1 = make_float4(c[4], c[5], c[6], c[7]); C2 = make_float4(c[8], c[9], c[10], c[11]);
            DT = make_float4(samp.dt[i], samp.dt[i + 1], samp.dt[i + 2], samp.dt[i + 3]);
        }
        float4 S, G0, G1, G2;
        diff_one(g0, g1, g2, A3.x, A3.y, A3.z, C2.y, C2.z, C2.w, DT.w, adj_T, S.w, G2.y, G2.z, G2.w);
        diff_one(g0, g1, g2, A2.x, A2.y, A2.z, C1.z, C1.w, C2.x, DT.z, adj_T, S.z, G1.z, G1.w, G2.x);
        diff_one(g0, g1, g2, A1.x, A1.y, A1.z, C0.w, C1.x, C1.y, DT.y, adj_T, S.y, G0.w, G1.x, G1.y);
        diff_one(g0, g1, g2, A0.x, A0.y, A0.z, C0.x, C0.y, C0.z, DT.x, adj_T, S.x, G0.x, G0.y, G0.z);
        if (kVec) {
            float4* o4 = reinterpret_cast<float4*>(grad_color + 3 * size_t(i));
            __stcs(o4, G0); __stcs(o4 + 1, G1); __stcs(o4 + 2, G2);
            __stcs(reinterpret_cast<float4*>(grad_sigma + i), S);
        } else {
            float* o = grad_color + 3 * size_t(i);
            o[0] = G0.x; o[1] = G0.y; o[2] = G0.z; o[3] = G0.w; o[4] = G1.x; o[5] = G1.y; o[6] = G1.z; o[7] = G1.w;
            o[8] = G2.x; o[9] = G2.y; o[10] = G2.z; o[11] = G2.w;
            grad_sigma[i] = S.x; grad_sigma[i + 1] = S.y; grad_sigma[i + 2] = S.z; grad_sigma[i + 3] = S.w;
        }
    }
    while (i > b) scalar(--i);
}

// ---- K5, warp-transposed: one warp = 32 rays, tiles of 16 samples through shared memory -------------------------------
// The recurrence is sequential along a ray, so a lane must own a ray; but a thread that walks its own ray touches memory in
// 16-64 byte pieces 4 KB away from its neighbours' (262 144 open DRAM pages at once: 0.9 TB/s).  Here the WARP loads the
// last 32 samples of each of its 32 rays with full-line accesses (aux 512 B, colour 384 B, dt 128 B per ray, all loads of a
// tile in flight together), transposes through shared memory, every lane runs the reference's recurrence over its ray's 32
// samples from shared memory (row stride 33: conflict-free), and the results leave the same way.  Arithmetic and its order
// are those of diff_kernel: the outputs are bit-identical.
constexpr int kDiffTile = 16;

__global__ void __launch_bounds__(32)
diff_tile_kernel(const float* __restrict__ dL_dI, int64_t stride_ray, int64_t stride_c, SampleArrays samp, const float* __restrict__ aux,
                 uint32_t n_rays, uint32_t n_samples, float* __restrict__ grad_sigma, float* __restrict__ grad_color, uint32_t* status,
                 bool aux_aligned) {
    __shared__ float sm[7][32][kDiffTile + 1];   // alpha, w, T_prev, c0 / out0, c1 / out1, c2 / out2, dt / d sigma
    const uint32_t lane = threadIdx.x, ray = blockIdx.x * 32u + lane;
    uint32_t b = 0, e = 0;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (ray < n_rays) {
        b = samp.ray_offset[ray];
        e = samp.ray_offset[ray + 1];
        if (e < b || e > n_samples) {
            atomicOr(status, kErrBadOffsets);
            e = b;
        }
        const float* gp = dL_dI + static_cast<int64_t>(ray) * stride_ray;
        g0 = gp[0]; g1 = gp[stride_c]; g2 = gp[2 * stride_c];
    }
    const uint32_t longest = __reduce_max_sync(0xffffffffu, e - b);
    float adj_T = 0.0f;
    for (uint32_t done = 0; done < longest; done += kDiffTile) {
        // ---- load: two rays per pass (lanes 0-15 / 16-31), a lane = one sample of its ray's tile [lo, hi)
        const uint32_t half = lane >> 4, l = lane & 15u;
#pragma unroll 8
        for (uint32_t rr = 0; rr < 32u; rr += 2u) {
            const uint32_t r = rr + half;
            const uint32_t rb = __shfl_sync(0xffffffffu, b, r), re = __shfl_sync(0xffffffffu, e, r);
            if (re - rb <= done) continue;
            const uint32_t hi = re - done, lo = hi - rb > kDiffTile ? hi - kDiffTile : rb, count = hi - lo;
            if (l < count) {
                const size_t i = static_cast<size_t>(lo) + l;
                if (aux_aligned) {
                    const float4 a = __ldcs(reinterpret_cast<const float4*>(aux) + i);
                    sm[0][r][l] = a.x; sm[1][r][l] = a.y; sm[2][r][l] = a.z;
                } else {
                    sm[0][r][l] = aux[4 * i]; sm[1][r][l] = aux[4 * i + 1]; sm[2][r][l] = aux[4 * i + 2];
                }
                sm[6][r][l] = __ldcs(samp.dt + i);
            }
            const float* c = samp.color + 3 * static_cast<size_t>(lo);
            for (uint32_t k = l; k < 3u * count; k += 16u) sm[3 + k % 3u][r][k / 3u] = __ldcs(c + k);
        }
        __syncwarp();
        // ---- the recurrence, lane = ray, newest sample first (diff_cpu.cpp:170-194)
        if (e - b > done) {
            const uint32_t hi = e - done, lo = hi - b > kDiffTile ? hi - kDiffTile : b, count = hi - lo;
            for (uint32_t j = count; j-- > 0u;) {
                float ds, o0, o1, o2;
                diff_one(g0, g1, g2, sm[0][lane][j], sm[1][lane][j], sm[2][lane][j], sm[3][lane][j], sm[4][lane][j], sm[5][lane][j],
                         sm[6][lane][j], adj_T, ds, o0, o1, o2);
                sm[3][lane][j] = o0; sm[4][lane][j] = o1; sm[5][lane][j] = o2; sm[6][lane][j] = ds;
            }
        }
        __syncwarp();
        // ---- store, two rays per pass again
#pragma unroll 8
        for (uint32_t rr = 0; rr < 32u; rr += 2u) {
            const uint32_t r = rr + half;
            const uint32_t rb = __shfl_sync(0xffffffffu, b, r), re = __shfl_sync(0xffffffffu, e, r);
            if (re - rb <= done) continue;
            const uint32_t hi = re - done, lo = hi - rb > kDiffTile ? hi - kDiffTile : rb, count = hi - lo;
            if (l < count) __stcs(grad_sigma + static_cast<size_t>(lo) + l, sm[6][r][l]);
            float* o = grad_color + 3 * static_cast<size_t>(lo);
            for (uint32_t k = l; k < 3u * count; k += 16u) __stcs(o + k, sm[3 + k % 3u][r][k / 3u]);
        }
        __syncwarp();
    }
}

// ---- K7: image composition (img_cpu.cpp:148-185) ----------------------------
__global__ void background_planes_kernel(ImagePlanes img, size_t pixels, float t_far) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < pixels;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        img.image[3 * i] = 0.f; img.image[3 * i + 1] = 0.f; img.image[3 * i + 2] = 0.f;
        img.trans[i] = 1.0f;
        img.opacity[i] = 0.0f;
        img.depth[i] = t_far;
        img.hitmask[i] = 0u;
    }
}

// Unique-pixel fast path.  A second ray on an already-claimed pixel raises
// kFlagDuplicatePixel and the host re-runs the order-preserving kernel below.
__global__ void compose_kernel(ImagePlanes img, size_t pixels, const uint32_t* __restrict__ pixel_ids,
                               IntegralArrays intl, uint32_t n_rays, uint32_t* status) {
    const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const uint32_t pid = pixel_ids != nullptr ? pixel_ids[ray] : 0u;
    if (pid >= pixels) {
        atomicOr(status, kErrBadPixel);
        return;
    }
    if (atomicExch(&img.hitmask[pid], 1u) != 0u) {
        atomicOr(status, kFlagDuplicatePixel);
        return;
    }
    img.image[3 * size_t(pid)] = intl.radiance[3 * ray];
    img.image[3 * size_t(pid) + 1] = intl.radiance[3 * ray + 1];
    img.image[3 * size_t(pid) + 2] = intl.radiance[3 * ray + 2];
    img.trans[pid] = intl.transmittance[ray];
    img.opacity[pid] = intl.opacity[ray];
    img.depth[pid] = intl.depth[ray];
}

// Repeated pixel ids (override rays): ray order matters for the float sums and
// products, so one thread walks the rays in order, exactly like the reference.
__global__ void compose_sequential_kernel(ImagePlanes img, size_t pixels, const uint32_t* __restrict__ pixel_ids,
                                          IntegralArrays intl, uint32_t n_rays) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (uint32_t ray = 0; ray < n_rays; ++ray) {
        const uint32_t pid = pixel_ids != nullptr ? pixel_ids[ray] : 0u;
        if (pid >= pixels) return;
        if (img.hitmask[pid] == 0u) {
            img.image[3 * size_t(pid)] = intl.radiance[3 * ray];
            img.image[3 * size_t(pid) + 1] = intl.radiance[3 * ray + 1];
            img.image[3 * size_t(pid) + 2] = intl.radiance[3 * ray + 2];
            img.trans[pid] = intl.transmittance[ray];
            img.opacity[pid] = intl.opacity[ray];
            img.depth[pid] = intl.depth[ray];
            img.hitmask[pid] = 1u;
        } else {
            img.image[3 * size_t(pid)] += intl.radiance[3 * ray];
            img.image[3 * size_t(pid) + 1] += intl.radiance[3 * ray + 1];
            img.image[3 * size_t(pid) + 2] += intl.radiance[3 * ray + 2];
            img.trans[pid] *= intl.transmittance[ray];
            img.opacity[pid] = 1.0f - img.trans[pid];
            img.depth[pid] = fminf(img.depth[pid], intl.depth[ray]);
        }
    }
}

// ---- sample -> grid scatter on materialised samples -------------------------
__global__ void scatter_kernel(ScatterParams sp, const float* __restrict__ positions,
                               const float* __restrict__ grad_sigma, const float* __restrict__ grad_color,
                               size_t n_samples) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n_samples) return;
    scatter_sample(sp, positions[3 * i], positions[3 * i + 1], positions[3 * i + 2],
                   make_float4(grad_color[3 * i], grad_color[3 * i + 1], grad_color[3 * i + 2], grad_sigma[i]));
}

// ---- K8: pack / unpack -------------------------------------------------------
// grid[z0+z][y0+y][x0+x] += box[z][y][x] (float4 per voxel), then box = 0: the hand-over of one all-reduced gradient box
// (hpx_backward_box) to the grid's gradient block; clearing in the same pass leaves the box ready for the next step.
__global__ void add_box_kernel(float4* __restrict__ grid, float4* __restrict__ box, int32_t nx, int32_t ny, int32_t x0,
                               int32_t y0, int32_t z0, int32_t bx, int32_t by, int32_t bz) {
    const size_t n = static_cast<size_t>(bx) * by * bz;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int32_t x = static_cast<int32_t>(i % bx);
        const size_t r = i / bx;
        const int32_t y = static_cast<int32_t>(r % by), z = static_cast<int32_t>(r / by);
        const float4 v = box[i];
        box[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
        float4* g = grid + (static_cast<size_t>(z0 + z) * ny + (y0 + y)) * nx + (x0 + x);
        float4 a = *g;
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        *g = a;
    }
}

// sigma_stride / color_stride: floats between consecutive voxels of the sources (1 / 3 for the ABI layouts, 4 for fields
// that view another packed grid)
__global__ void pack_grid_kernel(const float* __restrict__ sigma, size_t sigma_stride, const float* __restrict__ color,
                                 size_t color_stride, float4* __restrict__ packed, size_t voxels, bool keep_missing) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < voxels;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (keep_missing && (sigma == nullptr || color == nullptr)) v = packed[i];
        if (color != nullptr) { v.x = color[color_stride * i]; v.y = color[color_stride * i + 1]; v.z = color[color_stride * i + 2]; }
        if (sigma != nullptr) v.w = sigma[sigma_stride * i];
        packed[i] = v;
    }
}

// first..first+voxels of the REFERENCE order (z slowest, x fastest); the packed block may use any axis order (strides)
__global__ void unpack_grad_kernel(const float4* __restrict__ packed, float* __restrict__ sigma_grad,
                                   float* __restrict__ color_grad, size_t first, size_t voxels, uint32_t nx, uint32_t ny,
                                   uint32_t sx, uint32_t sy, uint32_t sz) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < voxels;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t v_ref = first + i;
        const uint32_t x = static_cast<uint32_t>(v_ref % nx);
        const size_t r = v_ref / nx;
        const uint32_t y = static_cast<uint32_t>(r % ny), z = static_cast<uint32_t>(r / ny);
        const float4 v = packed[static_cast<size_t>(x) * sx + static_cast<size_t>(y) * sy + static_cast<size_t>(z) * sz];
        if (sigma_grad != nullptr) sigma_grad[i] = v.w;
        if (color_grad != nullptr) { color_grad[3 * i] = v.x; color_grad[3 * i + 1] = v.y; color_grad[3 * i + 2] = v.z; }
    }
}

// The slabs [s0, s1) along `axis` of a gradient block with arbitrary axis order (strides sx, sy, sz in voxels), un-interleaved
// into the REFERENCE layout sigma_grad[V], color_grad[3V] (z slowest, x fastest) at their own positions there.
__global__ void unpack_grad_slabs_kernel(const float4* __restrict__ packed, float* __restrict__ sigma_grad,
                                         float* __restrict__ color_grad, int axis, uint32_t s0, uint32_t s1, uint32_t nx, uint32_t ny,
                                         uint32_t nz, uint32_t sx, uint32_t sy, uint32_t sz) {
    const uint32_t dx = axis == 0 ? s1 - s0 : nx, dy = axis == 1 ? s1 - s0 : ny, dz = axis == 2 ? s1 - s0 : nz;
    const uint32_t ox = axis == 0 ? s0 : 0u, oy = axis == 1 ? s0 : 0u, oz = axis == 2 ? s0 : 0u;
    const size_t total = static_cast<size_t>(dx) * dy * dz;
    for (size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; j < total; j += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint32_t x = static_cast<uint32_t>(j % dx) + ox;
        const size_t r = j / dx;
        const uint32_t y = static_cast<uint32_t>(r % dy) + oy, z = static_cast<uint32_t>(r / dy) + oz;
        const float4 v = packed[static_cast<size_t>(x) * sx + static_cast<size_t>(y) * sy + static_cast<size_t>(z) * sz];
        const size_t ref = (static_cast<size_t>(z) * ny + y) * nx + x;
        if (sigma_grad != nullptr) sigma_grad[ref] = v.w;
        if (color_grad != nullptr) { color_grad[3 * ref] = v.x; color_grad[3 * ref + 1] = v.y; color_grad[3 * ref + 2] = v.z; }
    }
}

}  // namespace

cudaError_t launch_unpack_grad_slabs(cudaStream_t s, const float4* packed, float* sigma_grad, float* color_grad, int axis, uint32_t s0,
                                     uint32_t s1, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t sx, uint32_t sy, uint32_t sz) {
    if (s1 <= s0) return cudaSuccess;
    unpack_grad_slabs_kernel<<<148 * 8, 256, 0, s>>>(packed, sigma_grad, color_grad, axis, s0, s1, nx, ny, nz, sx, sy, sz);
    return cudaGetLastError();
}

cudaError_t launch_rays(cudaStream_t s, const FrameParams& p, const RayArrays& out, uint32_t n_rays) {
    if (n_rays == 0) return cudaSuccess;
    rays_kernel<<<blocks_for(n_rays, 256), 256, 0, s>>>(p, out, n_rays);
    return cudaGetLastError();
}

size_t scan_scratch_bytes(uint32_t n_rays) {
    const size_t tiles = (static_cast<size_t>(n_rays) + kScanTile - 1) / kScanTile + 1;
    return static_cast<size_t>(n_rays) * sizeof(uint32_t) + 16 + tiles * sizeof(unsigned long long);
}

cudaError_t launch_count_and_scan(cudaStream_t s, const MarchParams& mp, const RayArrays& rays, uint32_t n_rays,
                                  uint32_t* ray_offset, unsigned long long* d_total, void* scratch) {
    if (n_rays == 0) {
        cudaMemsetAsync(ray_offset, 0, sizeof(uint32_t), s);
        return cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), s);
    }
    uint32_t* counts = static_cast<uint32_t*>(scratch);
    const size_t counts_bytes = (static_cast<size_t>(n_rays) * sizeof(uint32_t) + 15) & ~size_t(15);
    auto* tile_sums = reinterpret_cast<unsigned long long*>(static_cast<char*>(scratch) + counts_bytes);
    const uint32_t tiles = blocks_for(n_rays, kScanTile);
    count_kernel<<<blocks_for(n_rays, 256), 256, 0, s>>>(mp, rays, n_rays, counts);
    scan_tile_sums_kernel<<<tiles, kScanThreads, 0, s>>>(counts, n_rays, tile_sums);
    scan_spine_kernel<<<1, 1024, 0, s>>>(tile_sums, tiles, d_total);
    scan_apply_kernel<<<tiles, kScanThreads, 0, s>>>(counts, n_rays, tile_sums, d_total, ray_offset);
    return cudaGetLastError();
}

cudaError_t launch_uniform_offsets(cudaStream_t s, uint32_t* ray_offset, uint32_t n_rays, uint32_t per_ray) {
    uniform_offsets_kernel<<<blocks_for(static_cast<size_t>(n_rays) + 1, 256), 256, 0, s>>>(ray_offset, n_rays, per_ray);
    return cudaGetLastError();
}

cudaError_t launch_sample(cudaStream_t s, const MarchParams& mp, float plan_t_near, float plan_t_far,
                          const FieldPair& fields, const RayArrays& rays, uint32_t n_rays, const SampleArrays& samp,
                          bool integrate, const IntegralArrays& intl) {
    if (n_rays == 0) return cudaSuccess;
    const uint32_t blocks = blocks_for(n_rays, kThreads);
    const bool al = (reinterpret_cast<uintptr_t>(intl.aux) & 15u) == 0;
    static const bool per_thread = std::getenv("DVREN_SAMPLE_PER_THREAD") != nullptr;   // A/B timing: the thread-per-ray kernel
    if (!per_thread && mp.max_steps >= 16u) {   // long rays: stage a warp's samples in shared memory, write whole lines
        const uint32_t wb = blocks_for(n_rays, 32);
        if (mp.stratified) {
            if (integrate) sample_tile_kernel<true, true><<<wb, 32, 0, s>>>(mp, plan_t_near, plan_t_far, fields, rays, n_rays, samp, intl, al);
            else           sample_tile_kernel<true, false><<<wb, 32, 0, s>>>(mp, plan_t_near, plan_t_far, fields, rays, n_rays, samp, intl, al);
        } else {
            if (integrate) sample_tile_kernel<false, true><<<wb, 32, 0, s>>>(mp, plan_t_near, plan_t_far, fields, rays, n_rays, samp, intl, al);
            else           sample_tile_kernel<false, false><<<wb, 32, 0, s>>>(mp, plan_t_near, plan_t_far, fields, rays, n_rays, samp, intl, al);
        }
        return cudaGetLastError();
    }
    if (mp.stratified) {
        if (integrate) sample_kernel<true, true><<<blocks, kThreads, 0, s>>>(mp, plan_t_near, plan_t_far, fields, rays, n_rays, samp, intl, al);
        else           sample_kernel<true, false><<<blocks, kThreads, 0, s>>>(mp, plan_t_near, plan_t_far, fields, rays, n_rays, samp, intl, al);
    } else {
        if (integrate) sample_kernel<false, true><<<blocks, kThreads, 0, s>>>(mp, plan_t_near, plan_t_far, fields, rays, n_rays, samp, intl, al);
        else           sample_kernel<false, false><<<blocks, kThreads, 0, s>>>(mp, plan_t_near, plan_t_far, fields, rays, n_rays, samp, intl, al);
    }
    return cudaGetLastError();
}

cudaError_t launch_integrate(cudaStream_t s, float plan_t_near, float plan_t_far, const SampleArrays& samp,
                             uint32_t n_rays, uint32_t n_samples, const IntegralArrays& intl, uint32_t* d_status) {
    if (n_rays == 0) return cudaSuccess;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0u; };
    const bool aligned = al16(intl.aux);
    static const bool per_thread = std::getenv("DVREN_INT_PER_THREAD") != nullptr;   // A/B timing: the thread-per-ray kernels below
    if (!per_thread && n_samples / n_rays >= 16u) {
        integrate_tile_kernel<<<blocks_for(n_rays, 32), 32, 0, s>>>(plan_t_near, plan_t_far, samp, n_rays, n_samples, intl, aligned, d_status);
        return cudaGetLastError();
    }
    if (al16(samp.color) && al16(samp.sigma) && al16(samp.dt))
        integrate_kernel<true><<<blocks_for(n_rays, 64), 64, 0, s>>>(plan_t_near, plan_t_far, samp, n_rays, n_samples, intl, aligned, d_status);
    else
        integrate_kernel<false><<<blocks_for(n_rays, 64), 64, 0, s>>>(plan_t_near, plan_t_far, samp, n_rays, n_samples, intl, aligned, d_status);
    return cudaGetLastError();
}

cudaError_t launch_diff(cudaStream_t s, const float* dL_dI, int64_t stride_ray, int64_t stride_c,
                        const SampleArrays& samp, const float* aux, uint32_t n_rays, uint32_t n_samples,
                        float* grad_sigma, float* grad_color, uint32_t* d_status) {
    if (n_rays == 0 || n_samples == 0) return cudaSuccess;
    auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0u; };
    static const bool per_thread = std::getenv("DVREN_DIFF_PER_THREAD") != nullptr;   // A/B timing: the thread-per-ray kernels below
    if (!per_thread && n_samples / n_rays >= 16u) {   // long rays: the warp-transposed kernel (short rays leave its tiles mostly empty)
        diff_tile_kernel<<<blocks_for(n_rays, 32), 32, 0, s>>>(dL_dI, stride_ray, stride_c, samp, aux, n_rays, n_samples, grad_sigma,
                                                               grad_color, d_status, aligned(aux));
        return cudaGetLastError();
    }
    if (aligned(aux) && aligned(samp.color) && aligned(samp.dt) && aligned(grad_sigma) && aligned(grad_color))
        diff_kernel<true><<<blocks_for(n_rays, 64), 64, 0, s>>>(dL_dI, stride_ray, stride_c, samp, aux, n_rays, n_samples, grad_sigma,
                                                                grad_color, d_status);
    else
        diff_kernel<false><<<blocks_for(n_rays, 64), 64, 0, s>>>(dL_dI, stride_ray, stride_c, samp, aux, n_rays, n_samples,
                                                                 grad_sigma, grad_color, d_status);
    return cudaGetLastError();
}

cudaError_t launch_background(cudaStream_t s, const ImagePlanes& img, size_t pixels, float t_far) {
    if (pixels == 0) return cudaSuccess;
    const uint32_t blocks = static_cast<uint32_t>(min(static_cast<size_t>(148 * 8), (pixels + 255) / 256));
    background_planes_kernel<<<blocks, 256, 0, s>>>(img, pixels, t_far);
    return cudaGetLastError();
}

cudaError_t launch_compose(cudaStream_t s, const ImagePlanes& img, size_t pixels, const uint32_t* pixel_ids,
                           const IntegralArrays& intl, uint32_t n_rays, uint32_t* d_status) {
    if (n_rays == 0) return cudaSuccess;
    compose_kernel<<<blocks_for(n_rays, 256), 256, 0, s>>>(img, pixels, pixel_ids, intl, n_rays, d_status);
    return cudaGetLastError();
}

cudaError_t launch_compose_sequential(cudaStream_t s, const ImagePlanes& img, size_t pixels,
                                      const uint32_t* pixel_ids, const IntegralArrays& intl, uint32_t n_rays) {
    compose_sequential_kernel<<<1, 32, 0, s>>>(img, pixels, pixel_ids, intl, n_rays);
    return cudaGetLastError();
}

cudaError_t launch_add_box(cudaStream_t s, float4* grid, float4* box, int32_t nx, int32_t ny, const int32_t b[6]) {
    const size_t n = static_cast<size_t>(b[3]) * b[4] * b[5];
    if (n == 0) return cudaSuccess;
    add_box_kernel<<<148 * 8, 256, 0, s>>>(grid, box, nx, ny, b[0], b[1], b[2], b[3], b[4], b[5]);
    return cudaGetLastError();
}

cudaError_t launch_scatter(cudaStream_t s, const ScatterParams& sp, const float* positions, const float* grad_sigma,
                           const float* grad_color, size_t n_samples) {
    if (n_samples == 0) return cudaSuccess;
    scatter_kernel<<<blocks_for(n_samples, 256), 256, 0, s>>>(sp, positions, grad_sigma, grad_color, n_samples);
    return cudaGetLastError();
}

cudaError_t launch_pack_grid(cudaStream_t s, const float* sigma, const float* color, float4* packed, size_t voxels,
                             bool keep_missing) {
    if (voxels == 0) return cudaSuccess;
    const uint32_t blocks = static_cast<uint32_t>(min(static_cast<size_t>(148 * 8), (voxels + 255) / 256));
    pack_grid_kernel<<<blocks, 256, 0, s>>>(sigma, 1, color, 3, packed, voxels, keep_missing);
    return cudaGetLastError();
}

cudaError_t launch_pack_grid_strided(cudaStream_t s, const float* sigma, int32_t sigma_stride, const float* color,
                                     int32_t color_stride, float4* packed, size_t voxels, bool keep_missing) {
    if (voxels == 0) return cudaSuccess;
    const uint32_t blocks = static_cast<uint32_t>(min(static_cast<size_t>(148 * 8), (voxels + 255) / 256));
    pack_grid_kernel<<<blocks, 256, 0, s>>>(sigma, static_cast<size_t>(sigma_stride), color, static_cast<size_t>(color_stride),
                                            packed, voxels, keep_missing);
    return cudaGetLastError();
}

cudaError_t launch_unpack_grad(cudaStream_t s, const float4* packed, float* sigma_grad, float* color_grad, size_t first,
                               size_t voxels, uint32_t nx, uint32_t ny, uint32_t sx, uint32_t sy, uint32_t sz) {
    if (voxels == 0) return cudaSuccess;
    const uint32_t blocks = static_cast<uint32_t>(min(static_cast<size_t>(148 * 8), (voxels + 255) / 256));
    unpack_grad_kernel<<<blocks, 256, 0, s>>>(packed, sigma_grad, color_grad, first, voxels, nx, ny, sx, sy, sz);
    return cudaGetLastError();
}

}  // namespace dv
