// dv_device.cuh -- the arithmetic of the hot path as __device__ functions.
//
// PARITY RULES.  The CPU reference is built without FMA
// contraction, so every translation unit that includes this file is compiled
// with -fmad=false and IEEE division / square root (nvcc defaults).  Anything
// that feeds an index, a sample count or the transmittance stop test
// (t, position, grid coordinate, sigma lerps, alpha, T) follows the reference
// operation by operation; citations are file:line under /root/reference.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "dv_types.h"

namespace dv {

// ---------------------------------------------------------------------------
// ray generation -- hotpath/src/cpu/ray_cpu.cpp:189-216
// ---------------------------------------------------------------------------
struct Ray { float ox, oy, oz, dx, dy, dz; };

// Unnormalised pinhole direction pieces, kept for the camera adjoint.
struct RayAux { float qx, qy, qz, inv_len; };

__device__ __forceinline__ Ray make_ray(const CameraParams& c, uint32_t px, uint32_t py, RayAux* aux = nullptr) {
    const float u = static_cast<float>(px) + 0.5f;
    const float v = static_cast<float>(py) + 0.5f;
    float qx = (u - c.cx) / c.fx;
    float qy = (v - c.cy) / c.fy;
    const float qz = 1.0f;
    if (c.ortho) { qx = 0.0f; qy = 0.0f; }
    float wx = c.r00 * qx + c.r01 * qy + c.r02 * qz;
    float wy = c.r10 * qx + c.r11 * qy + c.r12 * qz;
    float wz = c.r20 * qx + c.r21 * qy + c.r22 * qz;
    const float len_sq = wx * wx + wy * wy + wz * wz;
    const float inv_len = 1.0f / sqrtf(fmaxf(len_sq, FLT_MIN));
    Ray r;
    r.dx = wx * inv_len; r.dy = wy * inv_len; r.dz = wz * inv_len;
    r.ox = c.ox; r.oy = c.oy; r.oz = c.oz;
    if (aux) { aux->qx = qx; aux->qy = qy; aux->qz = qz; aux->inv_len = inv_len; }
    return r;
}

// ---------------------------------------------------------------------------
// stratified jitter -- hotpath/src/cpu/samp_cpu.cpp:21-35
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t s) {
    s = (s ^ (s >> 30)) * 0xbf58476d1ce4e5b9ULL;
    s = (s ^ (s >> 27)) * 0x94d049bb133111ebULL;
    return s ^ (s >> 31);
}

// The reference divides a 52-bit integer by 2^52 in double and rounds to float.
// Rounding the integer straight to float and scaling by the exact power of two
// is the same single rounding, without fp64.
__device__ __forceinline__ float jitter_unit(uint64_t seed, uint64_t ray_index, uint32_t step) {
    const uint64_t s = mix64(seed ^ (ray_index << 32) ^ static_cast<uint64_t>(step));
    return __ull2float_rn(s & 0x000fffffffffffffULL) * 0x1p-52f;
}

// ---------------------------------------------------------------------------
// one iteration of the marching loop -- hotpath/src/cpu/samp_cpu.cpp:226-244
// returns 0 = emit sample, 1 = skip (continue), 2 = stop (break)
// ---------------------------------------------------------------------------
template <bool kStratified>
__device__ __forceinline__ int march_step(float tn, float tf, float dts, uint64_t seed, uint64_t ray_index,
                                          uint32_t step, float& t_out, float& dt_out) {
    const float base = tn + static_cast<float>(step) * dts;
    if (base >= tf) return 2;
    float jit = 0.5f;
    if (kStratified) {
        jit = jitter_unit(seed, ray_index, step);
        jit = jit < 0.0f ? 0.0f : (jit > 1.0f ? 1.0f : jit);
    }
    float t = base + jit * dts;
    if (t >= tf) t = nextafterf(tf, tn);
    const float seg_end = fminf(base + dts, tf);
    const float dta = seg_end - base;
    if (!(dta > 0.0f)) return 1;
    t_out = t;
    dt_out = dta;
    return 0;
}

// Sample time of one step from the frame's step table (dv_lean.h): the table holds everything that does
// not depend on the ray; only the stratified jitter is evaluated here.
template <bool kStratified>
__device__ __forceinline__ float step_time(const float4& tab, float tn, float tf, float dts, uint64_t seed,
                                           uint64_t ray_index, uint32_t step) {
    if (!kStratified) return tab.y;
    float jit = jitter_unit(seed, ray_index, step);
    jit = jit < 0.0f ? 0.0f : (jit > 1.0f ? 1.0f : jit);
    float t = tab.x + jit * dts;
    if (t >= tf) t = nextafterf(tf, tn);
    return t;
}

// Conservative range of step indices [k_lo, k_hi) whose interval [base, base + dt] can overlap [t_in, t_out]:
// loop bounds only -- the exact per-step test is still applied inside the range.
__device__ __forceinline__ void step_range(float tn, float dts, uint32_t count, float t_in, float t_out,
                                           uint32_t& k_lo, uint32_t& k_hi) {
    k_lo = count;
    k_hi = 0;
    if (!(t_out >= t_in)) return;
    const float qa = (t_in - tn) / dts, qb = (t_out - tn) / dts;
    const float a = qa - (4.0f + fabsf(qa) * 1e-6f), b = qb + (4.0f + fabsf(qb) * 1e-6f);
    const float fc = static_cast<float>(count);
    k_lo = a <= 0.0f ? 0u : (a >= fc ? count : static_cast<uint32_t>(a));
    k_hi = b <= 0.0f ? 0u : (b >= fc ? count : min(count, static_cast<uint32_t>(b) + 1u));
    if (k_hi < k_lo) { k_lo = count; k_hi = 0; }
}

// ---------------------------------------------------------------------------
// dense grid -- hotpath/src/cpu/grid_dense_cpu.cpp:56-119,143-175
// World bounds are the unit cube (hp_runtime.cpp:289-294), so the reference's
// (p - 0) / 1 is the identity bit for bit and is not spelled out.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float lerp_ref(float a, float b, float t) { return a + (b - a) * t; }

struct Cell {
    int32_t x0, y0, z0, x1, y1, z1;
    float tx, ty, tz;
};

// Returns false when the query contributes nothing (outside + ZERO policy).
__device__ __forceinline__ bool grid_coords(float px, float py, float pz, bool clamp_oob, int32_t nx, int32_t ny,
                                            int32_t nz, float& fx, float& fy, float& fz) {
    const bool outside = (px < 0.0f || px > 1.0f) || (py < 0.0f || py > 1.0f) || (pz < 0.0f || pz > 1.0f);
    if (clamp_oob) {
        px = px < 0.0f ? 0.0f : (px > 1.0f ? 1.0f : px);
        py = py < 0.0f ? 0.0f : (py > 1.0f ? 1.0f : py);
        pz = pz < 0.0f ? 0.0f : (pz > 1.0f ? 1.0f : pz);
    } else if (outside) {
        return false;
    }
    fx = px * static_cast<float>(nx - 1);
    fy = py * static_cast<float>(ny - 1);
    fz = pz * static_cast<float>(nz - 1);
    return true;
}

__device__ __forceinline__ Cell make_cell(float fx, float fy, float fz, int32_t nx, int32_t ny, int32_t nz) {
    Cell c;
    c.x0 = static_cast<int32_t>(floorf(fx));
    c.y0 = static_cast<int32_t>(floorf(fy));
    c.z0 = static_cast<int32_t>(floorf(fz));
    c.x1 = min(c.x0 + 1, nx - 1);
    c.y1 = min(c.y0 + 1, ny - 1);
    c.z1 = min(c.z0 + 1, nz - 1);
    c.tx = fx - static_cast<float>(c.x0);
    c.ty = fy - static_cast<float>(c.y0);
    c.tz = fz - static_cast<float>(c.z0);
    return c;
}

__device__ __forceinline__ float trilerp(float c000, float c100, float c010, float c110, float c001, float c101,
                                         float c011, float c111, float tx, float ty, float tz) {
    const float c00 = lerp_ref(c000, c100, tx);
    const float c10 = lerp_ref(c010, c110, tx);
    const float c01 = lerp_ref(c001, c101, tx);
    const float c11 = lerp_ref(c011, c111, tx);
    const float c0 = lerp_ref(c00, c10, ty);
    const float c1 = lerp_ref(c01, c11, ty);
    return lerp_ref(c0, c1, tz);
}

// Colour channels never feed an index, a count or the transmittance stop test, so the lean
// kernels may contract their lerps (results stay within 1 ulp per lerp of the reference's).
__device__ __forceinline__ float lerp_fma(float a, float b, float t) { return __fmaf_rn(b - a, t, a); }

__device__ __forceinline__ float trilerp_fma(float c000, float c100, float c010, float c110, float c001, float c101,
                                             float c011, float c111, float tx, float ty, float tz) {
    const float c00 = lerp_fma(c000, c100, tx);
    const float c10 = lerp_fma(c010, c110, tx);
    const float c01 = lerp_fma(c001, c101, tx);
    const float c11 = lerp_fma(c011, c111, tx);
    const float c0 = lerp_fma(c00, c10, ty);
    const float c1 = lerp_fma(c01, c11, ty);
    return lerp_fma(c0, c1, tz);
}

__device__ __forceinline__ size_t voxel_index(int32_t x, int32_t y, int32_t z, int32_t nx, int32_t ny) {
    return (static_cast<size_t>(z) * static_cast<size_t>(ny) + static_cast<size_t>(y)) * static_cast<size_t>(nx) +
           static_cast<size_t>(x);
}

// Packed grids are limited to 2^32 - 1 voxels (hpx_grid_create / field packing check it), so corner
// indices are 32-bit: one IMAD.WIDE per load instead of 64-bit index chains.
__device__ __forceinline__ uint32_t voxel_index32(int32_t x, int32_t y, int32_t z, int32_t nx, int32_t ny) {
    return (static_cast<uint32_t>(z) * static_cast<uint32_t>(ny) + static_cast<uint32_t>(y)) * static_cast<uint32_t>(nx) +
           static_cast<uint32_t>(x);
}

struct Corners { float4 v000, v100, v010, v110, v001, v101, v011, v111; };

// Storage of one packed voxel: float4 {r,g,b,sigma} (16 B, the default) or four IEEE halfs (8 B, hpx_grid_set_storage).
// Half storage changes what is STORED, not the arithmetic: a voxel is widened to float4 exactly on load and everything
// downstream is the same fp32 code, so a half grid renders exactly like an fp32 grid holding the rounded values.
struct HalfVoxel { uint2 bits; };   // {r, g} | {b, sigma}

__device__ __forceinline__ float4 load_voxel(const float4* __restrict__ g, uint32_t i) { return __ldg(g + i); }
__device__ __forceinline__ float4 load_voxel(const HalfVoxel* __restrict__ g, uint32_t i) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(g) + i);
    const float2 rg = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 bs = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    return make_float4(rg.x, rg.y, bs.x, bs.y);
}

template <class V>
__device__ __forceinline__ Corners load_corners(const V* __restrict__ g, const Cell& c, int32_t nx, int32_t ny) {
    const uint32_t i000 = voxel_index32(c.x0, c.y0, c.z0, nx, ny);
    const uint32_t ox = static_cast<uint32_t>(c.x1 - c.x0);                                       // 0 or 1
    const uint32_t oy = static_cast<uint32_t>(c.y1 - c.y0) * static_cast<uint32_t>(nx);           // 0 or one row
    const uint32_t oz = static_cast<uint32_t>(c.z1 - c.z0) * static_cast<uint32_t>(nx) * static_cast<uint32_t>(ny);
    Corners k;
#ifdef DV_EXP_NOGATHER   // timing experiment only (tools/quick_time.py): everything but the gathers
    {
        const float f = __uint_as_float(0x3f000000u | (i000 & 0xffffu));
        k.v000 = make_float4(f, f, f, f); k.v100 = make_float4(f + ox, f, f, f); k.v010 = make_float4(f, f + oy, f, f);
        k.v110 = k.v000; k.v001 = make_float4(f, f, f + oz, f); k.v101 = k.v100; k.v011 = k.v010; k.v111 = k.v001;
        return k;
    }
#endif
    k.v000 = load_voxel(g, i000);           k.v100 = load_voxel(g, i000 + ox);
    k.v010 = load_voxel(g, i000 + oy);      k.v110 = load_voxel(g, i000 + oy + ox);
    k.v001 = load_voxel(g, i000 + oz);      k.v101 = load_voxel(g, i000 + oz + ox);
    k.v011 = load_voxel(g, i000 + oz + oy); k.v111 = load_voxel(g, i000 + oz + oy + ox);
    return k;
}

// Packed {r,g,b,sigma} gather: 8 x 16-byte loads through the read-only path.
template <bool kLinear, bool kClamp, bool kExactColor = true, class V = float4>
__device__ __forceinline__ float4 sample_packed(const V* __restrict__ g, int32_t nx, int32_t ny, int32_t nz,
                                                float px, float py, float pz) {
    float fx, fy, fz;
    if (!grid_coords(px, py, pz, kClamp, nx, ny, nz, fx, fy, fz)) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (!kLinear) {
        const int32_t ix = static_cast<int32_t>(roundf(fx));
        const int32_t iy = static_cast<int32_t>(roundf(fy));
        const int32_t iz = static_cast<int32_t>(roundf(fz));
        return load_voxel(g, voxel_index32(ix, iy, iz, nx, ny));
    }
    const Cell c = make_cell(fx, fy, fz, nx, ny, nz);
    const Corners k = load_corners(g, c, nx, ny);
    const float4 v000 = k.v000, v100 = k.v100, v010 = k.v010, v110 = k.v110;
    const float4 v001 = k.v001, v101 = k.v101, v011 = k.v011, v111 = k.v111;
    float4 o;
    if (kExactColor) {
        o.x = trilerp(v000.x, v100.x, v010.x, v110.x, v001.x, v101.x, v011.x, v111.x, c.tx, c.ty, c.tz);
        o.y = trilerp(v000.y, v100.y, v010.y, v110.y, v001.y, v101.y, v011.y, v111.y, c.tx, c.ty, c.tz);
        o.z = trilerp(v000.z, v100.z, v010.z, v110.z, v001.z, v101.z, v011.z, v111.z, c.tx, c.ty, c.tz);
    } else {
        o.x = trilerp_fma(v000.x, v100.x, v010.x, v110.x, v001.x, v101.x, v011.x, v111.x, c.tx, c.ty, c.tz);
        o.y = trilerp_fma(v000.y, v100.y, v010.y, v110.y, v001.y, v101.y, v011.y, v111.y, c.tx, c.ty, c.tz);
        o.z = trilerp_fma(v000.z, v100.z, v010.z, v110.z, v001.z, v101.z, v011.z, v111.z, c.tx, c.ty, c.tz);
    }
    // sigma decides alpha, T and the stop test: always the reference's operation order
    o.w = trilerp(v000.w, v100.w, v010.w, v110.w, v001.w, v101.w, v011.w, v111.w, c.tx, c.ty, c.tz);
    return o;
}

// ---------------------------------------------------------------------------
// Lean-kernel sampler.  Same gathers; the arithmetic is arranged for the Blackwell packed-fp32 pipe:
//   (r,g) pairs: lerp = FADD2 + FFMA2          (colour may contract: it feeds no index, count or stop test)
//   (b,sigma) pairs: FADD2 + FMUL2 on the pair, then two SCALAR adds -- sigma keeps the reference's
//   a + (b - a) * t with every intermediate rounded.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2
//   even under -fmad=false (seen in SASS), which is why the final add is scalar.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// contracted lerp of an (r,g) pair
__device__ __forceinline__ uint64_t lerp2_fma(uint64_t a, uint64_t b, uint64_t t) { return fma2(sub2(b, a), t, a); }
// exact lerp of a (b,sigma) pair: packed subtract and multiply, scalar adds
__device__ __forceinline__ uint64_t lerp2_exact(uint64_t a, uint64_t b, uint64_t t) {
    const uint64_t m = mul2(sub2(b, a), t);
    float alo, ahi, mlo, mhi;
    unpack2(a, alo, ahi);
    unpack2(m, mlo, mhi);
    return pack2(alo + mlo, ahi + mhi);
}

__device__ __forceinline__ float4 trilerp_pairs(const Corners& k, float tx, float ty, float tz) {
    const uint64_t tx2 = pack2(tx, tx), ty2 = pack2(ty, ty), tz2 = pack2(tz, tz);
#define DV_RG(v) pack2((v).x, (v).y)
#define DV_BS(v) pack2((v).z, (v).w)
    const uint64_t rg = lerp2_fma(
        lerp2_fma(lerp2_fma(DV_RG(k.v000), DV_RG(k.v100), tx2), lerp2_fma(DV_RG(k.v010), DV_RG(k.v110), tx2), ty2),
        lerp2_fma(lerp2_fma(DV_RG(k.v001), DV_RG(k.v101), tx2), lerp2_fma(DV_RG(k.v011), DV_RG(k.v111), tx2), ty2), tz2);
    const uint64_t bs = lerp2_exact(
        lerp2_exact(lerp2_exact(DV_BS(k.v000), DV_BS(k.v100), tx2), lerp2_exact(DV_BS(k.v010), DV_BS(k.v110), tx2), ty2),
        lerp2_exact(lerp2_exact(DV_BS(k.v001), DV_BS(k.v101), tx2), lerp2_exact(DV_BS(k.v011), DV_BS(k.v111), tx2), ty2), tz2);
#undef DV_RG
#undef DV_BS
    float4 o;
    unpack2(rg, o.x, o.y);
    unpack2(bs, o.z, o.w);
    return o;
}

template <bool kClamp, class V>
__device__ __forceinline__ float4 sample_packed_lean(const V* __restrict__ g, int32_t nx, int32_t ny, int32_t nz,
                                                     float px, float py, float pz) {
    float fx, fy, fz;
    if (!grid_coords(px, py, pz, kClamp, nx, ny, nz, fx, fy, fz)) return make_float4(0.f, 0.f, 0.f, 0.f);
    const Cell c = make_cell(fx, fy, fz, nx, ny, nz);
    return trilerp_pairs(load_corners(g, c, nx, ny), c.tx, c.ty, c.tz);
}

// ---------------------------------------------------------------------------
// Empty-space skipping.  Brick = 8 x 8 x 8 trilinear cells; a cell (x0,y0,z0) reads voxels x0..x0+1 etc., all inside the
// region the brick's bits were computed over.  Bit 0 clear: sigma = 0 at all eight corners, so the reference's own
// arithmetic gives sigma = +0 exactly, alpha = 0, w = T * 0 = 0: the sample changes neither T nor the radiance, depth or
// stop decision (int_cpu.cpp:181-215) -- the forward pass may skip it whatever the colours are.  Bit 1 clear: the colours
// are zero too, so g . c = 0 and the backward pass needs no gather either; its d sigma = -adj_T T dt is still scattered.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t occupancy_bits(const uint32_t* __restrict__ occ, int32_t x0, int32_t y0, int32_t z0,
                                                   int32_t obx, int32_t oby) {
    const uint32_t b = (static_cast<uint32_t>(z0 >> 3) * static_cast<uint32_t>(oby) + static_cast<uint32_t>(y0 >> 3)) *
                           static_cast<uint32_t>(obx) + static_cast<uint32_t>(x0 >> 3);
    return (__ldg(occ + (b >> 4)) >> ((b & 15u) * 2u)) & 3u;
}

// Forward-pass sampler with skipping: false = the sample contributes nothing (outside + OOB zero, or an empty brick).
template <bool kClamp, class V>
__device__ __forceinline__ bool sample_packed_lean_occ(const V* __restrict__ g, const uint32_t* __restrict__ occ, int32_t obx,
                                                       int32_t oby, int32_t nx, int32_t ny, int32_t nz, float px, float py,
                                                       float pz, float4& v) {
    float fx, fy, fz;
    if (!grid_coords(px, py, pz, kClamp, nx, ny, nz, fx, fy, fz)) return false;
    const Cell c = make_cell(fx, fy, fz, nx, ny, nz);
    if ((occupancy_bits(occ, c.x0, c.y0, c.z0, obx, oby) & 1u) == 0u) return false;
    v = trilerp_pairs(load_corners(g, c, nx, ny), c.tx, c.ty, c.tz);
    return true;
}

// Same, also returning the trilinear cell {x0 | y0 << 10 | z0 << 20, tx, ty, tz}: with the unit bounding box the
// backward scatter (src/fields/dense_grid.cpp:206-246) maps a position to exactly this cell with these
// fractions (local = p, g = local * (n - 1), same clamp), so the recompute pass hands it over instead of
// deriving it a second time.  Outside + OOB-zero: key = 0xffffffff.
template <bool kClamp, class V>
__device__ __forceinline__ float4 sample_packed_lean_cell(const V* __restrict__ g, int32_t nx, int32_t ny, int32_t nz,
                                                          float px, float py, float pz, float4& cell) {
    float fx, fy, fz;
    if (!grid_coords(px, py, pz, kClamp, nx, ny, nz, fx, fy, fz)) {
        cell = make_float4(__uint_as_float(0xffffffffu), 0.f, 0.f, 0.f);
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const Cell c = make_cell(fx, fy, fz, nx, ny, nz);
    cell = make_float4(__uint_as_float(static_cast<uint32_t>(c.x0) | (static_cast<uint32_t>(c.y0) << 10) |
                                       (static_cast<uint32_t>(c.z0) << 20)),
                       c.tx, c.ty, c.tz);
    return trilerp_pairs(load_corners(g, c, nx, ny), c.tx, c.ty, c.tz);
}

// Generic single-grid query (separate sigma / colour arrays with their own
// resolution and policies, as hp_samp allows).  `ch` selects the channel.
__device__ __forceinline__ float sample_plain(const GridParams& g, float px, float py, float pz, int ch) {
    float fx, fy, fz;
    if (!grid_coords(px, py, pz, g.clamp != 0, g.nx, g.ny, g.nz, fx, fy, fz)) return 0.0f;
    const float* __restrict__ d = g.data;
    const size_t cs = static_cast<size_t>(g.channels);
    if (!g.linear) {
        const int32_t ix = static_cast<int32_t>(roundf(fx));
        const int32_t iy = static_cast<int32_t>(roundf(fy));
        const int32_t iz = static_cast<int32_t>(roundf(fz));
        return __ldg(d + voxel_index(ix, iy, iz, g.nx, g.ny) * cs + ch);
    }
    const Cell c = make_cell(fx, fy, fz, g.nx, g.ny, g.nz);
    auto at = [&](int32_t x, int32_t y, int32_t z) { return __ldg(d + voxel_index(x, y, z, g.nx, g.ny) * cs + ch); };
    return trilerp(at(c.x0, c.y0, c.z0), at(c.x1, c.y0, c.z0), at(c.x0, c.y1, c.z0), at(c.x1, c.y1, c.z0),
                   at(c.x0, c.y0, c.z1), at(c.x1, c.y0, c.z1), at(c.x0, c.y1, c.z1), at(c.x1, c.y1, c.z1), c.tx,
                   c.ty, c.tz);
}

// sigma and rgb of one position for an arbitrary field pair -> {r,g,b,sigma}
__device__ __forceinline__ float4 sample_fields(const FieldPair& f, float px, float py, float pz) {
    if (f.packed != nullptr) {
        const bool lin = f.sigma.linear != 0, clampo = f.sigma.clamp != 0;
        const int32_t nx = f.sigma.nx, ny = f.sigma.ny, nz = f.sigma.nz;
        if (lin) {
            return clampo ? sample_packed<true, true>(f.packed, nx, ny, nz, px, py, pz)
                          : sample_packed<true, false>(f.packed, nx, ny, nz, px, py, pz);
        }
        return clampo ? sample_packed<false, true>(f.packed, nx, ny, nz, px, py, pz)
                      : sample_packed<false, false>(f.packed, nx, ny, nz, px, py, pz);
    }
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f.sigma.present) o.w = sample_plain(f.sigma, px, py, pz, 0);
    if (f.color.present) {
        o.x = sample_plain(f.color, px, py, pz, 0);
        o.y = sample_plain(f.color, px, py, pz, 1);
        o.z = sample_plain(f.color, px, py, pz, 2);
    }
    return o;
}

// ---------------------------------------------------------------------------
// Conservative parameter interval in which a ray can be inside the unit cube.  With the OOB-zero
// policy every sample outside [0,1]^3 has sigma = rgb = 0 exactly (grid_dense_cpu.cpp:147-149), i.e.
// alpha = w = 0 and T unchanged: such samples only advance the depth cursor.  The interval is
// widened by far more than the rounding error of o + d*t so that no sample the reference evaluates
// inside the cube is ever skipped; samples inside the widened shell take the ordinary path.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cube_interval(const Ray& r, float& t_in, float& t_out) {
    t_in = -CUDART_INF_F;
    t_out = CUDART_INF_F;
    const float o[3] = {r.ox, r.oy, r.oz}, d[3] = {r.dx, r.dy, r.dz};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (fabsf(d[i]) > 1e-12f) {
            const float inv = 1.0f / d[i];
            const float a = (0.0f - o[i]) * inv, b = (1.0f - o[i]) * inv;
            t_in = fmaxf(t_in, fminf(a, b));
            t_out = fminf(t_out, fmaxf(a, b));
        } else if (o[i] < -1e-3f || o[i] > 1.001f) {
            t_in = CUDART_INF_F;   // parallel to the slab and clearly outside it: never inside
            t_out = -CUDART_INF_F;
        }
    }
    const float pad_in = 1e-3f * (1.0f + fabsf(t_in)), pad_out = 1e-3f * (1.0f + fabsf(t_out));
    t_in -= pad_in;
    t_out += pad_out;
}

// ---------------------------------------------------------------------------
// alpha -- hotpath/src/cpu/int_cpu.cpp:98-109 (+ the call-site clamp :188)
//
// The reference evaluates alpha = (float) clamp(-expm1(-(double)od), 0, 1) with glibc's fp64 expm1.
// alpha_of() reproduces that FLOAT bit for bit without calling a generic fp64 expm1 (which costs ~35
// fp64-pipe instructions per sample): od = k ln2 - r with k = rint(od log2 e), |r| <= 0.35, then
//   1 - exp(-od) = (1 - 2^-k) - 2^-k expm1(r),   expm1(r) = r + r^2 (1/2! + r/3! + ... + r^9/11!)
// in fp64 (12 DFMA).  The truncation error is < 2^-40 relative, far below half an fp32 ulp, and
// tests/test_oracle_pin.py::test_alpha_bit_exact_all_floats checks the restatement in
// oracle/dvren_oracle.c (orc_alpha_fast, same operations) against the libm form for EVERY float in
// [1e-4, 18]; tests/test_gpu_abi.py::test_alpha_matches_oracle_bitwise checks this device code.
// ---------------------------------------------------------------------------
// 1/11!, 1/10!, ..., 1/2!  (operands of the DFMA chain come straight from the constant bank)
static __constant__ double kExpm1Taylor[10] = {1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
                                               1.0 / 720.0,      1.0 / 120.0,     1.0 / 24.0,     1.0 / 6.0,     0.5};

__device__ __forceinline__ float alpha_of(float sigma, float dt) {
    const float od = sigma * dt;
    if (od <= 0.0f) return 0.0f;
    if (od < 1e-4f) {
        const float half = 0.5f * od;
        return od * (1.0f - half);
    }
#if defined(DV_EXP_FLOAT_EXPM1)   // timing experiment only: cost of the fp64 part
    return fminf(fmaxf(-expm1f(-od), 0.0f), 1.0f);
#elif defined(DV_ALPHA_LIBM)       // the generic CUDA expm1 (kept for A/B timing)
    double a = -expm1(-static_cast<double>(od));
    a = a < 0.0 ? 0.0 : (a > 1.0 ? 1.0 : a);
    return static_cast<float>(a);
#else
    if (od > 17.5f) return 1.0f;   // 1 - e^-17.5 rounds to 1.0f (e^-17.5 = 2.5e-8 < 2^-25)
    const float kf = rintf(od * 1.44269504f);
    const double r = __fma_rn(static_cast<double>(kf), 0.693147180559945309417, -static_cast<double>(od));
    double q = kExpm1Taylor[0];
#pragma unroll
    for (int i = 1; i < 10; ++i) q = __fma_rn(q, r, kExpm1Taylor[i]);
    const double p = __fma_rn(__dmul_rn(r, r), q, r);                       // expm1(r)
    const double s = __hiloint2double((1023 - __float2int_rn(kf)) << 20, 0);  // 2^-k
    return __double2float_rn(__fma_rn(-s, p, 1.0 - s));
#endif
}

// ---------------------------------------------------------------------------
// per-ray emission-absorption scan -- hotpath/src/cpu/int_cpu.cpp:181-215
// ---------------------------------------------------------------------------
struct RayAccum {
    float T = 1.0f, depth_w = 0.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f, t_cursor = 0.0f;
};

// Returns true when the ray stops (T <= 1e-4 after this sample).  kExact = false contracts the
// radiance / depth accumulations (they feed no decision); alpha, w and T never change.
template <bool kExact = true>
__device__ __forceinline__ bool integrate_sample(RayAccum& s, float dtv, float4 rgbs, float& alpha_out,
                                                 float& weight_out, float& T_before_out) {
    const float alpha = alpha_of(rgbs.w, dtv);
    const float T_before = s.T;
    const float w = T_before * alpha;
    const float mid = s.t_cursor + 0.5f * dtv;
    if (kExact) {
        s.cr += w * rgbs.x;
        s.cg += w * rgbs.y;
        s.cb += w * rgbs.z;
        s.depth_w += w * mid;
    } else {
        s.cr = __fmaf_rn(w, rgbs.x, s.cr);
        s.cg = __fmaf_rn(w, rgbs.y, s.cg);
        s.cb = __fmaf_rn(w, rgbs.z, s.cb);
        s.depth_w = __fmaf_rn(w, mid, s.depth_w);
    }
    s.T = T_before * fmaxf(1.0f - alpha, 0.0f);
    s.t_cursor += dtv;
    alpha_out = alpha; weight_out = w; T_before_out = T_before;
    return s.T <= kStopThreshold;
}

__device__ __forceinline__ void finish_ray(const RayAccum& s, float plan_t_far, float& opacity, float& depth) {
    opacity = 1.0f - s.T;
    depth = opacity > 1e-6f ? s.depth_w / opacity : plan_t_far;
}

// ---------------------------------------------------------------------------
// reverse adjoint of one sample -- hotpath/src/cpu/diff_cpu.cpp:170-194
// ---------------------------------------------------------------------------
__device__ __forceinline__ void adjoint_sample(float dot_gc, float alpha, float T_prev, float dtv, float& adj_T,
                                               float& dsigma) {
    const float adj_alpha = dot_gc * T_prev - adj_T * T_prev;
    const float adj_prev = dot_gc * alpha + adj_T * (1.0f - alpha);
    dsigma = adj_alpha * (dtv * (1.0f - alpha));
    adj_T = adj_prev;
}

// ---------------------------------------------------------------------------
// sample -> grid scatter -- src/fields/dense_grid.cpp:206-306
// g = {d r, d g, d b, d sigma} of the sample.  One 16-byte red per corner.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void red_add4(float4* addr, float4 v) {
#ifdef DV_EXP_NORED   // timing experiment only (tools/quick_time.py): everything but the reds
    if (v.x != 1234.56789f) return;
#endif
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// Index of voxel (x,y,z) in the scatter target (whole grid or box); false when a boxed target does not contain it.
__device__ __forceinline__ bool scatter_index(const ScatterParams& sp, int32_t x, int32_t y, int32_t z, uint32_t& idx) {
    if (sp.boxed) {
        const int32_t lx = x - sp.box_ox, ly = y - sp.box_oy, lz = z - sp.box_oz;
        if (lx < 0 || lx >= sp.box_nx || ly < 0 || ly >= sp.box_ny || lz < 0 || lz >= sp.box_nz) {
            atomicAdd(sp.box_miss, 1u);
            return false;
        }
    }
    // sp.grad / sp.fixed are biased by the box origin on the host, so grid coordinates index them directly
    idx = static_cast<uint32_t>(z) * sp.box_sz + static_cast<uint32_t>(y) * sp.box_sy + static_cast<uint32_t>(x) * sp.box_sx;
    return true;
}

// One corner contribution: a 16-byte float red, or four 64-bit integer reds of the value in units of the quantum.
__device__ __forceinline__ void scatter_add_fixed(unsigned long long* p, float4 v, float inv_q) {
    atomicAdd(p + 0, static_cast<unsigned long long>(__float2ll_rn(v.x * inv_q)));
    atomicAdd(p + 1, static_cast<unsigned long long>(__float2ll_rn(v.y * inv_q)));
    atomicAdd(p + 2, static_cast<unsigned long long>(__float2ll_rn(v.z * inv_q)));
    atomicAdd(p + 3, static_cast<unsigned long long>(__float2ll_rn(v.w * inv_q)));
}
__device__ __forceinline__ void scatter_add(const ScatterParams& sp, uint32_t voxel, float4 v, float inv_q) {
#ifndef DV_EXP_NO_FIXED
    if (sp.fixed != nullptr) scatter_add_fixed(sp.fixed + 4ull * voxel, v, inv_q);
    else
#endif
        red_add4(sp.grad + voxel, v);
}
__device__ __forceinline__ float scatter_inv_quantum(const ScatterParams& sp) {
    return sp.fixed != nullptr ? __ldg(sp.fixed_meta + 2) : 1.0f;
}

__device__ __forceinline__ void scatter_sample(const ScatterParams& sp, float px, float py, float pz, float4 g,
                                               float inv_q = 1.0f) {
    const float ex = sp.bmax[0] - sp.bmin[0], ey = sp.bmax[1] - sp.bmin[1], ez = sp.bmax[2] - sp.bmin[2];
    float lx = ex != 0.0f ? (px - sp.bmin[0]) / ex : 0.0f;
    float ly = ey != 0.0f ? (py - sp.bmin[1]) / ey : 0.0f;
    float lz = ez != 0.0f ? (pz - sp.bmin[2]) / ez : 0.0f;
    const bool outside = lx < 0.0f || lx > 1.0f || ly < 0.0f || ly > 1.0f || lz < 0.0f || lz > 1.0f;
    if (outside) {
        if (!sp.clamp) return;
        lx = fmaxf(0.0f, fminf(1.0f, lx));
        ly = fmaxf(0.0f, fminf(1.0f, ly));
        lz = fmaxf(0.0f, fminf(1.0f, lz));
    }
    const int32_t nx = sp.nx, ny = sp.ny, nz = sp.nz;
    const float gx = lx * static_cast<float>(max(nx - 1, 1));
    const float gy = ly * static_cast<float>(max(ny - 1, 1));
    const float gz = lz * static_cast<float>(max(nz - 1, 1));
    if (sp.nearest || nx == 1 || ny == 1 || nz == 1) {
        const int32_t ix = static_cast<int32_t>(roundf(gx));
        const int32_t iy = static_cast<int32_t>(roundf(gy));
        const int32_t iz = static_cast<int32_t>(roundf(gz));
        if (ix < 0 || ix >= nx || iy < 0 || iy >= ny || iz < 0 || iz >= nz) return;
        uint32_t idx;
        if (scatter_index(sp, ix, iy, iz, idx)) scatter_add(sp, idx, g, inv_q);
        return;
    }
    const Cell c = make_cell(gx, gy, gz, nx, ny, nz);
    const float ux = 1.0f - c.tx, uy = 1.0f - c.ty, uz = 1.0f - c.tz;
    // Rolled on purpose (corner order dx, dy, dz as in the reference): eight unrolled copies of the index / bounds /
    // red sequence push the per-ray backward over an instruction-cache cliff (+75 % on early-terminating volumes).
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        const bool hx = (k & 4) != 0, hy = (k & 2) != 0, hz = (k & 1) != 0;
        const int32_t ix = hx ? c.x1 : c.x0, iy = hy ? c.y1 : c.y0, iz = hz ? c.z1 : c.z0;
        if (ix < 0 || ix >= nx || iy < 0 || iy >= ny || iz < 0 || iz >= nz) continue;
        const float w = (hx ? c.tx : ux) * (hy ? c.ty : uy) * (hz ? c.tz : uz);
        uint32_t idx;
        if (scatter_index(sp, ix, iy, iz, idx))
            scatter_add(sp, idx, make_float4(g.x * w, g.y * w, g.z * w, g.w * w), inv_q);
    }
}

}  // namespace dv
