// dv_staged.h -- launch interface of the materialising kernels behind hp.h.
#pragma once

#include <cuda_runtime.h>

#include "dv_types.h"

namespace dv {

// Device views of the ABI bundles (contiguous, like the reference assumes).
struct RayArrays {
    float* origins = nullptr;     // [N][3]
    float* directions = nullptr;  // [N][3]
    float* t_near = nullptr;      // [N]
    float* t_far = nullptr;       // [N]
    uint32_t* pixel_ids = nullptr;
};

struct SampleArrays {
    float* positions = nullptr;   // [M][3]
    float* dt = nullptr;          // [M]
    uint32_t* ray_offset = nullptr;  // [N+1]
    float* sigma = nullptr;       // [M]
    float* color = nullptr;       // [M][3]
};

struct IntegralArrays {
    float* radiance = nullptr;       // [N][3]
    float* transmittance = nullptr;  // [N]
    float* opacity = nullptr;        // [N]
    float* depth = nullptr;          // [N]
    float* aux = nullptr;            // [M][4]
};

struct ImagePlanes {
    float* image = nullptr;
    float* trans = nullptr;
    float* opacity = nullptr;
    float* depth = nullptr;
    uint32_t* hitmask = nullptr;
};

// bits of the device status word the kernels raise
constexpr uint32_t kErrBadOffsets = 1u;    // ray_offset not monotone / out of range
constexpr uint32_t kErrBadPixel = 2u;      // pixel id outside the frame
constexpr uint32_t kFlagDuplicatePixel = 4u;

cudaError_t launch_rays(cudaStream_t s, const FrameParams& p, const RayArrays& out, uint32_t n_rays);

// counts[i] = samples ray i emits; then an exclusive scan into ray_offset[0..n]
// and the 64-bit total in *d_total.  scratch: scan_scratch_bytes(n) bytes.
size_t scan_scratch_bytes(uint32_t n_rays);
cudaError_t launch_count_and_scan(cudaStream_t s, const MarchParams& mp, const RayArrays& rays, uint32_t n_rays,
                                  uint32_t* ray_offset, unsigned long long* d_total, void* scratch);

// ray_offset[i] = i * per_ray for i in [0, n_rays]
cudaError_t launch_uniform_offsets(cudaStream_t s, uint32_t* ray_offset, uint32_t n_rays, uint32_t per_ray);

// hp_samp (integrate=false) / hp_samp_int_fused (integrate=true, `intl` filled)
cudaError_t launch_sample(cudaStream_t s, const MarchParams& mp, float plan_t_near, float plan_t_far,
                          const FieldPair& fields, const RayArrays& rays, uint32_t n_rays, const SampleArrays& samp,
                          bool integrate, const IntegralArrays& intl);

// hp_int over materialised samples
cudaError_t launch_integrate(cudaStream_t s, float plan_t_near, float plan_t_far, const SampleArrays& samp,
                             uint32_t n_rays, uint32_t n_samples, const IntegralArrays& intl, uint32_t* d_status);

// hp_diff: per-sample gradients
cudaError_t launch_diff(cudaStream_t s, const float* dL_dI, int64_t stride_ray, int64_t stride_c,
                        const SampleArrays& samp, const float* aux, uint32_t n_rays, uint32_t n_samples,
                        float* grad_sigma, float* grad_color, uint32_t* d_status);

// hp_img
cudaError_t launch_background(cudaStream_t s, const ImagePlanes& img, size_t pixels, float t_far);
cudaError_t launch_compose(cudaStream_t s, const ImagePlanes& img, size_t pixels, const uint32_t* pixel_ids,
                           const IntegralArrays& intl, uint32_t n_rays, uint32_t* d_status);
cudaError_t launch_compose_sequential(cudaStream_t s, const ImagePlanes& img, size_t pixels,
                                      const uint32_t* pixel_ids, const IntegralArrays& intl, uint32_t n_rays);

// DenseGridField::AccumulateSampleGradients on materialised samples
// grid gradient += box, box = 0 (see add_box_kernel)
cudaError_t launch_add_box(cudaStream_t s, float4* grid, float4* box, int32_t nx, int32_t ny, const int32_t box6[6]);

cudaError_t launch_scatter(cudaStream_t s, const ScatterParams& sp, const float* positions, const float* grad_sigma,
                           const float* grad_color, size_t n_samples);

// grid packing (upload path) and gradient un-interleave
// missing (null) component: zero, or left as it is when keep_missing
cudaError_t launch_pack_grid(cudaStream_t s, const float* sigma, const float* color, float4* packed, size_t voxels,
                             bool keep_missing);
cudaError_t launch_pack_grid_strided(cudaStream_t s, const float* sigma, int32_t sigma_stride, const float* color,
                                     int32_t color_stride, float4* packed, size_t voxels, bool keep_missing);
cudaError_t launch_unpack_grad_slabs(cudaStream_t s, const float4* packed, float* sigma_grad, float* color_grad, int axis, uint32_t s0,
                                     uint32_t s1, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t sx, uint32_t sy, uint32_t sz);
cudaError_t launch_unpack_grad(cudaStream_t s, const float4* packed, float* sigma_grad, float* color_grad, size_t first,
                               size_t voxels, uint32_t nx, uint32_t ny, uint32_t sx, uint32_t sy, uint32_t sz);

}  // namespace dv
