// dv_lean.h -- host-side launch interface of the non-materialising kernels.
#pragma once

#include <cuda_runtime.h>

#include "dv_types.h"

namespace dv {

// Packed {r,g,b,sigma} grid resident in HBM.
struct PackedGrid {
    const float4* values = nullptr;
    const void* half_values = nullptr;   // HalfVoxel[V] instead of `values` (hpx_grid_set_storage): linear OOB-zero grids only
    int32_t nx = 0, ny = 0, nz = 0;
    bool linear = true;
    bool clamp = false;
    // Empty-space skipping (hpx_grid_build_occupancy): 2 bits per brick of 8^3 trilinear cells, 16 bricks per word.
    // bit 0: some voxel a cell of the brick can read has sigma != 0; bit 1: some such voxel has ANY channel != 0.
    // nullptr: no skipping.  Skipped samples contribute exactly nothing, so results do not change (see dv_lean.cu).
    const uint32_t* occ = nullptr;
    int32_t obx = 0, oby = 0;
};

// Device buffers one frame reads and writes (all owned by hpx_frame).
struct LeanBuffers {
    float* image = nullptr;        // [H][W][3]
    float* trans = nullptr;        // [H][W]
    float* opacity = nullptr;      // [H][W]
    float* depth = nullptr;        // [H][W]
    uint32_t* hitmask = nullptr;   // [H][W]
    uint32_t* live = nullptr;      // [rays] samples integrated before the stop
    float* ckpt = nullptr;         // [ceil(K / kSegment)][ckpt_stride] transmittance at segment starts
    size_t ckpt_stride = 0;
    unsigned long long* live_total = nullptr;
    // per-step table shared by every generated ray (they all carry the plan's t_near / t_far):
    // {base = t_near + step * dt, t of the fixed-mode sample, dt_actual, depth cursor before the step}
    // (reference samp_cpu.cpp:227-241, int_cpu.cpp:170-211), built on the host by hpx_frame_create
    const float4* steps = nullptr;
    // Device-side completion signal of the backward kernels (hpx_backward_signalled): CTAs are dispatched in tile-row
    // order; a CTA whose (owned) tile row lies in group g adds 1 to group_done[g] when all its reds are issued and
    // fenced.  A stream can then wait for a whole group of image rows (cuStreamWaitValue32) and all-reduce the gradient
    // slabs that group finished while later rows are still running -- one launch, no per-group launch tails.
    unsigned int* group_done = nullptr;
    uint32_t group_count = 0;
    uint32_t group_end[16] = {};                        // exclusive end (owned tile rows) of each group
};

// Fills `table[uniform_count]` with the per-step values above (plain IEEE float arithmetic, no contraction).
void build_step_table(const MarchParams& mp, float4* table);

uint32_t lean_block_count(const RoiParams& roi);

// d_bounds: 6 ints pre-set to {INT_MAX x3, INT_MIN x3}; receives min / max cell coordinates of the in-cube samples.
cudaError_t launch_ray_bounds(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params, int32_t nx,
                              int32_t ny, int32_t nz, int* d_bounds);

// Per slab along `slab_axis`: the inclusive range of rows (along `row_axis`) the backward of this launch can touch;
// d_lo / d_hi: [slabs] ints pre-set to INT_MAX / INT_MIN.
cudaError_t launch_ray_slab_rows(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params, int32_t nx, int32_t ny,
                                 int32_t nz, int slab_axis, int row_axis, int* d_lo, int* d_hi);

// Deterministic (fixed-point) gradient accumulation, see ScatterParams::fixed.  meta = {bits of max|rgb|, bits of
// max|dL/dI|, 1 / quantum, quantum}.
cudaError_t launch_abs_max(cudaStream_t stream, const float* d_values, size_t n, uint32_t* d_out_bits, bool packed_rgb_only = false);
cudaError_t launch_fixed_scale(cudaStream_t stream, float* d_meta, float dt);
cudaError_t launch_fixed_to_float(cudaStream_t stream, unsigned long long* d_fixed, float4* d_grad, size_t voxels,
                                  const float* d_meta);

// Measurement helpers: live in-cube samples of the last forward (OOB-zero fields; *d_total is accumulated into), and the
// number of voxels with a non-zero gradient.
cudaError_t launch_cube_count(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params,
                              const LeanBuffers& buf, unsigned long long* d_total);
cudaError_t launch_touched_voxels(cudaStream_t stream, const float4* d_grad, size_t voxels, unsigned long long* d_total);

// Occupancy bits of a packed grid (PackedGrid::occ); d_counts[0] / [1] receive the number of bricks without bit 0 / bit 1.
// Half storage (dv_device.cuh HalfVoxel): conversion both ways, parameter update, |rgb| maximum for the deterministic quantum.
cudaError_t launch_convert_storage(cudaStream_t stream, const float4* f32, void* half, size_t voxels, bool to_half);
cudaError_t launch_pack_grid_half(cudaStream_t stream, const float* sigma, const float* color, void* half, size_t voxels);
cudaError_t launch_abs_max_half(cudaStream_t stream, const void* half, size_t voxels, uint32_t* d_out_bits);

// values_are_half: `values` is really HalfVoxel[V]
cudaError_t launch_build_occupancy(cudaStream_t stream, const float4* values, int32_t nx, int32_t ny, int32_t nz, uint32_t* d_occ,
                                   size_t occ_words, unsigned int* d_counts, bool values_are_half = false);

// Writes the frame's parameter block; the values travel as kernel arguments (no staging buffer, no host sync).
cudaError_t launch_upload_params(cudaStream_t stream, FrameParams* d_params, const FrameParams& h_params);

cudaError_t launch_lean_forward(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params,
                                const PackedGrid& grid, const LeanBuffers& out, bool fill_background);

// scatter_mode: how the backward issues its gradient reds
enum : int { kScatterAuto = 0, kScatterPerRay = 1, kScatterMerge = 2 };

// kScatterPerRay or kScatterMerge: what launch_lean_backward will run for this request
int resolve_scatter_mode(const FrameParams& h_params, const PackedGrid& grid, const ScatterParams& sp, int scatter_mode);

cudaError_t launch_lean_backward(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params,
                                 const PackedGrid& grid, const ScatterParams& sp, const float* d_dL_dI,
                                 const LeanBuffers& state, int scatter_mode = kScatterAuto,
                                 double* cam_partials = nullptr, float* d_cam16 = nullptr);
// cam_partials != nullptr: the MERGED kernel also evaluates the camera adjoint from the corners it has loaded anyway and
// accumulates the 16 camera gradients into d_cam16 (only valid when resolve_scatter_mode() == kScatterMerge).

// d_partials: [lean_block_count][16] doubles of scratch; d_cam16 is accumulated into.
cudaError_t launch_camera_adjoint(cudaStream_t stream, const FrameParams* d_params, const FrameParams& h_params,
                                  const PackedGrid& grid, const float* d_dL_dI, const uint32_t* d_live,
                                  const float4* d_steps, double* d_partials, float* d_cam16);

}  // namespace dv
