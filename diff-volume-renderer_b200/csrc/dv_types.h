// dv_types.h -- POD parameter blocks shared by host and device code.
//
// Everything a kernel needs travels in one of these structs.  Values that can
// change between launches of a captured CUDA graph (camera, seed, ray-index
// base) live in FrameParams, which kernels read through a device pointer.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector_types.h>

namespace dv {

// Pinhole / orthographic camera as the reference's hp_camera_desc resolves it
// (reference hotpath/src/cpu/ray_cpu.cpp:158-176).
struct CameraParams {
    float fx, fy, cx, cy;
    float r00, r01, r02, r10, r11, r12, r20, r21, r22;
    float ox, oy, oz;
    uint32_t ortho;
};

// Marching parameters (reference hotpath/src/cpu/samp_cpu.cpp:197-199).
struct MarchParams {
    float t_near, t_far, dt;
    uint32_t max_steps;
    uint32_t stratified;
    uint32_t uniform_count;  // samples every generated ray emits (host-computed)
    uint64_t seed;
    uint64_t ray_index_base; // added to the plan-local ray index in the jitter hash
};

struct RoiParams {
    uint32_t x, y, w, h;     // region marched
    uint32_t img_w, img_h;   // full frame
    // CTA tile rows (kTileH * kWarpsY pixel rows each) this launch owns: rows t with t % stride == phase.  stride 1 = all.
    // Interleaving the tile rows of one frame over the GPUs of a box gives every rank the same mix of short and long
    // rays (strong scaling, diff-volume-renderer_b200/python/sharding.py).
    uint32_t tile_row_stride, tile_row_phase;
    // Order in which the CTAs (dispatched in blockIdx order) take the (owned) tile rows -- see tile_row_of():
    // 0 first-to-last; 1 last-to-first: a band of a sharded frame whose rays get LONGER towards its last row (the upper half
    // of a perspective image) would otherwise finish with its most expensive rows and idle SMs; 2 centre-out (middle row,
    // one below, one above, ...): the launch ends with the cheap outer rows of a full frame AND the gradient slabs become
    // final from the centre outwards right from the start (hpx_backward_streamed copies them while the kernel runs).
    // + kTileOrderColumns: the CTAs walk the tiles COLUMN by column (all rows of a column, in the row order above), the
    // columns centre-out -- a band of a few tile rows whose rows all cost the same (the middle of a sharded frame) has its
    // cheap tiles at the left and right end of every row; taking the columns from the middle outwards ends the launch with
    // them instead of with a last, partly empty wave of full-length tiles.
    uint32_t tile_row_reverse;
};
constexpr uint32_t kTileOrderColumns = 4u;

// Tile row taken by the i-th dispatched row of `rows` (RoiParams::tile_row_reverse = mode).
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t tile_row_of(uint32_t i, uint32_t rows, uint32_t mode) {
    if (mode == 1u) return rows - 1u - i;
    if (mode == 2u) {
        const uint32_t c = rows / 2u;
        if (i >= 2u * c) return i;                          // rows is odd: the lower half has one row more
        return (i & 1u) ? c - 1u - (i >> 1) : c + (i >> 1);
    }
    return i;
}

// Tile (column, owned row) the CTA `block` of a launch of tiles_x * rows CTAs takes (RoiParams::tile_row_reverse = mode).
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void tile_of(uint32_t block, uint32_t tiles_x, uint32_t rows, uint32_t mode, uint32_t* out_col, uint32_t* out_row) {
    if (mode & kTileOrderColumns) {
        *out_col = tile_row_of(block / rows, tiles_x, 2u);
        *out_row = tile_row_of(block % rows, rows, mode & 3u);
    } else {
        *out_col = block % tiles_x;
        *out_row = tile_row_of(block / tiles_x, rows, mode & 3u);
    }
}

struct FrameParams {
    CameraParams cam;
    MarchParams march;
    RoiParams roi;
};

// One dense grid as a field of the reference holds it
// (reference hotpath/src/runtime/hp_internal.hpp:24-31, grid_dense_cpu.cpp:17-36).
struct GridParams {
    const float* data;       // [nz][ny][nx][channels]
    int32_t nx, ny, nz, channels;
    uint32_t linear;         // HP_INTERP_LINEAR
    uint32_t clamp;          // HP_OOB_CLAMP
    uint32_t present;
};

// sigma + colour fields of one call.  `packed` (float4 {r,g,b,sigma} per voxel)
// is non-null when both grids have the same resolution / interpolation / OOB.
struct FieldPair {
    GridParams sigma;
    GridParams color;
    const float4* packed;
};

// Scatter target: packed gradient grid with the DenseGridField's bbox mapping
// (reference src/fields/dense_grid.cpp:198-230).
struct ScatterParams {
    float4* grad;            // [nz][ny][nx] {d r, d g, d b, d sigma}
    int32_t nx, ny, nz;
    uint32_t nearest;        // interp == NEAREST
    uint32_t clamp;
    uint32_t unit_bbox;      // bbox == [0,1]^3: the scatter sees the cube the forward pass sees
    float bmin[3], bmax[3];
    // deterministic mode (HPX_BACKWARD_DETERMINISTIC): contributions are rounded to multiples of a power-of-two
    // quantum and added with 64-bit INTEGER reds -- integer addition commutes, so the sum does not depend on the
    // order in which warps arrive and the gradient is bitwise reproducible.  fixed == nullptr: float reds.
    unsigned long long* fixed;   // [4 * voxels]
    const float* fixed_meta;     // {.., .., 1 / quantum, quantum} written by fixed_scale_kernel
    // Scatter target: the grid's gradient block (origin 0; strides of its layout -- z slowest by default, any axis
    // slowest after hpx_grid_set_grad_layout) or a dense box [bz][by][bx] of the grid (hpx_backward_box):
    // voxel (x,y,z) lives at (x - box_ox) * box_sx + (y - box_oy) * box_sy + (z - box_oz) * box_sz.
    int32_t box_ox, box_oy, box_oz;
    int32_t box_nx, box_ny, box_nz;
    uint32_t box_sx, box_sy, box_sz;
    uint32_t boxed;              // 1: contributions outside the box are dropped and counted in *box_miss
    unsigned int* box_miss;
};

constexpr int kSegment = 8;        // samples between transmittance checkpoints
constexpr int kTileW = 8;          // warp tile = 8 x 4 pixels
constexpr int kTileH = 4;
constexpr int kWarpsX = 2;         // CTA tile = 16 x 8 pixels (4 warps)
constexpr int kWarpsY = 2;
constexpr int kLeanThreads = 32 * kWarpsX * kWarpsY;
constexpr float kStopThreshold = 1e-4f;

}  // namespace dv
