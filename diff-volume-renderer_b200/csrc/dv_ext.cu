// dv_ext.cu -- hp_b200.h: packed device grids, per-plan frame workspaces, the
// lean forward / backward entry points and CUDA-graph capture; plus the
// hp_graph_* entry points of hp.h built on the same machinery.
#include <algorithm>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <new>

#include "dv_objects.h"

using namespace dv;

#define DV_TRY(expr)                                     \
    do {                                                 \
        const hp_status dv_st__ = (expr);                \
        if (dv_st__ != HP_STATUS_SUCCESS) return dv_st__; \
    } while (0)

namespace {

constexpr size_t kCameraFloats = 16;
constexpr uint32_t kMaxRowGroups = 16;   // LeanBuffers::group_end
constexpr size_t kStageVoxels = size_t(1) << 25;  // 32 Mi voxels per staging chunk (512 MiB)

hp_status grid_alloc(hpx_grid* g) {
    if (g->voxels > 0xffffffffull) {   // the kernels index voxels with 32 bits (dv_device.cuh: voxel_index32)
        set_last_error("packed device grids are limited to 2^32 - 1 voxels");
        return HP_STATUS_UNSUPPORTED;
    }
    DV_CUDA(cudaMalloc(&g->d_values, std::max<size_t>(g->voxels, 1) * sizeof(float4)));
    return HP_STATUS_SUCCESS;
}

hp_status grid_ensure_grad(hpx_grid* g) {
    if (g->d_grad != nullptr) return HP_STATUS_SUCCESS;
    const size_t floats = g->voxels * 4 + kCameraFloats;
    DV_CUDA(cudaMalloc(&g->d_grad, floats * sizeof(float)));
    DV_CUDA(cudaMemsetAsync(g->d_grad, 0, floats * sizeof(float), g->ctx->stream));
    return HP_STATUS_SUCCESS;
}

PackedGrid packed_view(const hpx_grid& g) {
    PackedGrid p;
    p.values = g.d_values;
    p.half_values = g.d_half;
    p.nx = g.nx; p.ny = g.ny; p.nz = g.nz;
    p.linear = g.linear;
    p.clamp = g.clamp;
    if (g.occ_ready && g.occ_enabled) {
        p.occ = g.d_occ;
        p.obx = (g.nx + 7) / 8;
        p.oby = (g.ny + 7) / 8;
    }
    return p;
}

void set_bbox(hpx_grid* g, const float bmin[3], const float bmax[3]) {
    for (int i = 0; i < 3; ++i) {
        g->bmin[i] = bmin ? bmin[i] : 0.0f;
        g->bmax[i] = bmax ? bmax[i] : 1.0f;
    }
}

}  // namespace

namespace dv {
hp_status grid_slabs_to_host(hpx_grid* g, cudaStream_t stream, int32_t lo, int32_t hi, float* sigma_host, float* color_host,
                             cudaStream_t color_stream, cudaEvent_t ev) {
    if (g == nullptr || lo < 0 || hi < lo) return HP_STATUS_INVALID_ARGUMENT;
    if (hi == lo || (sigma_host == nullptr && color_host == nullptr)) return HP_STATUS_SUCCESS;
    const int axis = g->grad_slow_axis;
    const int32_t n_axis = axis == 0 ? g->nx : axis == 1 ? g->ny : g->nz;
    if (hi > n_axis) return HP_STATUS_INVALID_ARGUMENT;
    // (Writing the host arrays straight from the un-interleave kernel -- zero copy over PCIe -- was measured at 29 GB/s
    // against 40 GB/s for staging + pitched copies and 54 GB/s for one contiguous copy: profiles/README.md, round 2.)
    if (g->d_unpacked == nullptr || g->unpacked_voxels < g->voxels) {   // full-size staging in the reference layout
        DV_CUDA(cudaStreamSynchronize(g->ctx->stream));
        cudaFree(g->d_unpacked);
        g->d_unpacked = nullptr;
        DV_CUDA(cudaMalloc(&g->d_unpacked, std::max<size_t>(g->voxels, 1) * 16));
        g->unpacked_voxels = g->voxels;
    }
    float* d_sig = g->d_unpacked;
    float* d_col = g->d_unpacked + g->unpacked_voxels;
    const ScatterParams lay = scatter_params(*g);
    const uint32_t nx = static_cast<uint32_t>(g->nx), ny = static_cast<uint32_t>(g->ny), nz = static_cast<uint32_t>(g->nz);
    DV_CUDA(launch_unpack_grad_slabs(stream, reinterpret_cast<const float4*>(g->d_grad), sigma_host ? d_sig : nullptr,
                                     color_host ? d_col : nullptr, axis, static_cast<uint32_t>(lo), static_cast<uint32_t>(hi), nx, ny, nz,
                                     lay.box_sx, lay.box_sy, lay.box_sz));
    // the slabs' voxels in the reference layout [z][y][x]: axis z: one contiguous run; axis y: per z-plane one run of
    // (hi - lo) rows; axis x: per row one run of (hi - lo) voxels -- all three are ONE pitched copy per array
    auto copy = [&](float* host, const float* dev, size_t ch, cudaStream_t stream) -> cudaError_t {
        if (axis == 2) {
            const size_t off = static_cast<size_t>(lo) * ny * nx * ch;
            return cudaMemcpyAsync(host + off, dev + off, static_cast<size_t>(hi - lo) * ny * nx * ch * 4, cudaMemcpyDeviceToHost, stream);
        }
        if (axis == 1) {
            const size_t off = static_cast<size_t>(lo) * nx * ch, pitch = static_cast<size_t>(ny) * nx * ch * 4;
            return cudaMemcpy2DAsync(host + off, pitch, dev + off, pitch, static_cast<size_t>(hi - lo) * nx * ch * 4, nz, cudaMemcpyDeviceToHost, stream);
        }
        const size_t off = static_cast<size_t>(lo) * ch, pitch = static_cast<size_t>(nx) * ch * 4;
        return cudaMemcpy2DAsync(host + off, pitch, dev + off, pitch, static_cast<size_t>(hi - lo) * ch * 4, static_cast<size_t>(ny) * nz,
                                 cudaMemcpyDeviceToHost, stream);
    };
    if (color_host != nullptr && color_stream != nullptr && ev != nullptr) {
        DV_CUDA(cudaEventRecord(ev, stream));
        DV_CUDA(cudaStreamWaitEvent(color_stream, ev, 0));
        DV_CUDA(copy(color_host, d_col, 3, color_stream));
        if (sigma_host != nullptr) DV_CUDA(copy(sigma_host, d_sig, 1, stream));
        return HP_STATUS_SUCCESS;
    }
    if (sigma_host != nullptr) DV_CUDA(copy(sigma_host, d_sig, 1, stream));
    if (color_host != nullptr) DV_CUDA(copy(color_host, d_col, 3, stream));
    return HP_STATUS_SUCCESS;
}
hp_status frame_slab_rows(hpx_frame* f, const hpx_grid* g, int32_t* out_lo, int32_t* out_hi) {
    if (f == nullptr || g == nullptr || out_lo == nullptr || out_hi == nullptr || f->ctx != g->ctx) return HP_STATUS_INVALID_ARGUMENT;
    DeviceScope scope;
    DV_TRY(scope.enter(f->ctx));
    DV_CUDA(launch_upload_params(f->ctx->stream, f->d_params, f->h_params));
    f->params_dirty = false;
    const int a = g->grad_slow_axis, b = a == 2 ? 1 : 2;   // [z][y][x] -> rows along y; [y][z][x] and [x][z][y] -> rows along z
    const int32_t dims[3] = {g->nx, g->ny, g->nz};
    const int32_t slabs = dims[a];
    cudaStream_t s = f->ctx->stream;
    DeviceScratch scratch;
    int* d = static_cast<int*>(scratch.take(static_cast<size_t>(slabs) * 2 * sizeof(int)));
    if (d == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    std::vector<int> init(static_cast<size_t>(slabs) * 2);
    for (int32_t i = 0; i < slabs; ++i) { init[static_cast<size_t>(i)] = INT_MAX; init[static_cast<size_t>(slabs + i)] = INT_MIN; }
    DV_CUDA(cudaMemcpyAsync(d, init.data(), init.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    DV_CUDA(launch_ray_slab_rows(s, f->d_params, f->h_params, g->nx, g->ny, g->nz, a, b, d, d + slabs));
    DV_CUDA(cudaMemcpyAsync(init.data(), d, init.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
    DV_CUDA(cudaStreamSynchronize(s));
    for (int32_t i = 0; i < slabs; ++i) {
        const int lo = init[static_cast<size_t>(i)], hi = init[static_cast<size_t>(slabs + i)];
        out_lo[i] = lo == INT_MAX ? 0 : lo;
        out_hi[i] = lo == INT_MAX ? 0 : hi + 1;
    }
    return HP_STATUS_SUCCESS;
}

hp_status frame_rows_bounds(hpx_frame* f, const hpx_grid* g, uint32_t row0, uint32_t rows, int32_t out_box[6]) {
    if (f == nullptr || g == nullptr || out_box == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    const RoiParams saved = f->h_params.roi;
    if (row0 >= saved.h || rows == 0) {
        for (int i = 0; i < 6; ++i) out_box[i] = 0;
        return HP_STATUS_SUCCESS;
    }
    f->h_params.roi.y = saved.y + row0;
    f->h_params.roi.h = std::min(rows, saved.h - row0);
    f->params_dirty = true;
    const hp_status st = hpx_frame_bounds(f, g, out_box);
    f->h_params.roi = saved;
    f->params_dirty = true;
    return st;
}
}  // namespace dv

extern "C" {

// =============================================================================
// hpx_grid
// =============================================================================
HP_API hp_status hpx_grid_create(const hp_ctx* ctx, const hp_field* fs, const hp_field* fc, const float bbox_min[3],
                                 const float bbox_max[3], hpx_grid** out_grid) {
    DV_RANGE("hpx_grid_create");
    if (ctx == nullptr || out_grid == nullptr || (fs == nullptr && fc == nullptr)) return HP_STATUS_INVALID_ARGUMENT;
    if ((fs && fs->kind != FieldKind::kDenseSigma) || (fc && fc->kind != FieldKind::kDenseColor))
        return HP_STATUS_INVALID_ARGUMENT;
    const hp_field* ref = fs ? fs : fc;
    if (fs && fc) {
        // one packed voxel serves both lookups only if they index identically
        if (fs->nx != fc->nx || fs->ny != fc->ny || fs->nz != fc->nz || fs->interp != fc->interp ||
            fs->oob != fc->oob || fc->channels != 3)
            return HP_STATUS_UNSUPPORTED;
    }
    DV_ENTER(ctx);
    hpx_grid* g = new (std::nothrow) hpx_grid();
    if (g == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    g->ctx = ctx_retain(ctx);
    g->nx = ref->nx; g->ny = ref->ny; g->nz = ref->nz;
    g->linear = ref->interp == HP_INTERP_LINEAR;
    g->clamp = ref->oob == HP_OOB_CLAMP;
    g->voxels = static_cast<size_t>(g->nx) * g->ny * g->nz;
    set_bbox(g, bbox_min, bbox_max);
    hp_status st = grid_alloc(g);
    if (st == HP_STATUS_SUCCESS) {
        // (fields that already view another packed grid have stride 4: unpack through the generic strided pack)
        const cudaError_t e = launch_pack_grid_strided(ctx->stream, fs ? fs->d_data : nullptr, fs ? fs->stride : 1,
                                                       fc ? fc->d_data : nullptr, fc ? fc->stride : 3, g->d_values, g->voxels, false);
        if (e != cudaSuccess) st = cuda_fail(e, "pack_grid");
    }
    if (st != HP_STATUS_SUCCESS) {
        hpx_grid_release(g);
        return st;
    }
    *out_grid = g;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_grid_update(hpx_grid* g, const float* sigma, const float* color, hp_memspace memspace) {
    DV_RANGE("hpx_grid_update");
    if (g != nullptr) g->value_max_stale = true;
    if (g == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (sigma == nullptr && color == nullptr) return HP_STATUS_SUCCESS;
    DV_ENTER(g->ctx);
    cudaStream_t s = g->ctx->stream;
    if (g->d_occ != nullptr) {
        // new values: the occupancy bits no longer describe them.  Every brick becomes "occupied" (no skipping, same
        // results) until hpx_grid_build_occupancy is called again; a captured graph that reads the bits stays correct.
        DV_CUDA(cudaMemsetAsync(g->d_occ, 0xff, g->occ_words * sizeof(uint32_t), s));
    }
    const bool half = g->d_half != nullptr;
    char* const half_base = static_cast<char*>(g->d_half);
    if (memspace == HP_MEMSPACE_DEVICE) {
        if (half) DV_CUDA(launch_pack_grid_half(s, sigma, color, half_base, g->voxels));
        else DV_CUDA(launch_pack_grid(s, sigma, color, g->d_values, g->voxels, true));
        return HP_STATUS_SUCCESS;
    }
    // HOST: stream the arrays through a bounded staging buffer and interleave on the device
    const size_t chunk = std::min<size_t>(g->voxels, kStageVoxels);
    DeviceScratch scratch;
    float* d_sig = sigma ? static_cast<float*>(scratch.take(chunk * 4)) : nullptr;
    float* d_col = color ? static_cast<float*>(scratch.take(chunk * 12)) : nullptr;
    if ((sigma && !d_sig) || (color && !d_col)) return HP_STATUS_OUT_OF_MEMORY;
    for (size_t off = 0; off < g->voxels; off += chunk) {
        const size_t n = std::min(chunk, g->voxels - off);
        if (sigma) DV_CUDA(cudaMemcpyAsync(d_sig, sigma + off, n * 4, cudaMemcpyHostToDevice, s));
        if (color) DV_CUDA(cudaMemcpyAsync(d_col, color + 3 * off, n * 12, cudaMemcpyHostToDevice, s));
        if (half) DV_CUDA(launch_pack_grid_half(s, d_sig, d_col, half_base + off * 8, n));
        else DV_CUDA(launch_pack_grid(s, d_sig, d_col, g->d_values + off, n, true));
    }
    DV_CUDA(cudaStreamSynchronize(s));  // the caller may free its buffers; scratch is released
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_grid_create_raw(const hp_ctx* ctx, int32_t nx, int32_t ny, int32_t nz, const float* sigma,
                                     const float* color, hp_memspace memspace, uint32_t interp, uint32_t oob,
                                     const float bbox_min[3], const float bbox_max[3], hpx_grid** out_grid) {
    DV_RANGE("hpx_grid_create_raw");
    if (ctx == nullptr || out_grid == nullptr || nx <= 0 || ny <= 0 || nz <= 0) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(ctx);
    hpx_grid* g = new (std::nothrow) hpx_grid();
    if (g == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    g->ctx = ctx_retain(ctx);
    g->nx = nx; g->ny = ny; g->nz = nz;
    g->linear = interp != static_cast<uint32_t>(HP_INTERP_NEAREST);
    g->clamp = oob == static_cast<uint32_t>(HP_OOB_CLAMP);
    g->voxels = static_cast<size_t>(nx) * ny * nz;
    set_bbox(g, bbox_min, bbox_max);
    hp_status st = grid_alloc(g);
    if (st == HP_STATUS_SUCCESS) {
        const cudaError_t e = cudaMemsetAsync(g->d_values, 0, g->voxels * sizeof(float4), ctx->stream);
        if (e != cudaSuccess) st = cuda_fail(e, "cudaMemsetAsync(grid)");
    }
    if (st == HP_STATUS_SUCCESS) st = hpx_grid_update(g, sigma, color, memspace);
    if (st != HP_STATUS_SUCCESS) {
        hpx_grid_release(g);
        return st;
    }
    *out_grid = g;
    return HP_STATUS_SUCCESS;
}

// Empty-space skipping.  Builds the occupancy bits of the CURRENT values (2 bits per brick of 8^3 trilinear cells) and
// turns skipping on when `enable` != 0: the forward kernel then skips samples whose eight corners all have sigma = 0, the
// backward kernels skip the gather of samples whose corners are zero in every channel.  Such samples contribute exactly
// nothing (alpha = 0, w = 0; their d sigma is still scattered), so images, counts and gradients are unchanged.
// Linear OOB-zero fields.  Blocks until the bits are built; out_*: fraction of bricks without sigma / without any value.
HP_API hp_status hpx_grid_build_occupancy(hpx_grid* g, int32_t enable, float* out_empty_sigma, float* out_empty_all) {
    DV_RANGE("hpx_grid_build_occupancy");
    if (g == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(g->ctx);
    cudaStream_t s = g->ctx->stream;
    const size_t bricks = static_cast<size_t>((g->nx + 7) / 8) * ((g->ny + 7) / 8) * ((g->nz + 7) / 8);
    if (g->d_occ == nullptr) {
        g->occ_words = (bricks + 15) / 16;
        DV_CUDA(cudaMalloc(&g->d_occ, g->occ_words * sizeof(uint32_t)));
        DV_CUDA(cudaMalloc(&g->d_occ_counts, 2 * sizeof(unsigned int)));
    }
    DV_CUDA(launch_build_occupancy(s, g->d_half != nullptr ? static_cast<const float4*>(g->d_half) : g->d_values, g->nx, g->ny, g->nz,
                                   g->d_occ, g->occ_words, g->d_occ_counts, g->d_half != nullptr));
    unsigned int counts[2] = {0, 0};
    DV_CUDA(cudaMemcpyAsync(counts, g->d_occ_counts, sizeof(counts), cudaMemcpyDeviceToHost, s));
    DV_CUDA(cudaStreamSynchronize(s));
    g->occ_ready = true;
    g->occ_enabled = enable != 0 && g->linear && !g->clamp;
    if (out_empty_sigma) *out_empty_sigma = static_cast<float>(counts[0]) / static_cast<float>(bricks);
    if (out_empty_all) *out_empty_all = static_cast<float>(counts[1]) / static_cast<float>(bricks);
    return HP_STATUS_SUCCESS;
}

// Storage precision of the packed VALUES (the gradient block stays fp32).  HPX_STORAGE_F16 keeps four IEEE halfs per voxel
// (8 B instead of 16: half the gather bytes, half the HBM footprint -- a 1024^3 grid drops from 17.2 GB to 8.6 GB).  The
// values are rounded to half ONCE (round to nearest even); a voxel is widened back to fp32 exactly when it is loaded and
// all arithmetic stays the fp32 code of dv_device.cuh, so the grid renders and differentiates exactly like an fp32 grid
// that holds the rounded values (that grid, through the oracle, is the parity twin: tests/test_gpu_runtime.py).
// Linear OOB-zero grids with the unit scatter box and no adopted hp_fields (the staged hp.h calls read fp32 views).
HP_API hp_status hpx_grid_set_storage(hpx_grid* g, uint32_t storage) {
    DV_RANGE("hpx_grid_set_storage");
    if (g == nullptr || (storage != HPX_STORAGE_F32 && storage != HPX_STORAGE_F16)) return HP_STATUS_INVALID_ARGUMENT;
    const bool want_half = storage == HPX_STORAGE_F16;
    if (want_half == (g->d_half != nullptr)) return HP_STATUS_SUCCESS;
    if (want_half && (!g->linear || g->clamp || scatter_params(*g).unit_bbox == 0u || !g->views.empty())) {
        set_last_error("half storage needs a linear OOB-zero grid with the unit scatter box and no adopted fields");
        return HP_STATUS_UNSUPPORTED;
    }
    DV_ENTER(g->ctx);
    cudaStream_t s = g->ctx->stream;
    if (want_half) {
        DV_CUDA(cudaMalloc(&g->d_half, std::max<size_t>(g->voxels, 1) * 8));
        DV_CUDA(launch_convert_storage(s, g->d_values, g->d_half, g->voxels, true));
        DV_CUDA(cudaStreamSynchronize(s));
        cudaFree(g->d_values);
        g->d_values = nullptr;
    } else {
        DV_CUDA(cudaMalloc(&g->d_values, std::max<size_t>(g->voxels, 1) * sizeof(float4)));
        DV_CUDA(launch_convert_storage(s, g->d_values, g->d_half, g->voxels, false));
        DV_CUDA(cudaStreamSynchronize(s));
        cudaFree(g->d_half);
        g->d_half = nullptr;
    }
    g->value_max_stale = true;
    return HP_STATUS_SUCCESS;
}

// Turns skipping on / off without rebuilding (off: the kernels without the occupancy test run).
HP_API hp_status hpx_grid_set_occupancy(hpx_grid* g, int32_t enable) {
    DV_RANGE("hpx_grid_set_occupancy");
    if (g == nullptr || (enable != 0 && !g->occ_ready)) return HP_STATUS_INVALID_ARGUMENT;
    g->occ_enabled = enable != 0 && g->linear && !g->clamp;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_grid_zero_grad(hpx_grid* g) {
    DV_RANGE("hpx_grid_zero_grad");
    if (g == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(g->ctx);
    if (g->d_grad == nullptr) return grid_ensure_grad(g);
    DV_CUDA(cudaMemsetAsync(g->d_grad, 0, (g->voxels * 4 + kCameraFloats) * sizeof(float), g->ctx->stream));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_grid_grad_buffer(hpx_grid* g, float** out_device_ptr, size_t* out_floats) {
    DV_RANGE("hpx_grid_grad_buffer");
    if (g == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(g->ctx);
    DV_TRY(grid_ensure_grad(g));
    if (out_device_ptr) *out_device_ptr = g->d_grad;
    if (out_floats) *out_floats = g->voxels * 4 + kCameraFloats;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_grid_read_grad(hpx_grid* g, float* sigma_grad, float* color_grad, float* camera16,
                                    hp_memspace memspace) {
    DV_RANGE("hpx_grid_read_grad");
    if (g == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    return hpx_grid_read_grad_range(g, 0, g->voxels, sigma_grad, color_grad, camera16, memspace);
}

// Voxels [first, first + count) of the REFERENCE order (z slowest, x fastest): sigma_grad[count], color_grad[3 * count].
// Lets every rank of a sharded job hand ITS share of the summed gradient to the host over its own PCIe link.
HP_API hp_status hpx_grid_read_grad_range(hpx_grid* g, size_t first, size_t count, float* sigma_grad, float* color_grad,
                                          float* camera16, hp_memspace memspace) {
    DV_RANGE("hpx_grid_read_grad_range");
    if (g == nullptr || first > g->voxels || count > g->voxels - first) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(g->ctx);
    DV_TRY(grid_ensure_grad(g));
    cudaStream_t s = g->ctx->stream;
    const float4* packed = reinterpret_cast<const float4*>(g->d_grad);
    const ScatterParams lay = scatter_params(*g);   // strides of the gradient block
    const uint32_t unx = static_cast<uint32_t>(g->nx), uny = static_cast<uint32_t>(g->ny);
    if (memspace == HP_MEMSPACE_DEVICE) {
        DV_CUDA(launch_unpack_grad(s, packed, sigma_grad, color_grad, first, count, unx, uny, lay.box_sx, lay.box_sy, lay.box_sz));
        if (camera16) DV_CUDA(cudaMemcpyAsync(camera16, g->d_grad + g->voxels * 4, kCameraFloats * sizeof(float),
                                              cudaMemcpyDeviceToDevice, s));
        return HP_STATUS_SUCCESS;
    }
    // HOST: un-interleave chunk by chunk through a staging buffer the grid keeps (no per-call malloc)
    const size_t chunk = std::min<size_t>(std::max<size_t>(count, 1), kStageVoxels);
    if (g->d_unpacked == nullptr || g->unpacked_voxels < chunk) {
        cudaFree(g->d_unpacked);
        g->d_unpacked = nullptr;
        DV_CUDA(cudaMalloc(&g->d_unpacked, chunk * 16));
        g->unpacked_voxels = chunk;
    }
    float* d_sig = sigma_grad ? g->d_unpacked : nullptr;
    float* d_col = color_grad ? g->d_unpacked + g->unpacked_voxels : nullptr;
    for (size_t off = 0; off < count && (sigma_grad || color_grad); off += chunk) {
        const size_t n = std::min(chunk, count - off);
        DV_CUDA(launch_unpack_grad(s, packed, d_sig, d_col, first + off, n, unx, uny, lay.box_sx, lay.box_sy, lay.box_sz));
        if (sigma_grad) DV_CUDA(cudaMemcpyAsync(sigma_grad + off, d_sig, n * 4, cudaMemcpyDeviceToHost, s));
        if (color_grad) DV_CUDA(cudaMemcpyAsync(color_grad + 3 * off, d_col, n * 12, cudaMemcpyDeviceToHost, s));
    }
    if (camera16) DV_CUDA(cudaMemcpyAsync(camera16, g->d_grad + g->voxels * 4, kCameraFloats * sizeof(float),
                                          cudaMemcpyDeviceToHost, s));
    DV_CUDA(cudaStreamSynchronize(s));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_grid_accumulate_samples(hpx_grid* g, const float* positions, const float* grad_sigma,
                                             const float* grad_color, size_t count, hp_memspace memspace) {
    DV_RANGE("hpx_grid_accumulate_samples");
    if (g == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (count == 0) return HP_STATUS_SUCCESS;
    if (positions == nullptr || grad_sigma == nullptr || grad_color == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(g->ctx);
    DV_TRY(grid_ensure_grad(g));
    cudaStream_t s = g->ctx->stream;
    if (memspace == HP_MEMSPACE_DEVICE) {
        DV_CUDA(launch_scatter(s, scatter_params(*g), positions, grad_sigma, grad_color, count));
        return HP_STATUS_SUCCESS;
    }
    const size_t chunk = std::min<size_t>(count, size_t(1) << 24);
    DeviceScratch scratch;
    float* d_pos = static_cast<float*>(scratch.take(chunk * 12));
    float* d_gs = static_cast<float*>(scratch.take(chunk * 4));
    float* d_gc = static_cast<float*>(scratch.take(chunk * 12));
    if (!d_pos || !d_gs || !d_gc) return HP_STATUS_OUT_OF_MEMORY;
    for (size_t off = 0; off < count; off += chunk) {
        const size_t n = std::min(chunk, count - off);
        DV_CUDA(cudaMemcpyAsync(d_pos, positions + 3 * off, n * 12, cudaMemcpyHostToDevice, s));
        DV_CUDA(cudaMemcpyAsync(d_gs, grad_sigma + off, n * 4, cudaMemcpyHostToDevice, s));
        DV_CUDA(cudaMemcpyAsync(d_gc, grad_color + 3 * off, n * 12, cudaMemcpyHostToDevice, s));
        DV_CUDA(launch_scatter(s, scatter_params(*g), d_pos, d_gs, d_gc, n));
    }
    DV_CUDA(cudaStreamSynchronize(s));
    return HP_STATUS_SUCCESS;
}

HP_API void hpx_grid_release(hpx_grid* g) {
    DV_RANGE("hpx_grid_release");
    if (g == nullptr) return;
    for (hp_field* v : g->views) {   // adopted fields lose their values with the grid: later queries fail cleanly
        v->d_data = nullptr;
        v->alias_of = nullptr;
    }
    if (g->ctx != nullptr && g->ctx->ready) {
        DeviceScope scope;
        scope.enter(g->ctx);
        cudaStreamSynchronize(g->ctx->stream);
        cudaFree(g->d_values);
        cudaFree(g->d_half);
        cudaFree(g->d_grad);
        cudaFree(g->d_fixed);
        cudaFree(g->d_fixed_meta);
        cudaFree(g->d_occ);
        cudaFree(g->d_occ_counts);
        cudaFree(g->d_unpacked);
    }
    ctx_unref(g->ctx);
    delete g;
}

// One copy of the values in HBM: the two fields the grid was built from drop their own snapshots and become strided
// views of the packed {r,g,b,sigma} voxels, so hpx_grid_update is what the staged hp_samp / hp_graph paths see too.
HP_API hp_status hpx_grid_adopt_fields(hpx_grid* g, hp_field* fs, hp_field* fc) {
    DV_RANGE("hpx_grid_adopt_fields");
    if (g == nullptr || (fs == nullptr && fc == nullptr)) return HP_STATUS_INVALID_ARGUMENT;
    if ((fs && fs->kind != FieldKind::kDenseSigma) || (fc && fc->kind != FieldKind::kDenseColor)) return HP_STATUS_INVALID_ARGUMENT;
    if (g->d_half != nullptr) {
        set_last_error("a grid stored as halfs cannot back hp_fields (the staged hp.h calls read fp32 views)");
        return HP_STATUS_UNSUPPORTED;
    }
    for (hp_field* f : {fs, fc}) {
        if (f == nullptr) continue;
        if (f->ctx != g->ctx || f->nx != g->nx || f->ny != g->ny || f->nz != g->nz ||
            (f->interp == HP_INTERP_LINEAR) != g->linear || (f->oob == HP_OOB_CLAMP) != g->clamp ||
            (f->alias_of != nullptr && f->alias_of != g))
            return HP_STATUS_INVALID_ARGUMENT;
    }
    DV_ENTER(g->ctx);
    DV_CUDA(cudaStreamSynchronize(g->ctx->stream));   // the pack that read the snapshots has finished
    for (hp_field* f : {fs, fc}) {
        if (f == nullptr || f->alias_of == g) continue;
        if (f->owns_data) cudaFree(f->d_data);
        f->owns_data = false;
        f->alias_of = g;
        f->stride = 4;
        f->d_data = reinterpret_cast<float*>(g->d_values) + (f->kind == FieldKind::kDenseSigma ? 3 : 0);
        g->views.push_back(f);
    }
    return HP_STATUS_SUCCESS;
}

// =============================================================================
// hpx_frame: the per-plan workspace planner
// =============================================================================
static void* frame_take(hpx_frame* f, size_t bytes, hp_status* st) {
    if (*st != HP_STATUS_SUCCESS) return nullptr;
    void* p = nullptr;
    const cudaError_t e = cudaMalloc(&p, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) {
        *st = cuda_fail(e, "cudaMalloc(frame)");
        return nullptr;
    }
    f->allocations.push_back(p);
    f->device_bytes += bytes;
    return p;
}

HP_API hp_status hpx_frame_create(const hp_plan* plan, hpx_frame** out_frame) {
    DV_RANGE("hpx_frame_create");
    if (plan == nullptr || out_frame == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (!plan->gap_free) {
        set_last_error("plan skips marching steps (dt below float resolution); use the materialising hp_* path");
        return HP_STATUS_UNSUPPORTED;
    }
    DV_ENTER(plan->ctx);
    hpx_frame* f = new (std::nothrow) hpx_frame();
    if (f == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    f->plan = plan_retain(plan);
    f->ctx = ctx_retain(plan->ctx);
    f->h_params = frame_params_from_plan(*plan);
    const hp_plan_desc& d = plan->desc;
    const size_t pixels = static_cast<size_t>(d.width) * d.height;
    const size_t rays = static_cast<size_t>(d.roi.width) * d.roi.height;
    f->rays = rays;
    f->samples = static_cast<uint64_t>(rays) * plan->uniform_count;
    const size_t segments = (plan->uniform_count + kSegment - 1) / kSegment;
    hp_status st = HP_STATUS_SUCCESS;
    f->d_params = static_cast<FrameParams*>(frame_take(f, sizeof(FrameParams), &st));
    f->buf.image = static_cast<float*>(frame_take(f, pixels * 12, &st));
    f->buf.trans = static_cast<float*>(frame_take(f, pixels * 4, &st));
    f->buf.opacity = static_cast<float*>(frame_take(f, pixels * 4, &st));
    f->buf.depth = static_cast<float*>(frame_take(f, pixels * 4, &st));
    f->buf.hitmask = static_cast<uint32_t*>(frame_take(f, pixels * 4, &st));
    f->buf.live = static_cast<uint32_t*>(frame_take(f, rays * 4, &st));
    f->buf.ckpt_stride = (rays + 31) & ~static_cast<size_t>(31);
    f->buf.ckpt = static_cast<float*>(frame_take(f, segments * f->buf.ckpt_stride * 4, &st));
    f->buf.live_total = static_cast<unsigned long long*>(frame_take(f, sizeof(unsigned long long), &st));
    f->d_dL_dI = static_cast<float*>(frame_take(f, rays * 12, &st));
    f->d_box_miss = static_cast<unsigned int*>(frame_take(f, sizeof(unsigned int), &st));
    f->d_group_done = static_cast<unsigned int*>(frame_take(f, kMaxRowGroups * sizeof(unsigned int), &st));
    if (st == HP_STATUS_SUCCESS) {
        const cudaError_t e = cudaMemset(f->d_box_miss, 0, sizeof(unsigned int));
        if (e != cudaSuccess) st = cuda_fail(e, "cudaMemset(frame)");
    }
    float4* d_steps = static_cast<float4*>(frame_take(f, static_cast<size_t>(plan->uniform_count) * sizeof(float4), &st));
    f->buf.steps = d_steps;
    if (st == HP_STATUS_SUCCESS && plan->uniform_count != 0) {
        std::vector<float4> table(plan->uniform_count);
        build_step_table(f->h_params.march, table.data());
        const cudaError_t e = cudaMemcpy(d_steps, table.data(), table.size() * sizeof(float4), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) st = cuda_fail(e, "cudaMemcpy(step table)");
    }
    f->d_cam_partials = static_cast<double*>(frame_take(f, static_cast<size_t>(lean_block_count(f->h_params.roi)) * 16 * 8, &st));
    if (st == HP_STATUS_SUCCESS) {
        const cudaError_t e = cudaMallocHost(&f->h_pinned, sizeof(FrameParams) + sizeof(unsigned long long));
        if (e != cudaSuccess) st = cuda_fail(e, "cudaMallocHost(frame)");
    }
    if (st != HP_STATUS_SUCCESS) {
        hpx_frame_release(f);
        return st;
    }
    f->params_dirty = true;
    *out_frame = f;
    return HP_STATUS_SUCCESS;
}

HP_API size_t hpx_frame_bytes(const hpx_frame* f) { DV_RANGE("hpx_frame_bytes"); return f ? f->device_bytes : 0; }

// Host-only: the per-step table a frame of this plan uploads (no device needed) -- lets the CPU test suite pin the
// ray-independent part of the marching loop against the oracle.
HP_API hp_status hpx_plan_step_table(const hp_plan* plan, float* out_steps4, size_t capacity_steps, uint32_t* out_count) {
    DV_RANGE("hpx_plan_step_table");
    if (plan == nullptr || out_count == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    *out_count = plan->uniform_count;
    if (!plan->gap_free) return HP_STATUS_UNSUPPORTED;
    if (out_steps4 == nullptr) return HP_STATUS_SUCCESS;
    if (capacity_steps < plan->uniform_count) return HP_STATUS_OUT_OF_MEMORY;
    const FrameParams p = frame_params_from_plan(*plan);
    build_step_table(p.march, reinterpret_cast<float4*>(out_steps4));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_frame_set_view(hpx_frame* f, const hp_camera_desc* camera, uint64_t seed,
                                    uint64_t ray_index_base) {
    DV_RANGE("hpx_frame_set_view");
    if (f == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (camera != nullptr) {
        hp_plan_desc tmp = f->plan->desc;   // apply the plan-time camera defaults to the new camera too
        tmp.camera = *camera;
        DV_TRY(resolve_plan_desc(&tmp));
        f->h_params.cam = camera_params(tmp.camera);
    }
    f->h_params.march.seed = seed;
    f->h_params.march.ray_index_base = ray_index_base;
    f->params_dirty = true;
    return HP_STATUS_SUCCESS;
}

static hp_status frame_push_params(hpx_frame* f) {
    if (!f->params_dirty) return HP_STATUS_SUCCESS;
    // The block travels as a kernel ARGUMENT (copied at launch), so there is no staging buffer to protect and no host
    // synchronisation: a loop that changes the view before every graph replay never blocks (config 4: 64 views).
    DV_CUDA(launch_upload_params(f->ctx->stream, f->d_params, f->h_params));
    f->params_dirty = false;
    return HP_STATUS_SUCCESS;
}

static hp_status frame_check_grid(const hpx_frame* f, const hpx_grid* g) {
    if (f == nullptr || g == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    // ONE context = one stream: the grid's own memsets / packing (hpx_grid_zero_grad, hpx_grid_update, the lazy gradient
    // block) and the frame's kernels are ordered only because they share it
    if (f->ctx != g->ctx) {
        set_last_error("frame and grid belong to different contexts (streams): create both from the same hp_ctx");
        return HP_STATUS_INVALID_ARGUMENT;
    }
    return HP_STATUS_SUCCESS;
}

static hp_status enqueue_forward(hpx_frame* f, const hpx_grid* g) {
    cudaStream_t s = f->ctx->stream;
    DV_CUDA(cudaMemsetAsync(f->buf.live_total, 0, sizeof(unsigned long long), s));
    const RoiParams& roi = f->h_params.roi;
    const bool partial = roi.w != roi.img_w || roi.h != roi.img_h || roi.tile_row_stride > 1;
    DV_CUDA(launch_lean_forward(s, f->d_params, f->h_params, packed_view(*g), f->buf, partial));
    return HP_STATUS_SUCCESS;
}

struct GradBox {
    float* data = nullptr;   // [bz][by][bx][4], DEVICE
    int32_t o[3] = {0, 0, 0}, n[3] = {0, 0, 0};
};

// Fixed-point shadow of the gradient grid (HPX_BACKWARD_DETERMINISTIC).  Allocation and the first clear happen HERE,
// outside any stream capture: cudaMalloc inside a capture invalidates it.
static hp_status grid_ensure_fixed(hpx_grid* g) {
    if (g->d_fixed != nullptr) return HP_STATUS_SUCCESS;
    DV_CUDA(cudaMalloc(&g->d_fixed, std::max<size_t>(g->voxels, 1) * 4 * sizeof(unsigned long long)));
    DV_CUDA(cudaMalloc(&g->d_fixed_meta, 4 * sizeof(float)));
    DV_CUDA(cudaMemsetAsync(g->d_fixed, 0, std::max<size_t>(g->voxels, 1) * 4 * sizeof(unsigned long long), g->ctx->stream));
    DV_CUDA(cudaMemsetAsync(g->d_fixed_meta, 0, 4 * sizeof(float), g->ctx->stream));
    g->value_max_stale = true;
    return HP_STATUS_SUCCESS;
}

// capturing: the launches are being recorded into a CUDA graph that will be replayed after hpx_grid_update calls this
// function never sees -- so the |value| maximum is ALWAYS part of the graph and the host-side staleness flag is left alone.
static hp_status enqueue_backward(hpx_frame* f, hpx_grid* g, const float* d_dL_dI, uint32_t flags,
                                  const GradBox* box = nullptr, bool capturing = false) {
    cudaStream_t s = f->ctx->stream;
    if (flags & HPX_BACKWARD_ZERO) {
        if (box != nullptr) {   // the caller's box is the grid-gradient target; the grid keeps the camera slots
            DV_CUDA(cudaMemsetAsync(box->data, 0, static_cast<size_t>(box->n[0]) * box->n[1] * box->n[2] * 16, s));
            DV_CUDA(cudaMemsetAsync(g->d_grad + g->voxels * 4, 0, kCameraFloats * sizeof(float), s));
        } else {
            DV_CUDA(cudaMemsetAsync(g->d_grad, 0, (g->voxels * 4 + kCameraFloats) * sizeof(float), s));
        }
    }
    const int want = (flags & HPX_BACKWARD_SCATTER_MERGED)    ? kScatterMerge
                     : (flags & HPX_BACKWARD_SCATTER_PER_RAY) ? kScatterPerRay
                                                              : kScatterAuto;
    float* const d_cam16 = g->d_grad + g->voxels * 4;
    // grid + camera gradients in ONE pass when the merged kernel runs (it holds the corners the camera adjoint needs)
    const bool fuse_camera = (flags & HPX_BACKWARD_GRID) && (flags & HPX_BACKWARD_CAMERA) && g->linear &&
                             resolve_scatter_mode(f->h_params, packed_view(*g), scatter_params(*g), want) == kScatterMerge;
    if (flags & HPX_BACKWARD_GRID) {
        ScatterParams sp = scatter_params(*g);
        if (box != nullptr) {
            sp.box_ox = box->o[0]; sp.box_oy = box->o[1]; sp.box_oz = box->o[2];
            sp.box_nx = box->n[0]; sp.box_ny = box->n[1]; sp.box_nz = box->n[2];
            sp.box_sx = 1u;
            sp.box_sy = static_cast<uint32_t>(box->n[0]);
            sp.box_sz = static_cast<uint32_t>(box->n[0]) * static_cast<uint32_t>(box->n[1]);
            // biased base: voxel (x,y,z) of the GRID lives at grad[x + y * sy + z * sz] (never dereferenced outside the box)
            sp.grad = reinterpret_cast<float4*>(box->data) -
                      (static_cast<ptrdiff_t>(box->o[0]) + static_cast<ptrdiff_t>(box->o[1]) * sp.box_sy +
                       static_cast<ptrdiff_t>(box->o[2]) * sp.box_sz);
            sp.boxed = 1u;
            sp.box_miss = f->d_box_miss;
        }
        const bool deterministic = (flags & HPX_BACKWARD_DETERMINISTIC) != 0 && box == nullptr;
        if (deterministic) {
            if (g->d_fixed == nullptr) {
                if (capturing) {
                    set_last_error("deterministic backward: the fixed-point grid must exist before the capture begins");
                    return HP_STATUS_INTERNAL_ERROR;
                }
                DV_TRY(grid_ensure_fixed(g));
            }
            uint32_t* meta_bits = reinterpret_cast<uint32_t*>(g->d_fixed_meta);
            if (g->value_max_stale || capturing) {
                DV_CUDA(cudaMemsetAsync(meta_bits, 0, sizeof(uint32_t), s));
                if (g->d_half != nullptr) DV_CUDA(launch_abs_max_half(s, g->d_half, g->voxels, meta_bits));
                else DV_CUDA(launch_abs_max(s, reinterpret_cast<const float*>(g->d_values), g->voxels * 4, meta_bits, true));
                if (!capturing) g->value_max_stale = false;
            }
            DV_CUDA(cudaMemsetAsync(meta_bits + 1, 0, sizeof(uint32_t), s));
            DV_CUDA(launch_abs_max(s, d_dL_dI, static_cast<size_t>(f->h_params.roi.w) * f->h_params.roi.h * 3, meta_bits + 1));
            DV_CUDA(launch_fixed_scale(s, g->d_fixed_meta, f->h_params.march.dt));
            sp.fixed = g->d_fixed;
            sp.fixed_meta = g->d_fixed_meta;
        }
        DV_CUDA(launch_lean_backward(s, f->d_params, f->h_params, packed_view(*g), sp, d_dL_dI, f->buf, want,
                                     fuse_camera ? f->d_cam_partials : nullptr, fuse_camera ? d_cam16 : nullptr));
        if (deterministic)
            DV_CUDA(launch_fixed_to_float(s, g->d_fixed, reinterpret_cast<float4*>(g->d_grad), g->voxels, g->d_fixed_meta));
    }
    if ((flags & HPX_BACKWARD_CAMERA) && !fuse_camera)
        DV_CUDA(launch_camera_adjoint(s, f->d_params, f->h_params, packed_view(*g), d_dL_dI, f->buf.live,
                                      f->buf.steps, f->d_cam_partials, d_cam16));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_forward(hpx_frame* f, const hpx_grid* g) {
    DV_RANGE("hpx_forward");
    DV_TRY(frame_check_grid(f, g));
    DV_ENTER(f->ctx);
    DV_TRY(frame_push_params(f));
    DV_TRY(enqueue_forward(f, g));
    f->forward_done = true;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_backward(hpx_frame* f, hpx_grid* g, const float* dL_dI, hp_memspace memspace, uint32_t flags) {
    DV_RANGE("hpx_backward");
    DV_TRY(frame_check_grid(f, g));
    if (dL_dI == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (!f->forward_done) {
        set_last_error("hpx_backward needs the checkpoints of a preceding hpx_forward on the same frame");
        return HP_STATUS_INVALID_ARGUMENT;
    }
    DV_ENTER(f->ctx);
    DV_TRY(grid_ensure_grad(g));
    DV_TRY(frame_push_params(f));
    const float* d_g = dL_dI;
    if (memspace == HP_MEMSPACE_HOST) {
        DV_CUDA(cudaMemcpyAsync(f->d_dL_dI, dL_dI, static_cast<size_t>(f->h_params.roi.w) * f->h_params.roi.h * 12,
                                cudaMemcpyHostToDevice, f->ctx->stream));
        d_g = f->d_dL_dI;
    }
    return enqueue_backward(f, g, d_g, flags);
}

HP_API hp_status hpx_frame_set_interleave(hpx_frame* f, uint32_t stride, uint32_t phase) {
    DV_RANGE("hpx_frame_set_interleave");
    if (f == nullptr || stride == 0 || phase >= stride) return HP_STATUS_INVALID_ARGUMENT;
    RoiParams& roi = f->h_params.roi;
    roi.tile_row_stride = stride;
    roi.tile_row_phase = phase;
    // rays / samples this frame now marches (hpx_frame_counts)
    const uint32_t tile_h = kTileH * kWarpsY;
    uint64_t rows = 0;
    for (uint32_t t = phase, tiles = (roi.h + tile_h - 1) / tile_h; t < tiles; t += stride)
        rows += std::min<uint32_t>(tile_h, roi.h - t * tile_h);
    f->rays = rows * roi.w;
    f->samples = f->rays * f->plan->uniform_count;
    f->params_dirty = true;
    // a captured graph has the old CTA count / background decision baked in: replay must fail until it is recaptured
    if (f->graph_exec != nullptr || f->graph != nullptr) {
        DV_ENTER(f->ctx);
        DV_CUDA(cudaStreamSynchronize(f->ctx->stream));
        if (f->graph_exec) cudaGraphExecDestroy(f->graph_exec);
        if (f->graph) cudaGraphDestroy(f->graph);
        f->graph_exec = nullptr;
        f->graph = nullptr;
    }
    return HP_STATUS_SUCCESS;
}

// Order in which the CTAs of this frame's launches take its tile rows (RoiParams::tile_row_reverse).  Results do not
// depend on it.  Not combined with hpx_backward_signalled, whose row groups count in dispatch order.
HP_API hp_status hpx_frame_set_row_order(hpx_frame* f, int32_t order) {
    DV_RANGE("hpx_frame_set_row_order");
    if (f == nullptr || order < 0 || (order & ~(HPX_ORDER_COLUMNS | 3)) != 0 || (order & 3) == 3) return HP_STATUS_INVALID_ARGUMENT;
    static_assert(HPX_ORDER_COLUMNS == kTileOrderColumns, "hp_b200.h: HPX_ORDER_COLUMNS");
    f->h_params.roi.tile_row_reverse = static_cast<uint32_t>(order);
    f->params_dirty = true;
    return HP_STATUS_SUCCESS;
}

// Host-only: the tile (column, row) the i-th dispatched CTA of a launch of tiles_x * rows CTAs takes under `order`.
HP_API hp_status hpx_tile_order(uint32_t block, uint32_t tiles_x, uint32_t rows, int32_t order, uint32_t* out_col, uint32_t* out_row) {
    DV_RANGE("hpx_tile_order");
    if (out_col == nullptr || out_row == nullptr || tiles_x == 0 || rows == 0 || block / tiles_x >= rows || order < 0 ||
        (order & ~(HPX_ORDER_COLUMNS | 3)) != 0 || (order & 3) == 3)
        return HP_STATUS_INVALID_ARGUMENT;
    tile_of(block, tiles_x, rows, static_cast<uint32_t>(order), out_col, out_row);
    return HP_STATUS_SUCCESS;
}

// Host-only: the tile row the i-th dispatched row of `rows` takes under hpx_frame_set_row_order(order) (dv_types.h tile_row_of).
HP_API hp_status hpx_tile_row_order(uint32_t i, uint32_t rows, int32_t order, uint32_t* out_row) {
    DV_RANGE("hpx_tile_row_order");
    if (out_row == nullptr || i >= rows || order < 0 || order > 2) return HP_STATUS_INVALID_ARGUMENT;
    *out_row = tile_row_of(i, rows, static_cast<uint32_t>(order));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_frame_bounds(hpx_frame* f, const hpx_grid* g, int32_t out_box[6]) {
    DV_RANGE("hpx_frame_bounds");
    DV_TRY(frame_check_grid(f, g));
    if (out_box == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(f->ctx);
    DV_TRY(frame_push_params(f));
    cudaStream_t s = f->ctx->stream;
    DeviceScratch scratch;
    int* d_bounds = static_cast<int*>(scratch.take(6 * sizeof(int)));
    if (d_bounds == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    const int init[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    DV_CUDA(cudaMemcpyAsync(d_bounds, init, sizeof(init), cudaMemcpyHostToDevice, s));
    DV_CUDA(launch_ray_bounds(s, f->d_params, f->h_params, g->nx, g->ny, g->nz, d_bounds));
    int got[6];
    DV_CUDA(cudaMemcpyAsync(got, d_bounds, sizeof(got), cudaMemcpyDeviceToHost, s));
    DV_CUDA(cudaStreamSynchronize(s));
    const int32_t dims[3] = {g->nx, g->ny, g->nz};
    for (int i = 0; i < 3; ++i) {
        if (got[i] == INT_MAX) {   // no ray of this frame enters the cube: empty box at the origin
            out_box[i] = 0;
            out_box[3 + i] = 0;
            continue;
        }
        const int32_t lo = std::max(0, got[i] - 1), hi = std::min(dims[i] - 1, got[3 + i] + 2);   // +1 upper corner, +-1 slack
        out_box[i] = lo;
        out_box[3 + i] = hi - lo + 1;
    }
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_backward_box(hpx_frame* f, hpx_grid* g, const float* dL_dI, hp_memspace memspace, uint32_t flags,
                                  float* box_grad, const int32_t box[6]) {
    DV_RANGE("hpx_backward_box");
    DV_TRY(frame_check_grid(f, g));
    if (dL_dI == nullptr || box_grad == nullptr || box == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    for (int i = 0; i < 3; ++i)
        if (box[i] < 0 || box[3 + i] < 0) return HP_STATUS_INVALID_ARGUMENT;
    if (box[0] + box[3] > g->nx || box[1] + box[4] > g->ny || box[2] + box[5] > g->nz) return HP_STATUS_INVALID_ARGUMENT;
    if (!f->forward_done) return HP_STATUS_INVALID_ARGUMENT;
    const ScatterParams probe = scatter_params(*g);
    if (!probe.unit_bbox || probe.nearest) {
        set_last_error("hpx_backward_box needs a linear field whose scatter box is the unit cube");
        return HP_STATUS_UNSUPPORTED;
    }
    DV_ENTER(f->ctx);
    DV_TRY(grid_ensure_grad(g));
    DV_TRY(frame_push_params(f));
    const float* d_g = dL_dI;
    if (memspace == HP_MEMSPACE_HOST) {
        DV_CUDA(cudaMemcpyAsync(f->d_dL_dI, dL_dI, static_cast<size_t>(f->h_params.roi.w) * f->h_params.roi.h * 12,
                                cudaMemcpyHostToDevice, f->ctx->stream));
        d_g = f->d_dL_dI;
    }
    GradBox gb;
    gb.data = box_grad;
    for (int i = 0; i < 3; ++i) { gb.o[i] = box[i]; gb.n[i] = box[3 + i]; }
    if (static_cast<size_t>(gb.n[0]) * gb.n[1] * gb.n[2] == 0) return HP_STATUS_SUCCESS;   // nothing of this frame enters the cube
    return enqueue_backward(f, g, d_g, flags, &gb);
}

HP_API hp_status hpx_backward_signalled(hpx_frame* f, hpx_grid* g, const float* dL_dI, hp_memspace memspace, uint32_t flags,
                                        const uint32_t* group_end_rows, uint32_t n_groups, uint32_t** out_device_counters,
                                        uint32_t* out_expected) {
    DV_RANGE("hpx_backward_signalled");
    DV_TRY(frame_check_grid(f, g));
    if (dL_dI == nullptr || group_end_rows == nullptr || n_groups == 0 || n_groups > kMaxRowGroups || out_expected == nullptr)
        return HP_STATUS_INVALID_ARGUMENT;
    if (!f->forward_done) return HP_STATUS_INVALID_ARGUMENT;
    const RoiParams& roi = f->h_params.roi;
    const uint32_t tiles_x = (roi.w + kTileW * kWarpsX - 1) / (kTileW * kWarpsX);
    const uint32_t owned_rows = lean_block_count(roi) / std::max(1u, tiles_x);
    uint32_t prev = 0;
    for (uint32_t i = 0; i < n_groups; ++i) {
        if (group_end_rows[i] < prev) return HP_STATUS_INVALID_ARGUMENT;
        const uint32_t end = i + 1 == n_groups ? owned_rows : std::min(group_end_rows[i], owned_rows);
        out_expected[i] = (end - std::min(prev, end)) * tiles_x;
        prev = end;
    }
    DV_ENTER(f->ctx);
    DV_TRY(grid_ensure_grad(g));
    DV_TRY(frame_push_params(f));
    const float* d_g = dL_dI;
    if (memspace == HP_MEMSPACE_HOST) {
        DV_CUDA(cudaMemcpyAsync(f->d_dL_dI, dL_dI, static_cast<size_t>(roi.w) * roi.h * 12, cudaMemcpyHostToDevice, f->ctx->stream));
        d_g = f->d_dL_dI;
    }
    const LeanBuffers saved = f->buf;
    f->buf.group_done = f->d_group_done;
    f->buf.group_count = n_groups;
    for (uint32_t i = 0; i < n_groups; ++i) f->buf.group_end[i] = i + 1 == n_groups ? owned_rows : group_end_rows[i];
    const hp_status st = enqueue_backward(f, g, d_g, flags);
    f->buf = saved;
    if (out_device_counters) *out_device_counters = f->d_group_done;
    return st;
}

HP_API hp_status hpx_frame_reset_group_counters(hpx_frame* f, uint32_t** out_device_counters) {
    DV_RANGE("hpx_frame_reset_group_counters");
    if (f == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(f->ctx);
    DV_CUDA(cudaMemsetAsync(f->d_group_done, 0, kMaxRowGroups * sizeof(unsigned int), f->ctx->stream));
    if (out_device_counters) *out_device_counters = f->d_group_done;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_stream_wait_counter(const hp_ctx* stream_ctx, const uint32_t* device_counter, uint32_t value) {
    DV_RANGE("hpx_stream_wait_counter");
    if (stream_ctx == nullptr || device_counter == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(stream_ctx);
    typedef int (*wait_fn)(CUstream_st*, unsigned long long, uint32_t, unsigned int);   // cuStreamWaitValue32
    static wait_fn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        DV_CUDA(cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q));
        if (p == nullptr || q != cudaDriverEntryPointSuccess) {
            set_last_error("cuStreamWaitValue32 is not available from this driver");
            return HP_STATUS_UNSUPPORTED;
        }
        fn = reinterpret_cast<wait_fn>(p);
    }
    const int rc = fn(stream_ctx->stream, reinterpret_cast<unsigned long long>(device_counter), value, 0u /* GEQ */);
    if (rc != 0) {
        set_last_error("cuStreamWaitValue32 failed with driver error " + std::to_string(rc));
        return HP_STATUS_INTERNAL_ERROR;
    }
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_grid_set_grad_layout(hpx_grid* g, int32_t slow_axis, size_t* out_slab_floats, int32_t* out_slabs) {
    DV_RANGE("hpx_grid_set_grad_layout");
    if (g == nullptr || slow_axis < 0 || slow_axis > 2) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(g->ctx);
    DV_TRY(grid_ensure_grad(g));
    const uint32_t nx = static_cast<uint32_t>(g->nx), ny = static_cast<uint32_t>(g->ny), nz = static_cast<uint32_t>(g->nz);
    int32_t slabs = g->nz;
    if (slow_axis == 2) { g->gsx = 1; g->gsy = nx; g->gsz = nx * ny; slabs = g->nz; }        // [z][y][x]
    else if (slow_axis == 1) { g->gsx = 1; g->gsz = nx; g->gsy = nx * nz; slabs = g->ny; }   // [y][z][x]
    else { g->gsy = 1; g->gsz = ny; g->gsx = ny * nz; slabs = g->nx; }                        // [x][z][y]
    g->grad_slow_axis = slow_axis;
    // whatever was accumulated in the old order is meaningless in the new one
    DV_CUDA(cudaMemsetAsync(g->d_grad, 0, (g->voxels * 4 + kCameraFloats) * sizeof(float), g->ctx->stream));
    if (out_slab_floats) *out_slab_floats = slabs > 0 ? g->voxels / static_cast<size_t>(slabs) * 4 : 0;
    if (out_slabs) *out_slabs = slabs;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_grid_add_box(const hp_ctx* stream_ctx, hpx_grid* g, float* box_grad, const int32_t box[6]) {
    DV_RANGE("hpx_grid_add_box");
    if (stream_ctx == nullptr || g == nullptr || box_grad == nullptr || box == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    for (int i = 0; i < 3; ++i)
        if (box[i] < 0 || box[3 + i] < 0) return HP_STATUS_INVALID_ARGUMENT;
    if (box[0] + box[3] > g->nx || box[1] + box[4] > g->ny || box[2] + box[5] > g->nz) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(stream_ctx);
    if (stream_ctx->device != g->ctx->device) return HP_STATUS_INVALID_ARGUMENT;
    if (g->grad_slow_axis != 2) {
        set_last_error("hpx_grid_add_box needs the default gradient layout (z slowest)");
        return HP_STATUS_UNSUPPORTED;
    }
    DV_TRY(grid_ensure_grad(g));
    DV_CUDA(launch_add_box(stream_ctx->stream, reinterpret_cast<float4*>(g->d_grad), reinterpret_cast<float4*>(box_grad),
                           g->nx, g->ny, box));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_frame_box_misses(hpx_frame* f, uint32_t* out_count) {
    DV_RANGE("hpx_frame_box_misses");
    if (f == nullptr || out_count == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(f->ctx);
    unsigned int host = 0;
    DV_CUDA(cudaMemcpyAsync(&host, f->d_box_miss, sizeof(host), cudaMemcpyDeviceToHost, f->ctx->stream));
    DV_CUDA(cudaMemsetAsync(f->d_box_miss, 0, sizeof(host), f->ctx->stream));
    DV_CUDA(cudaStreamSynchronize(f->ctx->stream));
    *out_count = host;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_backward_scatter(const hpx_frame* f, const hpx_grid* g, uint32_t flags, uint32_t* out_flag) {
    DV_RANGE("hpx_backward_scatter");
    if (f == nullptr || g == nullptr || out_flag == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    const int want = (flags & HPX_BACKWARD_SCATTER_MERGED) ? kScatterMerge
                   : (flags & HPX_BACKWARD_SCATTER_PER_RAY) ? kScatterPerRay : kScatterAuto;
    const int got = resolve_scatter_mode(f->h_params, packed_view(*g), scatter_params(*g), want);
    *out_flag = got == kScatterMerge ? HPX_BACKWARD_SCATTER_MERGED : HPX_BACKWARD_SCATTER_PER_RAY;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_frame_image(const hpx_frame* f, hp_img_t* out) {
    DV_RANGE("hpx_frame_image");
    if (f == nullptr || out == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    const hp_plan_desc& d = f->plan->desc;
    const int64_t w = d.width, h = d.height;
    shape_tensor(out->image, HP_DTYPE_F32, HP_MEMSPACE_DEVICE, 3, h, w, 3);
    shape_tensor(out->trans, HP_DTYPE_F32, HP_MEMSPACE_DEVICE, 2, h, w);
    shape_tensor(out->opacity, HP_DTYPE_F32, HP_MEMSPACE_DEVICE, 2, h, w);
    shape_tensor(out->depth, HP_DTYPE_F32, HP_MEMSPACE_DEVICE, 2, h, w);
    shape_tensor(out->hitmask, HP_DTYPE_U32, HP_MEMSPACE_DEVICE, 2, h, w);
    out->image.data = f->buf.image;
    out->trans.data = f->buf.trans;
    out->opacity.data = f->buf.opacity;
    out->depth.data = f->buf.depth;
    out->hitmask.data = f->buf.hitmask;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_frame_read(hpx_frame* f, float* image, float* trans, float* opacity, float* depth,
                                uint32_t* hitmask) {
    DV_RANGE("hpx_frame_read");
    if (f == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(f->ctx);
    cudaStream_t s = f->ctx->stream;
    const size_t pixels = static_cast<size_t>(f->plan->desc.width) * f->plan->desc.height;
    if (image) DV_CUDA(cudaMemcpyAsync(image, f->buf.image, pixels * 12, cudaMemcpyDeviceToHost, s));
    if (trans) DV_CUDA(cudaMemcpyAsync(trans, f->buf.trans, pixels * 4, cudaMemcpyDeviceToHost, s));
    if (opacity) DV_CUDA(cudaMemcpyAsync(opacity, f->buf.opacity, pixels * 4, cudaMemcpyDeviceToHost, s));
    if (depth) DV_CUDA(cudaMemcpyAsync(depth, f->buf.depth, pixels * 4, cudaMemcpyDeviceToHost, s));
    if (hitmask) DV_CUDA(cudaMemcpyAsync(hitmask, f->buf.hitmask, pixels * 4, cudaMemcpyDeviceToHost, s));
    DV_CUDA(cudaStreamSynchronize(s));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_frame_counts(hpx_frame* f, hpx_counts* out) {
    DV_RANGE("hpx_frame_counts");
    if (f == nullptr || out == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(f->ctx);
    auto* h_live = reinterpret_cast<unsigned long long*>(f->h_pinned + 1);
    DV_CUDA(cudaMemcpyAsync(h_live, f->buf.live_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            f->ctx->stream));
    DV_CUDA(cudaStreamSynchronize(f->ctx->stream));
    out->rays = f->rays;
    out->samples = f->samples;
    out->live_samples = *h_live;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_frame_cube_samples(hpx_frame* f, const hpx_grid* g, uint64_t* out_samples) {
    DV_RANGE("hpx_frame_cube_samples");
    DV_TRY(frame_check_grid(f, g));
    if (out_samples == nullptr || !f->forward_done) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(f->ctx);
    cudaStream_t s = f->ctx->stream;
    auto* h = reinterpret_cast<unsigned long long*>(f->h_pinned + 1);
    if (g->clamp) {   // clamped fields gather at every live sample
        DV_CUDA(cudaMemcpyAsync(h, f->buf.live_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        DV_CUDA(cudaStreamSynchronize(s));
        *out_samples = *h;
        return HP_STATUS_SUCCESS;
    }
    DV_TRY(frame_push_params(f));
    DeviceScratch scratch;
    auto* d_total = static_cast<unsigned long long*>(scratch.take(sizeof(unsigned long long)));
    if (d_total == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    DV_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), s));
    DV_CUDA(launch_cube_count(s, f->d_params, f->h_params, f->buf, d_total));
    DV_CUDA(cudaMemcpyAsync(h, d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    DV_CUDA(cudaStreamSynchronize(s));
    *out_samples = *h;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_grid_touched_voxels(hpx_grid* g, uint64_t* out_voxels) {
    DV_RANGE("hpx_grid_touched_voxels");
    if (g == nullptr || out_voxels == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(g->ctx);
    DV_TRY(grid_ensure_grad(g));
    cudaStream_t s = g->ctx->stream;
    DeviceScratch scratch;
    auto* d_total = static_cast<unsigned long long*>(scratch.take(sizeof(unsigned long long)));
    if (d_total == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    DV_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), s));
    DV_CUDA(launch_touched_voxels(s, reinterpret_cast<const float4*>(g->d_grad), g->voxels, d_total));
    unsigned long long host = 0;
    DV_CUDA(cudaMemcpyAsync(&host, d_total, sizeof(host), cudaMemcpyDeviceToHost, s));
    DV_CUDA(cudaStreamSynchronize(s));
    *out_voxels = host;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_frame_grad_input(hpx_frame* f, float** out) {
    DV_RANGE("hpx_frame_grad_input");
    if (f == nullptr || out == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    *out = f->d_dL_dI;
    return HP_STATUS_SUCCESS;
}

// Real capture: every node is a kernel or memset on the context's stream; no
// allocation, no synchronisation and no host read inside the captured region.
HP_API hp_status hpx_frame_capture(hpx_frame* f, hpx_grid* g, uint32_t backward_flags) {
    DV_RANGE("hpx_frame_capture");
    DV_TRY(frame_check_grid(f, g));
    DV_ENTER(f->ctx);
    if (backward_flags != 0) DV_TRY(grid_ensure_grad(g));
    if ((backward_flags & HPX_BACKWARD_GRID) && (backward_flags & HPX_BACKWARD_DETERMINISTIC)) DV_TRY(grid_ensure_fixed(g));
    DV_TRY(frame_push_params(f));
    cudaStream_t s = f->ctx->stream;
    DV_CUDA(cudaStreamSynchronize(s));
    if (f->graph_exec) { cudaGraphExecDestroy(f->graph_exec); f->graph_exec = nullptr; }
    if (f->graph) { cudaGraphDestroy(f->graph); f->graph = nullptr; }
    DV_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    hp_status st = enqueue_forward(f, g);
    if (st == HP_STATUS_SUCCESS && backward_flags != 0) st = enqueue_backward(f, g, f->d_dL_dI, backward_flags, nullptr, true);
    const cudaError_t e = cudaStreamEndCapture(s, &f->graph);
    if (st != HP_STATUS_SUCCESS) return st;
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture");
    DV_CUDA(cudaGraphInstantiate(&f->graph_exec, f->graph, nullptr, nullptr, 0));
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hpx_frame_replay(hpx_frame* f) {
    DV_RANGE("hpx_frame_replay");
    if (f == nullptr || f->graph_exec == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(f->ctx);
    DV_TRY(frame_push_params(f));
    DV_CUDA(cudaGraphLaunch(f->graph_exec, f->ctx->stream));
    f->forward_done = true;
    return HP_STATUS_SUCCESS;
}

HP_API void hpx_frame_release(hpx_frame* f) {
    DV_RANGE("hpx_frame_release");
    if (f == nullptr) return;
    if (f->ctx != nullptr && f->ctx->ready) {
        DeviceScope scope;
        scope.enter(f->ctx);
        cudaStreamSynchronize(f->ctx->stream);
        if (f->stream_plan != nullptr && f->stream_plan_free != nullptr) f->stream_plan_free(f->stream_plan);
        if (f->graph_exec) cudaGraphExecDestroy(f->graph_exec);
        if (f->graph) cudaGraphDestroy(f->graph);
        for (void* p : f->allocations) cudaFree(p);
        if (f->h_pinned) cudaFreeHost(f->h_pinned);
    }
    plan_unref(f->plan);
    ctx_unref(f->ctx);
    delete f;
}

}  // extern "C"
