/*
 * hotpath/hp_b200.h -- additive B200 entry points next to hotpath/hp.h.
 *
 * hp.h (reference hotpath/include/hotpath/hp.h) fixes the drop-in surface; it
 * materialises every sample (32 B + 16 B aux each), returns per-sample
 * gradients only, has no device ordinal / stream, and creates fields from HOST
 * tensors only (reference hp_runtime.cpp:271-274).  The entry points below add
 * what a device-resident training loop needs without touching that surface
 * (SURVEY section 8b "additive extensions"):
 *
 *   hpx_ctx_ext     device ordinal + caller stream, passed via hp_ctx_desc.reserved
 *   hpx_grid        sigma+RGB packed {r,g,b,sigma} float4 grid in HBM, plus its
 *                   packed gradient grid and the 16-float camera gradient
 *   hpx_frame       per-plan device workspace (the Plan workspace planner):
 *                   image planes, per-ray state, transmittance checkpoints
 *   hpx_forward     ray generation + marching + integration + image compose in
 *                   ONE kernel, no per-sample global traffic
 *                   (replaces hp_ray -> hp_samp_int_fused -> hp_img,
 *                   reference src/render/renderer.cpp:259-365)
 *   hpx_backward    reverse march with recompute, warp-level scatter into the
 *                   packed gradient grid, optional camera adjoint
 *                   (replaces hp_diff + DenseGridField::AccumulateSampleGradients,
 *                   reference src/render/renderer.cpp:415-427,
 *                   src/fields/dense_grid.cpp:171-309)
 *   hpx_frame_capture / hpx_frame_replay   real CUDA-graph capture of the above
 *   hpx_frame_set_interleave / hpx_frame_bounds / hpx_grid_set_grad_layout / hpx_backward_signalled /
 *   hpx_stream_wait_counter / hpx_backward_box / hpx_grid_add_box
 *                   building blocks for strong scaling of ONE frame over several GPUs (interleaved tile
 *                   rows, contiguous gradient slabs, device-signalled row groups); hpx_comm_* / hpx_shard_* put them
 *                   together behind the C ABI (csrc/dv_comm.cu)
 *
 * All functions return hp_status; none blocks the host unless documented.
 * Work is enqueued on the context's stream.
 */
#ifndef DVREN_HOTPATH_HP_B200_H_
#define DVREN_HOTPATH_HP_B200_H_

#include "hotpath/hp.h"

#ifdef __cplusplus
extern "C" {
#endif

#define HPX_CTX_EXT_MAGIC 0x42323030u /* "B200" */

/* Optional extension of hp_ctx_desc (pass its address in .reserved). */
typedef struct hpx_ctx_ext {
    uint32_t magic;          /* HPX_CTX_EXT_MAGIC */
    int32_t  device_ordinal; /* -1: take it from preferred_device / current device */
    void*    stream;         /* cudaStream_t to enqueue on; NULL: library-owned stream */
} hpx_ctx_ext;

/* Extended form (magic HPX_CTX_EXT2_MAGIC): additionally keeps `reserve_sms` streaming multiprocessors of the GPU OUT of
 * this context's reach: the context's stream is created in a CUDA green context that owns the remaining SMs (rounded to
 * the hardware's partition granularity, 8 SMs on sm_90+), so kernels of this library never occupy the reserved ones.
 * A collective running next to a long rendering launch (hpx_shard_step) then always finds free SMs for its CTAs
 * instead of waiting for the launch to drain.  Requires stream == NULL (the library creates the stream).
 * hpx_ctx_sm_counts reports what was provisioned. */
#define HPX_CTX_EXT2_MAGIC 0x42323031u
typedef struct hpx_ctx_ext2 {
    uint32_t magic;          /* HPX_CTX_EXT2_MAGIC */
    int32_t  device_ordinal;
    void*    stream;
    uint32_t reserve_sms;    /* 0: use the whole GPU */
    uint32_t flags;          /* 0 */
} hpx_ctx_ext2;

typedef struct hpx_grid  hpx_grid;
typedef struct hpx_frame hpx_frame;

/* flags for hpx_backward / hpx_frame_capture */
#define HPX_BACKWARD_GRID    0x1u /* accumulate d/d sigma, d/d rgb into the grid's gradient */
#define HPX_BACKWARD_CAMERA  0x2u /* accumulate d/d c2w[12], d/d {fx,fy,cx,cy}               */
#define HPX_BACKWARD_ZERO    0x4u /* zero the gradient buffers first                         */
/* Scatter strategy of the grid backward (default: chosen from the pixel / voxel spacing ratio):  */
#define HPX_BACKWARD_SCATTER_PER_RAY 0x10u /* one lane = one ray, 8 reds per sample                      */
#define HPX_BACKWARD_SCATTER_MERGED  0x20u /* 2x2 pixel quads x 2 steps merged in registers before the reds */
/* Bitwise reproducible grid gradients: contributions are rounded to a power-of-two quantum (2^-44 of the largest
 * possible contribution) and accumulated with 64-bit integer reds, whose sum is independent of arrival order; costs
 * 32 B per voxel of extra HBM and 4 integer reds per corner instead of one 16-byte float red.  Without this flag
 * the float reds arrive in a different order from run to run (differences at the 1e-7 relative level). */
#define HPX_BACKWARD_DETERMINISTIC   0x40u

/* Per-frame counters (valid after the stream has been synchronised). */
typedef struct hpx_counts {
    uint64_t rays;          /* rays marched                                        */
    uint64_t samples;       /* reference sample_count: every emitted sample        */
    uint64_t live_samples;  /* samples integrated before the T <= 1e-4 stop        */
} hpx_counts;

/* ---- context helpers ---------------------------------------------------- */
HP_API hp_status hpx_ctx_synchronize(const hp_ctx* ctx);
HP_API hp_status hpx_ctx_device(const hp_ctx* ctx, int32_t* out_ordinal, void** out_stream);
/* SMs the context's kernels can run on / SMs of the GPU (differ when hpx_ctx_ext2.reserve_sms took effect). */
HP_API hp_status hpx_ctx_sm_counts(const hp_ctx* ctx, uint32_t* out_usable, uint32_t* out_total);
/* Device-side stage timing (CUDA events on the context's stream) for callers that do not link a CUDA runtime:
 * hpx_ctx_mark records event `slot` (0..15); hpx_ctx_elapsed_ms waits for `slot_end` and returns the GPU time between
 * two recorded slots.  dvren::Renderer fills its RenderStats from these (reference renderer.hpp:41-48 uses host clocks). */
HP_API hp_status hpx_ctx_mark(const hp_ctx* ctx, uint32_t slot);
HP_API hp_status hpx_ctx_elapsed_ms(const hp_ctx* ctx, uint32_t slot_begin, uint32_t slot_end, float* out_ms);
/* Blocking device-to-host copy on the context's stream, for binding languages without a CUDA
 * runtime of their own (reading back the DEVICE views hp_graph_execute / hpx_frame_image return). */
HP_API hp_status hpx_copy_to_host(const hp_ctx* ctx, void* host_dst, const void* device_src, size_t bytes);
/* Plain device memory on the context's GPU for callers that do not link a CUDA runtime. */
HP_API hp_status hpx_device_alloc(const hp_ctx* ctx, size_t bytes, void** out_device_ptr);
HP_API void      hpx_device_free(const hp_ctx* ctx, void* device_ptr);
HP_API hp_status hpx_copy_to_device(const hp_ctx* ctx, void* device_dst, const void* host_src, size_t bytes);
/* Page-lock / release a caller-owned host range (direct DMA for the HOST-memspace copies of this library). */
HP_API hp_status hpx_host_register(const hp_ctx* ctx, void* host_ptr, size_t bytes);
HP_API void      hpx_host_unregister(const hp_ctx* ctx, void* host_ptr);
/* Last CUDA/runtime error text recorded on this thread ("" if none). */
HP_API const char* hpx_last_error(void);

/* ---- packed device grid -------------------------------------------------- */
/* Build from two dense-grid fields of equal resolution, interpolation and OOB
 * policy (what DenseGridField creates).  bbox_* (NULL = unit cube) is used by
 * the backward scatter only, like the reference (SURVEY finding 8). */
HP_API hp_status hpx_grid_create(const hp_ctx* ctx, const hp_field* fs, const hp_field* fc,
                                 const float bbox_min[3], const float bbox_max[3],
                                 hpx_grid** out_grid);
/* Build from raw arrays: sigma[nz][ny][nx], color[nz][ny][nx][3] in `memspace`. */
HP_API hp_status hpx_grid_create_raw(const hp_ctx* ctx, int32_t nx, int32_t ny, int32_t nz,
                                     const float* sigma, const float* color, hp_memspace memspace,
                                     uint32_t interp, uint32_t oob, const float bbox_min[3],
                                     const float bbox_max[3], hpx_grid** out_grid);
/* One copy of the values in HBM: the fields the grid was built from (either may be NULL) give up their own snapshots and
 * become strided views of the packed voxels, so that hpx_grid_update is also what the staged hp_samp / hp_graph_* calls
 * on those fields read.  The fields must come from the grid's context and match its shape and policies; they must be
 * released before, or stop being used after, the grid is released. */
HP_API hp_status hpx_grid_adopt_fields(hpx_grid* grid, hp_field* fs, hp_field* fc);
/* Replace the values (parameter update); either pointer may be NULL to keep it. */
HP_API hp_status hpx_grid_update(hpx_grid* grid, const float* sigma, const float* color,
                                 hp_memspace memspace);
/* Empty-space skipping (reference roadmap: hotpath/DESIGN_SPECIFICATION.md:99, DESIGN_SPECIFICATION.md:238-239).
 * hpx_grid_build_occupancy computes, for the CURRENT values, 2 bits per brick of 8^3 trilinear cells -- "some corner a
 * cell of the brick reads has sigma != 0" and "... has any channel != 0" -- and, with enable != 0, makes the forward
 * kernel skip samples whose eight corners all have sigma = 0 and the backward kernels skip the gather of samples whose
 * corners are zero in every channel.  Such samples contribute EXACTLY nothing in the reference's arithmetic (sigma = +0,
 * alpha = 0, w = T * 0; the backward still scatters their d sigma = -adj_T T dt), so images, sample counts and gradients
 * do not change: it is a pure speed-up on sparse volumes.  Linear OOB-zero fields; blocks until the bits are built.
 * out_empty_sigma / out_empty_all (may be NULL): fraction of bricks the forward / the backward can skip.
 * hpx_grid_update marks every brick occupied again (no skipping until the next build); hpx_grid_set_occupancy switches
 * skipping on / off without rebuilding. */
HP_API hp_status hpx_grid_build_occupancy(hpx_grid* grid, int32_t enable, float* out_empty_sigma, float* out_empty_all);
HP_API hp_status hpx_grid_set_occupancy(hpx_grid* grid, int32_t enable);
/* Storage precision of the packed VALUES (the gradient block stays fp32).  HPX_STORAGE_F16: four IEEE halfs per voxel, 8 B
 * instead of 16 -- half the gather bytes and half the HBM footprint (1024^3: 8.6 GB instead of 17.2 GB).  The values are
 * rounded to half once; voxels are widened back to fp32 exactly on load and all arithmetic stays fp32, so the grid behaves
 * EXACTLY like an fp32 grid holding the rounded values: against that twin the usual tolerances hold (counts bit-exact,
 * image 1e-5, gradients 1e-4); against the unrounded grid the difference is the storage rounding (2^-11 relative per
 * value).  Linear OOB-zero grids with the unit scatter box and no adopted hp_fields. */
#define HPX_STORAGE_F32 0u
#define HPX_STORAGE_F16 1u
HP_API hp_status hpx_grid_set_storage(hpx_grid* grid, uint32_t storage);
HP_API hp_status hpx_grid_zero_grad(hpx_grid* grid);
/* Device view of the contiguous gradient block [4*V grid floats {r,g,b,sigma} | 16 camera floats]
 * -- the buffer a data-parallel caller all-reduces. */
HP_API hp_status hpx_grid_grad_buffer(hpx_grid* grid, float** out_device_ptr, size_t* out_floats);
/* Un-interleave into the reference layout sigma_grad[V], color_grad[3V] (+camera[16], may be NULL)
 * in `memspace`.  Blocks until done when the destination is HOST. */
HP_API hp_status hpx_grid_read_grad(hpx_grid* grid, float* sigma_grad, float* color_grad,
                                    float* camera16, hp_memspace memspace);
/* Same for voxels [first, first + count) of the reference order (sigma_grad[count], color_grad[3 * count]): every rank of
 * a sharded job can hand its share of the summed gradient to the host over its own PCIe link. */
HP_API hp_status hpx_grid_read_grad_range(hpx_grid* grid, size_t first, size_t count, float* sigma_grad, float* color_grad,
                                          float* camera16, hp_memspace memspace);
/* DenseGridField::AccumulateSampleGradients on the GPU (reference src/fields/dense_grid.cpp:171-309):
 * scatter per-sample gradients (positions (M,3), grad_sigma (M), grad_color (M,3) in `memspace`)
 * into the packed gradient grid with the grid's bbox / interpolation / OOB policy. */
HP_API hp_status hpx_grid_accumulate_samples(hpx_grid* grid, const float* positions, const float* grad_sigma,
                                             const float* grad_color, size_t count, hp_memspace memspace);
HP_API void      hpx_grid_release(hpx_grid* grid);

/* ---- per-plan frame workspace ------------------------------------------- */
HP_API hp_status hpx_frame_create(const hp_plan* plan, hpx_frame** out_frame);
/* Host-only (works without a GPU): the per-step table {base = t_near + k dt, fixed-mode sample time, dt_actual,
 * depth cursor before the step} that frames of this plan upload; out_steps4 may be NULL to query the count.
 * All rays generated from a plan share it (reference samp_cpu.cpp:227-241, int_cpu.cpp:170-211). */
HP_API hp_status hpx_plan_step_table(const hp_plan* plan, float* out_steps4, size_t capacity_steps, uint32_t* out_count);
/* Bytes of device memory the frame owns (workspace accounting). */
HP_API size_t    hpx_frame_bytes(const hpx_frame* frame);
/* Change camera / seed / global ray-index base without re-planning (graph friendly:
 * the values live in a device parameter block that captured kernels re-read). */
HP_API hp_status hpx_frame_set_view(hpx_frame* frame, const hp_camera_desc* camera, uint64_t seed,
                                    uint64_t ray_index_base);
HP_API hp_status hpx_forward(hpx_frame* frame, const hpx_grid* grid);
/* dL_dI: (rays,3) f32 per ray in plan order, HOST (copied on the stream) or DEVICE. */
HP_API hp_status hpx_backward(hpx_frame* frame, hpx_grid* grid, const float* dL_dI,
                              hp_memspace memspace, uint32_t flags);
/* hpx_backward + hpx_grid_read_grad(HOST) in one call, with the read-back running UNDER the backward kernel: the kernel
 * signals per group of image rows, a high-priority copy stream waits for each group (no SM is occupied by the wait),
 * un-interleaves the gradient slabs that group finished and copies them into the caller's HOST arrays in the reference
 * layout (sigma_grad[V], color_grad[3V], camera16; any may be NULL) while later rows still render.  The gradient block is
 * switched to the slab order of the world axis the image rows advance along (hpx_grid_set_grad_layout) and stays there.
 * Page-lock the arrays (hpx_host_register) for the copies to be asynchronous; they are complete when the context's stream
 * is (hpx_ctx_synchronize).  Needs HPX_BACKWARD_ZERO, a linear OOB-zero field with the unit scatter box and image rows that
 * advance along world y or z; in every other case it runs the two plain calls (same results, blocking).
 * dvren::Renderer::Backward uses it when RenderOptions::pin_result_buffers is set (reference renderer.cpp:415-444 copies
 * the whole gradient after the backward has finished). */
HP_API hp_status hpx_backward_streamed(hpx_frame* frame, hpx_grid* grid, const float* dL_dI, hp_memspace memspace, uint32_t flags,
                                       float* sigma_grad_host, float* color_grad_host, float* camera16_host);
/* ---- building blocks for ONE frame over several GPUs (put together by hpx_shard_* below, csrc/dv_comm.cu) ---------
 * hpx_frame_set_interleave: this frame marches only the CTA tile rows (8 pixel rows each) t of its ROI with
 *   t % stride == phase; ray indices, pixel ids and buffers stay those of the whole ROI.  stride 1 = everything.
 * hpx_frame_bounds: box {x0, y0, z0, nx, ny, nz} of grid voxels the frame's backward can touch (blocks until done).
 * hpx_backward_box: like hpx_backward, but the grid gradient goes to the caller's dense DEVICE box
 *   box_grad[nz][ny][nx][4] ({dr,dg,db,dsigma}) instead of the grid's own gradient block, so that a group of image
 *   rows can be all-reduced (contiguously) while the next group is still being rendered.  HPX_BACKWARD_ZERO clears the
 *   box.  Unit scatter bbox + linear fields only; HPX_BACKWARD_DETERMINISTIC is ignored. */
/* (A captured graph of the frame is dropped: hpx_frame_replay fails until hpx_frame_capture is called again.  The
 * scatter strategy and the fused-camera choice of a captured graph are those of the view it was captured with.) */
HP_API hp_status hpx_frame_set_interleave(hpx_frame* frame, uint32_t stride, uint32_t phase);
HP_API hp_status hpx_frame_bounds(hpx_frame* frame, const hpx_grid* grid, int32_t out_box[6]);
/* Order in which the frame's launches take its tile rows: 0 first-to-last, 1 last-to-first, 2 centre-out (middle row, one
 * below, one above, ...).  CTAs are dispatched in order, so the rows a launch ends with decide its tail; a band whose rays
 * get longer towards its last row should end with its first one.  + HPX_ORDER_COLUMNS: the tiles are taken column by
 * column (every row of a column in the row order above), the columns from the middle of the image outwards -- for a band
 * whose rows all cost the same but whose tiles get cheaper towards the left and right edge (the middle bands of a sharded
 * perspective frame), so that the launch ends with cheap tiles instead of a partly empty wave of full-length ones.
 * Results do not depend on the order.  The row groups of hpx_backward_signalled count in DISPATCH order (with
 * HPX_ORDER_COLUMNS every group completes only with the last column). */
#define HPX_ORDER_COLUMNS 4
HP_API hp_status hpx_frame_set_row_order(hpx_frame* frame, int32_t order);
/* Host-only: which tile row the i-th dispatched one of `rows` is under `order` 0..2 (a permutation of 0 .. rows - 1). */
HP_API hp_status hpx_tile_row_order(uint32_t i, uint32_t rows, int32_t order, uint32_t* out_row);
/* Host-only: the tile (column, row) the CTA `block` of a launch of tiles_x * rows CTAs takes under `order`
 * (0..2, optionally + HPX_ORDER_COLUMNS): a permutation of the tiles. */
HP_API hp_status hpx_tile_order(uint32_t block, uint32_t tiles_x, uint32_t rows, int32_t order, uint32_t* out_col, uint32_t* out_row);
HP_API hp_status hpx_backward_box(hpx_frame* frame, hpx_grid* grid, const float* dL_dI, hp_memspace memspace,
                                  uint32_t flags, float* box_grad, const int32_t box[6]);
/* Axis order of the gradient block: slow_axis 0 = x, 1 = y, 2 = z (default) becomes the slowest-varying one, so that a
 * range of voxel planes perpendicular to that axis ("slabs") is one CONTIGUOUS piece of hpx_grid_grad_buffer:
 *   slow z: [z][y][x]   slow y: [y][z][x]   slow x: [x][z][y]   (x {r,g,b,sigma} floats).
 * A caller that renders one frame in groups of image rows picks the world axis the rows run along; the slabs a
 * finished group will never touch again can then be all-reduced IN PLACE while the next group is rendered
 * (hpx_shard_*, hpx_backward_streamed).  Clears the gradient block.  hpx_grid_read_grad always returns the reference order.
 * out_slab_floats: floats per slab; out_slabs: number of slabs. */
HP_API hp_status hpx_grid_set_grad_layout(hpx_grid* grid, int32_t slow_axis, size_t* out_slab_floats, int32_t* out_slabs);
/* grid gradient[box] += box_grad, then box_grad = 0, enqueued on the stream of `stream_ctx` (any context of the grid's
 * device: the caller's side stream, so that the hand-over overlaps the rendering of the next group). */
HP_API hp_status hpx_grid_add_box(const hp_ctx* stream_ctx, hpx_grid* grid, float* box_grad, const int32_t box[6]);
/* Contributions hpx_backward_box found outside its box and dropped since the last call (0 with the boxes of
 * hpx_frame_bounds; reading clears the counter and synchronises). */
HP_API hp_status hpx_frame_box_misses(hpx_frame* frame, uint32_t* out_count);
/* hpx_backward with a DEVICE-SIDE completion signal per group of image rows: ONE launch whose CTAs run in tile-row order;
 * group i covers the frame's owned tile rows [group_end_rows[i-1], group_end_rows[i]) (the last group runs to the end).
 * Every CTA adds 1 to counter i when its reds are issued and fenced; out_expected[i] is the count that means "group i is
 * done".  hpx_stream_wait_counter makes ANOTHER stream (stream_ctx's) wait until a counter has reached a value
 * (cuStreamWaitValue32, no SM is occupied by the wait): that stream can all-reduce the gradient slabs a group finished
 * while the later rows are still running -- no per-group launches, hence no launch tails.
 * Protocol per step: hpx_frame_reset_group_counters (clears them on the frame's stream) -> record an event there and make
 * the waiting stream wait for it (so that it cannot see the previous step's counts) -> hpx_backward_signalled ->
 * hpx_stream_wait_counter + collective per group on the waiting stream. */
HP_API hp_status hpx_frame_reset_group_counters(hpx_frame* frame, uint32_t** out_device_counters);
HP_API hp_status hpx_backward_signalled(hpx_frame* frame, hpx_grid* grid, const float* dL_dI, hp_memspace memspace,
                                        uint32_t flags, const uint32_t* group_end_rows, uint32_t n_groups,
                                        uint32_t** out_device_counters, uint32_t* out_expected);
HP_API hp_status hpx_stream_wait_counter(const hp_ctx* stream_ctx, const uint32_t* device_counter, uint32_t value);
/* Which scatter strategy hpx_backward(flags) runs for this frame / grid pair: writes HPX_BACKWARD_SCATTER_PER_RAY
 * or HPX_BACKWARD_SCATTER_MERGED (kernel names lean_backward_kernel / lean_backward_merge_kernel in profiles). */
HP_API hp_status hpx_backward_scatter(const hpx_frame* frame, const hpx_grid* grid, uint32_t flags, uint32_t* out_flag);
/* Device views of the composed frame (shapes as hp_img_t). */
HP_API hp_status hpx_frame_image(const hpx_frame* frame, hp_img_t* out_views);
/* Copy the composed frame to HOST buffers (any may be NULL); synchronises. */
HP_API hp_status hpx_frame_read(hpx_frame* frame, float* image, float* trans, float* opacity,
                                float* depth, uint32_t* hitmask);
HP_API hp_status hpx_frame_counts(hpx_frame* frame, hpx_counts* out_counts);
/* Measurement helpers (bench.py roofline): live samples of the last hpx_forward that lie INSIDE the unit cube, i.e. the
 * samples that gather 8 corners and scatter 8 reds (a separate counting kernel that re-marches the rays without touching
 * the grid; blocks until done), and the number of voxels whose gradient the backward has written since the last zero
 * (V_touched of the compulsory-HBM-bytes formula; blocks until done). */
HP_API hp_status hpx_frame_cube_samples(hpx_frame* frame, const hpx_grid* grid, uint64_t* out_samples);
HP_API hp_status hpx_grid_touched_voxels(hpx_grid* grid, uint64_t* out_voxels);
/* Capture forward (+ backward when flags != 0, reading dL/dI from the frame's own
 * device buffer, see hpx_frame_grad_input) into a CUDA graph; replay launches it. */
HP_API hp_status hpx_frame_capture(hpx_frame* frame, hpx_grid* grid, uint32_t backward_flags);
HP_API hp_status hpx_frame_replay(hpx_frame* frame);
HP_API hp_status hpx_frame_grad_input(hpx_frame* frame, float** out_device_dL_dI);
HP_API void      hpx_frame_release(hpx_frame* frame);

/* ---- multi-GPU: one process (or thread) per GPU, NCCL over NVLink ---------------------------------------------------
 * The reference is single-device.  Rays are independent, so the path shards with no forward
 * collective; the one exchange is the sum of the packed gradient block.  NCCL is loaded at run time
 * (libnccl.so.2, or the path in DVREN_NCCL_LIBRARY): without it these calls return HP_STATUS_UNSUPPORTED and the rest
 * of the library is unaffected.
 *
 *   hpx_comm_unique_id      rank 0 creates the 128-byte rendezvous id and hands it to the other ranks by any means
 *   hpx_comm_create         rank `rank` of `world` on ctx's GPU; world == 1 is valid (no NCCL involved).  max_ctas > 0 caps
 *                           the CTAs of NCCL's kernels (use the SMs hpx_ctx_ext2.reserve_sms left free)
 *   hpx_grid_allreduce_grad data parallelism over views: sum of the whole gradient block after the local backward passes
 *   hpx_shard_*             ONE frame rendered by all ranks (strong scaling): rank r marches the CTA tile rows t with
 *                           t % world == r; the gradient block is laid out with the axis the image rows advance along as
 *                           its slowest one; ONE backward launch signals per group of rows and a high-priority side stream
 *                           all-reduces in place the slabs each finished group leaves behind while later rows still render.
 *                           group_weights: relative heights of the row groups (NULL = equal); keep the last one small,
 *                           only its slabs are reduced after the kernel.  Linear OOB-zero fields with the unit scatter box. */
#define HPX_COMM_ID_BYTES 128
typedef struct hpx_comm  hpx_comm;
typedef struct hpx_shard hpx_shard;
HP_API hp_status hpx_comm_unique_id(uint8_t out_id[HPX_COMM_ID_BYTES]);
HP_API hp_status hpx_comm_create(const hp_ctx* ctx, const uint8_t id[HPX_COMM_ID_BYTES], int32_t rank, int32_t world,
                                 int32_t max_ctas, hpx_comm** out_comm);
HP_API void      hpx_comm_release(hpx_comm* comm);
HP_API hp_status hpx_comm_info(const hpx_comm* comm, int32_t* out_rank, int32_t* out_world, int32_t* out_nccl_version);
/* In-place sum over the ranks, ordered after the work already on the context's stream; that stream continues after it. */
HP_API hp_status hpx_comm_allreduce(hpx_comm* comm, float* device_buf, size_t floats);
HP_API hp_status hpx_grid_allreduce_grad(hpx_comm* comm, hpx_grid* grid);
HP_API hp_status hpx_shard_create(hpx_comm* comm, const hp_plan* full_frame_plan, hpx_grid* grid, const float* group_weights,
                                  uint32_t n_groups, hpx_shard** out_shard);
/* [HPX_BACKWARD_ZERO: clear the gradient block, overlapped with the forward] -> forward -> signalled backward -> slab
 * all-reduces.  dL_dI_device: (rays of the WHOLE frame, 3) on the device.  Afterwards every rank holds the summed gradient. */
HP_API hp_status hpx_shard_step(hpx_shard* shard, const float* dL_dI_device, uint32_t flags);
/* The rank's frame (image planes of its tile rows, counts).  Owned by the shard. */
HP_API hp_status hpx_shard_frame(hpx_shard* shard, hpx_frame** out_frame);
/* Timing experiments: run the step without its collectives. */
HP_API hp_status hpx_shard_set_reduce(hpx_shard* shard, int32_t enabled);
/* Slowest axis of the gradient block, number of row groups, image rows and touched slab range [lo, hi) per group
 * (arrays of 16 / 32 entries; any pointer may be NULL). */
HP_API hp_status hpx_shard_layout(const hpx_shard* shard, int32_t* out_slow_axis, uint32_t* out_groups,
                                  uint32_t* out_group_rows, int32_t* out_slab_ranges);
/* ---- the same frame in contiguous, work-balanced bands with a SPARSE exchange ---------------------------------------
 * hpx_shard_create_bands: rank r renders one contiguous band of image rows; the bands are cut so that every rank has the
 * same marching work (in-cube samples of its rays).  The rays of a band stay inside a wedge of the volume, so with the
 * gradient block laid out slab by slab along the world axis the image rows advance along, a rank's backward touches one
 * contiguous slab range (about 2 / world of the grid) and nothing else.  Every slab has an OWNER (cut half way through
 * the overlap of neighbouring wedges).  hpx_shard_step then runs: clear wedge + owned slabs -> forward -> backward ->
 * the parts of the wedge the rank does not own go point-to-point to their owners, which add them in rank order.
 *   HPX_SHARD_RESULT_OWNED       stops there: rank r holds the finished sum of ITS slabs (hpx_shard_owned) -- a reduce-scatter,
 *                                the hand-over for a slab-sharded optimiser; a few hundred MB exchanged at 512^3 / 8 GPUs
 *   HPX_SHARD_RESULT_REPLICATED  additionally broadcasts every owned range: all ranks hold the whole summed gradient, as after
 *                                an all-reduce (hpx_grid_read_grad / hpx_grid_grad_buffer as usual)
 * Same field restrictions as hpx_shard_create.  All ranks must pass the same plan, grid shape and result mode. */
#define HPX_SHARD_RESULT_OWNED      1u
#define HPX_SHARD_RESULT_REPLICATED 2u
HP_API hp_status hpx_shard_create_bands(hpx_comm* comm, const hp_plan* full_frame_plan, hpx_grid* grid, uint32_t result,
                                        hpx_shard** out_shard);
HP_API hp_status hpx_shard_set_result(hpx_shard* shard, uint32_t result);
/* Collective (every rank calls it, after at least one hpx_shard_step).  Re-cuts the bands from MEASURED time: every rank
 * contributes the GPU time of its last step's forward + backward; bands that took longer than their share get fewer rows.
 * (The work estimate -- in-cube steps per ray -- does not see cache behaviour or idle lanes; two or three rounds during
 * warm-up bring the ranks within a few percent.)  *out_changed = 1 when the bands moved: this rank's frame, the wedges and
 * the owner cuts were rebuilt, so hpx_shard_frame / hpx_shard_owned / hpx_shard_bands must be queried again. */
HP_API hp_status hpx_shard_rebalance(hpx_shard* shard, int32_t* out_changed);
/* How the band exchange runs.  1 (default when every rank can map every other rank's gradient block -- peer access inside
 * one process, CUDA IPC between processes of one node): the library's OWN kernels over peer memory: after a stream-ordered
 * cross-GPU barrier every owner pulls the wedge parts of its slabs out of its neighbours' blocks over NVLink and adds them
 * in rank order (no staging copy, no separate add pass; the replicated result pulls the finished sums the same way).
 * 0 (environment DVREN_SHARD_EXCHANGE=nccl, or mapping failed on some rank): NCCL send/recv into a staging buffer + an add
 * kernel, NCCL broadcasts for the replicated result. */
HP_API hp_status hpx_shard_exchange_is_direct(const hpx_shard* shard, int32_t* out_direct);
/* The dispatch order (hpx_frame_set_row_order values) the library chose for this rank's band: the one of {rows first-to-last,
 * rows last-to-first, columns centre-out (+ either row order)} whose launches end soonest in a list-scheduling estimate of
 * the band's CTA tiles on the GPU's resident slots.  Environment DVREN_SHARD_TILE_ORDER=rows|columns overrides the choice
 * (timing comparisons). */
HP_API hp_status hpx_shard_tile_order(const hpx_shard* shard, int32_t* out_order);
/* Rank-local, no collective: settles that choice by MEASUREMENT -- this rank's band is rendered (forward + backward, no
 * exchange) under every candidate order and the fastest is kept, also across hpx_shard_rebalance.  The backward passes
 * accumulate into the gradient block: call it during warm-up, before hpx_shard_rebalance, and pass HPX_BACKWARD_ZERO on the
 * next step.  dL_dI / flags as for hpx_shard_step.  out_order (optional): the order now in place. */
HP_API hp_status hpx_shard_tune_order(hpx_shard* shard, const float* dL_dI_device, uint32_t flags, int32_t* out_order);
/* Host-only (works without a GPU): that choice for the band [row0, row0 + rows) of the plan's ROI on `slots` resident CTAs. */
HP_API hp_status hpx_plan_best_tile_order(const hp_plan* plan, uint32_t row0, uint32_t rows, uint32_t slots, int32_t* out_order,
                                          double* out_ends4);
/* Host-only (works without a GPU): the bands hpx_shard_create_bands cuts for `world` ranks -- first row inside the ROI and
 * rows per rank, multiples of the 8-row CTA tile; out_work (may be NULL): estimated marching work per band in samples. */
HP_API hp_status hpx_plan_balanced_bands(const hp_plan* plan, uint32_t world, uint32_t* out_row0, uint32_t* out_rows,
                                         double* out_work);
/* Host-only: the owner cuts the shard chooses for per-rank wedges [lo, hi) (2 * world ints) over n_slabs slabs and a result
 * mode; out_cuts: world + 1 ints, rank r owns slabs [cuts[r], cuts[r + 1]). */
HP_API hp_status hpx_plan_owner_cuts(uint32_t world, int32_t n_slabs, const int32_t* wedges, uint32_t result, int32_t* out_cuts);
/* Per rank (arrays of `world` entries; any pointer may be NULL): first image row inside the plan's ROI and rows of its band,
 * wedge [lo, hi) in slabs (2 * world ints), owner cuts (world + 1 ints: rank r owns slabs [cuts[r], cuts[r + 1])); floats
 * this rank sends / receives per step in the reduce-scatter part. */
HP_API hp_status hpx_shard_bands(const hpx_shard* shard, uint32_t* out_row0, uint32_t* out_rows, int32_t* out_wedges,
                                 int32_t* out_cuts, size_t* out_send_floats, size_t* out_recv_floats);
/* Device view of the slabs this rank owns inside the gradient block ({dr,dg,db,dsigma} per voxel, slab order of
 * hpx_grid_set_grad_layout with *out_slow_axis slowest): slabs [first, first + count) x slab_floats floats. */
HP_API hp_status hpx_shard_owned(const hpx_shard* shard, float** out_device_ptr, int32_t* out_first_slab, int32_t* out_slabs,
                                 size_t* out_slab_floats, int32_t* out_slow_axis);
/* Release order: a shard before its communicator, its grid and its plan's context (it keeps plain pointers to them and
 * puts the grid's gradient block back into the default order). */
HP_API void      hpx_shard_release(hpx_shard* shard);

#ifdef __cplusplus
}
#endif
#endif /* DVREN_HOTPATH_HP_B200_H_ */
