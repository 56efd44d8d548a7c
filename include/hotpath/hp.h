/*
 * hotpath/hp.h -- C ABI of the dvren hot path, B200-native implementation.
 *
 * This header is the drop-in boundary.  Every type below has the same name,
 * member order, member types and enumerator values as the interface it
 * replaces (reference: hotpath/include/hotpath/hp.h:20-216), so a translation
 * unit compiled against either header links against either library.  The
 * layout equivalence is checked by tests/test_abi_layout.py against a
 * committed sizeof/offsetof table (tests/golden/hp_abi_layout.json).
 *
 * Behavioural contract (what the library behind this header does):
 *   - every entry point returns an hp_status, never throws across the ABI;
 *   - tensors are unowned views (element strides); output tensors whose
 *     .data is NULL are bump-allocated from the caller's workspace with
 *     4-byte alignment, in the member order of the bundle;
 *   - each call dispatches on the memspace of one designated tensor
 *     (hp_ray/hp_samp/hp_samp_int_fused: rays->origins, hp_int: samp->sigma,
 *     hp_img: rays->pixel_ids, hp_diff: dL_dI).  HOST tensors are staged to
 *     the GPU, DEVICE tensors are used in place.  There is no CPU compute
 *     path in this library: without a usable CUDA device every compute entry
 *     point fails with HP_STATUS_UNSUPPORTED.
 *
 * B200-specific additive entry points (device fields, the non-materialising
 * fused forward, the recompute backward with grid scatter and camera adjoint,
 * streams, sharding) live in hotpath/hp_b200.h and never change this file.
 */
#ifndef DVREN_HOTPATH_HP_H_
#define DVREN_HOTPATH_HP_H_

#include <stddef.h>
#include <stdint.h>

#if defined(_MSC_VER) && defined(HP_BUILD_DLL)
#define HP_API __declspec(dllexport)
#elif defined(__GNUC__) && defined(HP_BUILD_DLL)
#define HP_API __attribute__((visibility("default")))
#else
#define HP_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- version (reference hp.h:20-28) ------------------------------------ */
#define HP_VERSION_MAJOR 0U
#define HP_VERSION_MINOR 1U
#define HP_VERSION_PATCH 0U

typedef struct hp_version { uint32_t major, minor, patch; } hp_version;

/* ---- enumerations (reference hp.h:30-70) -------------------------------- */
typedef enum hp_status {
    HP_STATUS_SUCCESS          = 0,
    HP_STATUS_INVALID_ARGUMENT = 1,
    HP_STATUS_OUT_OF_MEMORY    = 2,
    HP_STATUS_NOT_IMPLEMENTED  = 3,
    HP_STATUS_UNSUPPORTED      = 4,
    HP_STATUS_INTERNAL_ERROR   = 5
} hp_status;

typedef enum hp_memspace { HP_MEMSPACE_HOST = 0, HP_MEMSPACE_DEVICE = 1 } hp_memspace;

typedef enum hp_dtype {
    HP_DTYPE_F16 = 0, HP_DTYPE_BF16 = 1, HP_DTYPE_F32 = 2, HP_DTYPE_I32 = 3, HP_DTYPE_U32 = 4
} hp_dtype;

typedef enum hp_camera_model  { HP_CAMERA_PINHOLE = 0, HP_CAMERA_ORTHOGRAPHIC = 1 } hp_camera_model;
typedef enum hp_sampling_mode { HP_SAMPLING_FIXED = 0, HP_SAMPLING_STRATIFIED = 1 } hp_sampling_mode;
typedef enum hp_interp_mode   { HP_INTERP_NEAREST = 0, HP_INTERP_LINEAR = 1 } hp_interp_mode;
typedef enum hp_oob_policy    { HP_OOB_ZERO = 0, HP_OOB_CLAMP = 1 } hp_oob_policy;

/* ---- descriptors (reference hp.h:72-118) -------------------------------- */

/* Marching parameters: step length, step cap per ray, fixed (mid-point) or
 * stratified (hashed jitter) placement inside each step. */
typedef struct hp_sampling_desc {
    float            dt;
    uint32_t         max_steps;
    hp_sampling_mode mode;
} hp_sampling_desc;

/* Unowned strided view.  Strides are in elements, not bytes. */
typedef struct hp_tensor {
    void*       data;
    hp_dtype    dtype;
    hp_memspace memspace;
    uint32_t    rank;
    int64_t     shape[8];
    int64_t     stride[8];
} hp_tensor;

/* preferred_device: NULL / "" / "cuda" / "cuda:<ordinal>" selects the GPU.
 * reserved: NULL, or a pointer to an hpx_ctx_ext (see hp_b200.h). */
typedef struct hp_ctx_desc {
    uint32_t    flags;
    const char* preferred_device;
    const void* reserved;
} hp_ctx_desc;

/* Row-major K (fx=K[0], fy=K[4], cx=K[2], cy=K[5]) and 3x4 camera-to-world
 * (rotation in columns 0..2, translation in c2w[3], c2w[7], c2w[11]). */
typedef struct hp_camera_desc {
    hp_camera_model model;
    float           K[9];
    float           c2w[12];
    float           ortho_scale;
} hp_camera_desc;

typedef struct hp_roi_desc { uint32_t x, y, width, height; } hp_roi_desc;

typedef struct hp_plan_desc {
    uint32_t         width;
    uint32_t         height;
    float            t_near;
    float            t_far;
    uint32_t         max_rays;
    uint32_t         max_samples;
    uint64_t         seed;
    hp_camera_desc   camera;
    hp_roi_desc      roi;
    hp_sampling_desc sampling;
} hp_plan_desc;

/* ---- opaque handles (reference hp.h:120-122) ---------------------------- */
typedef struct hp_ctx   hp_ctx;
typedef struct hp_plan  hp_plan;
typedef struct hp_field hp_field;

/* ---- structure-of-arrays bundles (reference hp.h:124-160) --------------- */
typedef struct hp_rays_t {
    hp_tensor origins;      /* (N,3) f32                     */
    hp_tensor directions;   /* (N,3) f32, unit length        */
    hp_tensor t_near;       /* (N,)  f32                     */
    hp_tensor t_far;        /* (N,)  f32                     */
    hp_tensor pixel_ids;    /* (N,)  u32, y*width + x        */
} hp_rays_t;

typedef struct hp_samp_t {
    hp_tensor positions;    /* (M,3) f32                     */
    hp_tensor dt;           /* (M,)  f32                     */
    hp_tensor ray_offset;   /* (N+1,) u32 exclusive prefix   */
    hp_tensor sigma;        /* (M,)  f32                     */
    hp_tensor color;        /* (M,3) f32                     */
} hp_samp_t;

typedef struct hp_intl_t {
    hp_tensor radiance;      /* (N,3) f32                    */
    hp_tensor transmittance; /* (N,)  f32                    */
    hp_tensor opacity;       /* (N,)  f32                    */
    hp_tensor depth;         /* (N,)  f32                    */
    hp_tensor aux;           /* (M,4) f32: alpha, weight, T before, log T before */
} hp_intl_t;

typedef struct hp_img_t {
    hp_tensor image;        /* (H,W,3) f32                   */
    hp_tensor trans;        /* (H,W)   f32                   */
    hp_tensor opacity;      /* (H,W)   f32                   */
    hp_tensor depth;        /* (H,W)   f32                   */
    hp_tensor hitmask;      /* (H,W)   u32                   */
} hp_img_t;

typedef struct hp_grads_t {
    hp_tensor sigma;        /* (M,)  f32 per-sample          */
    hp_tensor color;        /* (M,3) f32 per-sample          */
    hp_tensor camera;       /* (3,4) f32 d/d c2w             */
} hp_grads_t;

/* ---- entry points (reference hp.h:162-190) ------------------------------ */
HP_API hp_version hp_get_version(void);

HP_API hp_status hp_ctx_create(const hp_ctx_desc* desc, hp_ctx** out_ctx);
HP_API void      hp_ctx_release(hp_ctx* ctx);
HP_API hp_status hp_ctx_get_desc(const hp_ctx* ctx, hp_ctx_desc* out_desc);

HP_API hp_status hp_plan_create(const hp_ctx* ctx, const hp_plan_desc* desc, hp_plan** out_plan);
HP_API void      hp_plan_release(hp_plan* plan);
HP_API hp_status hp_plan_get_desc(const hp_plan* plan, hp_plan_desc* out_desc);

HP_API hp_status hp_ray(const hp_plan* plan, const hp_rays_t* override_or_null, hp_rays_t* rays,
                        void* ws, size_t ws_bytes);
HP_API hp_status hp_samp(const hp_plan* plan, const hp_field* fs, const hp_field* fc,
                         const hp_rays_t* rays, hp_samp_t* samp, void* ws, size_t ws_bytes);
HP_API hp_status hp_int(const hp_plan* plan, const hp_samp_t* samp, hp_intl_t* intl,
                        void* ws, size_t ws_bytes);
HP_API hp_status hp_img(const hp_plan* plan, const hp_intl_t* intl, const hp_rays_t* rays,
                        hp_img_t* img, void* ws, size_t ws_bytes);
HP_API hp_status hp_diff(const hp_plan* plan, const hp_tensor* dL_dI, const hp_samp_t* samp,
                         const hp_intl_t* intl, hp_grads_t* grads, void* ws, size_t ws_bytes);

HP_API hp_status hp_field_create_grid_sigma(const hp_ctx* ctx, const hp_tensor* grid,
                                            uint32_t interp, uint32_t oob, hp_field** out_field);
HP_API hp_status hp_field_create_grid_color(const hp_ctx* ctx, const hp_tensor* grid,
                                            uint32_t interp, uint32_t oob, hp_field** out_field);
HP_API hp_status hp_field_create_hash_mlp(const hp_ctx* ctx, const hp_tensor* params,
                                          hp_field** out_field);
HP_API void      hp_field_release(hp_field* field);

HP_API hp_status hp_samp_int_fused(const hp_plan* plan, const hp_field* fs, const hp_field* fc,
                                   const hp_rays_t* rays, hp_samp_t* samp, hp_intl_t* intl,
                                   void* ws, size_t ws_bytes);

/* ---- CUDA graph pipeline (reference hp.h:192-216) ----------------------- */
#if defined(HP_WITH_CUDA)
HP_API hp_status hp_graph_create(const hp_plan* plan, const hp_field* fs, const hp_field* fc,
                                 size_t ws_ray_bytes, size_t ws_fused_bytes,
                                 size_t ws_img_bytes, size_t ws_diff_bytes,
                                 void** out_graph_handle);
HP_API hp_status hp_graph_capture(void* graph_handle, const hp_plan* plan, const hp_field* fs,
                                  const hp_field* fc, const hp_tensor* dL_dI);
HP_API hp_status hp_graph_execute(void* graph_handle, hp_rays_t* out_rays, hp_samp_t* out_samp,
                                  hp_intl_t* out_intl, hp_img_t* out_img, hp_grads_t* out_grads);
HP_API void      hp_graph_release(void* graph_handle);
#endif

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* DVREN_HOTPATH_HP_H_ */
