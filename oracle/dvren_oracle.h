/*
 * oracle/dvren_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's CPU hot path (single thread, no FMA
 * contraction).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * library (libdvren_hp.so) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_pin.py checks every function here
 * bit-for-bit against (a) golden vectors generated from the unmodified
 * reference compiled from /root/reference (oracle/_ref, recipe in
 * oracle/Makefile; fixtures under tests/golden/ with their generator
 * tests/golden/make_golden.py) and (b) the live oracle/_ref library whenever
 * it is present.  The camera-gradient function has no reference counterpart
 * (the reference returns zeros, diff_cpu.cpp:25,73-74) and is pinned by
 * central finite differences of the pinned forward instead.
 *
 * All `file:line` citations are relative to /root/reference.
 */
#ifndef DVREN_ORACLE_H_
#define DVREN_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#include "hotpath/hp.h" /* POD descriptors only (hp_plan_desc, enums) */

#ifdef __cplusplus
extern "C" {
#endif

/* Dense grid view as the reference's hp_field holds it (hp_internal.hpp:24-31).
 * sigma: data[nz][ny][nx]; colour: data[nz][ny][nx][channels]. */
typedef struct orc_grid {
    const float* data;
    int32_t nx, ny, nz, channels;
    uint32_t interp; /* hp_interp_mode */
    uint32_t oob;    /* hp_oob_policy  */
    float wmin[3], wmax[3];
} orc_grid;

/* hp_plan_create validation + defaulting, in place.  Returns an hp_status. */
int orc_plan_resolve(hp_plan_desc* desc);

/* hp_ray over the plan's ROI; arrays sized roi.width*roi.height. */
int orc_rays(const hp_plan_desc* desc, float* origins, float* directions, float* t_near,
             float* t_far, uint32_t* pixel_ids);

/* Number of samples the marching loop emits for one ray (no field access). */
uint32_t orc_ray_sample_count(const hp_plan_desc* desc, float ray_t_near, float ray_t_far);

/* Stratified jitter in [0,1] for (seed, ray index, step). */
float orc_jitter(uint64_t seed, uint64_t ray_index, uint32_t step);

/* Field queries. */
float orc_grid_sigma(const orc_grid* g, const float pos[3]);
void  orc_grid_color(const orc_grid* g, const float pos[3], float rgb[3]);

/* hp_samp.  gs or gc may be NULL (not both).  ray_index_base is added to the
 * ray index fed to the jitter hash (0 reproduces the reference; a shard of a
 * larger plan passes its first global ray index). */
int orc_sample(const hp_plan_desc* desc, const orc_grid* gs, const orc_grid* gc, size_t n_rays,
               const float* origins, const float* directions, const float* t_near,
               const float* t_far, uint64_t ray_index_base, size_t capacity, float* positions,
               float* dt, float* sigma, float* color, uint32_t* ray_offset, size_t* out_count);

/* Optical depth -> alpha. */
float orc_alpha(float sigma, float dt);
/* Same value through the fp64 polynomial the CUDA kernels use; and the exhaustive comparison. */
float orc_alpha_fast(float sigma, float dt);
uint64_t orc_alpha_fast_mismatches(float lo, float hi, uint32_t stride, uint64_t* out_checked);

/* hp_int.  aux may be NULL. */
int orc_integrate(const hp_plan_desc* desc, size_t n_rays, size_t n_samples, const float* dt,
                  const float* sigma, const float* color, const uint32_t* ray_offset,
                  float* radiance, float* transmittance, float* opacity, float* depth, float* aux);

/* hp_diff: per-sample gradients from per-ray dL/dI (strided). */
int orc_diff(size_t n_rays, size_t n_samples, const float* dL_dI, int64_t stride_ray,
             int64_t stride_c, const float* dt, const float* color, const uint32_t* ray_offset,
             const float* aux, float* grad_sigma, float* grad_color);

/* DenseGridField::AccumulateSampleGradients (accumulates, does not zero). */
int orc_scatter(const int32_t res[3], const float bbox_min[3], const float bbox_max[3],
                uint32_t interp, uint32_t oob, size_t n_samples, const float* positions,
                const float* grad_sigma, const float* grad_color, float* sigma_grad,
                float* color_grad);

/* hp_img. */
int orc_image(const hp_plan_desc* desc, size_t n_rays, const uint32_t* pixel_ids,
              const float* radiance, const float* transmittance, const float* opacity,
              const float* depth, float* image, float* trans, float* opac, float* depth_img,
              uint32_t* hitmask);

/* Whole path without materialising more than one ray of samples: identical
 * arithmetic and accumulation order to orc_rays -> orc_sample -> orc_integrate
 * -> orc_image [-> orc_diff -> orc_scatter].  dL_dI (N,3 contiguous, per ray in
 * plan order) may be NULL for forward only; sigma_grad/color_grad are
 * accumulated into (caller zeroes).  Per-ray outputs (any may be NULL) are
 * indexed by plan ray order; image outputs (any may be NULL) are full frames
 * that are first filled with the background like hp_img does. */
typedef struct orc_render_out {
    float* radiance; float* transmittance; float* opacity; float* depth; /* per ray */
    float* image; float* trans; float* opac; float* depth_img; uint32_t* hitmask; /* per pixel */
    float* sigma_grad; float* color_grad; /* [V], [3V], bbox scatter */
    uint64_t sample_count;      /* reference sample_count (all emitted samples) */
    uint64_t live_sample_count; /* samples integrated before the T<=1e-4 stop  */
} orc_render_out;

int orc_render(const hp_plan_desc* desc, const orc_grid* gs, const orc_grid* gc,
               uint64_t ray_index_base, const float* dL_dI, const int32_t res[3],
               const float bbox_min[3], const float bbox_max[3], orc_render_out* out);

/* orc_render with a float64 SHADOW of the gradient scatter over a box of voxels (test adjudication; no reference
 * counterpart).  The float32 outputs are untouched -- same arithmetic, same order.  The shadow receives, per voxel of
 * the box [box_o, box_o + box_n) (x fastest), the sum in double of the very float32 terms the reference adds
 * (`*_sum`) and the sum of their magnitudes (`*_abs`); any pointer may be NULL.  `misses` counts contributions that
 * fell outside the box.  Lets a test tell "two float32 summation orders of the same terms" from a defect. */
typedef struct orc_render_shadow {
    int32_t box_o[3], box_n[3];
    double* sigma_sum; double* sigma_abs;   /* [box voxels]     */
    double* color_sum; double* color_abs;   /* [3 * box voxels] */
    uint64_t misses;
} orc_render_shadow;

int orc_render_shadowed(const hp_plan_desc* desc, const orc_grid* gs, const orc_grid* gc,
                        uint64_t ray_index_base, const float* dL_dI, const int32_t res[3],
                        const float bbox_min[3], const float bbox_max[3], orc_render_out* out,
                        orc_render_shadow* shadow);

/* Analytic camera adjoint (double accumulation), SURVEY Appendix A.11.  No
 * reference counterpart.  out16 = d/d c2w[12] followed by d/d {fx,fy,cx,cy}. */
int orc_camera_grad(const hp_plan_desc* desc, const orc_grid* gs, const orc_grid* gc,
                    uint64_t ray_index_base, const float* dL_dI, double out16[16]);

/* Same, also returning per output an upper bound of the sum of the magnitudes of its terms (mag16, may be NULL). */
int orc_camera_grad_mag(const hp_plan_desc* desc, const orc_grid* gs, const orc_grid* gc,
                        uint64_t ray_index_base, const float* dL_dI, double out16[16], double mag16[16]);

#ifdef __cplusplus
}
#endif
#endif
