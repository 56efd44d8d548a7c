/*
 * oracle/dvren_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see the
 * header).  CPU restatement of the reference hot path; parity PINNED by
 * tests/test_oracle_pin.py.  Build: gcc -O2 -std=c11 -ffp-contract=off
 * (the reference objects contain no fused multiply-adds; SURVEY finding 7).
 *
 * Citations are `file:line` under /root/reference.
 */
#include "dvren_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* plan defaults: hotpath/src/runtime/hp_runtime.cpp:45-146                  */
/* ------------------------------------------------------------------------ */
int orc_plan_resolve(hp_plan_desc* p) {
    if (!p) return HP_STATUS_INVALID_ARGUMENT;
    if (p->width == 0 || p->height == 0) return HP_STATUS_INVALID_ARGUMENT;   /* :54-57 */
    if (!(p->t_far > p->t_near)) return HP_STATUS_INVALID_ARGUMENT;           /* :58-61 */

    hp_camera_desc* c = &p->camera;                                           /* :63-98 */
    if (c->model != HP_CAMERA_PINHOLE && c->model != HP_CAMERA_ORTHOGRAPHIC)
        c->model = HP_CAMERA_PINHOLE;
    int k_zero = 1, m_zero = 1;
    for (int i = 0; i < 9; ++i) if (c->K[i] != 0.0f) k_zero = 0;
    if (k_zero) {
        c->K[0] = c->K[4] = c->K[8] = 1.0f;
        c->K[2] = (float)p->width * 0.5f;
        c->K[5] = (float)p->height * 0.5f;
    }
    if (c->K[0] == 0.0f) c->K[0] = 1.0f;
    if (c->K[4] == 0.0f) c->K[4] = 1.0f;
    for (int i = 0; i < 12; ++i) if (c->c2w[i] != 0.0f) m_zero = 0;
    if (m_zero) c->c2w[0] = c->c2w[5] = c->c2w[10] = 1.0f;
    if (c->model == HP_CAMERA_ORTHOGRAPHIC && c->ortho_scale <= 0.0f) c->ortho_scale = 1.0f;

    hp_roi_desc* r = &p->roi;                                                 /* :100-111 */
    if (r->width == 0 || r->height == 0) {
        r->x = r->y = 0; r->width = p->width; r->height = p->height;
    }
    /* u32 wrap-around is part of the reference's arithmetic here */
    if ((uint32_t)(r->x + r->width) > p->width || (uint32_t)(r->y + r->height) > p->height)
        return HP_STATUS_INVALID_ARGUMENT;
    const uint64_t roi_rays = (uint64_t)r->width * (uint64_t)r->height;       /* :112-119 */
    if (p->max_rays == 0U)
        p->max_rays = (uint32_t)(roi_rays < UINT32_MAX ? roi_rays : UINT32_MAX);
    if (roi_rays > p->max_rays) return HP_STATUS_INVALID_ARGUMENT;

    hp_sampling_desc* s = &p->sampling;                                       /* :121-133 */
    if (!(s->dt > 0.0f)) {
        const float span = p->t_far - p->t_near;
        const float d = span > 0.0f ? span / 64.0f : 1.0f;
        s->dt = d > 0.0f ? d : 1.0f;
    }
    if (s->max_steps == 0U) s->max_steps = 64U;
    if (s->mode != HP_SAMPLING_FIXED && s->mode != HP_SAMPLING_STRATIFIED)
        s->mode = HP_SAMPLING_FIXED;

    if (p->max_samples == 0U) {                                               /* :135-142 */
        const uint64_t want = (uint64_t)p->max_rays * (uint64_t)s->max_steps;
        const uint64_t capped = want < UINT32_MAX ? want : UINT32_MAX;
        p->max_samples = capped == 0 ? p->max_rays : (uint32_t)capped;
    }
    if (p->max_samples < p->max_rays) return HP_STATUS_INVALID_ARGUMENT;
    return HP_STATUS_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* ray generation: hotpath/src/cpu/ray_cpu.cpp:158-226                       */
/* ------------------------------------------------------------------------ */
static void one_ray(const hp_plan_desc* p, uint32_t px, uint32_t py, float o[3], float d[3]) {
    const hp_camera_desc* c = &p->camera;
    const float fx = c->K[0], fy = c->K[4], cx = c->K[2], cy = c->K[5];       /* :159-162 */
    const float u = (float)px + 0.5f;                                          /* :189-190 */
    const float v = (float)py + 0.5f;
    float qx = (u - cx) / fx;                                                  /* :191-193 */
    float qy = (v - cy) / fy;
    float qz = 1.0f;
    if (c->model == HP_CAMERA_ORTHOGRAPHIC) { qx = 0.0f; qy = 0.0f; qz = 1.0f; } /* :195-199 */
    float wx = c->c2w[0] * qx + c->c2w[1] * qy + c->c2w[2] * qz;               /* :201-203 */
    float wy = c->c2w[4] * qx + c->c2w[5] * qy + c->c2w[6] * qz;
    float wz = c->c2w[8] * qx + c->c2w[9] * qy + c->c2w[10] * qz;
    const float len_sq = wx * wx + wy * wy + wz * wz;                          /* :205-208 */
    const float inv_len = 1.0f / sqrtf(len_sq > FLT_MIN ? len_sq : FLT_MIN);
    d[0] = wx * inv_len; d[1] = wy * inv_len; d[2] = wz * inv_len;             /* :209-211 */
    o[0] = c->c2w[3]; o[1] = c->c2w[7]; o[2] = c->c2w[11];                      /* :214-216 */
}

int orc_rays(const hp_plan_desc* p, float* origins, float* directions, float* t_near,
             float* t_far, uint32_t* pixel_ids) {
    if (!p) return HP_STATUS_INVALID_ARGUMENT;
    const hp_roi_desc r = p->roi;
    if ((uint64_t)r.width * r.height > p->max_rays) return HP_STATUS_INVALID_ARGUMENT; /* :133-135 */
    for (uint32_t ly = 0; ly < r.height; ++ly) {
        for (uint32_t lx = 0; lx < r.width; ++lx) {
            const size_t i = (size_t)ly * r.width + lx;                        /* :186-187 */
            const uint32_t px = r.x + lx, py = r.y + ly;
            float o[3], d[3];
            one_ray(p, px, py, o, d);
            if (origins)    { origins[3 * i] = o[0]; origins[3 * i + 1] = o[1]; origins[3 * i + 2] = o[2]; }
            if (directions) { directions[3 * i] = d[0]; directions[3 * i + 1] = d[1]; directions[3 * i + 2] = d[2]; }
            if (t_near) t_near[i] = p->t_near;                                 /* :222-223 */
            if (t_far)  t_far[i] = p->t_far;
            if (pixel_ids) pixel_ids[i] = py * p->width + px;                  /* :224 */
        }
    }
    return HP_STATUS_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* jitter: hotpath/src/cpu/samp_cpu.cpp:21-35                                */
/* ------------------------------------------------------------------------ */
static uint64_t mix64(uint64_t s) {
    s = (s ^ (s >> 30)) * 0xbf58476d1ce4e5b9ULL;
    s = (s ^ (s >> 27)) * 0x94d049bb133111ebULL;
    return s ^ (s >> 31);
}

float orc_jitter(uint64_t seed, uint64_t ray_index, uint32_t step) {
    uint64_t s = seed ^ (ray_index << 32) ^ (uint64_t)step;
    s = mix64(s);
    const double unit = (double)(s & 0x000fffffffffffffULL) / (double)0x0010000000000000ULL;
    return (float)unit;
}

/* ------------------------------------------------------------------------ */
/* dense grid: hotpath/src/cpu/grid_dense_cpu.cpp                            */
/* ------------------------------------------------------------------------ */
static float lerpf(float a, float b, float t) { return a + (b - a) * t; }     /* :52-54 */

static float fetch(const orc_grid* g, int32_t ix, int32_t iy, int32_t iz, int ch, int stride) {
    if (ix < 0 || ix >= g->nx || iy < 0 || iy >= g->ny || iz < 0 || iz >= g->nz) return 0.0f; /* :38-42 */
    return g->data[(size_t)((iz * g->ny + iy) * g->nx + ix) * (size_t)stride + (size_t)ch]; /* :44-50 */
}

/* :94-119 -- returns 1 when the query is outside and the policy is ZERO */
static int grid_coords(const orc_grid* g, const float pos[3], float f[3]) {
    int outside = 0;
    float local[3];
    for (int i = 0; i < 3; ++i) {
        const float extent = g->wmax[i] - g->wmin[i];
        const float c = extent != 0.0f ? (pos[i] - g->wmin[i]) / extent : 0.0f;
        local[i] = c;
        if (c < 0.0f || c > 1.0f) outside = 1;
    }
    if (g->oob == HP_OOB_CLAMP) {
        for (int i = 0; i < 3; ++i) {
            float v = local[i];
            v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
            local[i] = v;
        }
        outside = 0;
    }
    f[0] = local[0] * (float)(g->nx - 1);                                      /* :143-145 */
    f[1] = local[1] * (float)(g->ny - 1);
    f[2] = local[2] * (float)(g->nz - 1);
    return outside;
}

static float sample_channel(const orc_grid* g, const float f[3], int ch, int stride) {
    if (g->interp == HP_INTERP_NEAREST) {                                      /* :154-159 */
        return fetch(g, (int32_t)roundf(f[0]), (int32_t)roundf(f[1]), (int32_t)roundf(f[2]), ch, stride);
    }
    const int32_t x0 = (int32_t)floorf(f[0]), y0 = (int32_t)floorf(f[1]), z0 = (int32_t)floorf(f[2]); /* :58-60 */
    const int32_t x1 = x0 + 1 < g->nx - 1 ? x0 + 1 : g->nx - 1;                /* :61-63 */
    const int32_t y1 = y0 + 1 < g->ny - 1 ? y0 + 1 : g->ny - 1;
    const int32_t z1 = z0 + 1 < g->nz - 1 ? z0 + 1 : g->nz - 1;
    const float tx = f[0] - (float)x0, ty = f[1] - (float)y0, tz = f[2] - (float)z0; /* :64-66 */
    const float c000 = fetch(g, x0, y0, z0, ch, stride), c100 = fetch(g, x1, y0, z0, ch, stride);
    const float c010 = fetch(g, x0, y1, z0, ch, stride), c110 = fetch(g, x1, y1, z0, ch, stride);
    const float c001 = fetch(g, x0, y0, z1, ch, stride), c101 = fetch(g, x1, y0, z1, ch, stride);
    const float c011 = fetch(g, x0, y1, z1, ch, stride), c111 = fetch(g, x1, y1, z1, ch, stride);
    const float c00 = lerpf(c000, c100, tx), c10 = lerpf(c010, c110, tx);      /* :77-85 */
    const float c01 = lerpf(c001, c101, tx), c11 = lerpf(c011, c111, tx);
    const float c0 = lerpf(c00, c10, ty), c1 = lerpf(c01, c11, ty);
    return lerpf(c0, c1, tz);
}

float orc_grid_sigma(const orc_grid* g, const float pos[3]) {                  /* :125-162 */
    if (!g || !g->data || g->nx <= 0 || g->ny <= 0 || g->nz <= 0) return 0.0f;
    float f[3];
    if (grid_coords(g, pos, f)) return 0.0f;
    return sample_channel(g, f, 0, 1);
}

void orc_grid_color(const orc_grid* g, const float pos[3], float rgb[3]) {     /* :164-245 */
    rgb[0] = rgb[1] = rgb[2] = 0.0f;
    if (!g || !g->data || g->nx <= 0 || g->ny <= 0 || g->nz <= 0 || g->channels < 3) return;
    float f[3];
    if (grid_coords(g, pos, f)) return;
    for (int ch = 0; ch < 3; ++ch) rgb[ch] = sample_channel(g, f, ch, g->channels);
}

/* ------------------------------------------------------------------------ */
/* sampling: hotpath/src/cpu/samp_cpu.cpp:207-295                            */
/* ------------------------------------------------------------------------ */
typedef struct step_out { float t, dt; } step_out;

/* One iteration of the marching loop.  Returns 0 = emit, 1 = skip (continue),
 * 2 = stop (break).  (:226-244) */
static int march_step(const hp_plan_desc* p, float tn, float tf, uint64_t ray_index, uint32_t step,
                      step_out* o) {
    const float dts = p->sampling.dt;
    const float base = tn + (float)step * dts;
    if (base >= tf) return 2;
    float jit = 0.5f;
    if (p->sampling.mode == HP_SAMPLING_STRATIFIED) jit = orc_jitter(p->seed, ray_index, step);
    jit = jit < 0.0f ? 0.0f : (jit > 1.0f ? 1.0f : jit);
    float t = base + jit * dts;
    if (t >= tf) t = nextafterf(tf, tn);
    const float end = (base + dts) < tf ? (base + dts) : tf;
    const float dta = end - base;
    if (!(dta > 0.0f)) return 1;
    o->t = t; o->dt = dta;
    return 0;
}

uint32_t orc_ray_sample_count(const hp_plan_desc* p, float tn, float tf) {
    if (!(tf > tn)) return 0;                                                  /* :222-224 */
    uint32_t n = 0;
    /* the emit/skip/stop decision does not depend on the jitter value */
    hp_plan_desc q = *p;
    q.sampling.mode = HP_SAMPLING_FIXED;
    for (uint32_t step = 0; step < p->sampling.max_steps; ++step) {
        step_out so;
        const int r = march_step(&q, tn, tf, 0, step, &so);
        if (r == 2) break;
        if (r == 0) ++n;
    }
    return n;
}

int orc_sample(const hp_plan_desc* p, const orc_grid* gs, const orc_grid* gc, size_t n_rays,
               const float* origins, const float* directions, const float* t_near,
               const float* t_far, uint64_t ray_index_base, size_t capacity, float* positions,
               float* dt, float* sigma, float* color, uint32_t* ray_offset, size_t* out_count) {
    if (!p || (!gs && !gc)) return HP_STATUS_INVALID_ARGUMENT;                 /* :158-163 */
    if (n_rays > p->max_rays) return HP_STATUS_INVALID_ARGUMENT;               /* :172-175 */
    if (capacity == 0 && n_rays > 0) return HP_STATUS_INVALID_ARGUMENT;        /* :177-180 */
    memset(ray_offset, 0, (n_rays + 1) * sizeof(uint32_t));                    /* :189 */
    size_t total = 0;
    for (size_t ray = 0; ray < n_rays; ++ray) {
        ray_offset[ray] = (uint32_t)total;                                     /* :208 */
        const float* o = origins + 3 * ray;
        const float* d = directions + 3 * ray;
        const float tn = t_near[ray], tf = t_far[ray];
        if (!(tf > tn)) continue;                                              /* :222-224 */
        for (uint32_t step = 0; step < p->sampling.max_steps; ++step) {
            step_out so;
            const int r = march_step(p, tn, tf, ray_index_base + ray, step, &so);
            if (r == 2) break;
            if (r == 1) continue;
            if (total >= capacity) return HP_STATUS_INVALID_ARGUMENT;          /* :245-247 */
            float* pos = positions + 3 * total;
            pos[0] = o[0] + d[0] * so.t;                                       /* :250-252 */
            pos[1] = o[1] + d[1] * so.t;
            pos[2] = o[2] + d[2] * so.t;
            dt[total] = so.dt;
            sigma[total] = gs ? orc_grid_sigma(gs, pos) : 0.0f;                /* :255-270 */
            if (gc) orc_grid_color(gc, pos, color + 3 * total);                /* :272-289 */
            else color[3 * total] = color[3 * total + 1] = color[3 * total + 2] = 0.0f;
            ++total;
        }
    }
    ray_offset[n_rays] = (uint32_t)total;                                      /* :294 */
    if (out_count) *out_count = total;
    return HP_STATUS_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* integration: hotpath/src/cpu/int_cpu.cpp:98-109,160-226                   */
/* ------------------------------------------------------------------------ */
float orc_alpha(float sigma, float dt) {
    const float od = sigma * dt;
    if (od <= 0.0f) return 0.0f;
    if (od < 1e-4f) { const float half = 0.5f * od; return od * (1.0f - half); }
    const double a = -expm1(-(double)od);
    return (float)(a < 0.0 ? 0.0 : (a > 1.0 ? 1.0 : a));
}

/* The polynomial form of the same alpha that the CUDA kernels use (csrc/dv_device.cuh alpha_of), operation
 * for operation, so the claim "bit-identical to the libm form above" can be checked on the CPU for every
 * float (orc_alpha_fast_mismatches; tests/test_oracle_pin.py). */
float orc_alpha_fast(float sigma, float dt) {
    const float od = sigma * dt;
    if (od <= 0.0f) return 0.0f;
    if (od < 1e-4f) { const float half = 0.5f * od; return od * (1.0f - half); }
    if (od > 17.5f) return 1.0f;
    const float kf = rintf(od * 1.44269504f);
    const double r = fma((double)kf, 0.693147180559945309417, -(double)od);
    double q = 1.0 / 39916800.0;
    q = fma(q, r, 1.0 / 3628800.0);
    q = fma(q, r, 1.0 / 362880.0);
    q = fma(q, r, 1.0 / 40320.0);
    q = fma(q, r, 1.0 / 5040.0);
    q = fma(q, r, 1.0 / 720.0);
    q = fma(q, r, 1.0 / 120.0);
    q = fma(q, r, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    const double p = fma(r * r, q, r);
    const uint64_t bits = (uint64_t)(1023 - (int)kf) << 52;
    double s;
    memcpy(&s, &bits, sizeof s);
    return (float)fma(-s, p, 1.0 - s);
}

/* Number of floats od in [lo, hi] (every `stride`-th bit pattern) where the two forms differ. */
uint64_t orc_alpha_fast_mismatches(float lo, float hi, uint32_t stride, uint64_t* out_checked) {
    uint32_t a, b;
    memcpy(&a, &lo, 4);
    memcpy(&b, &hi, 4);
    uint64_t bad = 0, n = 0;
    for (uint64_t u = a; u <= b; u += (stride ? stride : 1)) {
        const uint32_t u32 = (uint32_t)u;
        float od;
        memcpy(&od, &u32, 4);
        const float x = orc_alpha(od, 1.0f), y = orc_alpha_fast(od, 1.0f);
        ++n;
        if (memcmp(&x, &y, 4) != 0) ++bad;
    }
    if (out_checked) *out_checked = n;
    return bad;
}

typedef struct ray_state { float T, depth_w, c[3], t_cursor; } ray_state;

/* One sample of the per-ray scan (:187-215).  Returns 1 when the ray stops. */
static int integrate_step(ray_state* s, float dtv, float sig, const float* col, float* aux_row) {
    float alpha = orc_alpha(sig, dtv);
    alpha = alpha < 0.0f ? 0.0f : (alpha > 1.0f ? 1.0f : alpha);
    const float T_before = s->T;
    const float w = T_before * alpha;
    s->c[0] += w * col[0]; s->c[1] += w * col[1]; s->c[2] += w * col[2];
    const float mid = s->t_cursor + 0.5f * dtv;
    s->depth_w += w * mid;
    if (aux_row) {
        aux_row[0] = alpha; aux_row[1] = w; aux_row[2] = T_before;
        aux_row[3] = logf(T_before > 1e-30f ? T_before : 1e-30f);
    }
    const float rem = (1.0f - alpha) > 0.0f ? (1.0f - alpha) : 0.0f;
    s->T *= rem;
    s->t_cursor += dtv;
    return s->T <= 1e-4f;
}

int orc_integrate(const hp_plan_desc* p, size_t n_rays, size_t n_samples, const float* dt,
                  const float* sigma, const float* color, const uint32_t* ray_offset,
                  float* radiance, float* transmittance, float* opacity, float* depth, float* aux) {
    if (!p) return HP_STATUS_INVALID_ARGUMENT;
    if (n_samples > p->max_samples || n_rays > p->max_rays) return HP_STATUS_INVALID_ARGUMENT; /* :137-140 */
    for (size_t r = 0; r < n_rays; ++r) {                                      /* :160-165 */
        radiance[3 * r] = radiance[3 * r + 1] = radiance[3 * r + 2] = 0.0f;
        transmittance[r] = 1.0f; opacity[r] = 0.0f; depth[r] = p->t_far;
    }
    if (aux) memset(aux, 0, n_samples * 4 * sizeof(float));                    /* :166-168 */
    for (size_t r = 0; r < n_rays; ++r) {
        const uint32_t b = ray_offset[r], e = ray_offset[r + 1];
        if (e < b || e > n_samples) return HP_STATUS_INVALID_ARGUMENT;         /* :176-178 */
        ray_state s = {1.0f, 0.0f, {0.0f, 0.0f, 0.0f}, p->t_near};
        for (uint32_t i = b; i < e; ++i)
            if (integrate_step(&s, dt[i], sigma[i], color + 3 * (size_t)i, aux ? aux + 4 * (size_t)i : NULL)) break;
        const float op = 1.0f - s.T;                                           /* :218-225 */
        radiance[3 * r] = s.c[0]; radiance[3 * r + 1] = s.c[1]; radiance[3 * r + 2] = s.c[2];
        transmittance[r] = s.T; opacity[r] = op;
        depth[r] = op > 1e-6f ? s.depth_w / op : p->t_far;
    }
    return HP_STATUS_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* backward: hotpath/src/cpu/diff_cpu.cpp:156-195                            */
/* ------------------------------------------------------------------------ */
static void diff_ray(const float g[3], uint32_t count, const float* dt, const float* color,
                     const float* aux, float* gs, float* gc) {
    float adj_T = 0.0f;
    for (uint32_t i = count; i-- > 0;) {
        const float alpha = aux[4 * i], w = aux[4 * i + 1], T_prev = aux[4 * i + 2];
        const float* c = color + 3 * i;
        const float dot = g[0] * c[0] + g[1] * c[1] + g[2] * c[2];            /* :177-179 */
        gc[3 * i] += g[0] * w; gc[3 * i + 1] += g[1] * w; gc[3 * i + 2] += g[2] * w; /* :181-183 */
        const float adj_alpha = dot * T_prev - adj_T * T_prev;                 /* :185 */
        const float adj_prev = dot * alpha + adj_T * (1.0f - alpha);           /* :186 */
        gs[i] += adj_alpha * (dt[i] * (1.0f - alpha));                         /* :188-189 */
        adj_T = adj_prev;
    }
}

int orc_diff(size_t n_rays, size_t n_samples, const float* dL_dI, int64_t stride_ray,
             int64_t stride_c, const float* dt, const float* color, const uint32_t* ray_offset,
             const float* aux, float* grad_sigma, float* grad_color) {
    memset(grad_sigma, 0, n_samples * sizeof(float));                          /* :76-79 */
    memset(grad_color, 0, n_samples * 3 * sizeof(float));
    if (n_samples == 0 || n_rays == 0) return HP_STATUS_SUCCESS;
    if (!aux) return HP_STATUS_INVALID_ARGUMENT;
    for (size_t r = 0; r < n_rays; ++r) {
        const uint32_t b = ray_offset[r], e = ray_offset[r + 1];
        if (e < b || e > n_samples) return HP_STATUS_INVALID_ARGUMENT;
        const float* gp = dL_dI + (int64_t)r * stride_ray;
        const float g[3] = {gp[0], gp[stride_c], gp[2 * stride_c]};
        diff_ray(g, e - b, dt + b, color + 3 * (size_t)b, aux + 4 * (size_t)b, grad_sigma + b,
                 grad_color + 3 * (size_t)b);
    }
    return HP_STATUS_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* sample -> grid scatter: src/fields/dense_grid.cpp:198-306                 */
/* ------------------------------------------------------------------------ */
/* Optional float64 shadow of the scatter (test adjudication only, no reference counterpart): the SAME float32
 * per-corner terms the reference adds, accumulated in double, plus the sum of their magnitudes, over a box of
 * voxels [o, o + n).  It separates "the float32 sum depends on the order of its terms" from real defects. */
static void shadow_add(orc_render_shadow* sh, int32_t ix, int32_t iy, int32_t iz, float ts, const float tc[3]) {
    const int32_t bx = ix - sh->box_o[0], by = iy - sh->box_o[1], bz = iz - sh->box_o[2];
    if (bx < 0 || bx >= sh->box_n[0] || by < 0 || by >= sh->box_n[1] || bz < 0 || bz >= sh->box_n[2]) {
        ++sh->misses;
        return;
    }
    const size_t v = ((size_t)bz * (size_t)sh->box_n[1] + (size_t)by) * (size_t)sh->box_n[0] + (size_t)bx;
    if (sh->sigma_sum) sh->sigma_sum[v] += (double)ts;
    if (sh->sigma_abs) sh->sigma_abs[v] += fabs((double)ts);
    for (int c = 0; c < 3; ++c) {
        if (sh->color_sum) sh->color_sum[3 * v + c] += (double)tc[c];
        if (sh->color_abs) sh->color_abs[3 * v + c] += fabs((double)tc[c]);
    }
}

static void scatter_one(const int32_t res[3], const float bmin[3], const float bmax[3],
                        uint32_t interp, uint32_t oob, const float pos[3], float gsig,
                        const float gcol[3], float* sg, float* cg, orc_render_shadow* sh) {
    const int32_t nx = res[0], ny = res[1], nz = res[2];
    float l[3];
    int outside = 0;
    for (int i = 0; i < 3; ++i) {                                              /* :211-217 */
        const float ext = bmax[i] - bmin[i];
        l[i] = ext != 0.0f ? (pos[i] - bmin[i]) / ext : 0.0f;
        if (l[i] < 0.0f || l[i] > 1.0f) outside = 1;
    }
    if (outside) {                                                             /* :219-226 */
        if (oob == HP_OOB_ZERO) return;
        for (int i = 0; i < 3; ++i) {
            const float lo = l[i] < 1.0f ? l[i] : 1.0f;   /* max(0, min(1, v)) */
            l[i] = lo > 0.0f ? lo : 0.0f;
        }
    }
    const float gx = l[0] * (float)(nx - 1 > 1 ? nx - 1 : 1);                  /* :228-230 */
    const float gy = l[1] * (float)(ny - 1 > 1 ? ny - 1 : 1);
    const float gz = l[2] * (float)(nz - 1 > 1 ? nz - 1 : 1);
    if (interp == HP_INTERP_NEAREST || nx == 1 || ny == 1 || nz == 1) {        /* :232-246 */
        const int32_t ix = (int32_t)roundf(gx), iy = (int32_t)roundf(gy), iz = (int32_t)roundf(gz);
        if (ix < 0 || ix >= nx || iy < 0 || iy >= ny || iz < 0 || iz >= nz) return;
        const size_t v = ((size_t)iz * (size_t)ny + (size_t)iy) * (size_t)nx + (size_t)ix;
        sg[v] += gsig;
        cg[3 * v] += gcol[0]; cg[3 * v + 1] += gcol[1]; cg[3 * v + 2] += gcol[2];
        if (sh) shadow_add(sh, ix, iy, iz, gsig, gcol);
        return;
    }
    const int32_t x0 = (int32_t)floorf(gx), y0 = (int32_t)floorf(gy), z0 = (int32_t)floorf(gz);
    const int32_t xs[2] = {x0, x0 + 1 < nx - 1 ? x0 + 1 : nx - 1};             /* :248-253 */
    const int32_t ys[2] = {y0, y0 + 1 < ny - 1 ? y0 + 1 : ny - 1};
    const int32_t zs[2] = {z0, z0 + 1 < nz - 1 ? z0 + 1 : nz - 1};
    const float tx = gx - (float)x0, ty = gy - (float)y0, tz = gz - (float)z0;
    const float wx[2] = {1.0f - tx, tx}, wy[2] = {1.0f - ty, ty}, wz[2] = {1.0f - tz, tz};
    for (int dx = 0; dx < 2; ++dx)                                             /* :286-304 */
        for (int dy = 0; dy < 2; ++dy)
            for (int dz = 0; dz < 2; ++dz) {
                const int32_t ix = xs[dx], iy = ys[dy], iz = zs[dz];
                if (ix < 0 || ix >= nx || iy < 0 || iy >= ny || iz < 0 || iz >= nz) continue;
                const float w = wx[dx] * wy[dy] * wz[dz];                      /* :259-266 */
                const size_t v = ((size_t)iz * (size_t)ny + (size_t)iy) * (size_t)nx + (size_t)ix;
                const float ts = gsig * w;
                const float tc[3] = {gcol[0] * w, gcol[1] * w, gcol[2] * w};
                sg[v] += ts;
                cg[3 * v] += tc[0]; cg[3 * v + 1] += tc[1]; cg[3 * v + 2] += tc[2];
                if (sh) shadow_add(sh, ix, iy, iz, ts, tc);
            }
}

int orc_scatter(const int32_t res[3], const float bbox_min[3], const float bbox_max[3],
                uint32_t interp, uint32_t oob, size_t n_samples, const float* positions,
                const float* grad_sigma, const float* grad_color, float* sigma_grad,
                float* color_grad) {
    if (res[0] <= 0 || res[1] <= 0 || res[2] <= 0) return HP_STATUS_INVALID_ARGUMENT;
    for (size_t i = 0; i < n_samples; ++i)
        scatter_one(res, bbox_min, bbox_max, interp, oob, positions + 3 * i, grad_sigma[i],
                    grad_color + 3 * i, sigma_grad, color_grad, NULL);
    return HP_STATUS_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* image composition: hotpath/src/cpu/img_cpu.cpp:154-185                    */
/* ------------------------------------------------------------------------ */
static void image_background(const hp_plan_desc* p, float* image, float* trans, float* opac,
                             float* depth_img, uint32_t* hitmask) {
    const size_t np = (size_t)p->width * p->height;                            /* :148-152 */
    for (size_t i = 0; i < np; ++i) {
        if (image) image[3 * i] = image[3 * i + 1] = image[3 * i + 2] = 0.0f;
        if (trans) trans[i] = 1.0f;
        if (opac) opac[i] = 0.0f;
        if (depth_img) depth_img[i] = p->t_far;
        if (hitmask) hitmask[i] = 0U;
    }
}

int orc_image(const hp_plan_desc* p, size_t n_rays, const uint32_t* pixel_ids,
              const float* radiance, const float* transmittance, const float* opacity,
              const float* depth, float* image, float* trans, float* opac, float* depth_img,
              uint32_t* hitmask) {
    const size_t np = (size_t)p->width * p->height;
    image_background(p, image, trans, opac, depth_img, hitmask);
    for (size_t r = 0; r < n_rays; ++r) {
        const uint32_t px = pixel_ids ? pixel_ids[r] : 0;
        if (px >= np) return HP_STATUS_INVALID_ARGUMENT;                       /* :156-158 */
        if (hitmask[px] == 0U) {                                               /* :162-169 */
            image[3 * (size_t)px] = radiance[3 * r]; image[3 * (size_t)px + 1] = radiance[3 * r + 1];
            image[3 * (size_t)px + 2] = radiance[3 * r + 2];
            trans[px] = transmittance[r]; opac[px] = opacity[r]; depth_img[px] = depth[r];
            hitmask[px] = 1U;
        } else {                                                               /* :170-177 */
            image[3 * (size_t)px] += radiance[3 * r]; image[3 * (size_t)px + 1] += radiance[3 * r + 1];
            image[3 * (size_t)px + 2] += radiance[3 * r + 2];
            trans[px] *= transmittance[r];
            opac[px] = 1.0f - trans[px];
            depth_img[px] = depth_img[px] < depth[r] ? depth_img[px] : depth[r];
        }
    }
    return HP_STATUS_SUCCESS;
}

/* ------------------------------------------------------------------------ */
/* whole path, one ray of samples at a time                                  */
/* ------------------------------------------------------------------------ */
int orc_render(const hp_plan_desc* p, const orc_grid* gs, const orc_grid* gc,
               uint64_t ray_index_base, const float* dL_dI, const int32_t res[3],
               const float bbox_min[3], const float bbox_max[3], orc_render_out* out) {
    return orc_render_shadowed(p, gs, gc, ray_index_base, dL_dI, res, bbox_min, bbox_max, out, NULL);
}

int orc_render_shadowed(const hp_plan_desc* p, const orc_grid* gs, const orc_grid* gc,
                        uint64_t ray_index_base, const float* dL_dI, const int32_t res[3],
                        const float bbox_min[3], const float bbox_max[3], orc_render_out* out,
                        orc_render_shadow* shadow) {
    if (!p || !out || (!gs && !gc)) return HP_STATUS_INVALID_ARGUMENT;
    const hp_roi_desc roi = p->roi;
    const uint32_t K = p->sampling.max_steps;
    float* pos = (float*)malloc((size_t)K * 3 * sizeof(float));
    float* dts = (float*)malloc((size_t)K * sizeof(float));
    float* sig = (float*)malloc((size_t)K * sizeof(float));
    float* col = (float*)malloc((size_t)K * 3 * sizeof(float));
    float* aux = (float*)malloc((size_t)K * 4 * sizeof(float));
    float* gsg = (float*)malloc((size_t)K * sizeof(float));
    float* gcl = (float*)malloc((size_t)K * 3 * sizeof(float));
    if (!pos || !dts || !sig || !col || !aux || !gsg || !gcl) return HP_STATUS_OUT_OF_MEMORY;
    image_background(p, out->image, out->trans, out->opac, out->depth_img, out->hitmask);
    uint64_t total = 0, live = 0;
    int status = HP_STATUS_SUCCESS;
    for (uint32_t ly = 0; ly < roi.height && status == HP_STATUS_SUCCESS; ++ly) {
        for (uint32_t lx = 0; lx < roi.width; ++lx) {
            const size_t ray = (size_t)ly * roi.width + lx;
            const uint32_t px = roi.x + lx, py = roi.y + ly;
            float o[3], d[3];
            one_ray(p, px, py, o, d);
            uint32_t n = 0;
            if (p->t_far > p->t_near) {
                for (uint32_t step = 0; step < K; ++step) {
                    step_out so;
                    const int r = march_step(p, p->t_near, p->t_far, ray_index_base + ray, step, &so);
                    if (r == 2) break;
                    if (r == 1) continue;
                    if (total + n >= p->max_samples) { status = HP_STATUS_INVALID_ARGUMENT; break; }
                    float* q = pos + 3 * n;
                    q[0] = o[0] + d[0] * so.t; q[1] = o[1] + d[1] * so.t; q[2] = o[2] + d[2] * so.t;
                    dts[n] = so.dt;
                    sig[n] = gs ? orc_grid_sigma(gs, q) : 0.0f;
                    if (gc) orc_grid_color(gc, q, col + 3 * n);
                    else col[3 * n] = col[3 * n + 1] = col[3 * n + 2] = 0.0f;
                    ++n;
                }
                if (status != HP_STATUS_SUCCESS) break;
            }
            memset(aux, 0, (size_t)n * 4 * sizeof(float));
            ray_state s = {1.0f, 0.0f, {0.0f, 0.0f, 0.0f}, p->t_near};
            uint32_t used = 0;
            for (uint32_t i = 0; i < n; ++i) {
                ++used;
                if (integrate_step(&s, dts[i], sig[i], col + 3 * i, aux + 4 * i)) break;
            }
            total += n; live += used;
            const float op = 1.0f - s.T;
            const float dep = op > 1e-6f ? s.depth_w / op : p->t_far;
            if (out->radiance) { out->radiance[3 * ray] = s.c[0]; out->radiance[3 * ray + 1] = s.c[1]; out->radiance[3 * ray + 2] = s.c[2]; }
            if (out->transmittance) out->transmittance[ray] = s.T;
            if (out->opacity) out->opacity[ray] = op;
            if (out->depth) out->depth[ray] = dep;
            const size_t pid = (size_t)py * p->width + px;
            if (out->image) { out->image[3 * pid] = s.c[0]; out->image[3 * pid + 1] = s.c[1]; out->image[3 * pid + 2] = s.c[2]; }
            if (out->trans) out->trans[pid] = s.T;
            if (out->opac) out->opac[pid] = op;
            if (out->depth_img) out->depth_img[pid] = dep;
            if (out->hitmask) out->hitmask[pid] = 1U;
            if (dL_dI && out->sigma_grad && out->color_grad && n > 0) {
                memset(gsg, 0, (size_t)n * sizeof(float));
                memset(gcl, 0, (size_t)n * 3 * sizeof(float));
                diff_ray(dL_dI + 3 * ray, n, dts, col, aux, gsg, gcl);
                for (uint32_t i = 0; i < n; ++i)
                    scatter_one(res, bbox_min, bbox_max, gs ? gs->interp : gc->interp,
                                gs ? gs->oob : gc->oob, pos + 3 * i, gsg[i], gcl + 3 * i,
                                out->sigma_grad, out->color_grad, shadow);
            }
        }
    }
    out->sample_count = total;
    out->live_sample_count = live;
    free(pos); free(dts); free(sig); free(col); free(aux); free(gsg); free(gcl);
    return status;
}

/* ------------------------------------------------------------------------ */
/* analytic camera adjoint (no reference counterpart)  */
/* ------------------------------------------------------------------------ */
/* value and d/d(position) of one channel of the trilinear interpolant */
static double channel_grad(const orc_grid* g, const float pos[3], int ch, int stride, double grad[3]) {
    grad[0] = grad[1] = grad[2] = 0.0;
    double scale[3];
    float l[3];
    int outside = 0, clamped[3] = {0, 0, 0};
    const int32_t n[3] = {g->nx, g->ny, g->nz};
    for (int i = 0; i < 3; ++i) {
        const float ext = g->wmax[i] - g->wmin[i];
        l[i] = ext != 0.0f ? (pos[i] - g->wmin[i]) / ext : 0.0f;
        scale[i] = ext != 0.0f ? (double)(n[i] - 1) / (double)ext : 0.0;
        if (l[i] < 0.0f || l[i] > 1.0f) { outside = 1; clamped[i] = 1; }
    }
    if (g->oob == HP_OOB_CLAMP) {
        for (int i = 0; i < 3; ++i) l[i] = l[i] < 0.0f ? 0.0f : (l[i] > 1.0f ? 1.0f : l[i]);
    } else if (outside) {
        return 0.0;
    }
    const float f[3] = {l[0] * (float)(n[0] - 1), l[1] * (float)(n[1] - 1), l[2] * (float)(n[2] - 1)};
    if (g->interp == HP_INTERP_NEAREST)
        return fetch(g, (int32_t)roundf(f[0]), (int32_t)roundf(f[1]), (int32_t)roundf(f[2]), ch, stride);
    int32_t i0[3], i1[3];
    double t[3];
    for (int i = 0; i < 3; ++i) {
        i0[i] = (int32_t)floorf(f[i]);
        i1[i] = i0[i] + 1 < n[i] - 1 ? i0[i] + 1 : n[i] - 1;
        t[i] = (double)f[i] - (double)i0[i];
    }
    double c[2][2][2];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int e = 0; e < 2; ++e)
        c[a][b][e] = fetch(g, a ? i1[0] : i0[0], b ? i1[1] : i0[1], e ? i1[2] : i0[2], ch, stride);
    const double wx[2] = {1.0 - t[0], t[0]}, wy[2] = {1.0 - t[1], t[1]}, wz[2] = {1.0 - t[2], t[2]};
    double val = 0.0;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int e = 0; e < 2; ++e) {
        val += c[a][b][e] * wx[a] * wy[b] * wz[e];
        grad[0] += c[a][b][e] * (a ? 1.0 : -1.0) * wy[b] * wz[e];
        grad[1] += c[a][b][e] * wx[a] * (b ? 1.0 : -1.0) * wz[e];
        grad[2] += c[a][b][e] * wx[a] * wy[b] * (e ? 1.0 : -1.0);
    }
    for (int i = 0; i < 3; ++i) grad[i] = clamped[i] ? 0.0 : grad[i] * scale[i];
    return val;
}

int orc_camera_grad(const hp_plan_desc* p, const orc_grid* gs, const orc_grid* gc,
                    uint64_t ray_index_base, const float* dL_dI, double out16[16]) {
    return orc_camera_grad_mag(p, gs, gc, ray_index_base, dL_dI, out16, NULL);
}

/* mag16 (may be NULL): the same chain rule applied to the per-sample MAGNITUDES |g_x|, |g_x t|: an upper bound of the
 * sum of |terms| behind each of the 16 outputs -- the scale float32 rounding of a different summation order refers to. */
int orc_camera_grad_mag(const hp_plan_desc* p, const orc_grid* gs, const orc_grid* gc,
                        uint64_t ray_index_base, const float* dL_dI, double out16[16], double mag16[16]) {
    if (!p || !gs || !gc || !dL_dI || !out16) return HP_STATUS_INVALID_ARGUMENT;
    for (int i = 0; i < 16; ++i) out16[i] = 0.0;
    if (mag16) for (int i = 0; i < 16; ++i) mag16[i] = 0.0;
    const hp_roi_desc roi = p->roi;
    const hp_camera_desc* cam = &p->camera;
    const uint32_t K = p->sampling.max_steps;
    double* alpha = (double*)malloc((size_t)K * sizeof(double));
    double* Tprev = (double*)malloc((size_t)K * sizeof(double));
    double* tval = (double*)malloc((size_t)K * sizeof(double));
    double* dtv = (double*)malloc((size_t)K * sizeof(double));
    double* cval = (double*)malloc((size_t)K * 3 * sizeof(double));
    double* gsv = (double*)malloc((size_t)K * 3 * sizeof(double));
    double* gcv = (double*)malloc((size_t)K * 9 * sizeof(double));
    if (!alpha || !Tprev || !tval || !dtv || !cval || !gsv || !gcv) return HP_STATUS_OUT_OF_MEMORY;
    for (uint32_t ly = 0; ly < roi.height; ++ly) for (uint32_t lx = 0; lx < roi.width; ++lx) {
        const size_t ray = (size_t)ly * roi.width + lx;
        const uint32_t px = roi.x + lx, py = roi.y + ly;
        float o[3], d[3];
        one_ray(p, px, py, o, d);
        uint32_t n = 0;
        float T = 1.0f;
        for (uint32_t step = 0; step < K; ++step) {
            step_out so;
            const int r = march_step(p, p->t_near, p->t_far, ray_index_base + ray, step, &so);
            if (r == 2) break;
            if (r == 1) continue;
            const float q[3] = {o[0] + d[0] * so.t, o[1] + d[1] * so.t, o[2] + d[2] * so.t};
            const double sv = channel_grad(gs, q, 0, 1, gsv + 3 * n);
            for (int ch = 0; ch < 3; ++ch) cval[3 * n + ch] = channel_grad(gc, q, ch, gc->channels, gcv + 9 * n + 3 * ch);
            float a = orc_alpha((float)sv, so.dt);
            a = a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a);
            alpha[n] = a; Tprev[n] = T; tval[n] = so.t; dtv[n] = so.dt;
            const float rem = (1.0f - a) > 0.0f ? (1.0f - a) : 0.0f;
            T *= rem;
            ++n;
            if (T <= 1e-4f) break;
        }
        const float* g = dL_dI + 3 * ray;
        double adj_T = 0.0, dLdo[3] = {0, 0, 0}, dLdd[3] = {0, 0, 0}, Ao[3] = {0, 0, 0}, Ad[3] = {0, 0, 0};
        for (uint32_t i = n; i-- > 0;) {
            const double dot = g[0] * cval[3 * i] + g[1] * cval[3 * i + 1] + g[2] * cval[3 * i + 2];
            const double w = Tprev[i] * alpha[i];
            const double adj_alpha = dot * Tprev[i] - adj_T * Tprev[i];
            const double gsig = adj_alpha * dtv[i] * (1.0 - alpha[i]);
            adj_T = dot * alpha[i] + adj_T * (1.0 - alpha[i]);
            for (int ax = 0; ax < 3; ++ax) {
                double gx = gsig * gsv[3 * i + ax];
                double ga = fabs(gx);
                for (int ch = 0; ch < 3; ++ch) {
                    gx += g[ch] * w * gcv[9 * i + 3 * ch + ax];
                    ga += fabs(g[ch] * w * gcv[9 * i + 3 * ch + ax]);
                }
                dLdo[ax] += gx;
                dLdd[ax] += gx * tval[i];
                Ao[ax] += ga;
                Ad[ax] += ga * fabs(tval[i]);
            }
        }
        /* d = v / |v|, v = R q */
        const double fx = cam->K[0], fy = cam->K[4], cx = cam->K[2], cy = cam->K[5];
        double q[3] = {((double)px + 0.5 - cx) / fx, ((double)py + 0.5 - cy) / fy, 1.0};
        const int ortho = cam->model == HP_CAMERA_ORTHOGRAPHIC;
        if (ortho) { q[0] = 0.0; q[1] = 0.0; }
        double v[3];
        for (int i = 0; i < 3; ++i) v[i] = cam->c2w[4 * i] * q[0] + cam->c2w[4 * i + 1] * q[1] + cam->c2w[4 * i + 2] * q[2];
        const double len = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        const double dn[3] = {v[0] / len, v[1] / len, v[2] / len};
        const double dd = dn[0] * dLdd[0] + dn[1] * dLdd[1] + dn[2] * dLdd[2];
        double dLdv[3];
        for (int i = 0; i < 3; ++i) dLdv[i] = (dLdd[i] - dn[i] * dd) / len;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) out16[4 * i + j] += dLdv[i] * q[j];
            out16[4 * i + 3] += dLdo[i];
        }
        if (!ortho) {
            double dLdq[2] = {0, 0};
            for (int i = 0; i < 3; ++i) { dLdq[0] += cam->c2w[4 * i] * dLdv[i]; dLdq[1] += cam->c2w[4 * i + 1] * dLdv[i]; }
            out16[12] += -q[0] / fx * dLdq[0];
            out16[13] += -q[1] / fy * dLdq[1];
            out16[14] += -dLdq[0] / fx;
            out16[15] += -dLdq[1] / fy;
        }
        if (mag16) {
            const double ad = fabs(dn[0]) * Ad[0] + fabs(dn[1]) * Ad[1] + fabs(dn[2]) * Ad[2];
            double Av[3];
            for (int i = 0; i < 3; ++i) Av[i] = (Ad[i] + fabs(dn[i]) * ad) / len;
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) mag16[4 * i + j] += Av[i] * fabs(q[j]);
                mag16[4 * i + 3] += Ao[i];
            }
            if (!ortho) {
                double Aq[2] = {0, 0};
                for (int i = 0; i < 3; ++i) { Aq[0] += fabs(cam->c2w[4 * i]) * Av[i]; Aq[1] += fabs(cam->c2w[4 * i + 1]) * Av[i]; }
                mag16[12] += fabs(q[0] / fx) * Aq[0];
                mag16[13] += fabs(q[1] / fy) * Aq[1];
                mag16[14] += Aq[0] / fabs(fx);
                mag16[15] += Aq[1] / fabs(fy);
            }
        }
    }
    free(alpha); free(Tprev); free(tval); free(dtv); free(cval); free(gsv); free(gcv);
    return HP_STATUS_SUCCESS;
}
