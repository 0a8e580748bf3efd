/*
 * pvp_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 kernels.  It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product path (libpvp_b200.so) never
 * links, imports or calls anything in here.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference's own code compiled unmodified
 * into oracle/_ref/ (see oracle/Makefile, oracle/ref_harness_*.cpp) by
 * tests/test_oracle_pin.py, and against tests/golden/ fixtures that were
 * generated from oracle/_ref by tests/golden/make_golden.py.
 *
 * Every function cites the reference lines it follows.  Arithmetic is written
 * operation-for-operation (fp32 for path A, fp64 for path B); build with
 * -ffp-contract=off so no FMA is ever formed (the reference build has none).
 *
 * Plain C11.  Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -shared -fPIC
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846 /* process_image.cpp:12-14 */
#endif

#define PVPO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* Path A: ray_voxel.cpp                                               */
/* ------------------------------------------------------------------ */

/* ray_voxel.cpp:60-62 */
static inline float o_deg2rad(float deg) { return deg * 3.14159265358979323846f / 180.0f; }

/* ray_voxel.cpp:64-70 */
static inline void o_normalize(const float v[3], float out[3])
{
    float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (len < 1e-12f) {
        out[0] = 0.f; out[1] = 0.f; out[2] = 0.f;
        return;
    }
    out[0] = v[0] / len; out[1] = v[1] / len; out[2] = v[2] / len;
}

/* ray_voxel.cpp:73-79 */
static inline void o_mat3_mul_vec3(const float M[9], const float v[3], float r[3])
{
    r[0] = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    r[1] = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    r[2] = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
}

/* ray_voxel.cpp:115-124 */
static void o_matmul3x3(const float A[9], const float B[9], float C[9])
{
    for (int row = 0; row < 3; ++row)
        for (int col = 0; col < 3; ++col)
            C[row * 3 + col] = A[row * 3 + 0] * B[0 * 3 + col] +
                               A[row * 3 + 1] * B[1 * 3 + col] +
                               A[row * 3 + 2] * B[2 * 3 + col];
}

/* ray_voxel.cpp:84-135 -- R = Rz(yaw) * Ry(roll) * Rx(pitch); note roll drives Ry, pitch drives Rx */
PVPO_API void pvpo_rotation_matrix(float yaw_deg, float pitch_deg, float roll_deg, float R[9])
{
    float y = o_deg2rad(yaw_deg), p = o_deg2rad(pitch_deg), r = o_deg2rad(roll_deg);
    float cy = cosf(y), sy = sinf(y);
    float Rz[9] = { cy, -sy, 0.f, sy, cy, 0.f, 0.f, 0.f, 1.f };
    float cr = cosf(r), sr = sinf(r);
    float Ry[9] = { cr, 0.f, sr, 0.f, 1.f, 0.f, -sr, 0.f, cr };
    float cp = cosf(p), sp = sinf(p);
    float Rx[9] = { 1.f, 0.f, 0.f, 0.f, cp, -sp, 0.f, sp, cp };
    float Rt[9];
    o_matmul3x3(Rz, Ry, Rt);
    o_matmul3x3(Rt, Rx, R);
}

/* ray_voxel.cpp:490-491 */
PVPO_API float pvpo_focal_length(float fov_degrees, int width)
{
    float fov_rad = o_deg2rad(fov_degrees);
    return (width * 0.5f) / tanf(fov_rad * 0.5f);
}

/* ray_voxel.cpp:236-255 (equal-size case; the size-mismatch early-out is host logic) */
PVPO_API void pvpo_detect_motion(const float *prev, const float *next, int64_t n, float threshold,
                                 uint8_t *changed, float *diff)
{
    for (int64_t i = 0; i < n; ++i) {
        float d = fabsf(prev[i] - next[i]);
        diff[i] = d;
        changed[i] = (d > threshold) ? 1 : 0;
    }
}

/* ray_voxel.cpp:506-515 */
PVPO_API void pvpo_pixel_ray(int u, int v, int width, int height, float focal_len, const float R[9],
                             float dir_out[3])
{
    float cam[3], n1[3], w[3];
    cam[0] = ((float)u - 0.5f * width);
    cam[1] = -((float)v - 0.5f * height);
    cam[2] = -focal_len;
    o_normalize(cam, n1);
    o_mat3_mul_vec3(R, n1, w);
    o_normalize(w, dir_out);
}

/* ray_voxel.cpp:266-272 */
static inline float o_safe_div(float num, float den)
{
    float eps = 1e-12f;
    if (fabsf(den) < eps) return INFINITY;
    return num / den;
}

/*
 * ray_voxel.cpp:274-395 restated as a visitor: calls emit(ctx, ix, iy, iz, t) for
 * every RayStep the reference would push_back, in order.  Returns the step count.
 */
typedef void (*o_emit_fn)(void *ctx, int ix, int iy, int iz, float t);

static inline int64_t o_cast_ray(const float cam[3], const float dir[3], int N, float voxel_size,
                                 const float center[3], o_emit_fn emit, void *ctx)
{
    float half_size = 0.5f * (N * voxel_size);                     /* :284 */
    float gmin[3] = { center[0] - half_size, center[1] - half_size, center[2] - half_size };
    float gmax[3] = { center[0] + half_size, center[1] + half_size, center[2] + half_size };

    float t_min = 0.f;                                             /* :292 */
    float t_max = INFINITY;                                        /* :293 */

    for (int i = 0; i < 3; ++i) {                                  /* :296-317 */
        float origin = cam[i], d = dir[i], mn = gmin[i], mx = gmax[i];
        if (fabsf(d) < 1e-12f) {
            if (origin < mn || origin > mx) return 0;
        } else {
            float t1 = (mn - origin) / d;
            float t2 = (mx - origin) / d;
            float t_near = fminf(t1, t2);
            float t_far = fmaxf(t1, t2);
            if (t_near > t_min) t_min = t_near;
            if (t_far < t_max) t_max = t_far;
            if (t_min > t_max) return 0;
        }
    }
    if (t_min < 0.f) t_min = 0.f;                                  /* :319 */

    float sx = cam[0] + t_min * dir[0];                            /* :322-324 */
    float sy = cam[1] + t_min * dir[1];
    float sz = cam[2] + t_min * dir[2];
    float fx = (sx - gmin[0]) / voxel_size;                        /* :325-327 */
    float fy = (sy - gmin[1]) / voxel_size;
    float fz = (sz - gmin[2]) / voxel_size;
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;                  /* :329-331 */
    if (ix < 0 || ix >= N || iy < 0 || iy >= N || iz < 0 || iz >= N) return 0;   /* :332-334 */

    int step_x = (dir[0] >= 0.f) ? 1 : -1;                         /* :337-339 */
    int step_y = (dir[1] >= 0.f) ? 1 : -1;
    int step_z = (dir[2] >= 0.f) ? 1 : -1;

    int nx_x = ix + (step_x > 0 ? 1 : 0);                          /* :345-347 */
    int nx_y = iy + (step_y > 0 ? 1 : 0);
    int nx_z = iz + (step_z > 0 ? 1 : 0);
    float next_bx = gmin[0] + nx_x * voxel_size;                   /* :341-351 */
    float next_by = gmin[1] + nx_y * voxel_size;
    float next_bz = gmin[2] + nx_z * voxel_size;

    float t_max_x = o_safe_div(next_bx - cam[0], dir[0]);          /* :353-355 */
    float t_max_y = o_safe_div(next_by - cam[1], dir[1]);
    float t_max_z = o_safe_div(next_bz - cam[2], dir[2]);
    float t_delta_x = o_safe_div(voxel_size, fabsf(dir[0]));       /* :357-359 */
    float t_delta_y = o_safe_div(voxel_size, fabsf(dir[1]));
    float t_delta_z = o_safe_div(voxel_size, fabsf(dir[2]));

    float t_current = t_min;                                       /* :361 */
    int64_t step_count = 0;

    while (t_current <= t_max) {                                   /* :365-392 */
        emit(ctx, ix, iy, iz, t_current);
        if (t_max_x < t_max_y && t_max_x < t_max_z) {
            ix += step_x; t_current = t_max_x; t_max_x += t_delta_x;
        } else if (t_max_y < t_max_z) {
            iy += step_y; t_current = t_max_y; t_max_y += t_delta_y;
        } else {
            iz += step_z; t_current = t_max_z; t_max_z += t_delta_z;
        }
        step_count++;
        if (ix < 0 || ix >= N || iy < 0 || iy >= N || iz < 0 || iz >= N) break;
    }
    return step_count;
}

struct o_list_ctx { int32_t *vox; float *t; int64_t cap; int64_t n; };

static void o_emit_list(void *c, int ix, int iy, int iz, float t)
{
    struct o_list_ctx *l = (struct o_list_ctx *)c;
    if (l->n < l->cap) {
        if (l->vox) { l->vox[3 * l->n + 0] = ix; l->vox[3 * l->n + 1] = iy; l->vox[3 * l->n + 2] = iz; }
        if (l->t) l->t[l->n] = t;
    }
    l->n++;
}

/* Visited-voxel list of one ray (ray_voxel.cpp:274-395).  Writes at most `cap`
 * steps into vox[3*cap] / t[cap] (either may be NULL); returns the true count. */
PVPO_API int64_t pvpo_cast_ray(const float cam[3], const float dir[3], int N, float voxel_size,
                               const float center[3], int32_t *vox, float *t, int64_t cap)
{
    struct o_list_ctx l = { vox, t, cap, 0 };
    return o_cast_ray(cam, dir, N, voxel_size, center, o_emit_list, &l);
}

struct o_acc_ctx {
    float *grid_f32; int64_t *grid_i64; int N; int slab_lo, slab_hi;
    float val; int64_t qval; int64_t written;
};

/* ray_voxel.cpp:523-529: idx = ix*N*N + iy*N + iz; grid[idx] += val (attenuation is computed
 * and discarded by the reference, :525-526, so it is not restated).  slab_[lo,hi) restricts
 * writes to x-planes owned by this grid shard; the grid pointer addresses plane slab_lo. */
static void o_emit_acc(void *c, int ix, int iy, int iz, float t)
{
    (void)t;
    struct o_acc_ctx *a = (struct o_acc_ctx *)c;
    if (ix < a->slab_lo || ix >= a->slab_hi) return;
    int64_t idx = ((int64_t)(ix - a->slab_lo) * a->N + iy) * a->N + iz;
    if (a->grid_f32) a->grid_f32[idx] += a->val;
    else a->grid_i64[idx] += a->qval;
    a->written++;
}

/* SURVEY 8f N4 (opt-in extension, NOT reference behaviour): the distance attenuation the reference computes at
 * ray_voxel.cpp:524-525 and then discards (:526 multiplies by 1.f instead).  Here it is applied:
 *   float dist = rs.distance; float attenuation = 1.f/(1.f + alpha*dist); float val = pix_val * attenuation;
 * Parity for this mode is UNPINNED: no build of the unmodified reference executes it. */
struct o_att_ctx {
    struct o_acc_ctx acc;
    float pix_val, alpha;
    int accum_mode, frac_bits;
};
static inline int64_t o_quantize(float val, int frac_bits);
static void o_emit_acc_att(void *c, int ix, int iy, int iz, float t)
{
    struct o_att_ctx *a = (struct o_att_ctx *)c;
    float dist = t;                                         /* :524 */
    float attenuation = 1.f / (1.f + a->alpha * dist);      /* :525 */
    a->acc.val = a->pix_val * attenuation;                  /* :526 with the factor applied */
    if (a->accum_mode != 0) a->acc.qval = o_quantize(a->acc.val, a->frac_bits);
    o_emit_acc(&a->acc, ix, iy, iz, t);
}

/* Fixed-point quantisation shared by oracle and kernels: q = rint(val * 2^frac_bits),
 * round-to-nearest-even; the scaling is exact in fp32 (power of two). */
static inline int64_t o_quantize(float val, int frac_bits)
{
    return llrintf(val * (float)((int64_t)1 << frac_bits));
}

/*
 * The loop body of main(), ray_voxel.cpp:484-531, for ONE frame pair whose two images are
 * already float arrays (the parity boundary sits after load_image_gray, whose noise is
 * unseeded -- SURVEY.md fact 2).  Pose = CURRENT frame (:486-491), passed as cam/R/focal.
 * accum_mode 0: grid is float32 (+=, reference order); 1: grid is int64 fixed point.
 * Returns 0; counts rays cast and grid writes.
 */
PVPO_API int pvpo_accumulate_pair(const float *prev, const float *curr, int width, int height,
                                  const float cam[3], const float R[9], float focal_len,
                                  int N, float voxel_size, const float center[3], float threshold,
                                  void *grid, int accum_mode, int frac_bits, int slab_lo, int slab_hi,
                                  uint64_t *n_rays, uint64_t *n_steps, uint64_t *n_written)
{
    uint64_t rays = 0, steps = 0, written = 0;
    for (int v = 0; v < height; ++v) {
        for (int u = 0; u < width; ++u) {
            float d = fabsf(prev[(int64_t)v * width + u] - curr[(int64_t)v * width + u]);   /* :250 */
            if (!(d > threshold)) continue;                        /* :252, :496 */
            float pix_val = d;                                     /* :500 */
            if (pix_val < 1e-3f) continue;                         /* :501 */
            float dir[3];
            pvpo_pixel_ray(u, v, width, height, focal_len, R, dir);   /* :506-515 */
            struct o_acc_ctx a;
            a.grid_f32 = accum_mode == 0 ? (float *)grid : NULL;
            a.grid_i64 = accum_mode == 0 ? NULL : (int64_t *)grid;
            a.N = N; a.slab_lo = slab_lo; a.slab_hi = slab_hi;
            a.val = pix_val * 1.f;                                 /* :526 */
            a.qval = accum_mode == 0 ? 0 : o_quantize(a.val, frac_bits);
            a.written = 0;
            steps += (uint64_t)o_cast_ray(cam, dir, N, voxel_size, center, o_emit_acc, &a);
            written += (uint64_t)a.written;
            rays++;
        }
    }
    if (n_rays) *n_rays += rays;
    if (n_steps) *n_steps += steps;
    if (n_written) *n_written += written;
    return 0;
}

/* pvpo_accumulate_pair with the opt-in distance attenuation (see o_emit_acc_att); alpha == 0 is NOT routed here. */
PVPO_API int pvpo_accumulate_pair_att(const float *prev, const float *curr, int width, int height,
                                      const float cam[3], const float R[9], float focal_len,
                                      int N, float voxel_size, const float center[3], float threshold, float alpha,
                                      void *grid, int accum_mode, int frac_bits, int slab_lo, int slab_hi,
                                      uint64_t *n_rays, uint64_t *n_steps, uint64_t *n_written)
{
    uint64_t rays = 0, steps = 0, written = 0;
    for (int v = 0; v < height; ++v) {
        for (int u = 0; u < width; ++u) {
            float d = fabsf(prev[(int64_t)v * width + u] - curr[(int64_t)v * width + u]);   /* :250 */
            if (!(d > threshold)) continue;                        /* :252, :496 */
            float pix_val = d;                                     /* :500 */
            if (pix_val < 1e-3f) continue;                         /* :501 */
            float dir[3];
            pvpo_pixel_ray(u, v, width, height, focal_len, R, dir);   /* :506-515 */
            struct o_att_ctx a;
            a.acc.grid_f32 = accum_mode == 0 ? (float *)grid : NULL;
            a.acc.grid_i64 = accum_mode == 0 ? NULL : (int64_t *)grid;
            a.acc.N = N; a.acc.slab_lo = slab_lo; a.acc.slab_hi = slab_hi;
            a.acc.val = 0.f; a.acc.qval = 0; a.acc.written = 0;
            a.pix_val = pix_val; a.alpha = alpha; a.accum_mode = accum_mode; a.frac_bits = frac_bits;
            steps += (uint64_t)o_cast_ray(cam, dir, N, voxel_size, center, o_emit_acc_att, &a);
            written += (uint64_t)a.acc.written;
            rays++;
        }
    }
    if (n_rays) *n_rays += rays;
    if (n_steps) *n_steps += steps;
    if (n_written) *n_written += written;
    return 0;
}

/* Same with 8-bit frames: the exact (noise-free) uint8 -> float conversion of
 * ray_voxel.cpp:213 without the unseeded noise of :215. */
PVPO_API int pvpo_accumulate_pair_u8(const uint8_t *prev, const uint8_t *curr, int width, int height,
                                     const float cam[3], const float R[9], float focal_len,
                                     int N, float voxel_size, const float center[3], float threshold,
                                     void *grid, int accum_mode, int frac_bits, int slab_lo, int slab_hi,
                                     uint64_t *n_rays, uint64_t *n_steps, uint64_t *n_written)
{
    int64_t n = (int64_t)width * height;
    float *p = (float *)malloc(sizeof(float) * (size_t)n), *c = (float *)malloc(sizeof(float) * (size_t)n);
    if (!p || !c) { free(p); free(c); return -1; }
    for (int64_t i = 0; i < n; ++i) { p[i] = (float)prev[i]; c[i] = (float)curr[i]; }
    int rc = pvpo_accumulate_pair(p, c, width, height, cam, R, focal_len, N, voxel_size, center, threshold,
                                  grid, accum_mode, frac_bits, slab_lo, slab_hi, n_rays, n_steps, n_written);
    free(p); free(c);
    return rc;
}

/* int64 fixed-point grid -> float32 grid (the .bin payload type, ray_voxel.cpp:552). */
PVPO_API void pvpo_fixed_to_f32(const int64_t *src, float *dst, int64_t n, int frac_bits)
{
    double inv = 1.0 / (double)((int64_t)1 << frac_bits);
    for (int64_t i = 0; i < n; ++i) dst[i] = (float)((double)src[i] * inv);
}

/* ------------------------------------------------------------------ */
/* Path B: process_image.cpp                                           */
/* ------------------------------------------------------------------ */

/* process_image.cpp:24-61 */
static inline int o_ray_aabb(const double o[3], const double d[3], const double bmin[3], const double bmax[3],
                             double *t_entry, double *t_exit)
{
    double tmin = -INFINITY, tmax = INFINITY;
    for (int i = 0; i < 3; ++i) {
        if (fabs(d[i]) > 1e-8) {
            double t1 = (bmin[i] - o[i]) / d[i];
            double t2 = (bmax[i] - o[i]) / d[i];
            if (t1 > t2) { double s = t1; t1 = t2; t2 = s; }
            tmin = (tmin < t1) ? t1 : tmin;      /* std::max(tmin, t1) */
            tmax = (t2 < tmax) ? t2 : tmax;      /* std::min(tmax, t2) */
            if (tmin > tmax) return 0;
        } else {
            if (o[i] < bmin[i] || o[i] > bmax[i]) return 0;
        }
    }
    *t_entry = tmin; *t_exit = tmax;
    return 1;
}

/* process_image.cpp:68-113; returns 0 when the point is outside (reference returns -1,-1,-1) */
static inline int o_point_to_voxel(const double p[3], const double ext[6], const int64_t n[3], int64_t idx[3])
{
    for (int a = 0; a < 3; ++a)
        if (!(ext[2 * a] <= p[a] && p[a] <= ext[2 * a + 1])) return 0;
    for (int a = 0; a < 3; ++a) {
        double norm = (p[a] - ext[2 * a]) / (ext[2 * a + 1] - ext[2 * a]);
        int64_t i = (int64_t)(norm * n[a]);
        if (i < 0) i = 0;
        if (i > n[a] - 1) i = n[a] - 1;
        idx[a] = i;
    }
    return 1;
}

PVPO_API int64_t pvpo_quantize_f64(double val, int frac_bits)
{
    return llrint(val * (double)((int64_t)1 << frac_bits));
}

/*
 * process_image.cpp:125-366.  All arrays are addressed through element strides, like
 * the reference's numpy unchecked<> accessors.  grid may be NULL / gshape product 0
 * (voxel stage skipped, :152).  accum_mode 0: fp64 += (reference); 1: int64 fixed point
 * on both grid and texture (texture reads for background subtraction are converted back).
 * Pixels are visited in the reference's serial order when num_threads == 1.
 */
PVPO_API int pvpo_process_image(const double *image, int64_t img_s0, int64_t img_s1,
                                const double earth_position[3], const double pointing_in[3], double fov,
                                int64_t image_width, int64_t image_height,
                                void *grid, const int64_t gshape[3], const int64_t gstride[3],
                                const double extent[6], double max_distance, int num_steps,
                                void *tex, int64_t tex_h, int64_t tex_w, int64_t tex_s0, int64_t tex_s1,
                                double center_ra_rad, double center_dec_rad,
                                double angular_width_rad, double angular_height_rad,
                                int update_celestial_sphere, int perform_background_subtraction,
                                int accum_mode, int frac_bits, int num_threads,
                                uint64_t *n_rays, uint64_t *n_samples)
{
    int grid_provided = grid != NULL && gshape[0] * gshape[1] * gshape[2] > 0;      /* :152 */
    double focal_length = (image_width / 2.0) / tan(fov / 2.0);                      /* :186 */
    double cx = image_width / 2.0, cy = image_height / 2.0;                          /* :189-190 */

    double pd[3] = { pointing_in[0], pointing_in[1], pointing_in[2] };
    double z_norm = sqrt(pd[0] * pd[0] + pd[1] * pd[1] + pd[2] * pd[2]);             /* :194-199 */
    pd[0] /= z_norm; pd[1] /= z_norm; pd[2] /= z_norm;

    double up[3] = { 0.0, 0.0, 1.0 };                                                /* :202-211 */
    if ((fabs(pd[0] - up[0]) < 1e-8 && fabs(pd[1] - up[1]) < 1e-8 && fabs(pd[2] - up[2]) < 1e-8) ||
        (fabs(pd[0] + up[0]) < 1e-8 && fabs(pd[1] + up[1]) < 1e-8 && fabs(pd[2] + up[2]) < 1e-8)) {
        up[0] = 0.0; up[1] = 1.0; up[2] = 0.0;
    }
    double za[3] = { pd[0], pd[1], pd[2] }, xa[3], ya[3];                            /* :214-228 */
    xa[0] = up[1] * za[2] - up[2] * za[1];
    xa[1] = up[2] * za[0] - up[0] * za[2];
    xa[2] = up[0] * za[1] - up[1] * za[0];
    double x_norm = sqrt(xa[0] * xa[0] + xa[1] * xa[1] + xa[2] * xa[2]);
    xa[0] /= x_norm; xa[1] /= x_norm; xa[2] /= x_norm;
    ya[0] = za[1] * xa[2] - za[2] * xa[1];
    ya[1] = za[2] * xa[0] - za[0] * xa[2];
    ya[2] = za[0] * xa[1] - za[1] * xa[0];

    const double scale = (double)((int64_t)1 << frac_bits);
    uint64_t rays = 0, samples = 0;
    (void)num_threads;

#ifdef _OPENMP
#pragma omp parallel for num_threads(num_threads > 0 ? num_threads : 1) reduction(+ : rays, samples)
#endif
    for (int64_t i = 0; i < image_height; ++i) {                                     /* :231-232 */
        for (int64_t j = 0; j < image_width; ++j) {
            double brightness = image[i * img_s0 + j * img_s1];                      /* :236 */
            if (!(brightness > 0)) continue;                                         /* :238 */

            double x_cam = (j - cx), y_cam = (i - cy), z_cam = focal_length;         /* :241-243 */
            double norm = sqrt(x_cam * x_cam + y_cam * y_cam + z_cam * z_cam);
            double dc[3] = { x_cam / norm, y_cam / norm, z_cam / norm };
            double dw[3];
            dw[0] = xa[0] * dc[0] + ya[0] * dc[1] + za[0] * dc[2];                   /* :250-252 */
            dw[1] = xa[1] * dc[0] + ya[1] * dc[1] + za[1] * dc[2];
            dw[2] = xa[2] * dc[0] + ya[2] * dc[1] + za[2] * dc[2];
            double dir_norm = sqrt(dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2]);   /* :255-260 */
            dw[0] /= dir_norm; dw[1] /= dir_norm; dw[2] /= dir_norm;

            double dx = dw[0], dy = dw[1], dz = dw[2];
            double r = sqrt(dx * dx + dy * dy + dz * dz);                            /* :267 */
            double dec = asin(dz / r);
            double ra = atan2(dy, dx);
            if (ra < 0) ra += 2 * M_PI;                                              /* :270 */
            double ra_offset = ra - center_ra_rad, dec_offset = dec - center_dec_rad;
            if (ra_offset > M_PI) ra_offset -= 2 * M_PI;                             /* :277-278 */
            if (ra_offset < -M_PI) ra_offset += 2 * M_PI;
            int within = (fabs(ra_offset) <= angular_width_rad / 2) &&
                         (fabs(dec_offset) <= angular_height_rad / 2);               /* :281-282 */
            double u = (ra_offset + angular_width_rad / 2) / angular_width_rad * tex_w;     /* :285-286 */
            double v = (dec_offset + angular_height_rad / 2) / angular_height_rad * tex_h;
            int64_t u_idx = (int64_t)u, v_idx = (int64_t)v;                          /* :288-293 */
            if (u_idx < 0) u_idx = 0;
            if (u_idx > tex_w - 1) u_idx = tex_w - 1;
            if (v_idx < 0) v_idx = 0;
            if (v_idx > tex_h - 1) v_idx = tex_h - 1;

            double background = 0.0;                                                 /* :296-300 */
            if (within) {
                if (accum_mode == 0) background = ((double *)tex)[v_idx * tex_s0 + u_idx * tex_s1];
                else background = (double)((int64_t *)tex)[v_idx * tex_s0 + u_idx * tex_s1] / scale;
            }
            if (perform_background_subtraction) {                                    /* :302-307 */
                brightness -= background;
                if (brightness <= 0) continue;
            }
            int64_t qb = accum_mode ? pvpo_quantize_f64(brightness, frac_bits) : 0;
            if (update_celestial_sphere && within) {                                 /* :310-314 */
                if (accum_mode == 0) {
#ifdef _OPENMP
#pragma omp atomic
#endif
                    ((double *)tex)[v_idx * tex_s0 + u_idx * tex_s1] += brightness;
                } else {
#ifdef _OPENMP
#pragma omp atomic
#endif
                    ((int64_t *)tex)[v_idx * tex_s0 + u_idx * tex_s1] += qb;
                }
            }
            rays++;
            if (!grid_provided) continue;                                            /* :317 */

            double step_size = max_distance / num_steps;                             /* :323 */
            double t_entry, t_exit;
            if (!o_ray_aabb(earth_position, dw, (const double[3]){ extent[0], extent[2], extent[4] },
                            (const double[3]){ extent[1], extent[3], extent[5] }, &t_entry, &t_exit))
                continue;                                                            /* :332 */
            t_entry = (t_entry < 0.0) ? 0.0 : t_entry;                               /* :334 std::max */
            t_exit = (max_distance < t_exit) ? max_distance : t_exit;                /* :335 std::min */
            if (!(t_entry <= t_exit)) continue;                                      /* :337 */
            int s_entry = (int)(t_entry / step_size);                                /* :339-340 */
            int s_exit = (int)(t_exit / step_size);
            for (int s = s_entry; s <= s_exit; ++s) {                                /* :342 */
                double d = s * step_size;                                            /* :344-347 */
                double p[3] = { earth_position[0] + d * dw[0], earth_position[1] + d * dw[1],
                                earth_position[2] + d * dw[2] };
                int64_t idx[3];
                if (!o_point_to_voxel(p, extent, gshape, idx)) continue;             /* :349-354 */
                int64_t off = idx[0] * gstride[0] + idx[1] * gstride[1] + idx[2] * gstride[2];
                if (accum_mode == 0) {
#ifdef _OPENMP
#pragma omp atomic
#endif
                    ((double *)grid)[off] += brightness;                             /* :356-357 */
                } else {
#ifdef _OPENMP
#pragma omp atomic
#endif
                    ((int64_t *)grid)[off] += qb;
                }
                samples++;
            }
        }
    }
    if (n_rays) *n_rays += rays;
    if (n_samples) *n_samples += samples;
    return 0;
}

/* Number of worker threads the OpenMP runtime would use (reported as cpu_baseline.cores). */
PVPO_API int pvpo_max_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
