// pvp_device.cuh -- device-side building blocks shared by the path-A kernels.
//
// Bit-exactness contract (BASELINE.json north_star; SURVEY.md section 7 "hard parts"):
// every fp32 operation of the reference's ray set-up and DDA (ray_voxel.cpp:64-79,
// :266-395, :506-515) is issued as an IEEE round-to-nearest, un-fused operation in the
// reference's order.  All arithmetic goes through __f*_rn intrinsics, which ptxas never
// contracts into FMA, so parity does not depend on -fmad=false (the TU is compiled with it
// anyway).  No fast-math.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/pvp_b200.h"

namespace pvp {

// Grid constants derived once on the host with the reference's fp32 expressions
// (ray_voxel.cpp:284-290) -- x86-64 host code has no FMA contraction either.
struct GridP {
    int n;
    int slab_lo, slab_hi;
    float vs;
    float gmin[3];
    float gmax[3];
    float fixed_scale;  // 2^frac_bits
    int frac_shift;     // frac_bits (0 for float32 grids)
    float alpha;        // opt-in distance attenuation (0 = off, the reference's behaviour)
    unsigned mul_n, mul_nn;  // floor(2^32 / n), floor(2^32 / n^2): division of a voxel offset by a multiply (udiv_by)
    int one;            // 1, but not known to the compiler: the z stride of rays_step_mad (a multiply by a literal 1 would be turned back into an add)
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// ray_voxel.cpp:64-70
__device__ __forceinline__ void normalize3(float &x, float &y, float &z)
{
    float len = __fsqrt_rn(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z)));
    if (len < 1e-12f) {
        x = 0.f; y = 0.f; z = 0.f;
        return;
    }
    x = fdiv(x, len); y = fdiv(y, len); z = fdiv(z, len);
}

// ray_voxel.cpp:506-515: pixel (u,v) -> unit world direction.  R row-major (ray_voxel.cpp:73-79).
__device__ __forceinline__ void pixel_ray(int u, int v, int width, int height, float focal, const float *__restrict__ R,
                                          float &dx, float &dy, float &dz)
{
    float x = fsub((float)u, fmul(0.5f, (float)width));
    float y = -fsub((float)v, fmul(0.5f, (float)height));
    float z = -focal;
    normalize3(x, y, z);
    float wx = fadd(fadd(fmul(R[0], x), fmul(R[1], y)), fmul(R[2], z));
    float wy = fadd(fadd(fmul(R[3], x), fmul(R[4], y)), fmul(R[5], z));
    float wz = fadd(fadd(fmul(R[6], x), fmul(R[7], y)), fmul(R[8], z));
    normalize3(wx, wy, wz);
    dx = wx; dy = wy; dz = wz;
}

// ray_voxel.cpp:266-272
__device__ __forceinline__ float safe_div(float num, float den)
{
    return (fabsf(den) < 1e-12f) ? CUDART_INF_F : fdiv(num, den);
}

// ray_voxel.cpp:329-331: int(float) as x86 cvttss2si does it -- NaN and out-of-range map to
// INT_MIN (CUDA's cvt.rzi saturates and sends NaN to 0), so such starts are rejected by :332.
__device__ __forceinline__ int trunc_like_x86(float f)
{
    if (!(f > -2147483904.f && f < 2147483648.f)) return INT32_MIN;
    return (int)f;
}

// State of one ray between set-up and the walk (ray_voxel.cpp:292-362).
struct Ray {
    float t_cur, t_end;          // t_current (:361) and the box exit t_max (:293)
    float tmx, tmy, tmz;         // t_max_{x,y,z} (:353-355)
    float tdx, tdy, tdz;         // t_delta_{x,y,z} (:357-359)
    int ix, iy, iz;              // current voxel (:329-331)
    int sx, sy, sz;              // step_{x,y,z} (:337-339)
};

// ray_voxel.cpp:296-362.  Returns false when the reference returns an empty step list.
__device__ __forceinline__ bool ray_setup(float ox, float oy, float oz, float dx, float dy, float dz, const GridP &g, Ray &r)
{
    float t_min = 0.f, t_max = CUDART_INF_F;
#define PVP_SLAB(o, d, mn, mx)                                  \
    if (fabsf(d) < 1e-12f) {                                    \
        if (o < mn || o > mx) return false;                     \
    } else {                                                    \
        float t1 = fdiv(fsub(mn, o), d);                        \
        float t2 = fdiv(fsub(mx, o), d);                        \
        float t_near = fminf(t1, t2);                           \
        float t_far = fmaxf(t1, t2);                            \
        if (t_near > t_min) t_min = t_near;                     \
        if (t_far < t_max) t_max = t_far;                       \
        if (t_min > t_max) return false;                        \
    }
    PVP_SLAB(ox, dx, g.gmin[0], g.gmax[0])
    PVP_SLAB(oy, dy, g.gmin[1], g.gmax[1])
    PVP_SLAB(oz, dz, g.gmin[2], g.gmax[2])
#undef PVP_SLAB
    if (t_min < 0.f) t_min = 0.f;  // :319

    float sxw = fadd(ox, fmul(t_min, dx));  // :322-324
    float syw = fadd(oy, fmul(t_min, dy));
    float szw = fadd(oz, fmul(t_min, dz));
    int ix = trunc_like_x86(fdiv(fsub(sxw, g.gmin[0]), g.vs));  // :325-331
    int iy = trunc_like_x86(fdiv(fsub(syw, g.gmin[1]), g.vs));
    int iz = trunc_like_x86(fdiv(fsub(szw, g.gmin[2]), g.vs));
    const int N = g.n;
    if (ix < 0 || ix >= N || iy < 0 || iy >= N || iz < 0 || iz >= N) return false;  // :332-334

    r.sx = (dx >= 0.f) ? 1 : -1;  // :337-339 (-0.0f >= 0 is true, as in the reference)
    r.sy = (dy >= 0.f) ? 1 : -1;
    r.sz = (dz >= 0.f) ? 1 : -1;
    float nbx = fadd(g.gmin[0], fmul((float)(ix + (r.sx > 0 ? 1 : 0)), g.vs));  // :341-351
    float nby = fadd(g.gmin[1], fmul((float)(iy + (r.sy > 0 ? 1 : 0)), g.vs));
    float nbz = fadd(g.gmin[2], fmul((float)(iz + (r.sz > 0 ? 1 : 0)), g.vs));
    r.tmx = safe_div(fsub(nbx, ox), dx);  // :353-355 -- measured from the camera, not the entry point
    r.tmy = safe_div(fsub(nby, oy), dy);
    r.tmz = safe_div(fsub(nbz, oz), dz);
    r.tdx = safe_div(g.vs, fabsf(dx));  // :357-359
    r.tdy = safe_div(g.vs, fabsf(dy));
    r.tdz = safe_div(g.vs, fabsf(dz));
    r.ix = ix; r.iy = iy; r.iz = iz;
    r.t_cur = t_min;  // :361
    r.t_end = t_max;
    return true;
}

// One iteration of the walk, ray_voxel.cpp:375-391, AFTER the current voxel has been emitted.
// Returns false when the ray left the grid (:389-391).  Tie-break: x only if strictly smaller
// than both, else y if strictly smaller than z, else z.
__device__ __forceinline__ bool ray_advance(Ray &r, int N)
{
    if (r.tmx < r.tmy && r.tmx < r.tmz) {
        r.ix += r.sx; r.t_cur = r.tmx; r.tmx = fadd(r.tmx, r.tdx);
        return (unsigned)r.ix < (unsigned)N;
    } else if (r.tmy < r.tmz) {
        r.iy += r.sy; r.t_cur = r.tmy; r.tmy = fadd(r.tmy, r.tdy);
        return (unsigned)r.iy < (unsigned)N;
    } else {
        r.iz += r.sz; r.t_cur = r.tmz; r.tmz = fadd(r.tmz, r.tdz);
        return (unsigned)r.iz < (unsigned)N;
    }
}

// |prev - next| of ray_voxel.cpp:250 for both frame element types.
__device__ __forceinline__ float motion_diff(uint8_t a, uint8_t b) { return fabsf(fsub((float)a, (float)b)); }
__device__ __forceinline__ float motion_diff(float a, float b) { return fabsf(fsub(a, b)); }

// The two filters between a pixel and a ray: changed (:252, :496) and pix_val >= 1e-3 (:501).
__device__ __forceinline__ bool motion_casts_ray(float d, float threshold) { return (d > threshold) && !(d < 1e-3f); }

// The factor the reference computes and discards (ray_voxel.cpp:524-526), applied only when alpha > 0.
__device__ __forceinline__ float attenuate(float pix_val, float alpha, float dist)
{
    return fmul(pix_val, fdiv(1.f, fadd(1.f, fmul(alpha, dist))));
}

// Fixed-point quantisation: rint(val * 2^frac_bits), nearest-even.
__device__ __forceinline__ long long quantize_fixed(float val, float scale) { return __float2ll_rn(fmul(val, scale)); }

}  // namespace pvp
