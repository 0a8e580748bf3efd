// pvp_image.cu -- path B of libpvp_b200.so: process_image_cpp (process_image.cpp:125-366)
// as an sm_100a kernel.
//
//   K3  process_image_kernel   one ray per bright pixel: camera ray in fp64 (:241-260), RA/Dec
//                              and sky-texture lookup / accumulate (:263-314), AABB clip (:24-61)
//                              and the fixed-step march s = s_entry..s_exit (:339-359) with
//                              red.global.add.f64 (or int64 fixed point) into the voxel grid.
//
// All fp64 arithmetic that decides a voxel index is issued through __d*_rn intrinsics in the
// reference's operation order (no FMA contraction), so visited voxels are bit-reproducible.
// asin / atan2 come from CUDA's libdevice (<= 2 ulp from glibc): a texture index can differ only
// when u or v lands within an ulp of an integer.  The camera basis and focal
// length are computed on the host with glibc, exactly as the reference does.
#include <math.h>

#include <algorithm>
#include <cmath>

#include <math_constants.h>

#include "pvp_internal.h"

namespace pvp {

struct ImageP {
    double o[3];           // earth_position
    double xa[3], ya[3], za[3];  // camera basis (:214-228)
    double focal, cx, cy;  // :186-190
    long long W, H, is0, is1;
    long long gn[3], gs[3];
    double ext[6];
    double max_distance, step_size;
    long long th, tw, ts0, ts1;
    double c_ra, c_dec, half_w, half_h, ang_w, ang_h;
    int update_tex, bg_sub, grid_provided;
    double scale;          // 2^frac_bits
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// static_cast<ssize_t>(double) as x86-64 cvttsd2si: NaN / out of range -> INT64_MIN.
__device__ __forceinline__ long long trunc64_like_x86(double f)
{
    if (!(f >= -9223372036854775808.0 && f < 9223372036854775808.0)) return INT64_MIN;
    return (long long)f;
}
__device__ __forceinline__ int trunc32_like_x86(double f)
{
    if (!(f > -2147483649.0 && f < 2147483648.0)) return INT32_MIN;
    return (int)f;
}

template <int MODE> struct Cell;
template <> struct Cell<PVP_ACCUM_F64> {
    using T = double;
    using V = double;
    __device__ static __forceinline__ V make(double b, double) { return b; }
    __device__ static __forceinline__ void add(T *p, V v) { atomicAdd(p, v); }
    __device__ static __forceinline__ double read(const T *p, double) { return *p; }
};
template <> struct Cell<PVP_ACCUM_FIXED> {
    using T = unsigned long long;
    using V = unsigned long long;
    __device__ static __forceinline__ V make(double b, double scale) { return (unsigned long long)__double2ll_rn(dmul(b, scale)); }
    __device__ static __forceinline__ void add(T *p, V v) { atomicAdd(p, v); }
    __device__ static __forceinline__ double read(const T *p, double scale) { return ddiv((double)(long long)*p, scale); }
};

constexpr int K3_THREADS = 128;

template <int MODE>
__global__ void __launch_bounds__(K3_THREADS)
process_image_kernel(const double *__restrict__ image, typename Cell<MODE>::T *__restrict__ grid,
                     typename Cell<MODE>::T *__restrict__ tex, ImageP p, Ctl *__restrict__ ctl)
{
    using C = Cell<MODE>;
    unsigned long long my_rays = 0, my_samples = 0;
    const long long n_px = p.W * p.H;
    for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < n_px; px += (long long)gridDim.x * blockDim.x) {
        const long long i = px / p.W, j = px - i * p.W;
        double brightness = image[i * p.is0 + j * p.is1];  // :236
        if (!(brightness > 0)) continue;                   // :238

        const double x_cam = dsub((double)j, p.cx), y_cam = dsub((double)i, p.cy), z_cam = p.focal;  // :241-243
        const double norm = __dsqrt_rn(dadd(dadd(dmul(x_cam, x_cam), dmul(y_cam, y_cam)), dmul(z_cam, z_cam)));
        const double c0 = ddiv(x_cam, norm), c1 = ddiv(y_cam, norm), c2 = ddiv(z_cam, norm);
        double w0 = dadd(dadd(dmul(p.xa[0], c0), dmul(p.ya[0], c1)), dmul(p.za[0], c2));  // :250-252
        double w1 = dadd(dadd(dmul(p.xa[1], c0), dmul(p.ya[1], c1)), dmul(p.za[1], c2));
        double w2 = dadd(dadd(dmul(p.xa[2], c0), dmul(p.ya[2], c1)), dmul(p.za[2], c2));
        const double dn = __dsqrt_rn(dadd(dadd(dmul(w0, w0), dmul(w1, w1)), dmul(w2, w2)));  // :255-260
        w0 = ddiv(w0, dn); w1 = ddiv(w1, dn); w2 = ddiv(w2, dn);

        const double r = __dsqrt_rn(dadd(dadd(dmul(w0, w0), dmul(w1, w1)), dmul(w2, w2)));   // :267
        const double dec = asin(ddiv(w2, r));
        double ra = atan2(w1, w0);
        if (ra < 0) ra = dadd(ra, 2 * M_PI);  // :270
        double ra_off = dsub(ra, p.c_ra), dec_off = dsub(dec, p.c_dec);
        if (ra_off > M_PI) ra_off = dsub(ra_off, 2 * M_PI);  // :277-278
        if (ra_off < -M_PI) ra_off = dadd(ra_off, 2 * M_PI);
        const bool within = (fabs(ra_off) <= p.half_w) && (fabs(dec_off) <= p.half_h);  // :281-282
        const double u = dmul(ddiv(dadd(ra_off, p.half_w), p.ang_w), (double)p.tw);      // :285-286
        const double v = dmul(ddiv(dadd(dec_off, p.half_h), p.ang_h), (double)p.th);
        long long u_idx = trunc64_like_x86(u), v_idx = trunc64_like_x86(v);              // :288-293
        u_idx = min(max(u_idx, 0ll), p.tw - 1);
        v_idx = min(max(v_idx, 0ll), p.th - 1);
        typename C::T *texel = tex + (v_idx * p.ts0 + u_idx * p.ts1);

        if (p.bg_sub) {  // :296-307 (the reference also loads the texel when it is not used)
            const double background = within ? C::read(texel, p.scale) : 0.0;
            brightness = dsub(brightness, background);
            if (brightness <= 0) continue;
        }
        const typename C::V val = C::make(brightness, p.scale);
        if (p.update_tex && within) C::add(texel, val);  // :310-314
        ++my_rays;
        if (!p.grid_provided) continue;  // :317

        // ray_aabb_intersection, :24-61
        double tmin = -CUDART_INF, tmax = CUDART_INF;
        bool hit = true;
        const double dir[3] = {w0, w1, w2};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (fabs(dir[a]) > 1e-8) {
                double t1 = ddiv(dsub(p.ext[2 * a], p.o[a]), dir[a]);
                double t2 = ddiv(dsub(p.ext[2 * a + 1], p.o[a]), dir[a]);
                if (t1 > t2) { double s = t1; t1 = t2; t2 = s; }
                tmin = (tmin < t1) ? t1 : tmin;  // std::max(tmin, t1)
                tmax = (t2 < tmax) ? t2 : tmax;  // std::min(tmax, t2)
                if (tmin > tmax) hit = false;
            } else if (p.o[a] < p.ext[2 * a] || p.o[a] > p.ext[2 * a + 1]) {
                hit = false;
            }
        }
        if (!hit) continue;
        const double t_entry = (tmin < 0.0) ? 0.0 : tmin;                      // :334
        const double t_exit = (p.max_distance < tmax) ? p.max_distance : tmax;  // :335
        if (!(t_entry <= t_exit)) continue;                                     // :337
        const int s_entry = trunc32_like_x86(ddiv(t_entry, p.step_size));       // :339-340
        const int s_exit = trunc32_like_x86(ddiv(t_exit, p.step_size));

        const double ex0 = dsub(p.ext[1], p.ext[0]), ex1 = dsub(p.ext[3], p.ext[2]), ex2 = dsub(p.ext[5], p.ext[4]);
        for (long long s = s_entry; s <= s_exit; ++s) {  // :342
            const double d = dmul((double)(int)s, p.step_size);  // :344
            const double qx = dadd(p.o[0], dmul(d, w0)), qy = dadd(p.o[1], dmul(d, w1)), qz = dadd(p.o[2], dmul(d, w2));
            // point_to_voxel_indices, :68-113 (inclusive bounds, clamp to n-1)
            if (!(p.ext[0] <= qx && qx <= p.ext[1] && p.ext[2] <= qy && qy <= p.ext[3] && p.ext[4] <= qz && qz <= p.ext[5])) continue;
            long long xi = trunc64_like_x86(dmul(ddiv(dsub(qx, p.ext[0]), ex0), (double)p.gn[0]));
            long long yi = trunc64_like_x86(dmul(ddiv(dsub(qy, p.ext[2]), ex1), (double)p.gn[1]));
            long long zi = trunc64_like_x86(dmul(ddiv(dsub(qz, p.ext[4]), ex2), (double)p.gn[2]));
            xi = min(max(xi, 0ll), p.gn[0] - 1);
            yi = min(max(yi, 0ll), p.gn[1] - 1);
            zi = min(max(zi, 0ll), p.gn[2] - 1);
            C::add(grid + (xi * p.gs[0] + yi * p.gs[1] + zi * p.gs[2]), val);  // :356-357
            ++my_samples;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_rays += __shfl_xor_sync(0xffffffffu, my_rays, o);
        my_samples += __shfl_xor_sync(0xffffffffu, my_samples, o);
    }
    if ((threadIdx.x & 31) == 0 && (my_rays | my_samples)) {
        atomicAdd(&ctl->img_rays, my_rays);
        atomicAdd(&ctl->img_samples, my_samples);
    }
}

// Host part of process_image.cpp:186-228 (focal length, principal point, camera basis) in fp64
// with glibc, every product/sum kept un-fused via volatile temporaries.
static int make_imagep(const pvp_image_args *a, ImageP *p)
{
    PVP_REQUIRE(a != nullptr, "pvp_process_image: args is NULL");
    PVP_REQUIRE(a->image_width > 0 && a->image_height > 0, "pvp_process_image: bad image size");
    PVP_REQUIRE(a->tex_shape[0] > 0 && a->tex_shape[1] > 0, "pvp_process_image: empty texture");
    PVP_REQUIRE(a->accum_mode == PVP_ACCUM_F64 || a->accum_mode == PVP_ACCUM_FIXED, "pvp_process_image: bad accum_mode");
    PVP_REQUIRE(a->accum_mode != PVP_ACCUM_FIXED || (a->frac_bits >= 0 && a->frac_bits <= 52), "pvp_process_image: bad frac_bits");
    PVP_REQUIRE(a->num_steps != 0, "pvp_process_image: num_steps is 0");
    p->focal = (a->image_width / 2.0) / tan(a->fov / 2.0);
    p->cx = a->image_width / 2.0;
    p->cy = a->image_height / 2.0;
    volatile double pd[3] = {a->pointing_direction[0], a->pointing_direction[1], a->pointing_direction[2]};
    volatile double s = pd[0] * pd[0];
    s = s + pd[1] * pd[1];
    s = s + pd[2] * pd[2];
    const double zn = sqrt(s);
    pd[0] = pd[0] / zn; pd[1] = pd[1] / zn; pd[2] = pd[2] / zn;
    double up[3] = {0.0, 0.0, 1.0};
    if ((fabs(pd[0] - up[0]) < 1e-8 && fabs(pd[1] - up[1]) < 1e-8 && fabs(pd[2] - up[2]) < 1e-8) ||
        (fabs(pd[0] + up[0]) < 1e-8 && fabs(pd[1] + up[1]) < 1e-8 && fabs(pd[2] + up[2]) < 1e-8)) {
        up[0] = 0.0; up[1] = 1.0; up[2] = 0.0;
    }
    const double z[3] = {pd[0], pd[1], pd[2]};
    volatile double x[3], y[3], m1, m2;
    m1 = up[1] * z[2]; m2 = up[2] * z[1]; x[0] = m1 - m2;
    m1 = up[2] * z[0]; m2 = up[0] * z[2]; x[1] = m1 - m2;
    m1 = up[0] * z[1]; m2 = up[1] * z[0]; x[2] = m1 - m2;
    s = x[0] * x[0];
    s = s + x[1] * x[1];
    s = s + x[2] * x[2];
    const double xn = sqrt(s);
    x[0] = x[0] / xn; x[1] = x[1] / xn; x[2] = x[2] / xn;
    m1 = z[1] * x[2]; m2 = z[2] * x[1]; y[0] = m1 - m2;
    m1 = z[2] * x[0]; m2 = z[0] * x[2]; y[1] = m1 - m2;
    m1 = z[0] * x[1]; m2 = z[1] * x[0]; y[2] = m1 - m2;
    for (int k = 0; k < 3; ++k) { p->o[k] = a->earth_position[k]; p->xa[k] = x[k]; p->ya[k] = y[k]; p->za[k] = z[k]; }
    p->W = a->image_width; p->H = a->image_height; p->is0 = a->image_stride[0]; p->is1 = a->image_stride[1];
    long long cells = 1;
    for (int k = 0; k < 3; ++k) { p->gn[k] = a->grid_shape[k]; p->gs[k] = a->grid_stride[k]; cells *= a->grid_shape[k]; }
    p->grid_provided = cells > 0;
    for (int k = 0; k < 6; ++k) p->ext[k] = a->extent[k];
    p->max_distance = a->max_distance;
    p->step_size = a->max_distance / a->num_steps;  // :323
    p->th = a->tex_shape[0]; p->tw = a->tex_shape[1]; p->ts0 = a->tex_stride[0]; p->ts1 = a->tex_stride[1];
    p->c_ra = a->center_ra_rad; p->c_dec = a->center_dec_rad;
    p->ang_w = a->angular_width_rad; p->ang_h = a->angular_height_rad;
    p->half_w = a->angular_width_rad / 2; p->half_h = a->angular_height_rad / 2;
    p->update_tex = a->update_celestial_sphere != 0; p->bg_sub = a->perform_background_subtraction != 0;
    p->scale = (double)(1ll << (a->accum_mode == PVP_ACCUM_FIXED ? a->frac_bits : 0));
    return PVP_OK;
}

static int launch_image(pvp_ctx *ctx, const double *image_dev, void *grid_dev, void *tex_dev, const ImageP &p, int accum_mode)
{
    const long long n_px = p.W * p.H;
    int blocks = (int)std::min<long long>((n_px + K3_THREADS - 1) / K3_THREADS, (long long)ctx->sm_count * 16);
    LaunchScope scope(ctx, 3);
    if (accum_mode == PVP_ACCUM_F64)
        process_image_kernel<PVP_ACCUM_F64><<<blocks, K3_THREADS, 0, ctx->stream>>>(image_dev, (double *)grid_dev, (double *)tex_dev, p, ctx->ctl_dev);
    else
        process_image_kernel<PVP_ACCUM_FIXED><<<blocks, K3_THREADS, 0, ctx->stream>>>(image_dev, (unsigned long long *)grid_dev,
                                                                                       (unsigned long long *)tex_dev, p, ctx->ctl_dev);
    PVP_CUDA(cudaGetLastError());
    return PVP_OK;
}

static int fetch_image_stats(pvp_ctx *ctx, pvp_image_stats *stats)
{
    if (!stats) return PVP_OK;
    unsigned long long h[2];
    PVP_CUDA(cudaMemcpyAsync(h, &ctx->ctl_dev->img_rays, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    PVP_CUDA(cudaMemsetAsync(&ctx->ctl_dev->img_rays, 0, sizeof(h), ctx->stream));
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    stats->n_rays += h[0]; stats->n_samples += h[1];
    return PVP_OK;
}

// Highest element offset reachable through (shape, stride) -- sizes the host<->device copies.
static long long span_elems(const int64_t *shape, const int64_t *stride, int nd)
{
    long long hi = 0;
    for (int k = 0; k < nd; ++k) {
        if (shape[k] <= 0) return 0;
        if (stride[k] < 0) return -1;
        hi += (shape[k] - 1) * stride[k];
    }
    return hi + 1;
}

}  // namespace pvp

using namespace pvp;

extern "C" {

int pvp_process_image(pvp_ctx *ctx, const double *image_dev, void *grid_dev, void *tex_dev, const pvp_image_args *args,
                      pvp_image_stats *stats)
{
    PVP_REQUIRE(ctx && image_dev && tex_dev, "pvp_process_image: NULL argument");
    ImageP p;
    int rc = make_imagep(args, &p);
    if (rc) return rc;
    if (!grid_dev) p.grid_provided = 0;
    DeviceGuard guard(ctx->device);
    // the counters describe THIS call: clear whatever earlier calls without `stats` left behind
    if (stats) PVP_CUDA(cudaMemsetAsync(&ctx->ctl_dev->img_rays, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if ((rc = launch_image(ctx, image_dev, grid_dev, tex_dev, p, args->accum_mode))) return rc;
    return fetch_image_stats(ctx, stats);
}

int pvp_process_image_host(pvp_ctx *ctx, const double *image, double *grid, double *tex, const pvp_image_args *args,
                           pvp_image_stats *stats)
{
    PVP_REQUIRE(ctx && image && tex && args, "pvp_process_image_host: NULL argument");
    PVP_REQUIRE(args->accum_mode == PVP_ACCUM_F64, "pvp_process_image_host: host arrays are float64 (accum_mode 0)");
    const int64_t ishape[2] = {args->image_height, args->image_width};
    const long long n_img = span_elems(ishape, args->image_stride, 2);
    const long long n_tex = span_elems(args->tex_shape, args->tex_stride, 2);
    const long long n_grid = grid ? span_elems(args->grid_shape, args->grid_stride, 3) : 0;
    PVP_REQUIRE(n_img > 0 && n_tex > 0 && n_grid >= 0, "pvp_process_image_host: negative strides are not supported");
    DeviceGuard guard(ctx->device);
    double *d_img = nullptr, *d_tex = nullptr, *d_grid = nullptr;
    int rc = PVP_OK;
    auto cleanup = [&]() { cudaFree(d_img); cudaFree(d_tex); cudaFree(d_grid); };
#define PVP_TRY(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #expr, __FILE__, __LINE__); } \
    } while (0)
    PVP_TRY(cudaMalloc(&d_img, sizeof(double) * n_img));
    PVP_TRY(cudaMalloc(&d_tex, sizeof(double) * n_tex));
    if (n_grid) PVP_TRY(cudaMalloc(&d_grid, sizeof(double) * n_grid));
    PVP_TRY(cudaMemcpyAsync(d_img, image, sizeof(double) * n_img, cudaMemcpyHostToDevice, ctx->stream));
    PVP_TRY(cudaMemcpyAsync(d_tex, tex, sizeof(double) * n_tex, cudaMemcpyHostToDevice, ctx->stream));
    if (n_grid) PVP_TRY(cudaMemcpyAsync(d_grid, grid, sizeof(double) * n_grid, cudaMemcpyHostToDevice, ctx->stream));
    if (stats) PVP_TRY(cudaMemsetAsync(&ctx->ctl_dev->img_rays, 0, 2 * sizeof(unsigned long long), ctx->stream));
    rc = pvp_process_image(ctx, d_img, d_grid, d_tex, args, nullptr);
    if (rc) { cleanup(); return rc; }
    PVP_TRY(cudaMemcpyAsync(tex, d_tex, sizeof(double) * n_tex, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_grid) PVP_TRY(cudaMemcpyAsync(grid, d_grid, sizeof(double) * n_grid, cudaMemcpyDeviceToHost, ctx->stream));
    rc = fetch_image_stats(ctx, stats);
    PVP_TRY(cudaStreamSynchronize(ctx->stream));
#undef PVP_TRY
    cleanup();
    return rc;
}

}  // extern "C"
