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
    long long row0, row1;  // rows [row0, row1) of the image are processed (a row shard of a multi-GPU run; the whole image by default)
    long long gn[3], gs[3];
    double ext[6];
    double max_distance, step_size;
    long long th, tw, ts0, ts1;
    double c_ra, c_dec, half_w, half_h, ang_w, ang_h;
    int update_tex, bg_sub, grid_provided;
    double scale;          // 2^frac_bits
};

// checked build (see pvp_internal.h): every grid / texture offset handed to a RED must lie inside the span the caller's shape and
// strides describe; the spans and the violation counters live in this translation unit
#ifdef PVP_CHECKED
__device__ unsigned long long g_img_span_grid, g_img_span_tex, g_img_bad[2];
#define PVP_IMG_CHECK_GRID(off) ((unsigned long long)(off) < g_img_span_grid ? true : (atomicAdd(&g_img_bad[0], 1ull), false))
#define PVP_IMG_CHECK_TEX(off) ((unsigned long long)(off) < g_img_span_tex ? true : (atomicAdd(&g_img_bad[1], 1ull), false))
#else
#define PVP_IMG_CHECK_GRID(off) true
#define PVP_IMG_CHECK_TEX(off) true
#endif

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

// World direction of pixel (i, j)'s ray, process_image.cpp:241-260.  A function of its own because the four-pixel kernel evaluates it
// again (same operations, same bits) where it needs a direction that it does not keep in registers across its march loop.
__device__ __forceinline__ void pixel_dir(const ImageP &p, long long i, long long j, double &w0, double &w1, double &w2)
{
    const double x_cam = dsub((double)j, p.cx), y_cam = dsub((double)i, p.cy), z_cam = p.focal;  // :241-243
    const double norm = __dsqrt_rn(dadd(dadd(dmul(x_cam, x_cam), dmul(y_cam, y_cam)), dmul(z_cam, z_cam)));
    const double c0 = ddiv(x_cam, norm), c1 = ddiv(y_cam, norm), c2 = ddiv(z_cam, norm);
    w0 = dadd(dadd(dmul(p.xa[0], c0), dmul(p.ya[0], c1)), dmul(p.za[0], c2));  // :250-252
    w1 = dadd(dadd(dmul(p.xa[1], c0), dmul(p.ya[1], c1)), dmul(p.za[1], c2));
    w2 = dadd(dadd(dmul(p.xa[2], c0), dmul(p.ya[2], c1)), dmul(p.za[2], c2));
    const double dn = __dsqrt_rn(dadd(dadd(dmul(w0, w0), dmul(w1, w1)), dmul(w2, w2)));  // :255-260
    w0 = ddiv(w0, dn); w1 = ddiv(w1, dn); w2 = ddiv(w2, dn);
}

// The march bounds of a ray, process_image.cpp:24-61 (ray_aabb_intersection) and :334-340.  Like pixel_dir a function of the pixel
// alone (no side effects), so the four-pixel kernel can evaluate it a second time instead of keeping the bounds in registers.
__device__ __forceinline__ bool pixel_range(const ImageP &p, double w0, double w1, double w2, int &s_entry, int &s_exit)
{
    double tmin = -CUDART_INF, tmax = CUDART_INF;
    bool hit = true;
    const double dir[3] = {w0, w1, w2};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (fabs(dir[a]) > 1e-8) {
            double t1 = ddiv(dsub(p.ext[2 * a], p.o[a]), dir[a]);
            double t2 = ddiv(dsub(p.ext[2 * a + 1], p.o[a]), dir[a]);
            if (t1 > t2) { double sw = t1; t1 = t2; t2 = sw; }
            tmin = (tmin < t1) ? t1 : tmin;  // std::max(tmin, t1)
            tmax = (t2 < tmax) ? t2 : tmax;  // std::min(tmax, t2)
            if (tmin > tmax) hit = false;
        } else if (p.o[a] < p.ext[2 * a] || p.o[a] > p.ext[2 * a + 1]) {
            hit = false;
        }
    }
    if (!hit) return false;
    const double t_entry = (tmin < 0.0) ? 0.0 : tmin;                      // :334
    const double t_exit = (p.max_distance < tmax) ? p.max_distance : tmax;  // :335
    if (!(t_entry <= t_exit)) return false;                                 // :337
    s_entry = trunc32_like_x86(ddiv(t_entry, p.step_size));                 // :339-340
    s_exit = trunc32_like_x86(ddiv(t_exit, p.step_size));
    return true;
}

// One pixel up to the march bounds (process_image.cpp:236-340): brightness filter, fp64 camera ray, RA/Dec, sky-texture
// lookup / background subtraction / accumulate, AABB clip.  Returns true when the pixel marches samples
// s_entry..s_exit; `counted` tells whether it reached the ray stage (statistics).
template <int MODE>
__device__ __forceinline__ bool pixel_setup(const double *__restrict__ image, typename Cell<MODE>::T *__restrict__ tex, const ImageP &p,
                                            long long px, double &w0, double &w1, double &w2, typename Cell<MODE>::V &val,
                                            int &s_entry, int &s_exit, bool &counted)
{
    using C = Cell<MODE>;
    counted = false;
    const long long i = px / p.W, j = px - i * p.W;
    double brightness = image[i * p.is0 + j * p.is1];  // :236
    if (!(brightness > 0)) return false;               // :238

    pixel_dir(p, i, j, w0, w1, w2);  // :241-260

    if (p.update_tex || p.bg_sub) {  // RA/Dec and the texel have no other consumer (:267-314); warp-uniform
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
            if (brightness <= 0) return false;
        }
        val = C::make(brightness, p.scale);
        if (p.update_tex && within && PVP_IMG_CHECK_TEX(v_idx * p.ts0 + u_idx * p.ts1)) C::add(texel, val);  // :310-314
    } else {
        val = C::make(brightness, p.scale);
    }
    counted = true;
    if (!p.grid_provided) return false;  // :317

    return pixel_range(p, w0, w1, w2, s_entry, s_exit);
}

// Generic kernel: one thread per pixel, any grid size / strides / extent; one RED per sample.
template <int MODE>
__global__ void __launch_bounds__(K3_THREADS)
process_image_kernel(const double *__restrict__ image, typename Cell<MODE>::T *__restrict__ grid,
                     typename Cell<MODE>::T *__restrict__ tex, ImageP p, Ctl *__restrict__ ctl)
{
    using C = Cell<MODE>;
    unsigned long long my_rays = 0, my_samples = 0;
    const long long px_end = p.W * p.row1;
    for (long long px = p.W * p.row0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; px < px_end; px += (long long)gridDim.x * blockDim.x) {
        double w0, w1, w2;
        typename C::V val;
        int s_entry, s_exit;
        bool counted;
        const bool marches = pixel_setup<MODE>(image, tex, p, px, w0, w1, w2, val, s_entry, s_exit, counted);
        my_rays += counted;
        if (!marches) continue;

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
            if (PVP_IMG_CHECK_GRID(xi * p.gs[0] + yi * p.gs[1] + zi * p.gs[2])) C::add(grid + (xi * p.gs[0] + yi * p.gs[1] + zi * p.gs[2]), val);  // :356-357
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

// ---- lock-step kernel ----------------------------------------------------------------------------------------------
// The 32 lanes of a warp hold 32 consecutive pixels and march the SAME sample index s together, so neighbouring
// rays sit at the same depth and mostly in the same voxel: runs of equal voxel offsets among consecutive lanes are
// summed with a segmented warp scan and one RED is issued per run (the mechanism of K2).  Per sample the index
// arithmetic is 32 bit and the three divisions by the grid extent use a reciprocal prepared on the host:
//   q0 = RN(a*y); r = a - q0*b (exact, FMA); q1 = RN(q0 + r*y); r = a - q1*b; q = RN(q1 + r*y)   with y = RN(1/b)
// which is the correctly rounded a/b (Markstein's theorem; q1 is already faithful) unless b's significand is all
// ones -- the host checks that, and the exponent range, before choosing this kernel (fast_ok), so the voxel index
// is the reference's bit for bit.  FMA appears only inside this division, never in the reference's own expressions.
struct ImageFastP {
    double b[3], y[3], nd[3];  // extent, RN(1/extent), (double)n
    double k20[3];             // RN(n / extent) * 2^20: the screening factor of voxel_index()
    int n1[3];                 // n - 1
    unsigned gs[3];            // element strides
    double margin[3];          // 32 units of 2^-20 voxel in world units (safe_range)
    int affine_ok;             // the safe-range march may take the scaled coordinate from one FMA per axis (sample_keys2<2>)
    double x0[3];              // (o - ext_lo) * k20 + (2^52 - 5): the constant term of that FMA, the same for every pixel (Affine::x0)
};

__device__ __forceinline__ double div_by_extent(double a, double b, double y)
{
    double q = dmul(a, y);
    double r = __fma_rn(-q, b, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-q, b, a);
    return __fma_rn(r, y, q);
}

// Voxel index of a coordinate offset a in [0, b] along axis k: min(trunc(RN(RN(a / b) * n)), n - 1), process_image.cpp:94-104.
// Screening: x = a * K with K = RN(n/b) * 2^20 approximates (a*n/b) * 2^20 to a relative 2.3e-16, as does the reference's own value;
// for n <= 2^24 both lie within 4e-9 of the real number a*n/b, i.e. within 0.005 units of 2^-20.  y = x + (2^52 - 5) holds
// r - 5 in its low mantissa bits, r = x rounded to an integer (within one unit of x).  When the 20 fraction bits of r are in
// [5, 0xFFFFA], x is at least 4 units away from every multiple of 2^20, so r >> 20 == (r - 5) >> 20 IS the reference's index: one
// DMUL, one DADD and a funnel shift instead of the five-operation exact division and a 64-bit conversion.  Otherwise (about 10
// samples in a million, a ray that runs exactly along voxel faces, or x < 5, where y drops below 2^52 and its fraction field reads
// >= 0xFFFF6) the exact sequence decides.  CLAMP = false: the caller knows a < b with a margin (safe_range), so the index is < n.
// the screened index of voxel_index(); `unsafe` collects "the exact sequence must decide" over several calls so that one
// branch serves all the indices of a sample (six with two pixels per lane)
__device__ __forceinline__ int voxel_index_screened(double a, const ImageFastP &f, int k, bool &unsafe)
{
    const double y = dadd(dmul(a, f.k20[k]), 0x1p52 - 5.0);
    const unsigned lo = (unsigned)__double2loint(y), hi = (unsigned)__double2hiint(y);
    unsafe = unsafe || (lo & 0xFFFFFu) >= 0xFFFF6u;
    return (int)__funnelshift_r(lo, hi, 20);
}
__device__ __forceinline__ int voxel_index_exact(double a, const ImageFastP &f, int k)
{
    return __double2int_rz(dmul(div_by_extent(a, f.b[k], f.y[k]), f.nd[k]));
}
template <bool CLAMP = true>
__device__ __forceinline__ int voxel_index(double a, const ImageFastP &f, int k)
{
    bool unsafe = false;
    int idx = voxel_index_screened(a, f, k, unsafe);
    if (unsafe) idx = voxel_index_exact(a, f, k);
    return CLAMP ? min(idx, f.n1[k]) : idx;
}

// Self-test of div_by_extent against the IEEE division it replaces (pvp_selftest_div_by_extent): numerators are drawn
// where the march draws them -- uniformly in [0, b], and within a few ulps of every k*b/n boundary.  The quotient must be
// bit-identical for a >= 2^-960 (below that the exact remainder underflows and the quotient, < 2^-460, can only select
// voxel 0); the voxel index must be identical always.
__device__ __forceinline__ bool div_differs(double a, double b, double y, double nd, int n1, const ImageFastP &f)
{
    const double q1 = div_by_extent(a, b, y), q2 = ddiv(a, b);
    const int want = min(__double2int_rz(dmul(q2, nd)), n1);
    const bool idx_differs = min(__double2int_rz(dmul(q1, nd)), n1) != want || voxel_index(a, f, 0) != want;
    return idx_differs || (a >= 0x1p-960 && __double_as_longlong(q1) != __double_as_longlong(q2));
}

__global__ void div_selftest_kernel(ImageFastP f, long long n_random, int n_grid, unsigned long long seed,
                                    unsigned long long *__restrict__ mismatches)
{
    const double b = f.b[0], y = f.y[0];
    unsigned long long bad = 0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < n_random; i += nth) {
        unsigned long long x = (unsigned long long)i * 0x9e3779b97f4a7c15ull + seed;
        x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
        const double a = dmul((double)(x >> 11) * 0x1p-53, b);
        bad += div_differs(a, b, y, (double)n_grid, n_grid - 1, f);
    }
    for (long long i = tid; i < (long long)(n_grid + 1) * 129; i += nth) {
        const long long k = i / 129;
        const int d = (int)(i % 129) - 64;
        double a = dmul(ddiv((double)k, (double)n_grid), b);
        a = __longlong_as_double(__double_as_longlong(a) + d);  // +-64 ulps around the boundary
        if (!(a >= 0.0 && a <= b)) continue;
        bad += div_differs(a, b, y, (double)n_grid, n_grid - 1, f);
    }
    if (bad) atomicAdd(mismatches, bad);
}

constexpr unsigned IMG_NONE = 0xffffffffu;
constexpr int K3L_THREADS = 256;

struct ScanMasks { unsigned w1, w2, w4, w8, w16, next; };
__device__ __forceinline__ ScanMasks scan_masks(unsigned lane)
{
    ScanMasks m;
    auto win = [lane](unsigned d) { return lane + 1 >= d ? (((1u << d) - 1u) << (lane + 1 - d)) : ((2u << lane) - 1u); };
    m.w1 = 1u << lane; m.w2 = win(2); m.w4 = win(4); m.w8 = win(8);
    m.w16 = lane + 1 >= 16 ? (0xffffu << (lane - 15)) : ((2u << lane) - 1u);
    m.next = lane == 31 ? 1u : (1u << (lane + 1));
    // opaque: otherwise the compiler rebuilds the masks from the lane number inside the march loop (~20 instructions per iteration
    // in the float64 kernel) instead of keeping six registers
    asm volatile("" : "+r"(m.w1), "+r"(m.w2), "+r"(m.w4), "+r"(m.w8), "+r"(m.w16), "+r"(m.next));
    return m;
}

__device__ __forceinline__ bool img_run_head(unsigned key)
{
    unsigned prev, inrange;
    asm volatile("{\n\t.reg .pred q;\n\tshfl.sync.up.b32 %0|q, %2, 1, 0, 0xffffffff;\n\tselp.u32 %1, 1, 0, q;\n\t}" : "=r"(prev), "=r"(inrange) : "r"(key));
    return !inrange || key != prev;
}

__device__ __forceinline__ void img_red(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void img_red(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }
__device__ __forceinline__ void img_red(unsigned long long *p, unsigned v) { atomicAdd(p, (unsigned long long)v); }

// position of the sample at distance d along one ray (process_image.cpp:345-347) and its offsets from the low faces
struct SamplePos { double qx, qy, qz, ax, ay, az; };
__device__ __forceinline__ SamplePos sample_pos(const ImageP &p, double d, double w0, double w1, double w2)
{
    SamplePos r;
    r.qx = dadd(p.o[0], dmul(d, w0)); r.qy = dadd(p.o[1], dmul(d, w1)); r.qz = dadd(p.o[2], dmul(d, w2));
    r.ax = dsub(r.qx, p.ext[0]); r.ay = dsub(r.qy, p.ext[2]); r.az = dsub(r.qz, p.ext[4]);
    return r;
}

// in = lo <= s <= hi and the inclusive bounds test of :85 -- one chained predicate (the compiler's own form is a chain of
// eight selects) -- then key = in ? off : IMG_NONE
__device__ __forceinline__ unsigned tested_key(const ImageP &p, int s, int lo, int hi, const SamplePos &q, unsigned off)
{
    unsigned key;
    asm("{\n\t.reg .pred q;\n\t"
        "setp.ge.s32 q, %1, %2;\n\t"
        "setp.le.and.s32 q, %1, %3, q;\n\t"
        "setp.le.and.f64 q, %4, %5, q;\n\t"
        "setp.le.and.f64 q, %5, %6, q;\n\t"
        "setp.le.and.f64 q, %7, %8, q;\n\t"
        "setp.le.and.f64 q, %8, %9, q;\n\t"
        "setp.le.and.f64 q, %10, %11, q;\n\t"
        "setp.le.and.f64 q, %11, %12, q;\n\t"
        "selp.u32 %0, %13, 0xffffffff, q;\n\t}"
        : "=r"(key)
        : "r"(s), "r"(lo), "r"(hi), "d"(p.ext[0]), "d"(q.qx), "d"(p.ext[1]), "d"(p.ext[2]), "d"(q.qy), "d"(p.ext[3]), "d"(p.ext[4]), "d"(q.qz),
          "d"(p.ext[5]), "r"(off));
    return key;
}

// voxel offset of sample s of one ray, or IMG_NONE (outside [lo,hi], or outside the extent: process_image.cpp:85, :349-352)
__device__ __forceinline__ unsigned sample_key(const ImageP &p, const ImageFastP &f, int s, double w0, double w1, double w2, int lo, int hi)
{
    const SamplePos q = sample_pos(p, dmul((double)s, p.step_size), w0, w1, w2);  // :344
    // inside the extent the scaled coordinate lies in [0, n]: int truncation + upper clamp == :94-104
    bool unsafe = false;
    int xi = voxel_index_screened(q.ax, f, 0, unsafe), yi = voxel_index_screened(q.ay, f, 1, unsafe), zi = voxel_index_screened(q.az, f, 2, unsafe);
    if (unsafe) { xi = voxel_index_exact(q.ax, f, 0); yi = voxel_index_exact(q.ay, f, 1); zi = voxel_index_exact(q.az, f, 2); }
    xi = min(xi, f.n1[0]); yi = min(yi, f.n1[1]); zi = min(zi, f.n1[2]);
    return tested_key(p, s, lo, hi, q, (unsigned)xi * f.gs[0] + (unsigned)yi * f.gs[1] + (unsigned)zi * f.gs[2]);
}

// run-sum update of one scan level as a predicated add (the compiler's own form for integers is a select and an add)
__device__ __forceinline__ void scan_add(bool take, unsigned &sum, unsigned t)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q add.u32 %0, %0, %1;\n\t}" : "+r"(sum) : "r"(t), "r"((unsigned)take));
}
__device__ __forceinline__ void scan_add(bool take, unsigned long long &sum, unsigned long long t)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q add.u64 %0, %0, %1;\n\t}" : "+l"(sum) : "l"(t), "r"((unsigned)take));
}
__device__ __forceinline__ void scan_add(bool take, double &sum, double t)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q add.rn.f64 %0, %0, %1;\n\t}" : "+d"(sum) : "d"(t), "r"((unsigned)take));
}

// One warp, one group of 32 pixels: samples wlo..whi in lock-step.  S is the type the run sums are formed in.
template <typename S, typename CellT>
__device__ __forceinline__ unsigned march_lockstep(CellT *__restrict__ grid, const ImageP &p, const ImageFastP &f, const ScanMasks &lm,
                                                   double w0, double w1, double w2, S val, int lo, int hi, int wlo, int whi)
{
    const unsigned FULL = 0xffffffffu;
    unsigned n_in = 0;
    auto key_of = [&](int s) -> unsigned { return sample_key(p, f, s, w0, w1, w2, lo, hi); };
    unsigned key = key_of(wlo);
    unsigned heads = __ballot_sync(FULL, img_run_head(key));
    for (int s = wlo;; ++s) {
        // next sample's key first: its arithmetic and shuffle overlap the scan of the current one
        unsigned key_n = IMG_NONE;
        if (s != whi) key_n = key_of(s + 1);
        const bool head_n = img_run_head(key_n);
        const bool live = key != IMG_NONE;
        n_in += live;
        S sum = live ? val : (S)0, t;
        t = __shfl_up_sync(FULL, sum, 1);  if (!(heads & lm.w1))  sum += t;
        t = __shfl_up_sync(FULL, sum, 2);  if (!(heads & lm.w2))  sum += t;
        t = __shfl_up_sync(FULL, sum, 4);  if (!(heads & lm.w4))  sum += t;
        t = __shfl_up_sync(FULL, sum, 8);  if (!(heads & lm.w8))  sum += t;
        t = __shfl_up_sync(FULL, sum, 16); if (!(heads & lm.w16)) sum += t;
        if ((heads & lm.next) && live && PVP_IMG_CHECK_GRID(key)) img_red(grid + key, sum);  // :356-357, once per run
        if (s == whi) break;
        heads = __ballot_sync(FULL, head_n);
        key = key_n;
    }
    return n_in;
}

// A sub-range [slo, shi] of a ray's samples [lo, hi] on which the inclusive bounds test of process_image.cpp:85 provably passes,
// so that the march can skip its six fp64 compares there.  The computed sample position q = RN(o + RN(RN(s*step) * w)) is
// within e = 2^-48 (|o| + |d_max w| + |ext|) of the exact x(s) = o + s*step*w per axis (three roundings, each below 2^-52 of
// those magnitudes), and x(s) is linear in s: if at two samples the computed position keeps a margin of 2e to all six faces,
// the exact position keeps e there and hence everywhere between them, and every computed position between them is inside.
// The candidates are the first / last three samples of the ray; a ray that grazes a face gets an empty range (tests everywhere).
__device__ __forceinline__ void safe_range(const ImageP &p, const ImageFastP &f, double w0, double w1, double w2, int lo, int hi, int &slo, int &shi)
{
    slo = INT32_MAX; shi = INT32_MIN;
    if (lo > hi) return;
    const double w[3] = {w0, w1, w2};
    const double dmax = fabs((double)hi * p.step_size) + fabs((double)lo * p.step_size);
    double e[3];
#pragma unroll
    // 2e, plus 32 units of 2^-20 voxel: the scaled coordinate of a safe sample is then >= 32, so sample_keys2<2>'s y = RN(s dx + x0),
    // within 1.6 units of x(s) + 2^52 - 5, cannot drop below 2^52 (where its mantissa would no longer hold r - 5)
    for (int k = 0; k < 3; ++k) e[k] = 0x1p-47 * (fabs(p.o[k]) + dmax * fabs(w[k]) + fabs(p.ext[2 * k]) + fabs(p.ext[2 * k + 1])) + f.margin[k];
    auto inside = [&](int s) {
        const double d = (double)s * p.step_size;
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double q = p.o[k] + d * w[k];
            ok = ok && (p.ext[2 * k] + e[k] <= q) && (q <= p.ext[2 * k + 1] - e[k]);
        }
        return ok;
    };
    int a = lo, b = hi;
    for (int tries = 0; tries < 3 && a <= hi && !inside(a); ++tries) ++a;
    for (int tries = 0; tries < 3 && b >= lo && !inside(b); ++tries) --b;
    if (a <= b && inside(a) && inside(b)) { slo = a; shi = b; }
}

// Two pixels per lane (vertical neighbours, the lean2 idea of path A): pixel B's sample sits in pixel A's voxel most of the
// time and then joins A's value before the scan, so one scan serves 64 samples; otherwise B issues its own RED.
struct March2 {
    unsigned ka, kb, heads, n_in;
};

// Per-axis affine form of the scaled voxel coordinate of a ray's samples, x(s) = (o + s*step*w - ext_lo) * K + (2^52 - 5)
// = s * dx + x0, for the safe range.
struct Affine { double dx[3], x0[3]; };
__device__ __forceinline__ Affine affine_of(const ImageP &p, const ImageFastP &f, double w0, double w1, double w2)
{
    const double w[3] = {w0, w1, w2};
    Affine a;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a.dx[k] = (p.step_size * w[k]) * f.k20[k];
        a.x0[k] = (p.o[k] - p.ext[2 * k]) * f.k20[k] + (0x1p52 - 5.0);
    }
    return a;
}

// keys of sample s of both rays.  MODE 0: with the range and bounds tests.  MODE 1: for samples where both are known to pass
// (safe_range: also a margin to the upper faces, so no clamp).  MODE 2: as 1, and the scaled coordinate of voxel_index_screened --
// whose screening only needs it to within a few units of 2^-20 -- comes from ONE FMA per axis, y = RN(s * dx + x0), instead of
// the reference's five roundings: with A = |o| + max_distance + |ext| and K = n/extent * 2^20, y is within 1 + 2^-51 A K units of
// the real x(s) + 2^52 - 5 (rounding of dx, x0 and the FMA itself) and the reference's own RN(RN(a/b) n) within 2^-50 A K + 2^-8
// units of x(s); the host allows this mode when A K <= 2^49 (make_fastp), so the two differ by less than 2.1 units and the window of
// voxel_index_screened (5 units to either side of a multiple of 2^20) still proves trunc() equal; safe_range keeps x(s) >= 32, so y >= 2^52.  Flagged samples take the exact
// sequence on the reference's own operands.  One distance, one "exact division needed" branch for all six indices.
template <int MODE>
__device__ __forceinline__ void sample_keys2(const ImageP &p, const ImageFastP &f, int s, double a0, double a1, double a2, int lo_a, int hi_a,
                                             double b0, double b1, double b2, int lo_b, int hi_b, const Affine &fa, const Affine &fb,
                                             unsigned &ka, unsigned &kb)
{
    int xa, ya, za, xb, yb, zb;
    if (MODE == 2) {
        const double sd = (double)s;
        const double y[6] = {__fma_rn(sd, fa.dx[0], fa.x0[0]), __fma_rn(sd, fa.dx[1], fa.x0[1]), __fma_rn(sd, fa.dx[2], fa.x0[2]),
                             __fma_rn(sd, fb.dx[0], fb.x0[0]), __fma_rn(sd, fb.dx[1], fb.x0[1]), __fma_rn(sd, fb.dx[2], fb.x0[2])};
        int idx[6];
        bool unsafe = false;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const unsigned lo = (unsigned)__double2loint(y[k]), hi = (unsigned)__double2hiint(y[k]);
            unsafe = unsafe || (lo & 0xFFFFFu) >= 0xFFFF6u;
            idx[k] = (int)__funnelshift_r(lo, hi, 20);
        }
        xa = idx[0]; ya = idx[1]; za = idx[2]; xb = idx[3]; yb = idx[4]; zb = idx[5];
        if (unsafe) {
            const double d = dmul(sd, p.step_size);  // :344
            const SamplePos qa = sample_pos(p, d, a0, a1, a2), qb = sample_pos(p, d, b0, b1, b2);
            xa = voxel_index_exact(qa.ax, f, 0); ya = voxel_index_exact(qa.ay, f, 1); za = voxel_index_exact(qa.az, f, 2);
            xb = voxel_index_exact(qb.ax, f, 0); yb = voxel_index_exact(qb.ay, f, 1); zb = voxel_index_exact(qb.az, f, 2);
        }
        ka = lo_a <= hi_a ? (unsigned)xa * f.gs[0] + (unsigned)ya * f.gs[1] + (unsigned)za * f.gs[2] : IMG_NONE;
        kb = lo_b <= hi_b ? (unsigned)xb * f.gs[0] + (unsigned)yb * f.gs[1] + (unsigned)zb * f.gs[2] : IMG_NONE;
        return;
    }
    const double d = dmul((double)s, p.step_size);  // :344
    const SamplePos qa = sample_pos(p, d, a0, a1, a2), qb = sample_pos(p, d, b0, b1, b2);
    bool unsafe = false;
    xa = voxel_index_screened(qa.ax, f, 0, unsafe); ya = voxel_index_screened(qa.ay, f, 1, unsafe); za = voxel_index_screened(qa.az, f, 2, unsafe);
    xb = voxel_index_screened(qb.ax, f, 0, unsafe); yb = voxel_index_screened(qb.ay, f, 1, unsafe); zb = voxel_index_screened(qb.az, f, 2, unsafe);
    if (unsafe) {  // ~10 per million indices: the exact sequence is the reference's value for every one of them
        xa = voxel_index_exact(qa.ax, f, 0); ya = voxel_index_exact(qa.ay, f, 1); za = voxel_index_exact(qa.az, f, 2);
        xb = voxel_index_exact(qb.ax, f, 0); yb = voxel_index_exact(qb.ay, f, 1); zb = voxel_index_exact(qb.az, f, 2);
    }
    if (MODE == 0) {
        xa = min(xa, f.n1[0]); ya = min(ya, f.n1[1]); za = min(za, f.n1[2]);
        xb = min(xb, f.n1[0]); yb = min(yb, f.n1[1]); zb = min(zb, f.n1[2]);
    }
    const unsigned off_a = (unsigned)xa * f.gs[0] + (unsigned)ya * f.gs[1] + (unsigned)za * f.gs[2];
    const unsigned off_b = (unsigned)xb * f.gs[0] + (unsigned)yb * f.gs[1] + (unsigned)zb * f.gs[2];
    if (MODE == 0) {
        ka = tested_key(p, s, lo_a, hi_a, qa, off_a);
        kb = tested_key(p, s, lo_b, hi_b, qb, off_b);
    } else {
        ka = lo_a <= hi_a ? off_a : IMG_NONE;  // a lane without a ray stays silent
        kb = lo_b <= hi_b ? off_b : IMG_NONE;
    }
}

template <bool COUNT, typename S, typename CellT>
__device__ __forceinline__ void march2_accumulate(CellT *__restrict__ grid, const ScanMasks &lm, March2 &m, S val_a, S val_b)
{
    const unsigned FULL = 0xffffffffu;
    const bool live_a = m.ka != IMG_NONE, live_b = m.kb != IMG_NONE;
    if (COUNT) m.n_in += (unsigned)live_a + (unsigned)live_b;
    const bool same = live_b && m.kb == m.ka;
    const bool b_alone = live_b && !same;
    S sum = (live_a ? val_a : (S)0) + (same ? val_b : (S)0), t;
    t = __shfl_up_sync(FULL, sum, 1);  scan_add(!(m.heads & lm.w1), sum, t);
    t = __shfl_up_sync(FULL, sum, 2);  scan_add(!(m.heads & lm.w2), sum, t);
    t = __shfl_up_sync(FULL, sum, 4);  scan_add(!(m.heads & lm.w4), sum, t);
    t = __shfl_up_sync(FULL, sum, 8);  scan_add(!(m.heads & lm.w8), sum, t);
    t = __shfl_up_sync(FULL, sum, 16); scan_add(!(m.heads & lm.w16), sum, t);
    if ((m.heads & lm.next) && live_a && PVP_IMG_CHECK_GRID(m.ka)) img_red(grid + m.ka, sum);
    if (b_alone && PVP_IMG_CHECK_GRID(m.kb)) img_red(grid + m.kb, val_b);
}

// iterations s .. s_stop-1; each first computes the keys of sample s+1 (their arithmetic and shuffle overlap the scan of
// sample s) -- with the full tests (MODE 0), or without them where s+1 lies in the warp's safe range (MODE 1 / 2, sample_keys2)
template <int MODE, typename S, typename CellT>
__device__ __forceinline__ void march2_span(CellT *__restrict__ grid, const ImageP &p, const ImageFastP &f, const ScanMasks &lm, March2 &m, int &s, int s_stop,
                                            double a0, double a1, double a2, S val_a, int lo_a, int hi_a,
                                            double b0, double b1, double b2, S val_b, int lo_b, int hi_b)
{
    const unsigned FULL = 0xffffffffu;
    if (s >= s_stop) return;
    Affine fa, fb;
    if (MODE == 2) { fa = affine_of(p, f, a0, a1, a2); fb = affine_of(p, f, b0, b1, b2); }
    if (MODE != 0) {
        // not counted per iteration in the safe range: sample s (keys computed earlier) by its keys, samples s+1 .. s_stop-1 -- computed
        // AND accumulated by this span -- are inside for every ray the lane has; sample s_stop is counted by whoever accumulates it
        m.n_in += (unsigned)(m.ka != IMG_NONE) + (unsigned)(m.kb != IMG_NONE);
        m.n_in += (unsigned)(s_stop - 1 - s) * ((unsigned)(lo_a <= hi_a) + (unsigned)(lo_b <= hi_b));
    }
    for (; s < s_stop; ++s) {
        unsigned ka_n, kb_n;
        sample_keys2<MODE>(p, f, s + 1, a0, a1, a2, lo_a, hi_a, b0, b1, b2, lo_b, hi_b, fa, fb, ka_n, kb_n);
        const bool head_n = img_run_head(ka_n);
        march2_accumulate<MODE == 0, S>(grid, lm, m, val_a, val_b);
        m.heads = __ballot_sync(FULL, head_n);
        m.ka = ka_n; m.kb = kb_n;
    }
}

// [mlo, mhi]: the samples inside every active ray's safe range (empty: mlo > mhi)
template <typename S, typename CellT>
__device__ __forceinline__ unsigned march_lockstep2(CellT *__restrict__ grid, const ImageP &p, const ImageFastP &f, const ScanMasks &lm,
                                                    double a0, double a1, double a2, S val_a, int lo_a, int hi_a,
                                                    double b0, double b1, double b2, S val_b, int lo_b, int hi_b, int wlo, int whi, int mlo, int mhi)
{
    const unsigned FULL = 0xffffffffu;
    March2 m;
    m.n_in = 0;
    const Affine none = {};
    sample_keys2<0>(p, f, wlo, a0, a1, a2, lo_a, hi_a, b0, b1, b2, lo_b, hi_b, none, none, m.ka, m.kb);
    m.heads = __ballot_sync(FULL, img_run_head(m.ka));
    int s = wlo;
    march2_span<0, S>(grid, p, f, lm, m, s, min(whi, mlo - 1), a0, a1, a2, val_a, lo_a, hi_a, b0, b1, b2, val_b, lo_b, hi_b);   // s+1 < mlo
    if (f.affine_ok) march2_span<2, S>(grid, p, f, lm, m, s, min(whi, mhi), a0, a1, a2, val_a, lo_a, hi_a, b0, b1, b2, val_b, lo_b, hi_b);  // mlo <= s+1 <= mhi
    else march2_span<1, S>(grid, p, f, lm, m, s, min(whi, mhi), a0, a1, a2, val_a, lo_a, hi_a, b0, b1, b2, val_b, lo_b, hi_b);
    march2_span<0, S>(grid, p, f, lm, m, s, whi, a0, a1, a2, val_a, lo_a, hi_a, b0, b1, b2, val_b, lo_b, hi_b);
    march2_accumulate<true, S>(grid, lm, m, val_a, val_b);  // sample whi
    return m.n_in;
}

#ifndef PVP_K3_MINB
#define PVP_K3_MINB 3  // 80 registers: measured 4096^2 -> 512^3 fixed point 624 G samples/s, against 517 G with ptxas' own choice (94 registers, 2 CTAs per SM)
#endif
template <int MODE>
__global__ void __launch_bounds__(K3L_THREADS, PVP_K3_MINB)
process_image_lockstep2_kernel(const double *__restrict__ image, typename Cell<MODE>::T *__restrict__ grid,
                               typename Cell<MODE>::T *__restrict__ tex, ImageP p, ImageFastP f, Ctl *__restrict__ ctl)
{
    using C = Cell<MODE>;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const ScanMasks lm = scan_masks(lane);
    // a warp's tile: 32 columns x 2 rows; lane l holds the two pixels of column l
    const long long tiles_x = (p.W + 31) / 32, tiles_y = (p.row1 - p.row0 + 1) / 2;
    const unsigned long long n_groups = (unsigned long long)(tiles_x * tiles_y);
    unsigned long long my_rays = 0, my_samples = 0;
    for (;;) {
        unsigned long long group = 0;
        if (lane == 0) group = atomicAdd(&ctl->head, 1ull);
        group = __shfl_sync(FULL, group, 0);
        if (group >= n_groups) break;
        const long long ty = (long long)(group / (unsigned long long)tiles_x), tx = (long long)group - ty * tiles_x;
        const long long pj = tx * 32 + lane, pi = p.row0 + ty * 2;
        double a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
        typename C::V val_a = 0, val_b = 0;
        int lo_a = INT32_MAX, hi_a = INT32_MIN, lo_b = INT32_MAX, hi_b = INT32_MIN;  // empty ranges
        if (pj < p.W) {
            int s_entry, s_exit;
            bool counted;
            if (pixel_setup<MODE>(image, tex, p, pi * p.W + pj, a0, a1, a2, val_a, s_entry, s_exit, counted) && s_entry <= s_exit) { lo_a = s_entry; hi_a = s_exit; }
            my_rays += counted;
            if (pi + 1 < p.row1) {
                if (pixel_setup<MODE>(image, tex, p, (pi + 1) * p.W + pj, b0, b1, b2, val_b, s_entry, s_exit, counted) && s_entry <= s_exit) { lo_b = s_entry; hi_b = s_exit; }
                my_rays += counted;
            }
        }
        const int wlo = __reduce_min_sync(FULL, min(lo_a, lo_b)), whi = __reduce_max_sync(FULL, max(hi_a, hi_b));
        if (wlo > whi) continue;
        // the samples on which every active ray of the warp provably passes its range and bounds tests
        int slo_a, shi_a, slo_b, shi_b;
        safe_range(p, f, a0, a1, a2, lo_a, hi_a, slo_a, shi_a);
        safe_range(p, f, b0, b1, b2, lo_b, hi_b, slo_b, shi_b);
        const int mlo = __reduce_max_sync(FULL, max(lo_a <= hi_a ? slo_a : INT32_MIN, lo_b <= hi_b ? slo_b : INT32_MIN));
        const int mhi = __reduce_min_sync(FULL, min(lo_a <= hi_a ? shi_a : INT32_MAX, lo_b <= hi_b ? shi_b : INT32_MAX));
        if (MODE == PVP_ACCUM_FIXED) {
            // run sums of 64 values fit 32 bits when every value is below 2^26
            if (__all_sync(FULL, (unsigned long long)val_a < (1ull << 26) && (unsigned long long)val_b < (1ull << 26)))
                my_samples += march_lockstep2<unsigned>((unsigned long long *)grid, p, f, lm, a0, a1, a2, (unsigned)val_a, lo_a, hi_a, b0, b1, b2, (unsigned)val_b,
                                                        lo_b, hi_b, wlo, whi, mlo, mhi);
            else
                my_samples += march_lockstep2<unsigned long long>((unsigned long long *)grid, p, f, lm, a0, a1, a2, (unsigned long long)val_a, lo_a, hi_a, b0, b1,
                                                                  b2, (unsigned long long)val_b, lo_b, hi_b, wlo, whi, mlo, mhi);
        } else {
            my_samples += march_lockstep2<double>((double *)grid, p, f, lm, a0, a1, a2, (double)val_a, lo_a, hi_a, b0, b1, b2, (double)val_b, lo_b, hi_b, wlo, whi,
                                                  mlo, mhi);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_rays += __shfl_xor_sync(FULL, my_rays, o);
        my_samples += __shfl_xor_sync(FULL, my_samples, o);
    }
    if (lane == 0 && (my_rays | my_samples)) {
        atomicAdd(&ctl->img_rays, my_rays);
        atomicAdd(&ctl->img_samples, my_samples);
    }
}

// ---- NP (three) pixels per lane (the lean4 idea of path A) ---------------------------------------------------------------
// A warp's tile is 32 columns x NP rows; lane l holds the NP pixels of column l.  The per-iteration overhead of the march loop (run
// scan, run-head ballot and shuffle, loop control) is paid once per 32 * NP samples.  Inside a lane the pixels are chained bottom-up:
// pixel j's sample joins pixel j-1's where both sit in the same voxel, else it issues its own RED; pixel 0 carries the lane's sum into
// the run scan, whose runs are cut at every CAP-th lane (fewer scan levels).  To hold four pixels in registers the loop over the safe
// range keeps only what its one FMA per axis needs (the increments dx; the constant term x0 is the same for every pixel and comes from
// the host: ImageFastP::x0); the ray directions and march bounds, which only the tested spans at either end of the march and the ~10 samples per
// million that take the exact division read, live in shared memory (LaneRays).  Requires ImageFastP::affine_ok (else the two-pixel kernel).
// Measured, 4096^2 -> 512^3 int64, G samples/s looking along z / y (tools/image_probe.py; two-pixel kernel: 631 / 654;
// profiles/r2_k3n_tuning.log): pixels per lane x CTA shape x run cap -- 4 x (256 x 2, 124 registers) x 8: 656; 4 x (256 x 3, 80 registers) x 8:
// 734 / ...; 3 x (128 x 6, 80) x 8: 746 / 798; 3 x (128 x 7, 72) x 8: 757 / 809; 3 x (128 x 7, 72) x 4: 775 / 829 <- kept; 3 x (128 x 8, 64) x 4:
// 759 / 807; 2 x (128 x 8, 64) x 4: 771 / 787; 4 x (256 x 3) x 4: 756 / 812; run cap 16: 710-716.
#ifndef PVP_K3N_NP
#define PVP_K3N_NP 3        // pixels per lane (rows of a warp's 32-column tile)
#endif
#ifndef PVP_K3N_CAP
#define PVP_K3N_CAP 4       // longest run of lanes the scan merges
#endif
#ifndef PVP_K3N_THREADS
#define PVP_K3N_THREADS 128
#endif
#ifndef PVP_K3N_MINB
#define PVP_K3N_MINB 7      // 72 registers, 28 warps per SM
#endif
constexpr int K3N_THREADS = PVP_K3N_THREADS;

template <int CAP>
__device__ __forceinline__ ScanMasks scan_masks_cap(unsigned lane)
{
    ScanMasks m;
    auto win = [lane](unsigned d) { return lane + 1 >= d ? (((1u << d) - 1u) << (lane + 1 - d)) : ((2u << lane) - 1u); };
    m.w1 = 1u << lane; m.w2 = win(2); m.w4 = win(4); m.w8 = win(8);
    m.w16 = lane + 1 >= 16 ? (0xffffu << (lane - 15)) : ((2u << lane) - 1u);
    m.next = lane == 31 ? 1u : (1u << (lane + 1));
    if (CAP < 32) {  // bit 0 of the head ballot is always set: a window that reaches a cut gets it too, the lane below a cut issues through it
        const unsigned in_group = lane % CAP;
        if (1 > in_group) m.w1 |= 1u;
        if (2 > in_group) m.w2 |= 1u;
        if (4 > in_group) m.w4 |= 1u;
        if (8 > in_group) m.w8 |= 1u;
        if (16 > in_group) m.w16 |= 1u;
        if (in_group == CAP - 1) m.next |= 1u;
    }
    asm volatile("" : "+r"(m.w1), "+r"(m.w2), "+r"(m.w4), "+r"(m.w8), "+r"(m.w16), "+r"(m.next));
    return m;
}

template <int NP> struct MarchN {
    unsigned k[NP];
    unsigned heads, n_in;
};

// What the tested spans and the exact-division fallback need of a lane's pixels -- ray directions and march bounds -- lives in shared
// memory, one column per thread (every thread reads back only what it wrote: no barrier, no bank conflict), so that the safe-range
// loop holds four pixels in 96 registers: the FMA increments dx, the values, the keys.
template <int NP> struct LaneRays {
    double (*w)[K3N_THREADS];   // [NP * 3][threads]
    int (*lo)[K3N_THREADS];     // [NP][threads]
    int (*hi)[K3N_THREADS];
};

// keys of sample s of the lane's NP pixels.  TESTED: with the range and bounds tests (the spans at either end of a march).  Otherwise
// (inside every ray's safe range): one FMA per axis, y = RN(s * dx + x0) -- see sample_keys2<2> for why its screened truncation is the
// reference's index; a flagged sample takes the exact sequence on the reference's own operands.
template <bool TESTED, int NP>
__device__ __forceinline__ void sample_keysN(const ImageP &p, const ImageFastP &f, int s, const LaneRays<NP> &lr, const double (&dx)[NP][3],
                                             unsigned has, unsigned (&k)[NP])
{
    const unsigned t = threadIdx.x;
    if (TESTED) {
        const double d = dmul((double)s, p.step_size);  // :344
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const SamplePos q = sample_pos(p, d, lr.w[3 * j][t], lr.w[3 * j + 1][t], lr.w[3 * j + 2][t]);
            bool unsafe = false;
            int xi = voxel_index_screened(q.ax, f, 0, unsafe), yi = voxel_index_screened(q.ay, f, 1, unsafe), zi = voxel_index_screened(q.az, f, 2, unsafe);
            if (unsafe) { xi = voxel_index_exact(q.ax, f, 0); yi = voxel_index_exact(q.ay, f, 1); zi = voxel_index_exact(q.az, f, 2); }
            xi = min(xi, f.n1[0]); yi = min(yi, f.n1[1]); zi = min(zi, f.n1[2]);
            k[j] = tested_key(p, s, lr.lo[j][t], lr.hi[j][t], q, (unsigned)xi * f.gs[0] + (unsigned)yi * f.gs[1] + (unsigned)zi * f.gs[2]);
        }
        return;
    }
    const double sd = (double)s;
    int idx[NP][3];
    bool unsafe = false;
#pragma unroll
    for (int j = 0; j < NP; ++j)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double y = __fma_rn(sd, dx[j][a], f.x0[a]);
            const unsigned l = (unsigned)__double2loint(y), h = (unsigned)__double2hiint(y);
            unsafe = unsafe || (l & 0xFFFFFu) >= 0xFFFF6u;
            idx[j][a] = (int)__funnelshift_r(l, h, 20);
        }
    if (unsafe) {
        const double d = dmul(sd, p.step_size);  // :344
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            if (!(has & (1u << j))) continue;   // no ray: the key below is IMG_NONE whatever the indices
            const SamplePos q = sample_pos(p, d, lr.w[3 * j][t], lr.w[3 * j + 1][t], lr.w[3 * j + 2][t]);
            idx[j][0] = voxel_index_exact(q.ax, f, 0); idx[j][1] = voxel_index_exact(q.ay, f, 1); idx[j][2] = voxel_index_exact(q.az, f, 2);
        }
    }
#pragma unroll
    for (int j = 0; j < NP; ++j)
        k[j] = (has & (1u << j)) ? (unsigned)idx[j][0] * f.gs[0] + (unsigned)idx[j][1] * f.gs[1] + (unsigned)idx[j][2] * f.gs[2] : IMG_NONE;
}

template <bool COUNT, int NP, int CAP, typename S, typename CellT>
__device__ __forceinline__ void marchN_accumulate(CellT *__restrict__ grid, const ScanMasks &lm, MarchN<NP> &m, const S (&val)[NP])
{
    const unsigned FULL = 0xffffffffu;
    S sv[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const bool live = m.k[j] != IMG_NONE;
        if (COUNT) m.n_in += (unsigned)live;
        sv[j] = live ? val[j] : (S)0;
    }
#pragma unroll
    for (int j = NP - 1; j > 0; --j) {   // the lane's chain
        const bool live = m.k[j] != IMG_NONE;
        const bool same = live && m.k[j] == m.k[j - 1];
        if (live && !same && PVP_IMG_CHECK_GRID(m.k[j])) img_red(grid + m.k[j], sv[j]);
        scan_add(same, sv[j - 1], sv[j]);
    }
    S sum = sv[0], t;
    t = __shfl_up_sync(FULL, sum, 1);  scan_add(!(m.heads & lm.w1), sum, t);
    t = __shfl_up_sync(FULL, sum, 2);  scan_add(!(m.heads & lm.w2), sum, t);
    if (CAP > 4)  { t = __shfl_up_sync(FULL, sum, 4);  scan_add(!(m.heads & lm.w4), sum, t); }
    if (CAP > 8)  { t = __shfl_up_sync(FULL, sum, 8);  scan_add(!(m.heads & lm.w8), sum, t); }
    if (CAP > 16) { t = __shfl_up_sync(FULL, sum, 16); scan_add(!(m.heads & lm.w16), sum, t); }
    if ((m.heads & lm.next) && m.k[0] != IMG_NONE && PVP_IMG_CHECK_GRID(m.k[0])) img_red(grid + m.k[0], sum);
}

// iterations s .. s_stop-1; each first computes the keys of sample s+1 (their arithmetic and shuffle overlap the scan of sample s)
template <bool TESTED, int NP, int CAP, typename S, typename CellT>
__device__ __forceinline__ void marchN_span(CellT *__restrict__ grid, const ImageP &p, const ImageFastP &f, const ScanMasks &lm, MarchN<NP> &m, int &s, int s_stop,
                                            const LaneRays<NP> &lr, const double (&dx)[NP][3], const S (&val)[NP], unsigned has)
{
    const unsigned FULL = 0xffffffffu;
    if (s >= s_stop) return;
    if (!TESTED) {
        // not counted per iteration in the safe range: sample s (keys computed earlier) by its keys; samples s+1 .. s_stop-1 -- computed
        // AND accumulated by this span -- are inside for every ray the lane has; sample s_stop is counted by whoever accumulates it
#pragma unroll
        for (int j = 0; j < NP; ++j) m.n_in += (unsigned)(m.k[j] != IMG_NONE);
        m.n_in += (unsigned)(s_stop - 1 - s) * (unsigned)__popc(has);
    }
    for (; s < s_stop; ++s) {
        unsigned kn[NP];
        sample_keysN<TESTED, NP>(p, f, s + 1, lr, dx, has, kn);
        const bool head_n = img_run_head(kn[0]);
        marchN_accumulate<TESTED, NP, CAP, S>(grid, lm, m, val);
        m.heads = __ballot_sync(FULL, head_n);
#pragma unroll
        for (int j = 0; j < NP; ++j) m.k[j] = kn[j];
    }
}

// [mlo, mhi]: the samples inside every active ray's safe range (empty: mlo > mhi); has: bit j = pixel j marches
template <int NP, int CAP, typename S, typename CellT>
__device__ __forceinline__ unsigned march_lockstepN(CellT *__restrict__ grid, const ImageP &p, const ImageFastP &f, const ScanMasks &lm,
                                                    const LaneRays<NP> &lr, const double (&dx)[NP][3], const S (&val)[NP], unsigned has,
                                                    int wlo, int whi, int mlo, int mhi)
{
    const unsigned FULL = 0xffffffffu;
    MarchN<NP> m;
    m.n_in = 0;
    sample_keysN<true, NP>(p, f, wlo, lr, dx, has, m.k);
    m.heads = __ballot_sync(FULL, img_run_head(m.k[0]));
    int s = wlo;
    marchN_span<true, NP, CAP, S>(grid, p, f, lm, m, s, min(whi, mlo - 1), lr, dx, val, has);    // s+1 < mlo
    marchN_span<false, NP, CAP, S>(grid, p, f, lm, m, s, min(whi, mhi), lr, dx, val, has);       // mlo <= s+1 <= mhi
    marchN_span<true, NP, CAP, S>(grid, p, f, lm, m, s, whi, lr, dx, val, has);
    marchN_accumulate<true, NP, CAP, S>(grid, lm, m, val);  // sample whi
    return m.n_in;
}

template <int MODE>
__global__ void __launch_bounds__(K3N_THREADS, PVP_K3N_MINB)
process_image_lockstepN_kernel(const double *__restrict__ image, typename Cell<MODE>::T *__restrict__ grid,
                               typename Cell<MODE>::T *__restrict__ tex, ImageP p, ImageFastP f, Ctl *__restrict__ ctl)
{
    using C = Cell<MODE>;
    constexpr int NP = PVP_K3N_NP, CAP = PVP_K3N_CAP;
    __shared__ double s_w[NP * 3][K3N_THREADS];
    __shared__ int s_lo[NP][K3N_THREADS], s_hi[NP][K3N_THREADS];
    const LaneRays<NP> lr = {s_w, s_lo, s_hi};
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31, t = threadIdx.x;
    const ScanMasks lm = scan_masks_cap<CAP>(lane);
    const long long tiles_x = (p.W + 31) / 32, tiles_y = (p.row1 - p.row0 + NP - 1) / NP;
    const unsigned long long n_groups = (unsigned long long)(tiles_x * tiles_y);
    unsigned long long my_rays = 0, my_samples = 0;
    for (;;) {
        unsigned long long group = 0;
        if (lane == 0) group = atomicAdd(&ctl->head, 1ull);
        group = __shfl_sync(FULL, group, 0);
        if (group >= n_groups) break;
        const long long ty = (long long)(group / (unsigned long long)tiles_x), tx = (long long)group - ty * tiles_x;
        const long long pj = tx * 32 + lane, pi = p.row0 + ty * NP;
        unsigned has = 0;
        double dx[NP][3];
        typename C::V val[NP];
        int lane_lo = INT32_MAX, lane_hi = INT32_MIN, lane_mlo = INT32_MIN, lane_mhi = INT32_MAX;
        bool narrow = true;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            double w0 = 0.0, w1 = 0.0, w2 = 0.0;
            int lo = INT32_MAX, hi = INT32_MIN;  // empty range
            val[j] = 0;
            if (pj < p.W && pi + j < p.row1) {
                int s_entry, s_exit;
                bool counted;
                if (pixel_setup<MODE>(image, tex, p, (pi + j) * p.W + pj, w0, w1, w2, val[j], s_entry, s_exit, counted) && s_entry <= s_exit) {
                    lo = s_entry; hi = s_exit;
                    has |= 1u << j;
                }
                my_rays += counted;
            }
            s_w[3 * j][t] = w0; s_w[3 * j + 1][t] = w1; s_w[3 * j + 2][t] = w2;
            s_lo[j][t] = lo; s_hi[j][t] = hi;
            dx[j][0] = dmul(dmul(p.step_size, w0), f.k20[0]);   // Affine::dx
            dx[j][1] = dmul(dmul(p.step_size, w1), f.k20[1]);
            dx[j][2] = dmul(dmul(p.step_size, w2), f.k20[2]);
            int slo, shi;
            safe_range(p, f, w0, w1, w2, lo, hi, slo, shi);
            lane_lo = min(lane_lo, lo); lane_hi = max(lane_hi, hi);
            if (lo <= hi) { lane_mlo = max(lane_mlo, slo); lane_mhi = min(lane_mhi, shi); }
            if (MODE == PVP_ACCUM_FIXED) narrow = narrow && (unsigned long long)val[j] < (1ull << 26);
        }
        const int wlo = __reduce_min_sync(FULL, lane_lo), whi = __reduce_max_sync(FULL, lane_hi);
        if (wlo > whi) continue;
        const int mlo = __reduce_max_sync(FULL, lane_mlo), mhi = __reduce_min_sync(FULL, lane_mhi);
        if (MODE == PVP_ACCUM_FIXED) {
            // a run has at most CAP * NP values: its sum fits 32 bits when every value is below 2^26
            static_assert(CAP * NP <= 64, "32-bit run sums");
            if (__all_sync(FULL, narrow)) {
                unsigned v32[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) v32[j] = (unsigned)val[j];
                my_samples += march_lockstepN<NP, CAP, unsigned>((unsigned long long *)grid, p, f, lm, lr, dx, v32, has, wlo, whi, mlo, mhi);
            } else {
                unsigned long long v64[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) v64[j] = (unsigned long long)val[j];
                my_samples += march_lockstepN<NP, CAP, unsigned long long>((unsigned long long *)grid, p, f, lm, lr, dx, v64, has, wlo, whi, mlo, mhi);
            }
        } else {
            double vd[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) vd[j] = (double)val[j];
            my_samples += march_lockstepN<NP, CAP, double>((double *)grid, p, f, lm, lr, dx, vd, has, wlo, whi, mlo, mhi);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_rays += __shfl_xor_sync(FULL, my_rays, o);
        my_samples += __shfl_xor_sync(FULL, my_samples, o);
    }
    if (lane == 0 && (my_rays | my_samples)) {
        atomicAdd(&ctl->img_rays, my_rays);
        atomicAdd(&ctl->img_samples, my_samples);
    }
}

template <int MODE, int TILE_H>
__global__ void __launch_bounds__(K3L_THREADS)
process_image_lockstep_kernel(const double *__restrict__ image, typename Cell<MODE>::T *__restrict__ grid,
                              typename Cell<MODE>::T *__restrict__ tex, ImageP p, ImageFastP f, Ctl *__restrict__ ctl)
{
    using C = Cell<MODE>;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const ScanMasks lm = scan_masks(lane);
    constexpr int TILE_W = 32 / TILE_H;
    // a warp's 32 pixels form a compact tile of the image, TILE_H rows x 32/TILE_H columns, consecutive lanes being
    // vertical neighbours (the row-group order of K1): the rays of a tile diverge less than those of a 32 x 1 strip
    const long long tiles_x = (p.W + TILE_W - 1) / TILE_W, tiles_y = (p.row1 - p.row0 + TILE_H - 1) / TILE_H;
    const unsigned long long n_groups = (unsigned long long)(tiles_x * tiles_y);
    unsigned long long my_rays = 0, my_samples = 0;
    for (;;) {
        unsigned long long group = 0;
        if (lane == 0) group = atomicAdd(&ctl->head, 1ull);
        group = __shfl_sync(FULL, group, 0);
        if (group >= n_groups) break;
        const long long ty = (long long)(group / (unsigned long long)tiles_x), tx = (long long)group - ty * tiles_x;
        const long long pi = p.row0 + ty * TILE_H + (lane % TILE_H), pj = tx * TILE_W + (lane / TILE_H);
        const bool in_image = pi < p.row1 && pj < p.W;
        const long long px = pi * p.W + pj;
        double w0 = 0, w1 = 0, w2 = 0;
        typename C::V val = 0;
        int lo = INT32_MAX, hi = INT32_MIN;  // empty range
        if (in_image) {
            int s_entry, s_exit;
            bool counted;
            if (pixel_setup<MODE>(image, tex, p, px, w0, w1, w2, val, s_entry, s_exit, counted) && s_entry <= s_exit) { lo = s_entry; hi = s_exit; }
            my_rays += counted;
        }
        const int wlo = __reduce_min_sync(FULL, lo), whi = __reduce_max_sync(FULL, hi);
        if (wlo > whi) continue;
        if (MODE == PVP_ACCUM_FIXED) {
            // run sums of 32 lanes fit 32 bits when every value is below 2^27 (brightness <= 1 at frac_bits <= 27)
            if (__all_sync(FULL, (unsigned long long)val < (1ull << 27)))
                my_samples += march_lockstep<unsigned>((unsigned long long *)grid, p, f, lm, w0, w1, w2, (unsigned)val, lo, hi, wlo, whi);
            else
                my_samples += march_lockstep<unsigned long long>((unsigned long long *)grid, p, f, lm, w0, w1, w2, (unsigned long long)val, lo, hi, wlo, whi);
        } else {
            my_samples += march_lockstep<double>((double *)grid, p, f, lm, w0, w1, w2, (double)val, lo, hi, wlo, whi);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_rays += __shfl_xor_sync(FULL, my_rays, o);
        my_samples += __shfl_xor_sync(FULL, my_samples, o);
    }
    if (lane == 0 && (my_rays | my_samples)) {
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
    PVP_REQUIRE(a->row_begin >= 0 && a->row_end >= 0 && a->row_end <= a->image_height && (a->row_end == 0 || a->row_begin <= a->row_end),
                "pvp_process_image: bad row range [%lld, %lld) for %lld rows", (long long)a->row_begin, (long long)a->row_end, (long long)a->image_height);
    p->row0 = a->row_begin; p->row1 = a->row_end == 0 && a->row_begin == 0 ? a->image_height : a->row_end;
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

// Conditions under which the lock-step kernel reproduces the reference's voxel index bit for bit (see its header).
static bool make_fastp(const ImageP &p, ImageFastP *f)
{
    if (!p.grid_provided) return false;
    unsigned long long span = 0;
    for (int k = 0; k < 3; ++k) {
        volatile double b = p.ext[2 * k + 1] - p.ext[2 * k];  // :94-96, the divisor
        int e = 0;
        const double m = frexp(b, &e);  // b = m * 2^e, m in [0.5, 1)
        if (!(b > 0.0) || !std::isfinite(b) || e < -500 || e > 500) return false;
        if (m == 1.0 - 0x1p-53) return false;  // significand all ones: the one case Markstein's theorem excludes
        if (!(fabs(p.ext[2 * k]) < 0x1p500 && fabs(p.ext[2 * k + 1]) < 0x1p500)) return false;
        if (p.gn[k] < 1 || p.gn[k] > (1ll << 24) || p.gs[k] < 0 || p.gs[k] > 0xffffffffll) return false;  // 2^24: voxel_index()
        f->b[k] = b;
        volatile double y = 1.0 / b;  // correctly rounded (IEEE division, x86-64 SSE2)
        f->y[k] = y;
        f->nd[k] = (double)p.gn[k];
        volatile double kq = (double)p.gn[k] / b;  // RN(n / b); the scaling by 2^20 is exact
        f->k20[k] = kq * 1048576.0;
        f->margin[k] = 32.0 / f->k20[k];
        f->n1[k] = (int)(p.gn[k] - 1);
        f->gs[k] = (unsigned)p.gs[k];
        span += (unsigned long long)(p.gn[k] - 1) * (unsigned long long)p.gs[k];
    }
    // sample_keys2<2>: A K <= 2^49 with A = |o| + max_distance + |ext_lo| + |ext_hi| + extent (bounds every magnitude of the sample position)
    f->affine_ok = std::isfinite(p.max_distance) && p.max_distance >= 0.0;
    for (int k = 0; k < 3; ++k) {
        const double A = fabs(p.o[k]) + p.max_distance + fabs(p.ext[2 * k]) + fabs(p.ext[2 * k + 1]) + f->b[k];
        if (!(A * f->k20[k] <= 0x1p49)) f->affine_ok = 0;
    }
    for (int k = 0; k < 3; ++k) {   // the expression of affine_of, un-fused like the device's (-fmad=false)
        volatile double d = p.o[k] - p.ext[2 * k];
        volatile double m = d * f->k20[k];
        volatile double x0 = m + (0x1p52 - 5.0);
        f->x0[k] = x0;
    }
    return span < 0xffffffffull;  // voxel offsets are 32-bit keys, 0xffffffff is reserved
}

// image_variant: 0 = auto (three-pixel lock-step where fast_ok and affine_ok, two-pixel where only fast_ok), 1 = generic per-pixel kernel,
// 2 = one-pixel lock-step, 3 = two-pixel lock-step, 4 = three-pixel lock-step (the lock-step kernels fall back to the generic one if !fast_ok)
static int launch_image(pvp_ctx *ctx, const double *image_dev, void *grid_dev, void *tex_dev, const ImageP &p, int accum_mode)
{
    const long long rows = p.row1 - p.row0, n_px = p.W * rows;
    if (rows <= 0) return PVP_OK;  // an empty row shard
    LaunchScope scope(ctx, 3);
    ImageFastP f;
    if (ctx->opt_image_variant != 1 && make_fastp(p, &f)) {
        PVP_CUDA(cudaMemsetAsync(&ctx->ctl_dev->head, 0, sizeof(unsigned long long), ctx->stream));
        int per_sm = 0;
        const int th = ctx->opt_image_tile_h == 1 ? 1 : ctx->opt_image_tile_h == 4 ? 4 : ctx->opt_image_tile_h == 8 ? 8 : 2;
        const bool f64 = accum_mode == PVP_ACCUM_F64;
        if ((ctx->opt_image_variant == 4 || ctx->opt_image_variant == 0) && f.affine_ok) {  // default where the one-FMA coordinate is allowed: three pixels per lane
            const void *k4 = f64 ? (const void *)process_image_lockstepN_kernel<PVP_ACCUM_F64> : (const void *)process_image_lockstepN_kernel<PVP_ACCUM_FIXED>;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k4, K3N_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
            const long long groups4 = ((p.W + 31) / 32) * ((rows + PVP_K3N_NP - 1) / PVP_K3N_NP);
            const int blocks4 = (int)std::min<long long>((groups4 + K3N_THREADS / 32 - 1) / (K3N_THREADS / 32), (long long)ctx->sm_count * per_sm);
            void *args4[] = {(void *)&image_dev, (void *)&grid_dev, (void *)&tex_dev, (void *)&p, (void *)&f, (void *)&ctx->ctl_dev};
            PVP_CUDA(cudaLaunchKernel(k4, dim3((unsigned)blocks4), dim3(K3N_THREADS), args4, 0, ctx->stream));
            PVP_CUDA(cudaGetLastError());
            return PVP_OK;
        }
        if (ctx->opt_image_variant != 2) {  // two pixels per lane (image_variant 3; the default where !affine_ok)
            const void *k2 = f64 ? (const void *)process_image_lockstep2_kernel<PVP_ACCUM_F64> : (const void *)process_image_lockstep2_kernel<PVP_ACCUM_FIXED>;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2, K3L_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
            const long long groups2 = ((p.W + 31) / 32) * ((rows + 1) / 2);
            const int blocks2 = (int)std::min<long long>((groups2 + K3L_THREADS / 32 - 1) / (K3L_THREADS / 32), (long long)ctx->sm_count * per_sm);
            void *args2[] = {(void *)&image_dev, (void *)&grid_dev, (void *)&tex_dev, (void *)&p, (void *)&f, (void *)&ctx->ctl_dev};
            PVP_CUDA(cudaLaunchKernel(k2, dim3((unsigned)blocks2), dim3(K3L_THREADS), args2, 0, ctx->stream));
            PVP_CUDA(cudaGetLastError());
            return PVP_OK;
        }
        const void *k = th == 1 ? (f64 ? (const void *)process_image_lockstep_kernel<PVP_ACCUM_F64, 1> : (const void *)process_image_lockstep_kernel<PVP_ACCUM_FIXED, 1>)
                      : th == 4 ? (f64 ? (const void *)process_image_lockstep_kernel<PVP_ACCUM_F64, 4> : (const void *)process_image_lockstep_kernel<PVP_ACCUM_FIXED, 4>)
                      : th == 8 ? (f64 ? (const void *)process_image_lockstep_kernel<PVP_ACCUM_F64, 8> : (const void *)process_image_lockstep_kernel<PVP_ACCUM_FIXED, 8>)
                                : (f64 ? (const void *)process_image_lockstep_kernel<PVP_ACCUM_F64, 2> : (const void *)process_image_lockstep_kernel<PVP_ACCUM_FIXED, 2>);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, K3L_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
        const long long groups = ((p.W + 32 / th - 1) / (32 / th)) * ((rows + th - 1) / th);
        const int blocks = (int)std::min<long long>((groups + K3L_THREADS / 32 - 1) / (K3L_THREADS / 32), (long long)ctx->sm_count * per_sm);
        void *args[] = {(void *)&image_dev, (void *)&grid_dev, (void *)&tex_dev, (void *)&p, (void *)&f, (void *)&ctx->ctl_dev};
        PVP_CUDA(cudaLaunchKernel(k, dim3((unsigned)blocks), dim3(K3L_THREADS), args, 0, ctx->stream));
        PVP_CUDA(cudaGetLastError());
        return PVP_OK;
    }
    int blocks = (int)std::min<long long>((n_px + K3_THREADS - 1) / K3_THREADS, (long long)ctx->sm_count * 16);
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
#ifdef PVP_CHECKED
    {
        unsigned long long spans[2] = {(unsigned long long)std::max(0ll, span_elems(args->grid_shape, args->grid_stride, 3)),
                                       (unsigned long long)std::max(0ll, span_elems(args->tex_shape, args->tex_stride, 2))};
        if (ctx->opt_inject_fault == 2) spans[0] /= 2;   // a test of the check itself
        const unsigned long long zero[2] = {0, 0};
        PVP_CUDA(cudaMemcpyToSymbolAsync(g_img_span_grid, &spans[0], sizeof(unsigned long long), 0, cudaMemcpyHostToDevice, ctx->stream));
        PVP_CUDA(cudaMemcpyToSymbolAsync(g_img_span_tex, &spans[1], sizeof(unsigned long long), 0, cudaMemcpyHostToDevice, ctx->stream));
        PVP_CUDA(cudaMemcpyToSymbolAsync(g_img_bad, zero, sizeof(zero), 0, cudaMemcpyHostToDevice, ctx->stream));
        PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    }
#endif
    if ((rc = launch_image(ctx, image_dev, grid_dev, tex_dev, p, args->accum_mode))) return rc;
#ifdef PVP_CHECKED
    {
        unsigned long long bad[2] = {0, 0};
        PVP_CUDA(cudaStreamSynchronize(ctx->stream));
        PVP_CUDA(cudaMemcpyFromSymbol(bad, g_img_bad, sizeof(bad)));
        if (bad[0] | bad[1]) {
            set_error("pvp_process_image: checked build caught out-of-bounds offsets: voxel grid %llu, texture %llu", bad[0], bad[1]);
            return PVP_ERR_CHECK;
        }
    }
#endif
    return fetch_image_stats(ctx, stats);
}

int pvp_selftest_div_by_extent(pvp_ctx *ctx, double extent, int64_t n_random, int32_t n_grid, uint64_t seed, uint64_t *mismatches,
                               int32_t *eligible)
{
    PVP_REQUIRE(ctx && mismatches && eligible && n_random >= 0 && n_grid >= 1, "pvp_selftest_div_by_extent: bad argument");
    ImageP p = {};
    p.grid_provided = 1;
    for (int k = 0; k < 3; ++k) { p.ext[2 * k] = 0.0; p.ext[2 * k + 1] = extent; p.gn[k] = n_grid; }
    p.gs[0] = (long long)n_grid * n_grid; p.gs[1] = n_grid; p.gs[2] = 1;
    ImageFastP f;
    *eligible = make_fastp(p, &f) ? 1 : 0;
    *mismatches = 0;
    if (!*eligible) return PVP_OK;  // the lock-step kernel would not be used for this extent
    DeviceGuard guard(ctx->device);
    unsigned long long *d_bad = nullptr;
    PVP_CUDA(cudaMalloc(&d_bad, sizeof(*d_bad)));
    cudaMemsetAsync(d_bad, 0, sizeof(*d_bad), ctx->stream);
    ++ctx->prof.launches_other;
    div_selftest_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(f, n_random, n_grid, seed, d_bad);
    cudaError_t e = cudaMemcpyAsync(mismatches, d_bad, sizeof(*d_bad), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_bad);
    PVP_CUDA(e);
    return PVP_OK;
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
