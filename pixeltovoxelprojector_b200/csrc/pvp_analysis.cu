// pvp_analysis.cu -- the consumer side of the .bin: the two data-parallel steps of
// voxelmotionviewer.py's extract_top_percentile_z_up (:56-96) with the grid left in HBM.
//
//   order statistics   np.percentile(flat, q) needs the k-th and (k+1)-th smallest of N^3 floats (:72).
//                      Radix select on the order-preserving 32-bit image of the floats: 4 histogram passes
//                      of 8 bits (each one coalesced read of the grid, HBM-bound: 4 B per cell per pass), the
//                      bucket choice stays on the device, plus one min-reduction pass when the (k+1)-th value
//                      is not a duplicate of the k-th.  The interpolation between the two (numpy's "linear"
//                      method) is two flops and is done by the host mirror with numpy's own formula.
//   extraction         np.argwhere(grid > thresh) in C order + intensities + world coordinates (:77-95): ordered
//                      stream compaction in two passes (per-CTA counts, scan, ranked writes).
//
// NaNs: numpy's percentile of data containing NaN is NaN; a voxel grid produced by this library holds none, and
// the select treats them by their bit pattern (above +inf).
#include <string.h>

#include <algorithm>

#include "pvp_internal.h"

namespace pvp {

__device__ __forceinline__ uint32_t float_key(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone: a < b  <=>  key(a) < key(b); -0 < +0 adjacent
}

struct SelectState {
    unsigned long long hist[256];
    unsigned long long k;       // rank still to be found inside the current prefix bucket
    unsigned long long below;   // cells strictly below the current prefix bucket
    unsigned long long equal;   // after the last pass: cells equal to the selected value
    uint32_t prefix;            // high bits fixed so far
    uint32_t next_key;          // min key strictly above the selected one (0xffffffff if none)
};

constexpr int AN_THREADS = 256;

// histogram of byte `shift/8` over the cells whose higher bytes equal st->prefix
__global__ void __launch_bounds__(AN_THREADS)
select_hist_kernel(const float *__restrict__ x, long long n, int shift, SelectState *__restrict__ st)
{
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t prefix = st->prefix;
    const uint32_t himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
    const long long n4 = n / 4, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const bool vec = ((uintptr_t)x % 16) == 0;
    // voxel grids are mostly zeros: a thread counts runs of one bin in a register and flushes on change, so the
    // shared-memory atomics do not serialise on the zero bin
    unsigned cur = 0xffffffffu, cnt = 0;
    auto take = [&](float f) {
        const uint32_t k = float_key(f);
        if ((k & himask) != prefix) return;
        const unsigned b = (k >> shift) & 0xffu;
        if (b != cur) {
            if (cnt) atomicAdd(&h[cur], cnt);
            cur = b; cnt = 0;
        }
        ++cnt;
    };
    if (vec) {
        for (long long i = tid; i < n4; i += nth) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
            take(v.x); take(v.y); take(v.z); take(v.w);
        }
        for (long long i = n4 * 4 + tid; i < n; i += nth) take(__ldg(x + i));
    } else {
        for (long long i = tid; i < n; i += nth) take(__ldg(x + i));
    }
    if (cnt) atomicAdd(&h[cur], cnt);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// choose the bucket holding rank k, extend the prefix, clear the histogram for the next pass
__global__ void select_pick_kernel(SelectState *st, int shift)
{
    if (threadIdx.x != 0) return;
    unsigned long long k = st->k, acc = 0;
    int b = 0;
    for (; b < 255; ++b) {
        if (acc + st->hist[b] > k) break;
        acc += st->hist[b];
    }
    st->k = k - acc;
    st->below += acc;
    st->prefix |= (uint32_t)b << shift;
    if (shift == 0) st->equal = st->hist[b];
    for (int i = 0; i < 256; ++i) st->hist[i] = 0;
}

__global__ void __launch_bounds__(AN_THREADS)
select_next_kernel(const float *__restrict__ x, long long n, SelectState *__restrict__ st)
{
    const uint32_t sel = st->prefix;
    uint32_t best = 0xffffffffu;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t k = float_key(__ldg(x + i));
        if (k > sel && k < best) best = k;
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0 && best != 0xffffffffu) atomicMin(&st->next_key, best);
}

// ---- ordered extraction ------------------------------------------------------------------------------------
constexpr int EX_PER_THREAD = 8;
constexpr int EX_PER_CTA = AN_THREADS * EX_PER_THREAD;

__global__ void __launch_bounds__(AN_THREADS)
extract_count_kernel(const float *__restrict__ x, long long n, double thresh, unsigned long long *__restrict__ counts)
{
    const long long base = (long long)blockIdx.x * EX_PER_CTA + (long long)threadIdx.x * EX_PER_THREAD;
    int c = 0;
#pragma unroll
    for (int j = 0; j < EX_PER_THREAD; ++j)
        if (base + j < n && (double)__ldg(x + base + j) > thresh) ++c;
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int w[AN_THREADS / 32];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < AN_THREADS / 32; ++i) t += w[i];
        counts[blockIdx.x] = (unsigned long long)t;
    }
}

// exclusive scan of the per-CTA counts in place (one CTA; counts[n_blocks] receives the total)
__global__ void __launch_bounds__(1024)
extract_scan_kernel(unsigned long long *__restrict__ counts, long long n_blocks)
{
    __shared__ unsigned long long part[1024];
    const long long per = (n_blocks + 1023) / 1024, lo = (long long)threadIdx.x * per, hi = min(lo + per, n_blocks);
    unsigned long long s = 0;
    for (long long i = lo; i < hi; ++i) s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < 1024; ++i) { const unsigned long long t = part[i]; part[i] = acc; acc += t; }
        counts[n_blocks] = acc;
    }
    __syncthreads();
    unsigned long long acc = part[threadIdx.x];
    for (long long i = lo; i < hi; ++i) { const unsigned long long t = counts[i]; counts[i] = acc; acc += t; }
}

// voxelmotionviewer.py:85-95: index (i0,i1,i2) -> z = i0, y = i1, x = i2; world = grid_min + (idx + 0.5) * voxel_size
__global__ void __launch_bounds__(AN_THREADS)
extract_write_kernel(const float *__restrict__ x, long long n, int N, double thresh, const unsigned long long *__restrict__ offsets,
                     long long cap, double vs, double gmin0, double gmin1, double gmin2, double *__restrict__ points,
                     float *__restrict__ intens, int *__restrict__ coords)
{
    const long long base = (long long)blockIdx.x * EX_PER_CTA + (long long)threadIdx.x * EX_PER_THREAD;
    float v[EX_PER_THREAD];
    int c = 0;
#pragma unroll
    for (int j = 0; j < EX_PER_THREAD; ++j) {
        v[j] = base + j < n ? __ldg(x + base + j) : 0.f;
        if (base + j < n && (double)v[j] > thresh) ++c;
    }
    // rank of this thread's first hit inside the CTA: warp scan + warp totals
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
    }
    __shared__ int w[AN_THREADS / 32];
    if (lane == 31) w[warp] = incl;
    __syncthreads();
    int before = 0;
    for (unsigned i = 0; i < warp; ++i) before += w[i];
    long long pos = (long long)offsets[blockIdx.x] + before + (incl - c);
#pragma unroll
    for (int j = 0; j < EX_PER_THREAD; ++j) {
        if (base + j < n && (double)v[j] > thresh) {
            if (pos < cap) {
                const long long idx = base + j;
                const int i0 = (int)(idx / ((long long)N * N)), rem = (int)(idx - (long long)i0 * N * N), i1 = rem / N, i2 = rem - i1 * N;
                if (coords) { coords[3 * pos] = i0; coords[3 * pos + 1] = i1; coords[3 * pos + 2] = i2; }
                if (points) {
                    points[3 * pos] = __dadd_rn(gmin0, __dmul_rn(__dadd_rn((double)i2, 0.5), vs));
                    points[3 * pos + 1] = __dadd_rn(gmin1, __dmul_rn(__dadd_rn((double)i1, 0.5), vs));
                    points[3 * pos + 2] = __dadd_rn(gmin2, __dmul_rn(__dadd_rn((double)i0, 0.5), vs));
                }
                if (intens) intens[pos] = v[j];
            }
            ++pos;
        }
    }
}

}  // namespace pvp

using namespace pvp;

extern "C" {

int pvp_grid_order_stats(pvp_ctx *ctx, const float *grid_dev, int64_t n, int64_t k, float *v_k, float *v_k1)
{
    PVP_REQUIRE(ctx && grid_dev && v_k && v_k1, "pvp_grid_order_stats: NULL argument");
    PVP_REQUIRE(n > 0 && k >= 0 && k < n, "pvp_grid_order_stats: need 0 <= k < n (k=%lld, n=%lld)", (long long)k, (long long)n);
    DeviceGuard guard(ctx->device);
    SelectState *st = nullptr;
    PVP_CUDA(cudaMalloc(&st, sizeof(SelectState)));
    SelectState h = {};
    h.k = (unsigned long long)k;
    h.next_key = 0xffffffffu;
    cudaError_t e = cudaMemcpyAsync(st, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream);
    const int blocks = (int)std::min<long long>((n / 4 + AN_THREADS) / AN_THREADS, (long long)ctx->sm_count * 8);
    for (int shift = 24; shift >= 0 && e == cudaSuccess; shift -= 8) {
        ctx->prof.launches_other += 2;
        select_hist_kernel<<<blocks, AN_THREADS, 0, ctx->stream>>>(grid_dev, n, shift, st);
        select_pick_kernel<<<1, 32, 0, ctx->stream>>>(st, shift);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, st, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    // the (k+1)-th smallest: a duplicate of the k-th unless k is the last cell holding that value
    bool need_next = e == cudaSuccess && k + 1 < n && (unsigned long long)(k + 1) >= h.below + h.equal;
    if (need_next) {
        ++ctx->prof.launches_other;
        select_next_kernel<<<blocks, AN_THREADS, 0, ctx->stream>>>(grid_dev, n, st);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h, st, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(st);
    PVP_CUDA(e);
    auto to_float = [](uint32_t key) { uint32_t u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key; float f; memcpy(&f, &u, 4); return f; };
    *v_k = to_float(h.prefix);
    *v_k1 = need_next && h.next_key != 0xffffffffu ? to_float(h.next_key) : *v_k;
    return PVP_OK;
}

int pvp_grid_extract_above(pvp_ctx *ctx, const float *grid_dev, int32_t n_side, double thresh, double voxel_size, const double grid_min[3],
                           int64_t cap, double *points_dev, float *intens_dev, int32_t *coords_dev, int64_t *count)
{
    PVP_REQUIRE(ctx && grid_dev && count, "pvp_grid_extract_above: NULL argument");
    PVP_REQUIRE(n_side > 0 && n_side <= 2048 && cap >= 0, "pvp_grid_extract_above: bad size");
    PVP_REQUIRE(points_dev == nullptr || grid_min != nullptr, "pvp_grid_extract_above: grid_min is NULL");
    DeviceGuard guard(ctx->device);
    const long long n = (long long)n_side * n_side * n_side;
    const long long n_blocks = (n + EX_PER_CTA - 1) / EX_PER_CTA;
    unsigned long long *counts = nullptr;
    PVP_CUDA(cudaMalloc(&counts, sizeof(unsigned long long) * (n_blocks + 1)));
    ctx->prof.launches_other += 2;
    extract_count_kernel<<<(unsigned)n_blocks, AN_THREADS, 0, ctx->stream>>>(grid_dev, n, thresh, counts);
    extract_scan_kernel<<<1, 1024, 0, ctx->stream>>>(counts, n_blocks);
    cudaError_t e = cudaGetLastError();
    unsigned long long total = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, counts + n_blocks, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && cap > 0 && (points_dev || intens_dev || coords_dev)) {
        ++ctx->prof.launches_other;
        const double g0 = grid_min ? grid_min[0] : 0.0, g1 = grid_min ? grid_min[1] : 0.0, g2 = grid_min ? grid_min[2] : 0.0;
        extract_write_kernel<<<(unsigned)n_blocks, AN_THREADS, 0, ctx->stream>>>(grid_dev, n, n_side, thresh, counts, cap, voxel_size, g0, g1, g2,
                                                                                  points_dev, intens_dev, coords_dev);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(counts);
    PVP_CUDA(e);
    *count = (int64_t)total;
    return PVP_OK;
}

}  // extern "C"
