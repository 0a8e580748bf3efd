// pvp_motion.cu -- path A of libpvp_b200.so: the per-frame motion pipeline of ray_voxel.cpp
// (:484-531) as sm_100a kernels.
//
//   K1  motion_compact_kernel   diff / threshold (:236-255, :496-503) fused with compaction of
//                               the motion pixels into an SoA ray queue.  HBM-bound: 2 B/pixel
//                               (u8) read with 16 B vector loads, ~10 B written per motion pixel.
//   K2  dda_accumulate_kernel   persistent warps pull 32-entry chunks from the queue, build the
//                               camera ray (:506-515), run the Amanatides-Woo walk (:274-395) and
//                               accumulate (:523-529) with RED atomics.  Bound by L2 atomic
//                               throughput (grid L2-resident) or HBM sector traffic (not resident).
//
// Stage-level kernels (detect_motion / pixel_rays / cast_rays) expose the same device functions
// for the bit-exact parity tests.
#include <algorithm>
#include <vector>

#include "pvp_device.cuh"
#include "pvp_internal.h"

namespace pvp {

#ifdef PVP_CHECKED
#define PVP_CHECKED_NOTE_PAIRS(ctl) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) (ctl)->n_pairs = gridDim.y; } while (0)
#else
#define PVP_CHECKED_NOTE_PAIRS(ctl) do { } while (0)
#endif

// checked build: is queue slot i inside the queue, and is what it holds a pixel of the frame, a pair of the batch and a number?
__device__ __forceinline__ bool queue_slot_ok(Ctl *ctl, const Queue &q, unsigned long long i) { return PVP_DEV_CHECK(ctl, PVP_CHK_QUEUE_READ, i < q.capacity); }
__device__ __forceinline__ bool queue_entry_ok(Ctl *ctl, uint32_t pix, unsigned pair, float diff, int width, int height)
{
    return PVP_DEV_CHECK(ctl, PVP_CHK_ENTRY, pix < (uint32_t)width * (uint32_t)height && pair < ctl->n_pairs && diff == diff);
}

// ------------------------------------------------------------------------------------------
// K1: diff / threshold / compact
// ------------------------------------------------------------------------------------------

constexpr int K1_THREADS = 256;

template <typename T> struct FrameVec;
template <> struct FrameVec<uint8_t> { static constexpr int PX = 16; using V = uint4; };
template <> struct FrameVec<float> { static constexpr int PX = 4; using V = float4; };

template <typename T>
__device__ __forceinline__ void load_px(const T *__restrict__ p, long long first, long long n_pix, bool vec_ok,
                                        T (&out)[FrameVec<T>::PX])
{
    constexpr int PX = FrameVec<T>::PX;
    if (vec_ok && first + PX <= n_pix) {
        typename FrameVec<T>::V v = __ldg(reinterpret_cast<const typename FrameVec<T>::V *>(p + first));
        memcpy(out, &v, sizeof(v));
    } else {
#pragma unroll
        for (int k = 0; k < PX; ++k) out[k] = (first + k < n_pix) ? __ldg(p + first + k) : T(0);
    }
}

// grid = (ceil(n_pix / (K1_THREADS*PX)), n_pairs).  Queue order: pair-major inside a CTA's
// reservation, pixel order inside a CTA, so 32 consecutive entries are (almost always) 32
// neighbouring pixels of one frame -- coherent rays for K2's warps.
template <typename T>
__global__ void __launch_bounds__(K1_THREADS)
motion_compact_kernel(const T *__restrict__ frames, long long frame_stride, const pvp_pair *__restrict__ pairs,
                      long long n_pix, float threshold, bool vec_ok, Queue q, Ctl *__restrict__ ctl)
{
    constexpr int PX = FrameVec<T>::PX;
    const int pair = blockIdx.y;
    PVP_CHECKED_NOTE_PAIRS(ctl);
    const T *prev = frames + (long long)pairs[pair].prev * frame_stride;
    const T *curr = frames + (long long)pairs[pair].curr * frame_stride;
    const long long first = ((long long)blockIdx.x * K1_THREADS + threadIdx.x) * PX;

    T a[PX], b[PX];
    float d[PX];
    unsigned mask = 0;
    if (first < n_pix) {
        load_px<T>(prev, first, n_pix, vec_ok, a);
        load_px<T>(curr, first, n_pix, vec_ok, b);
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            d[k] = motion_diff(a[k], b[k]);
            if (first + k < n_pix && motion_casts_ray(d[k], threshold)) mask |= 1u << k;
        }
    }
    const int cnt = __popc(mask);

    // CTA-wide exclusive scan of cnt: ballot fast-out, shuffle scan per warp, smem across warps.
    __shared__ unsigned warp_tot[K1_THREADS / 32];
    __shared__ unsigned long long cta_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = cnt;
    if (__ballot_sync(0xffffffffu, cnt != 0) != 0) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
#pragma unroll
        for (int w = 0; w < K1_THREADS / 32; ++w) { unsigned c = warp_tot[w]; warp_tot[w] = tot; tot += c; }
        cta_base = tot ? atomicAdd(&ctl->count, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    if (cnt == 0) return;
    unsigned long long pos = cta_base + warp_tot[warp] + (unsigned)(incl - cnt);
    // the host sizes the queue for the worst case (n_pairs * n_pix), so pos < capacity always
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        if ((mask & (1u << k)) && PVP_DEV_CHECK(ctl, PVP_CHK_QUEUE_WRITE, pos < q.capacity)) {
            q.pix[pos] = (uint32_t)(first + k);
            q.diff[pos] = d[k];
            q.pair[pos] = (uint16_t)pair;
            ++pos;
        }
    }
}

// Row-group ordering: a thread handles 4 columns of R consecutive rows and emits its entries column by column, top pixel
// first, threads in column order -- so 32 (or 64) consecutive queue entries form a compact 2-D tile of the image
// (32/R columns x R rows) instead of a 32 x 1 strip.  The rays of a tile diverge less -- fewer distinct voxels per
// lock-step iteration, more even ray lengths -- and for lean2 entries 2i / 2i+1 are vertical neighbours.  Same filters,
// same entries as motion_compact_kernel; only the order differs.  The CTA's entries are staged in shared memory in
// that order and copied out with coalesced stores (a thread writing its own run of entries would scatter every store
// instruction over 32 lines: 157 us per launch of four 1080p pairs, against 10 B x 8 M entries = 12 us of HBM time).
constexpr int K1T_COLS = 4;                                   // columns per thread: one 32-bit (uint8) / 128-bit (float) load per row
constexpr int K1T_MAX_ENTRIES = 4096;                          // per CTA: 32 KiB of staging
template <int R> struct K1Tile {
    static constexpr int THREADS = (K1T_MAX_ENTRIES / (K1T_COLS * R)) < 256 ? (K1T_MAX_ENTRIES / (K1T_COLS * R)) : 256;
    static constexpr int ITER = R <= 2 ? 4 : R == 4 ? 2 : 1;  // blocks per CTA: 8 rows x 2 frames of loads in flight per thread
};

template <typename T> struct RowSeg;  // 4 pixels of a row as loaded: one 32-bit word (uint8) / one float4
template <> struct RowSeg<uint8_t> { using V = unsigned; };
template <> struct RowSeg<float> { using V = float4; };

template <typename T>
__device__ __forceinline__ typename RowSeg<T>::V load_cols(const T *__restrict__ row, int u0, int width, bool vec_ok)
{
    using V = typename RowSeg<T>::V;
    if (vec_ok && u0 + K1T_COLS <= width) return __ldg(reinterpret_cast<const V *>(row + u0));
    T px[K1T_COLS];
#pragma unroll
    for (int k = 0; k < K1T_COLS; ++k) px[k] = (u0 + k < width) ? __ldg(row + u0 + k) : T(0);
    V v;
    memcpy(&v, px, sizeof(v));
    return v;
}

// A CTA handles ITER consecutive blocks of THREADS segments and issues all their loads first (sparse motion: most blocks are
// still, the kernel is then a pure read of the frames and wants bytes in flight, not many small CTAs).
template <typename T, int R>
__global__ void __launch_bounds__(K1Tile<R>::THREADS)
motion_compact_rowgroup_kernel(const T *__restrict__ frames, long long frame_stride, const pvp_pair *__restrict__ pairs,
                               int width, int height, float threshold, bool vec_ok, bool snake, Queue q, Ctl *__restrict__ ctl)
{
    constexpr int THREADS = K1Tile<R>::THREADS, C = K1T_COLS, ITER = K1Tile<R>::ITER;
    const int pair = blockIdx.y;
    PVP_CHECKED_NOTE_PAIRS(ctl);
    const T *prev = frames + (long long)pairs[pair].prev * frame_stride;
    const T *curr = frames + (long long)pairs[pair].curr * frame_stride;
    const int segs = (width + C - 1) / C, row_groups = (height + R - 1) / R;

    typename RowSeg<T>::V a[ITER][R] = {}, b[ITER][R] = {};
    int rg[ITER], u0[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const unsigned seg_id = (blockIdx.x * ITER + it) * THREADS + threadIdx.x;  // < 2^31: the host bounds width * height below 2^32
        rg[it] = (int)(seg_id / (unsigned)segs);
        u0[it] = (int)(seg_id - (unsigned)rg[it] * (unsigned)segs) * C;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int v = R * rg[it] + r;
            if (rg[it] < row_groups && v < height) {
                a[it][r] = load_cols<T>(prev + (long long)v * width, u0[it], width, vec_ok);
                b[it][r] = load_cols<T>(curr + (long long)v * width, u0[it], width, vec_ok);
            }
        }
    }

    __shared__ unsigned warp_tot[THREADS / 32];
    __shared__ unsigned long long cta_base;
    __shared__ uint32_t s_pix[THREADS * C * R];
    __shared__ float s_diff[THREADS * C * R];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned masks[ITER];  // bit r * C + k: the pixel casts a ray
    unsigned any = 0;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        unsigned mask = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int v = R * rg[it] + r;
            if (rg[it] < row_groups && v < height) {
                T pa[C], pb[C];
                memcpy(pa, &a[it][r], sizeof(pa));
                memcpy(pb, &b[it][r], sizeof(pb));
#pragma unroll
                for (int k = 0; k < C; ++k)
                    if (u0[it] + k < width && motion_casts_ray(motion_diff(pa[k], pb[k]), threshold)) mask |= 1u << (r * C + k);
            }
        }
        masks[it] = mask;
        any |= mask;
    }
    if (!__syncthreads_or(any != 0)) return;  // nothing moves in any of the CTA's blocks: one barrier and out
    bool staged = false;                      // a previous block of this CTA has used the shared arrays (uniform)
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const unsigned mask = masks[it];
        const int cnt = __popc(mask);
        int incl = cnt;
        if (__ballot_sync(0xffffffffu, cnt != 0) != 0) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
        }
        if (staged) __syncthreads();  // the previous block's copy-out has read warp_tot / the staging arrays
        staged = true;
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        unsigned before = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) {
            const unsigned c = warp_tot[w];
            if (w < warp) before += c;
            tot += c;
        }
        if (tot == 0) continue;  // a still block: nothing to reserve, nothing to write (uniform over the CTA)
        if (threadIdx.x == 0) cta_base = atomicAdd(&ctl->count, (unsigned long long)tot);
        unsigned pos = before + (unsigned)(incl - cnt);
#pragma unroll
        for (int k = 0; k < C; ++k) {
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                // snake order (one ray per lane on tall tiles): odd columns bottom-up, so the lanes on either side of a column
                // change are horizontal neighbours too and their runs can merge
                const int r = (snake && (k & 1)) ? R - 1 - rr : rr;
                if ((mask & (1u << (r * C + k))) && PVP_DEV_CHECK(ctl, PVP_CHK_STAGE, pos < (unsigned)(THREADS * C * R))) {
                    T pa[C], pb[C];
                    memcpy(pa, &a[it][r], sizeof(pa));
                    memcpy(pb, &b[it][r], sizeof(pb));
                    s_pix[pos] = (uint32_t)((R * rg[it] + r) * width + u0[it] + k);
                    s_diff[pos] = motion_diff(pa[k], pb[k]);
                    ++pos;
                }
            }
        }
        __syncthreads();
        // the host sizes the queue for the worst case (n_pairs * n_pix), so base + tot <= capacity always
        const unsigned long long base = cta_base;
        for (unsigned i = threadIdx.x; i < tot; i += THREADS) {
            if (!PVP_DEV_CHECK(ctl, PVP_CHK_QUEUE_WRITE, base + i < q.capacity && i < (unsigned)(THREADS * C * R))) continue;
            q.pix[base + i] = s_pix[i];
            q.diff[base + i] = s_diff[i];
            q.pair[base + i] = (uint16_t)pair;
        }
    }
}

// Z-order queue: the order that keeps a warp's rays together at EVERY motion density.  A thread owns a 4 x 8 pixel block and emits
// its motion pixels column by column in snake order (the order that merges best when everything moves: entries 2i / 2i+1 are
// vertical neighbours, which is what lean2 pairs up, and the lanes on either side of a column change are horizontal neighbours);
// the threads of a CTA follow the Z-order (Morton order) of their blocks inside a 64 x 64 pixel footprint (thread bit 0 -> x,
// bit 1 -> y, ...).  So ANY run of consecutive queue entries is a compact 2-D cluster of the image: 64 entries are an 8 x 8 square
// when everything moves, a ~25 x 25 pixel neighbourhood when 10 % of the pixels move -- not a 160 x 4 strip of a row tile.  Rays of
// one neighbourhood visit the same voxels: fewer distinct voxels per lock-step iteration (fewer REDs) and more even ray lengths
// (measured, G voxel-steps/s, row tiles of 4 rows -> Z-order: 10 % iid motion 274 -> 493, 2 % 136 -> 288; dense 842 -> 846).
// A full Morton order INSIDE the blocks as well loses the snake's run merging on dense motion (0.219 REDs per voxel step against
// 0.186: 824 G).  Same filters, same entries as motion_compact_kernel; only the order differs.
// The kernel also counts the motion pixels whose vertical partner (rows 2j / 2j+1 of the block) moves too: the share of rays that
// lean2 can pair decides between lean2 and lean on the device (launch_k2).
// R = rows of a thread's block: 8 (16 x 8 blocks per footprint, 128 threads) or 4 (16 x 16 blocks, 256 threads).  A dense 64-entry
// chunk of lean2 is then an 8 x 8 or a 16 x 4 pixel tile; like the round-1 row tiles the host takes 4 rows while an x-plane of the
// grid stays within 2 MiB (wide, flat tiles merge slightly better: 730 against 711 G voxel-steps/s on a moving object) and 8 rows
// above (neighbouring image columns land in neighbouring x-planes: the RED misses of a wide tile dominate, 1024^3: 457 -> 600 G).
constexpr int K1M_FOOT = 64;
template <int R> struct K1Morton { static constexpr int THREADS = 16 * (K1M_FOOT / R); };

// bits 0, 2, 4, 6 of t -> x (block column, 0..15); bits 1, 3, 5 (, 7) -> y (block row)
__device__ __forceinline__ void morton_block(unsigned t, int &bx, int &by)
{
    bx = (int)((t & 1u) | ((t >> 1) & 2u) | ((t >> 2) & 4u) | ((t >> 3) & 8u));
    by = (int)(((t >> 1) & 1u) | ((t >> 2) & 2u) | ((t >> 3) & 4u) | ((t >> 4) & 8u));
}

template <typename T, int R>
__global__ void __launch_bounds__(K1Morton<R>::THREADS)
motion_compact_morton_kernel(const T *__restrict__ frames, long long frame_stride, const pvp_pair *__restrict__ pairs,
                             int width, int height, float threshold, bool vec_ok, Queue q, Ctl *__restrict__ ctl)
{
    constexpr int THREADS = K1Morton<R>::THREADS, C = K1T_COLS;
    const int pair = blockIdx.y;
    PVP_CHECKED_NOTE_PAIRS(ctl);
    const T *prev = frames + (long long)pairs[pair].prev * frame_stride;
    const T *curr = frames + (long long)pairs[pair].curr * frame_stride;
    const int foot_per_row = (width + K1M_FOOT - 1) / K1M_FOOT;
    const int fy = (int)(blockIdx.x / (unsigned)foot_per_row), fx = (int)(blockIdx.x - (unsigned)fy * (unsigned)foot_per_row);
    int bx, by;
    morton_block(threadIdx.x, bx, by);
    const int u0 = fx * K1M_FOOT + bx * C, v0 = fy * K1M_FOOT + by * R;

    typename RowSeg<T>::V a[R] = {}, b[R] = {};
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int v = v0 + r;
        if (u0 < width && v < height) {
            a[r] = load_cols<T>(prev + (long long)v * width, u0, width, vec_ok);
            b[r] = load_cols<T>(curr + (long long)v * width, u0, width, vec_ok);
        }
    }
    __shared__ unsigned warp_tot[THREADS / 32];
    __shared__ unsigned long long cta_base;
    __shared__ uint32_t s_pix[THREADS * C * R];
    __shared__ float s_diff[THREADS * C * R];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned mask = 0;  // bit r * C + k: the pixel casts a ray
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int v = v0 + r;
        if (u0 < width && v < height) {
            T pa[C], pb[C];
            memcpy(pa, &a[r], sizeof(pa));
            memcpy(pb, &b[r], sizeof(pb));
#pragma unroll
            for (int k = 0; k < C; ++k)
                if (u0 + k < width && motion_casts_ray(motion_diff(pa[k], pb[k]), threshold)) mask |= 1u << (r * C + k);
        }
    }
    if (!__syncthreads_or(mask != 0)) return;  // nothing moves in the CTA's footprint: one barrier and out
    const int cnt = __popc(mask);
    // motion pixels in even rows whose lower neighbour moves too (rows are C bits apart in the mask): each is one pair of lean2
    int paired = __popc(mask & (mask >> C) & 0x0f0f0f0fu);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) paired += __shfl_xor_sync(0xffffffffu, paired, o);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) {
        const unsigned c = warp_tot[w];
        if (w < warp) before += c;
        tot += c;
    }
    if (threadIdx.x == 0) cta_base = atomicAdd(&ctl->count, (unsigned long long)tot);
    if (lane == 0 && paired) atomicAdd(&ctl->n_paired, (unsigned long long)paired);
    unsigned pos = before + (unsigned)(incl - cnt);
    // A footprint in which most pixels move is emitted strip by strip instead (block rows of 16 blocks, left to right): there every
    // block is full, consecutive blocks of a strip are seamless for the snake (a block ends at the top of its last column, the next
    // one starts at the top of its first), and the Z-order's jumps between blocks would only break runs (dense motion: 0.200 -> 0.186
    // REDs per voxel step).  Same entries, other positions: the exclusive prefix is taken over the threads in strip order.
    if (tot * 2 >= (unsigned)(THREADS * C * R)) {   // uniform over the CTA
        __shared__ unsigned s_cnt[THREADS];
        const unsigned rank = (unsigned)(by * 16 + bx);
        __syncthreads();                 // warp_tot has been read by everyone
        s_cnt[rank] = (unsigned)cnt;
        __syncthreads();
        int inc2 = (int)s_cnt[threadIdx.x];
        const int own = inc2;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc2, o);
            if (lane >= o) inc2 += t;
        }
        if (lane == 31) warp_tot[warp] = (unsigned)inc2;
        __syncthreads();
        unsigned before2 = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w)
            if (w < warp) before2 += warp_tot[w];
        s_cnt[threadIdx.x] = before2 + (unsigned)(inc2 - own);   // exclusive prefix of strip position threadIdx.x
        __syncthreads();
        pos = s_cnt[rank];
    }
    // inside the block: column by column, odd columns bottom-up (snake)
#pragma unroll
    for (int k = 0; k < C; ++k) {
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const int r = (k & 1) ? R - 1 - rr : rr;
            if ((mask & (1u << (r * C + k))) && PVP_DEV_CHECK(ctl, PVP_CHK_STAGE, pos < (unsigned)(THREADS * C * R))) {
                T pa[C], pb[C];
                memcpy(pa, &a[r], sizeof(pa));
                memcpy(pb, &b[r], sizeof(pb));
                s_pix[pos] = (uint32_t)(v0 + r) * (uint32_t)width + (uint32_t)(u0 + k);   // < 2^32: the host bounds width * height
                s_diff[pos] = motion_diff(pa[k], pb[k]);
                ++pos;
            }
        }
    }
    __syncthreads();
    // the host sizes the queue for the worst case (n_pairs * n_pix), so base + tot <= capacity always
    const unsigned long long base = cta_base;
    for (unsigned i = threadIdx.x; i < tot; i += THREADS) {
        if (!PVP_DEV_CHECK(ctl, PVP_CHK_QUEUE_WRITE, base + i < q.capacity && i < (unsigned)(THREADS * C * R))) continue;
        q.pix[base + i] = s_pix[i];
        q.diff[base + i] = s_diff[i];
        q.pair[base + i] = (uint16_t)pair;
    }
}

// detect_motion as a stand-alone stage (ray_voxel.cpp:236-255): mask + diff image.
template <typename T>
__global__ void detect_motion_kernel(const T *__restrict__ prev, const T *__restrict__ next, long long n, float threshold,
                                     uint8_t *__restrict__ changed, float *__restrict__ diff)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float d = motion_diff(prev[i], next[i]);
        diff[i] = d;
        changed[i] = (d > threshold) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------
// K2: ray set-up + DDA + accumulate
// ------------------------------------------------------------------------------------------

constexpr int K2_THREADS = 256;

template <int MODE> struct Accum;
template <> struct Accum<PVP_ACCUM_F32> {
    using cell = float;
    using val = float;
    __device__ static __forceinline__ val make(float diff, const GridP &) { return diff; }
    __device__ static __forceinline__ void add(cell *p, val v) { atomicAdd(p, v); }  // RED.E.ADD.F32
};
template <> struct Accum<PVP_ACCUM_FIXED> {
    using cell = unsigned long long;
    using val = unsigned long long;
    __device__ static __forceinline__ val make(float diff, const GridP &g) { return (unsigned long long)quantize_fixed(diff, g.fixed_scale); }
    __device__ static __forceinline__ void add(cell *p, val v) { atomicAdd(p, v); }  // RED.E.ADD.64
};

// Variant 0: one ray per lane, one RED per voxel step.
template <int MODE>
__global__ void __launch_bounds__(K2_THREADS)
dda_accumulate_kernel(Queue q, const pvp_pair *__restrict__ pairs, GridP g, typename Accum<MODE>::cell *__restrict__ grid,
                      int width, int height, Ctl *__restrict__ ctl)
{
    using A = Accum<MODE>;
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long count = ctl->count;
    const int N = g.n;
    const long long sx_stride = (long long)N * N;
    unsigned long long my_steps = 0, my_written = 0;

    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&ctl->head, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
        const unsigned long long i = base + lane;
        if (i < count && queue_slot_ok(ctl, q, i)) {
            const uint32_t pix = __ldg(q.pix + i);
            const float diff = __ldg(q.diff + i);
            const unsigned pair_id = __ldg(q.pair + i);
            if (!queue_entry_ok(ctl, pix, pair_id, diff, width, height)) continue;
            const pvp_pair *P = pairs + pair_id;
            const int v = (int)(pix / (uint32_t)width), u = (int)(pix - (uint32_t)v * (uint32_t)width);
            float dx, dy, dz;
            pixel_ray(u, v, width, height, __ldg(&P->focal), P->R, dx, dy, dz);
            Ray r;
            if (ray_setup(__ldg(&P->cam[0]), __ldg(&P->cam[1]), __ldg(&P->cam[2]), dx, dy, dz, g, r)) {
                const typename A::val val = A::make(diff, g);
                while (r.t_cur <= r.t_end) {
                    ++my_steps;
                    if (r.ix >= g.slab_lo && r.ix < g.slab_hi) {
                        const long long cell = (long long)(r.ix - g.slab_lo) * sx_stride + (long long)r.iy * N + r.iz;
                        if (PVP_DEV_CHECK(ctl, PVP_CHK_GRID, cell >= 0 && cell < (long long)(g.slab_hi - g.slab_lo) * sx_stride && r.iy >= 0 && r.iy < N && r.iz >= 0 && r.iz < N))
                            A::add(grid + cell, g.alpha > 0.f ? A::make(attenuate(diff, g.alpha, r.t_cur), g) : val);
                        ++my_written;
                    }
                    if (!ray_advance(r, N)) break;
                }
            }
        }
    }
    // counters: one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
        my_written += __shfl_xor_sync(0xffffffffu, my_written, o);
    }
    if (lane == 0 && my_steps) {
        atomicAdd(&ctl->n_steps, my_steps);
        atomicAdd(&ctl->n_written, my_written);
        atomicAdd(&ctl->n_atomics, my_written);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->n_rays, count);
}

// ---- warp-aggregated variants --------------------------------------------------------------
// Neighbouring pixels cast nearly identical rays: in lockstep, the 32 lanes of a warp that holds 32
// consecutive motion pixels of one image row touch only ~4.5 distinct voxels per DDA iteration
// (measured on the 1080p -> 512^3 workload).  The SM issues REDs at ~1.5 cycles per LANE
// whether or not lanes collide (tools/atomic_roofline.py), so summing equal-voxel lanes in registers
// first divides the atomic traffic by ~7.  The DDA arithmetic per lane is unchanged (bit-exact visited
// sets); only the order of the additions differs, which is exact for integer-valued addends and for
// fixed point.

// Lock-step state of one lane's ray: branch-free advance with the reference's tie-break
// (ray_voxel.cpp:375-387) and per-axis "steps left before leaving the grid" counters (:389-391).
struct Walk {
    float t_cur, t_end, tmx, tmy, tmz, tdx, tdy, tdz;
    unsigned idx;             // linear voxel offset inside the stored slab
    int dix, diy, diz;        // signed strides of one x / y / z step
    int remx, remy, remz;     // steps left along each axis before the index leaves [0, N)
    int ix, sx;               // x index and x step (slab test only)
};

__device__ __forceinline__ void walk_init(Walk &w, const Ray &r, const GridP &g)
{
    const int N = g.n;
    w.t_cur = r.t_cur; w.t_end = r.t_end;
    w.tmx = r.tmx; w.tmy = r.tmy; w.tmz = r.tmz; w.tdx = r.tdx; w.tdy = r.tdy; w.tdz = r.tdz;
    w.idx = (unsigned)((r.ix - g.slab_lo) * N + r.iy) * (unsigned)N + (unsigned)r.iz;  // wraps harmlessly outside the slab
    w.dix = r.sx * N * N; w.diy = r.sy * N; w.diz = r.sz;
    w.remx = r.sx > 0 ? N - 1 - r.ix : r.ix;
    w.remy = r.sy > 0 ? N - 1 - r.iy : r.iy;
    w.remz = r.sz > 0 ? N - 1 - r.iz : r.iz;
    w.ix = r.ix; w.sx = r.sx;
}

// One DDA step; returns whether the ray is still inside the grid AND t_current <= t_max (:365, :389).
__device__ __forceinline__ bool walk_advance(Walk &w)
{
    const bool cx = (w.tmx < w.tmy) & (w.tmx < w.tmz);
    const bool cy = !cx & (w.tmy < w.tmz);
    const bool cz = !(cx | cy);
    w.t_cur = cx ? w.tmx : (cy ? w.tmy : w.tmz);
    const float nx = fadd(w.tmx, w.tdx), ny = fadd(w.tmy, w.tdy), nz = fadd(w.tmz, w.tdz);
    w.tmx = cx ? nx : w.tmx; w.tmy = cy ? ny : w.tmy; w.tmz = cz ? nz : w.tmz;
    w.idx += (unsigned)(cx ? w.dix : (cy ? w.diy : w.diz));
    w.ix += cx ? w.sx : 0;
    w.remx -= (int)cx; w.remy -= (int)cy; w.remz -= (int)cz;
    return ((w.remx | w.remy | w.remz) >= 0) & (w.t_cur <= w.t_end);
}

template <int MODE> struct AggVal;
template <> struct AggVal<PVP_ACCUM_F32> {
    using T = float;
    __device__ static __forceinline__ T shfl_up(T v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
    __device__ static __forceinline__ T add(T a, T b) { return fadd(a, b); }
};
template <> struct AggVal<PVP_ACCUM_FIXED> {
    using T = unsigned long long;
    __device__ static __forceinline__ T shfl_up(T v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
    __device__ static __forceinline__ T add(T a, T b) { return a + b; }
};

constexpr unsigned KEY_NONE = 0xffffffffu;

// Variant 1: lock-step DDA, runs of equal voxels among consecutive lanes are summed with a segmented
// warp scan (5 shuffle steps, any value type) and the last lane of each run issues one RED.
// Variant 2: groups found with match.any (arbitrary lane sets); integer-valued addends are summed with
// redux.sync, anything else falls back to the scan.  SLAB: the grid pointer holds only x-planes [lo,hi).
template <int MODE, int VARIANT, bool SLAB>
__global__ void __launch_bounds__(K2_THREADS)
dda_accumulate_agg_kernel(Queue q, const pvp_pair *__restrict__ pairs, GridP g, typename Accum<MODE>::cell *__restrict__ grid,
                          int width, int height, Ctl *__restrict__ ctl)
{
    using A = Accum<MODE>;
    using V = AggVal<MODE>;
    const unsigned lane = threadIdx.x & 31;
    const unsigned le_mask = 0xffffffffu >> (31 - lane);
    const unsigned long long count = ctl->count;
    unsigned long long my_steps = 0, my_written = 0, my_atomics = 0;

    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&ctl->head, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
        const unsigned long long i = base + lane;
        bool active = false;
        Walk w;
        typename V::T val = 0;
        if (i < count && queue_slot_ok(ctl, q, i) && queue_entry_ok(ctl, __ldg(q.pix + i), __ldg(q.pair + i), __ldg(q.diff + i), width, height)) {
            const uint32_t pix = __ldg(q.pix + i);
            const float diff = __ldg(q.diff + i);
            const pvp_pair *P = pairs + __ldg(q.pair + i);
            const int v = (int)(pix / (uint32_t)width), u = (int)(pix - (uint32_t)v * (uint32_t)width);
            float dx, dy, dz;
            pixel_ray(u, v, width, height, __ldg(&P->focal), P->R, dx, dy, dz);
            Ray r;
            active = ray_setup(__ldg(&P->cam[0]), __ldg(&P->cam[1]), __ldg(&P->cam[2]), dx, dy, dz, g, r);
            if (active) {
                walk_init(w, r, g);
                val = A::make(diff, g);
                active = w.t_cur <= w.t_end;  // :365 on entry
            }
        }
        while (__any_sync(0xffffffffu, active)) {
            const bool writes = active && (!SLAB || (w.ix >= g.slab_lo && w.ix < g.slab_hi));
            const unsigned key = writes ? w.idx : KEY_NONE;
            typename V::T sum = writes ? val : (typename V::T)0;
            bool issue;
            if (VARIANT == 2) {
                const unsigned grp = __match_any_sync(0xffffffffu, key);
                issue = writes && (lane == (unsigned)(__ffs(grp) - 1));
                if (MODE == PVP_ACCUM_F32) {
                    // integer-valued addends (uint8 frames): exact integer warp reduction
                    const int iv = __float2int_rz((float)sum);
                    const bool integral = ((float)iv == (float)sum);
                    if (__all_sync(0xffffffffu, integral)) {
                        sum = (typename V::T)__reduce_add_sync(grp, iv);
                    } else {  // arbitrary floats: leader gathers its group's values
                        typename V::T acc = 0;
                        unsigned m = grp;
                        for (unsigned todo = __reduce_max_sync(0xffffffffu, (unsigned)__popc(grp)); todo > 0; --todo) {
                            const int src = m ? __ffs(m) - 1 : (int)lane;
                            const typename V::T t = __shfl_sync(0xffffffffu, sum, src);
                            if (m) { acc = V::add(acc, t); m &= m - 1; }
                        }
                        sum = acc;
                    }
                } else {
                    const unsigned long long qv = (unsigned long long)sum;
                    if (__all_sync(0xffffffffu, qv < (1ull << 26))) {
                        sum = (typename V::T)__reduce_add_sync(grp, (unsigned)qv);
                    } else {
                        typename V::T acc = 0;
                        unsigned m = grp;
                        for (unsigned todo = __reduce_max_sync(0xffffffffu, (unsigned)__popc(grp)); todo > 0; --todo) {
                            const int src = m ? __ffs(m) - 1 : (int)lane;
                            const typename V::T t = __shfl_sync(0xffffffffu, sum, src);
                            if (m) { acc = V::add(acc, t); m &= m - 1; }
                        }
                        sum = acc;
                    }
                }
            } else {
                const unsigned prev_key = __shfl_up_sync(0xffffffffu, key, 1);
                const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev_key);
                const int my_head = 31 - __clz(heads & le_mask);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const typename V::T t = V::shfl_up(sum, d);
                    if ((int)lane - d >= my_head) sum = V::add(sum, t);
                }
                issue = writes && (lane == 31 || ((heads >> (lane + 1)) & 1u));
            }
            if (issue && PVP_DEV_CHECK(ctl, PVP_CHK_GRID, key < (unsigned)(g.slab_hi - g.slab_lo) * (unsigned)g.n * (unsigned)g.n)) { A::add(grid + key, sum); ++my_atomics; }
            my_steps += active;
            my_written += writes;
            if (active) active = walk_advance(w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
        my_written += __shfl_xor_sync(0xffffffffu, my_written, o);
        my_atomics += __shfl_xor_sync(0xffffffffu, my_atomics, o);
    }
    if (lane == 0 && my_steps) {
        atomicAdd(&ctl->n_steps, my_steps);
        atomicAdd(&ctl->n_written, my_written);
        atomicAdd(&ctl->n_atomics, my_atomics);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->n_rays, count);
}

// ---- variant 4: "lean" lock-step kernel ---------------------------------------------------------
// Same algorithm as variant 1 (per-lane DDA in lock-step, runs of equal voxels among consecutive lanes
// summed by a segmented warp scan, one RED per run) with the inner loop cut from 78 to 48 SASS
// instructions, 5 instead of 6 shuffles, and re-balanced between the ALU and FMA pipes (ncu on variant 1:
// ALU pipe 84 %, FMA 16 %).  Measured 1080p -> 512^3: 514 G voxel-steps/s against 361 (profiles/):
//   * axis choice through one FMNMX3: m = min(tmx,tmy,tmz); z iff tmz == m, else y iff tmy == m, else x.
//     For NaN-free t values this is exactly the reference's strict-< chain with ties to the later axis
//     (ray_voxel.cpp:375-387).  The only way t_max += t_delta can make a NaN is a -inf t_max, i.e. an overflow
//     of (boundary - camera) / dir with |dir| >= 1e-12: the host only selects this kernel when every camera and
//     grid coordinate of the batch is below 1e25 in magnitude (lean_geometry_ok); anything else runs variant 3,
//     which keeps the reference's compare structure.
//   * no per-step bookkeeping of the grid bounds: one t limit per ray (axis_limit) below which a step provably neither
//     leaves the grid nor passes t_end -- one compare per step; the reference's exact tests run in the slow path only
//     (step_leaves).  The first version kept three float "steps left" counters per ray: 6 predicated FADDs, a 3-input
//     min and two compares per step.  A ray's step count is the chunk's iteration counter when it parks.
//   * no per-iteration select of key/value: a lane whose ray ended is parked in a state that keeps it
//     harmless (key NONE, value 0, zero strides, t frozen) by a warp-uniform slow path that runs only in
//     iterations where some ray ends (~5 %), which also decides the loop exit.
//   * scan predicates are single LOP3s of the run-head ballot against lane-constant window masks.
//   * software pipelining: the DDA step runs first and the shuffle of the NEW keys (next iteration's run
//     heads) is issued before the scan of the current voxel, off the scan's dependent chain; the first scan
//     step needs no shuffle at all because the value of the lane below is constant while no lane parks.
struct LaneMasks {
    unsigned w1, w2, w4, w8, w16, next;
};

// CAP < 32 (8 or 16): runs are cut at every CAP-th lane, so the scan needs log2(CAP) levels only.  The cut costs no instruction: bit 0 of
// the head ballot is always set, so a window that contains a multiple of CAP simply gets bit 0 as well (its test is then always "head
// inside"), and the lane below a cut issues through the same bit.
template <int CAP = 32>
__device__ __forceinline__ LaneMasks lane_masks(unsigned lane)
{
    LaneMasks m;
    // window_d = lanes (lane-d+1 .. lane); a head inside it means lane-d belongs to another run.  For
    // lane < d the window reaches bit 0, which is always a head, so the predicate is false as required.
    auto win = [lane](unsigned d) { return lane + 1 >= d ? (((1u << d) - 1u) << (lane + 1 - d)) : ((2u << lane) - 1u); };
    m.w1 = 1u << lane; m.w2 = win(2); m.w4 = win(4); m.w8 = win(8);
    asm volatile("" : "+r"(m.w1));  // opaque: otherwise "heads & w1" is compiled as ((heads >> lane) & 1) != 1, three instructions instead of one LOP3
    m.w16 = lane + 1 >= 16 ? (0xffffu << (lane - 15)) : ((2u << lane) - 1u);
    m.next = lane == 31 ? 1u : (1u << (lane + 1));  // lane 31 always ends a run: bit 0 is always set
    if (CAP < 32) {
        const unsigned in_group = lane % CAP;  // a window of d lanes reaches a cut iff d > in_group
        if (1 > in_group) m.w1 |= 1u;
        if (2 > in_group) m.w2 |= 1u;
        if (4 > in_group) m.w4 |= 1u;
        if (8 > in_group) m.w8 |= 1u;
        if (16 > in_group) m.w16 |= 1u;
        if (in_group == CAP - 1) m.next |= 1u;
        // opaque: under register pressure the compiler otherwise re-derives the masks from the lane number inside the loop (27 extra instructions)
        asm volatile("" : "+r"(m.w1), "+r"(m.w2), "+r"(m.w4), "+r"(m.next));
    }
    return m;
}

__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <int MODE, bool NARROW> struct LeanVal;
template <bool NARROW> struct LeanVal<PVP_ACCUM_F32, NARROW> {
    using T = float;
    __device__ static __forceinline__ T make(float diff, const GridP &) { return diff; }
    __device__ static __forceinline__ T add(T a, T b) { return fadd(a, b); }
    __device__ static __forceinline__ float cell_value(T v, const GridP &) { return v; }
};
template <> struct LeanVal<PVP_ACCUM_FIXED, false> {
    using T = unsigned long long;
    __device__ static __forceinline__ T make(float diff, const GridP &g) { return (unsigned long long)quantize_fixed(diff, g.fixed_scale); }
    __device__ static __forceinline__ T add(T a, T b) { return a + b; }
    __device__ static __forceinline__ unsigned long long cell_value(T v, const GridP &) { return v; }
};
// NARROW ("integral"): uint8 frames -- every |prev - curr| is an integer <= 255, so its fixed-point image is diff << frac_bits
// exactly and a run sum of up to 64 of them (< 2^14) is exact in fp32.  The scan then runs on the plain diffs with FADDs
// (FMA pipe, like the float32 mode) and only the lane that issues the RED converts: (u64)sum << frac_bits.
template <> struct LeanVal<PVP_ACCUM_FIXED, true> {
    using T = float;
    __device__ static __forceinline__ T make(float diff, const GridP &) { return diff; }
    __device__ static __forceinline__ T add(T a, T b) { return fadd(a, b); }
    __device__ static __forceinline__ unsigned long long cell_value(T v, const GridP &g) { return (unsigned long long)__float2uint_rz(v) << g.frac_shift; }
};


// red.global.add with the issue decision as a PTX predicate (the address is formed outside the predicate)
__device__ __forceinline__ void red_if(bool p, float *addr, float v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q red.global.add.f32 [%1], %2;\n\t}" ::"r"((unsigned)p), "l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_if(bool p, unsigned long long *addr, unsigned long long v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q red.global.add.u64 [%1], %2;\n\t}" ::"r"((unsigned)p), "l"(addr), "l"(v) : "memory");
}

// key of the lane below (lane 0: none) and whether this lane starts a run, from one SHFL.UP and its range predicate
__device__ __forceinline__ bool run_head(unsigned key)
{
    unsigned prev, inrange;
    asm volatile("{\n\t.reg .pred q;\n\tshfl.sync.up.b32 %0|q, %2, 1, 0, 0xffffffff;\n\tselp.u32 %1, 1, 0, q;\n\t}" : "=r"(prev), "=r"(inrange) : "r"(key));
    return !inrange || key != prev;
}

#ifndef PVP_LEAN_THREADS
#define PVP_LEAN_THREADS 256  // measured at 1080p -> 512^3: 256-thread CTAs 503 G steps/s, 128 and 64 threads 489 G (same resident warps)
#endif
#ifndef PVP_RUN_CAP
#define PVP_RUN_CAP 32  // longest run of lanes that the lean2 / lean4 scans merge (32 = no cut; 8 / 16: fewer scan levels, a few more REDs)
#endif
#ifndef PVP_LEAN2_MINB
#define PVP_LEAN2_MINB 3  // resident CTAs of the default lean2 instantiation (80 registers at 256 threads)
#endif
constexpr int LEAN_THREADS = PVP_LEAN_THREADS;  // warps are independent; the CTA size only sets how registers divide an SM
constexpr int LEAN_CTAS_A = 7 * 128 / LEAN_THREADS > 0 ? 7 * 128 / LEAN_THREADS : 1, LEAN_CTAS_B = 6 * 128 / LEAN_THREADS, LEAN_CTAS_C = 8 * 128 / LEAN_THREADS;

// The step that takes an axis out of [0, N) is its (rem+1)-th, taken when t_max of that axis -- tm0 after `rem` rounded
// additions of td -- is the minimum.  A rigorous LOWER bound of that value: each addition errs by at most 2^-24 of a
// partial sum, all partial sums lie within |tm0| + rem*td, rem < 2^11, so the iterated sum is within 2^-13 (|tm0| + rem*td)
// of tm0 + rem*td; the slack used is 2^-11 of it.  While the step's t (the minimum of the three t_max) is below the bound of
// every axis and below t_end, the ray provably stays inside and alive -- no per-step bookkeeping; from the bound on (the
// last step or two of a ray) the exact test runs in the warp-uniform slow path (step_leaves).
__device__ __forceinline__ float axis_limit(float tm0, float td, int rem)
{
    if (!(td < CUDART_INF_F)) return CUDART_INF_F;  // parallel axis (safe_div, :266-272): never chosen before the others run out
    const float span = fmul((float)rem, td);
    const float bound = fsub(fadd(tm0, span), fmul(fadd(fabsf(tm0), span), 0x1p-11f));
    return bound == bound ? bound : -CUDART_INF_F;  // overflow (inf - inf): no proof, every step of this ray takes the exact test
}

// x / d for d >= 2 with m = floor(2^32 / d): umulhi(x, m) is the quotient or one below it (x * m / 2^32 > x / d - 1), one
// correction step; d == 1 has m = 2^32, which does not fit, and is handled by the caller's geometry (N == 1: every offset is 0).
__device__ __forceinline__ unsigned udiv_by(unsigned x, unsigned d, unsigned m)
{
    if (d == 1) return x;
    unsigned q = __umulhi(x, m);
    if (x - q * d >= d) ++q;
    return q;
}

// Does the step along the chosen axis (cz / cy / else x) leave [0, N)?  Decodes the voxel the step starts from.  Slab kernels
// hold x relative to the slab (the slab's cells number <= 2^31, so the offset is a valid int; the floor division also takes a negative one).
template <bool SLAB>
__device__ __forceinline__ bool step_leaves(unsigned cur, bool cz, bool cy, int dix, int diy, int diz, const GridP &g)
{
    const int N = g.n, NN = N * N;
    int ix;
    unsigned rest;
    if (SLAB) {
        const int rel = (int)cur;
        const int q = rel >= 0 ? (int)udiv_by((unsigned)rel, (unsigned)NN, g.mul_nn) : -(int)udiv_by((unsigned)(NN - 1 - rel), (unsigned)NN, g.mul_nn);  // floor
        ix = q + g.slab_lo;
        rest = (unsigned)(rel - q * NN);
    } else {
        ix = (int)udiv_by(cur, (unsigned)NN, g.mul_nn);
        rest = cur - (unsigned)ix * (unsigned)NN;
    }
    const int iy = (int)udiv_by(rest, (unsigned)N, g.mul_n), iz = (int)(rest - (unsigned)iy * (unsigned)N);
    if (cz) return diz > 0 ? iz == N - 1 : iz == 0;
    if (cy) return diy > 0 ? iy == N - 1 : iy == 0;
    // an x step: out of the grid, or -- slab kernels -- past the far face of the slab: x only moves one way, the ray cannot
    // write into this slab again, so this rank stops walking it (per-slab clipping that keeps the walk itself untouched)
    if (SLAB) return dix > 0 ? ix + 1 >= g.slab_hi : ix - 1 < g.slab_lo;
    return dix > 0 ? ix == N - 1 : ix == 0;
}

// Slab kernels: can the ray write into x-planes [slab_lo, slab_hi) at all, and how many x steps until it has passed them?
// x moves one way (sx); the last x index the walk can reach is within one voxel of the box exit point, estimated with two
// voxels of margin.  Returns false when the ray need not be walked by this rank; else rem_x = x steps that stay on this side of
// the slab's far face (the step after them ends the walk here: step_leaves).
__device__ __forceinline__ bool slab_reach(const Ray &r, float cam_x, float dir_x, const GridP &g, int &rem_x)
{
    const float x_exit = (cam_x + r.t_end * dir_x - g.gmin[0]) / g.vs;  // in voxels; inf / NaN compare false below: not skipped
    if (r.sx > 0) {
        if (r.ix >= g.slab_hi || x_exit + 2.f < (float)g.slab_lo) return false;
        rem_x = g.slab_hi - 1 - r.ix;
    } else {
        if (r.ix < g.slab_lo || x_exit - 2.f >= (float)g.slab_hi) return false;
        rem_x = r.ix - g.slab_lo;
    }
    return true;
}

// The device-side choice between the two lock-step kernels of a batch (both are launched; the other one exits at once).
// min_rays: lean2 needs that many rays (its 64-ray chunks leave SMs idle on small batches); paired_pct: ... and at least that share
// of the rays in vertical pairs (counted by K1), i.e. locally dense motion -- with scattered motion the second ray of a lane rarely
// shares the first one's voxel, every lane issues two REDs and lean does the same work better (10 % iid motion: 493 against 444 G
// voxel-steps/s; a moving object covering 3.6 % of the frame is locally dense: 730 against 566 G the other way round).
// lean4 (four rays per lane, 128-ray chunks) takes the batches that lean2 would take from min_rays4 rays up (float32 cells, small planes:
// the host passes ~0 otherwise): dense 1080p 966 against 829 G voxel-steps/s, 60 % iid motion 855 against 794, but a moving object of
// 75 k rays per pair 685 against 706.  Returns 1 (lean), 2 (lean2) or 4 (lean4).
// Forced variants: (min_rays, paired_pct, min_rays4) = (~0, 0, ~0) lean, (0, 0, ~0) lean2, (0, 0, 0) lean4.  (count == 0 falls to lean2 or
// lean4 under the forced settings and to lean otherwise; with an empty queue it does not matter.)
__device__ __forceinline__ int pick_lockstep(const Ctl *ctl, unsigned long long count, unsigned long long min_rays, unsigned long long paired_pct,
                                             unsigned long long min_rays4)
{
    if (count < min_rays) return 1;
    if (paired_pct != 0 && 200ull * ctl->n_paired < paired_pct * count) return 1;
    return count >= min_rays4 ? 4 : 2;
}

// One thread, between K1 and the lock-step K2 kernels of a batch: hands the queue to the kernel pick_lockstep names (Ctl::sel_count) and
// resets the three chunk heads.  The others see an empty queue.
__global__ void select_lockstep_kernel(Ctl *__restrict__ ctl, unsigned long long min_rays, unsigned long long paired_pct, unsigned long long min_rays4)
{
    const unsigned long long count = ctl->count;
    const int pick = pick_lockstep(ctl, count, min_rays, paired_pct, min_rays4);
    ctl->sel_count[0] = pick == 1 ? count : 0ull;
    ctl->sel_count[1] = pick == 2 ? count : 0ull;
    ctl->sel_count[2] = pick == 4 ? count : 0ull;
    ctl->sel_head[0] = ctl->sel_head[1] = ctl->sel_head[2] = 0ull;
}

// Slab kernels: FAST-FORWARD a ray that starts outside the slab to the moment it enters it, without walking the voxels in between
// and without changing a single rounding of the reference's walk.  The walk is a 3-way merge of three independent sequences: axis
// a's t_max takes the values S_a[0] = t_max_a, S_a[j+1] = fl(S_a[j] + t_delta_a) whatever the other axes do, and the loop always
// steps the axis with the smallest head, ties to the later axis (ray_voxel.cpp:375-387: x only if strictly below y and z, y only if
// strictly below z).  Hence the k-th x step, taken at t = X = S_x[k-1], comes exactly after every y step with S_y[j] <= X and every z
// step with S_z[j] <= X -- three short scalar loops (one rounded add and one compare per skipped voxel step instead of a full
// lock-step iteration) give the state at the slab's near face: voxel, the three t_max, t_current = X.  The ray never gets there if
// X > t_end (the loop condition :365 fails first: the chosen t values are non-decreasing, X is the largest before the entry) or if y
// or z would have left [0, N) before (:388-391).  x moves one way along a ray, so "entering the slab" is its k-th x step with
// k = distance from the start plane to the slab's near plane.  rem_x becomes the x steps that stay inside the slab.
__device__ __forceinline__ bool slab_enter(Ray &r, float cam_x, float dir_x, const GridP &g, int &rem_x)
{
    if (!slab_reach(r, cam_x, dir_x, g, rem_x)) return false;
    const int kx = r.sx > 0 ? g.slab_lo - r.ix : r.ix - (g.slab_hi - 1);
    if (kx <= 0) return true;  // starts inside the slab
    float X = r.tmx;
    for (int j = 1; j < kx; ++j) X = fadd(X, r.tdx);
    if (!(X <= r.t_end)) return false;  // ends before it reaches the slab (also: x never steps, t_max_x = inf)
    const int N = g.n;
    const int rem_y = r.sy > 0 ? N - 1 - r.iy : r.iy, rem_z = r.sz > 0 ? N - 1 - r.iz : r.iz;
    int ky = 0, kz = 0;
    float y = r.tmy, z = r.tmz;
    while (y <= X) {
        if (++ky > rem_y) return false;  // that y step leaves the grid: the walk ends before the slab
        y = fadd(y, r.tdy);
    }
    while (z <= X) {
        if (++kz > rem_z) return false;
        z = fadd(z, r.tdz);
    }
    r.ix += r.sx * kx; r.iy += r.sy * ky; r.iz += r.sz * kz;
    r.tmx = fadd(X, r.tdx); r.tmy = y; r.tmz = z;
    r.t_cur = X;
    rem_x -= kx;
    return true;
}

template <int MODE, bool SLAB, bool NARROW, int MINB>
__global__ void __launch_bounds__(LEAN_THREADS, MINB)
dda_accumulate_lean_kernel(Queue q, const pvp_pair *__restrict__ pairs, GridP g, typename Accum<MODE>::cell *__restrict__ grid,
                           int width, int height, Ctl *__restrict__ ctl)
{
    using A = Accum<MODE>;
    using V = LeanVal<MODE, NARROW>;
    using T = typename V::T;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const LaneMasks lm = lane_masks(lane);
    const unsigned long long count = ctl->sel_count[0];  // the whole queue, or 0 when another lock-step kernel takes the batch (select_lockstep_kernel)
    const int N = g.n;
    const unsigned slab_cells = (unsigned)(g.slab_hi - g.slab_lo) * (unsigned)N * (unsigned)N;
    unsigned steps32 = 0, written32 = 0, atomics32 = 0;  // per lane; see lean2

    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&ctl->sel_head[0], 32ull);
        base = __shfl_sync(FULL, base, 0);
        if (base >= count) break;
        const unsigned long long i = base + lane;

        // parked state (also the state of lanes past the end of the queue and of rays that miss the grid)
        float t_end = CUDART_INF_F, lim = CUDART_INF_F, tmx = 0.f, tmy = 0.f, tmz = 0.f, tdx = 0.f, tdy = 0.f, tdz = 0.f;
        float walked = 0.f;  // iterations of this chunk so far == voxels emitted by every ray that is still alive
        unsigned idx = KEY_NONE;
        int dix = 0, diy = 0, diz = 0;
        T val = 0;
        float n_issued = 0.f, n_inslab = 0.f;

        if (i < count && queue_slot_ok(ctl, q, i) && queue_entry_ok(ctl, __ldg(q.pix + i), __ldg(q.pair + i), __ldg(q.diff + i), width, height)) {
            const uint32_t pix = __ldg(q.pix + i);
            const float diff = __ldg(q.diff + i);
            const pvp_pair *P = pairs + __ldg(q.pair + i);
            const int v = (int)(pix / (uint32_t)width), u = (int)(pix - (uint32_t)v * (uint32_t)width);
            float dx, dy, dz;
            pixel_ray(u, v, width, height, __ldg(&P->focal), P->R, dx, dy, dz);
            Ray r;
            int rem_x = 0;
            if (ray_setup(__ldg(&P->cam[0]), __ldg(&P->cam[1]), __ldg(&P->cam[2]), dx, dy, dz, g, r) && r.t_cur <= r.t_end &&
                (!SLAB || slab_enter(r, __ldg(&P->cam[0]), dx, g, rem_x))) {
                if (!SLAB) rem_x = r.sx > 0 ? N - 1 - r.ix : r.ix;
                t_end = r.t_end;
                tmx = r.tmx; tmy = r.tmy; tmz = r.tmz; tdx = r.tdx; tdy = r.tdy; tdz = r.tdz;
                idx = (unsigned)((r.ix - g.slab_lo) * N + r.iy) * (unsigned)N + (unsigned)r.iz;  // negative x offsets wrap past slab_cells
                dix = r.sx * N * N; diy = r.sy * N; diz = r.sz;
                lim = fminf(fminf(axis_limit(r.tmx, r.tdx, rem_x), axis_limit(r.tmy, r.tdy, r.sy > 0 ? N - 1 - r.iy : r.iy)),
                            fminf(axis_limit(r.tmz, r.tdz, r.sz > 0 ? N - 1 - r.iz : r.iz), r.t_end));
                val = V::make(diff, g);
            }
        }
        if (__all_sync(FULL, tdx == 0.f)) continue;  // (a live ray has t_delta > 0; not "idx == KEY_NONE": in slab kernels a live voxel offset can wrap to it)

        // Loop-carried across iterations so that they are off the scan's critical path: the run heads of the
        // CURRENT voxel keys (computed right after the previous DDA step) and, for whole-grid kernels, the value
        // of the lane below (constant while no lane parks).
        unsigned heads = 0;
        T pval = 0;
        if (!SLAB) {
            heads = __ballot_sync(FULL, run_head(idx));  // bit 0 is always set
            pval = __shfl_up_sync(FULL, val, 1);
        }

        for (;;) {
            // ---- one DDA step (ray_voxel.cpp:375-391), first: the new keys' shuffle overlaps the scan below ----
            const unsigned cur = idx;
            const float m = fmin3(tmx, tmy, tmz);
            const bool cz = (tmz == m);
            const bool cy = !cz && (tmy == m);
            const bool cx = !cz && (tmy != m);
            if (cx) tmx = fadd(tmx, tdx);
            if (cy) tmy = fadd(tmy, tdy);
            if (cz) tmz = fadd(tmz, tdz);
            idx += (unsigned)(cz ? diz : (cy ? diy : dix));
            walked += 1.f;
            const bool sure = m < lim;  // below the ray's limit the step can neither leave the grid nor pass t_end (axis_limit)
            bool next_head = false;
            if (!SLAB) next_head = run_head(idx);
            // ---- accumulate the voxel the step left: one RED per run of equal keys -------------------------
            unsigned key = cur;
            T sum, t;
            if (SLAB) {
                const bool in = cur < slab_cells;  // parked lanes (KEY_NONE) are outside every slab
                key = in ? cur : KEY_NONE;
                sum = in ? val : (T)0;
                if (in) n_inslab += 1.f;
                heads = __ballot_sync(FULL, run_head(key));
                t = __shfl_up_sync(FULL, sum, 1);
                if (!(heads & lm.w1)) sum = V::add(sum, t);
            } else {
                sum = (heads & lm.w1) ? val : V::add(val, pval);
            }
            t = __shfl_up_sync(FULL, sum, 2);  if (!(heads & lm.w2))  sum = V::add(sum, t);
            t = __shfl_up_sync(FULL, sum, 4);  if (!(heads & lm.w4))  sum = V::add(sum, t);
            t = __shfl_up_sync(FULL, sum, 8);  if (!(heads & lm.w8))  sum = V::add(sum, t);
            t = __shfl_up_sync(FULL, sum, 16); if (!(heads & lm.w16)) sum = V::add(sum, t);
            const bool issue = (heads & lm.next) && key != KEY_NONE;
            red_if(issue && PVP_DEV_CHECK(ctl, PVP_CHK_GRID, key < slab_cells), grid + key, V::cell_value(sum, g));
            if (issue) n_issued += 1.f;
            if (!SLAB) heads = __ballot_sync(FULL, next_head);
            // ---- rays that ended: park the lane; leave when the whole warp is parked ---------------------
            if (__any_sync(FULL, !sure)) {
                // exact end-of-ray test (:388-391, :365) for the lanes at their limit.  (live is "tdx != 0", not "idx != KEY_NONE": a
                // ray that leaves the grid downwards from linear offset N^2-1 / N-1 / 0 wraps to exactly 0xffffffff)
                const bool park = !sure && tdx != 0.f && (step_leaves<SLAB>(cur, cz, cy, dix, diy, diz, g) || !(m <= t_end));
                steps32 += park ? (unsigned)walked : 0u;
                idx = park ? KEY_NONE : idx; val = park ? (T)0 : val;
                dix = park ? 0 : dix; diy = park ? 0 : diy; diz = park ? 0 : diz;
                tmx = park ? 0.f : tmx;  // the step's t of a parked lane stays at 0 < lim = inf; tmy, tmz, t_end are never read again
                tdx = park ? 0.f : tdx; tdy = park ? 0.f : tdy; tdz = park ? 0.f : tdz;
                lim = park ? CUDART_INF_F : lim;
                if (__all_sync(FULL, tdx == 0.f)) break;
                if (!SLAB) {
                    heads = __ballot_sync(FULL, run_head(idx));
                    pval = __shfl_up_sync(FULL, val, 1);
                }
            }
        }
        atomics32 += (unsigned)n_issued;
        if (SLAB) written32 += (unsigned)n_inslab;

    }
    unsigned long long my_steps = steps32, my_written = SLAB ? written32 : steps32, my_atomics = atomics32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_steps += __shfl_xor_sync(FULL, my_steps, o);
        my_written += __shfl_xor_sync(FULL, my_written, o);
        my_atomics += __shfl_xor_sync(FULL, my_atomics, o);
    }
    if (lane == 0 && my_steps) {
        atomicAdd(&ctl->n_steps, my_steps);
        atomicAdd(&ctl->n_written, my_written);
        atomicAdd(&ctl->n_atomics, my_atomics);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->n_rays, count);
}

// ---- variant 5: "lean2", two rays per lane -------------------------------------------------------------------------
// A warp takes 64 consecutive queue entries; lane l walks entries 2l (ray A) and 2l+1 (ray B), which the tile
// ordering of K1 makes vertical neighbours (u,v) / (u,v+1).  Both rays advance one DDA step per iteration.  B's voxel
// is A's voxel ~89 % of the time: then B's value simply joins A's before the run scan, so ONE scan and ONE set of run
// REDs serve 64 voxel steps instead of 32 -- the warp's rays form a 16 x 4 (32 x 2, 8 x 8) tile, and a tile touches fewer
// distinct voxels than a 64 x 1 strip.  Where B sits in another voxel it issues its own RED.  Per ray the arithmetic, the step
// order and the counters are those of variant 4 (same bit-exact visited sets).
struct RayS {
    float t_end, lim;  // box exit t_max (:293); lim: below it a step can neither leave the grid nor pass t_end (see axis_limit)
    float tmx, tmy, tmz, tdx, tdy, tdz;
    float t_cur;  // RayStep::distance of the current voxel (only read by the attenuating kernels)
    unsigned idx;
    int dix, diy, diz;
};

// A live ray has positive (or infinite) t_delta on every axis; a parked lane has zeros.
__device__ __forceinline__ bool rays_live(const RayS &r) { return r.tdx != 0.f; }

__device__ __forceinline__ void rays_park(RayS &r)
{
    r.t_end = CUDART_INF_F; r.lim = CUDART_INF_F; r.tmx = r.tmy = r.tmz = 0.f; r.tdx = r.tdy = r.tdz = 0.f;
    r.t_cur = 0.f;
    r.idx = KEY_NONE; r.dix = r.diy = r.diz = 0;
}
// parking a ray that has been walking: only what keeps it silent and "sure" for the rest of the chunk (the step's t stays
// at 0 < lim = inf, the offset at KEY_NONE); t_end, tmy, tmz are never read again for a parked ray
__device__ __forceinline__ void rays_retire(RayS &r)
{
    r.lim = CUDART_INF_F; r.tmx = 0.f; r.tdx = r.tdy = r.tdz = 0.f; r.t_cur = 0.f;
    r.idx = KEY_NONE; r.dix = r.diy = r.diz = 0;
}

// queue entry -> walking state; returns the value to accumulate (0 when the ray casts no step)
// SIGNS: dix / diy / diz hold the step signs (+-1) instead of the offset strides; the step multiplies them by n^2 / n / 1 (rays_step_mad)
template <typename V, bool SLAB, int SIGNS = 0>
__device__ __forceinline__ typename V::T rays_setup(RayS &s, const Queue &q, unsigned long long i, unsigned long long count,
                                                    const pvp_pair *__restrict__ pairs, const GridP &g, int width, int height, float &raw_diff, Ctl *ctl)
{
    rays_park(s);
    raw_diff = 0.f;
    if (i >= count || !queue_slot_ok(ctl, q, i)) return 0;
    const uint32_t pix = __ldg(q.pix + i);
    const float diff = __ldg(q.diff + i);
    if (!queue_entry_ok(ctl, pix, __ldg(q.pair + i), diff, width, height)) return 0;
    const pvp_pair *P = pairs + __ldg(q.pair + i);
    const int v = (int)(pix / (uint32_t)width), u = (int)(pix - (uint32_t)v * (uint32_t)width);
    float dx, dy, dz;
    pixel_ray(u, v, width, height, __ldg(&P->focal), P->R, dx, dy, dz);
    Ray r;
    if (!(ray_setup(__ldg(&P->cam[0]), __ldg(&P->cam[1]), __ldg(&P->cam[2]), dx, dy, dz, g, r) && r.t_cur <= r.t_end)) return 0;
    const int N = g.n;
    int rem_x = r.sx > 0 ? N - 1 - r.ix : r.ix;
    if (SLAB && !slab_enter(r, __ldg(&P->cam[0]), dx, g, rem_x)) return 0;  // never writes into this slab: not walked here; else fast-forwarded to its near face
    s.t_end = r.t_end; s.t_cur = r.t_cur;
    raw_diff = diff;
    s.tmx = r.tmx; s.tmy = r.tmy; s.tmz = r.tmz; s.tdx = r.tdx; s.tdy = r.tdy; s.tdz = r.tdz;
    s.idx = (unsigned)((r.ix - g.slab_lo) * N + r.iy) * (unsigned)N + (unsigned)r.iz;
    s.dix = SIGNS ? r.sx : r.sx * N * N; s.diy = SIGNS == 1 ? r.sy : r.sy * N; s.diz = r.sz;
    s.lim = fminf(fminf(axis_limit(r.tmx, r.tdx, rem_x), axis_limit(r.tmy, r.tdy, r.sy > 0 ? N - 1 - r.iy : r.iy)),
                  fminf(axis_limit(r.tmz, r.tdz, r.sz > 0 ? N - 1 - r.iz : r.iz), r.t_end));
    return V::make(diff, g);
}

// The axis choice of ray_voxel.cpp:375-387 for m = min(tmx, tmy, tmz) -- z iff tmz == m, else y iff tmy == m, else x -- as three
// one-instruction predicates, and the step as predicated adds: t_max of the chosen axis += its t_delta, the voxel offset += its
// stride.  (The compiler's own form of the same C++ was selects: 13 instructions per ray step, most on the ALU pipe; this is 10.)
__device__ __forceinline__ void dda_advance(float m, float &tmx, float &tmy, float &tmz, unsigned &idx, float tdx, float tdy, float tdz,
                                            int dix, int diy, int diz)
{
    asm("{\n\t.reg .pred pz, py, px;\n\t"
        "setp.eq.f32 pz, %2, %4;\n\t"
        "setp.eq.and.f32 py, %1, %4, !pz;\n\t"
        "setp.neu.and.f32 px, %1, %4, !pz;\n\t"
        "@px add.rn.f32 %0, %0, %5;\n\t"
        "@py add.rn.f32 %1, %1, %6;\n\t"
        "@pz add.rn.f32 %2, %2, %7;\n\t"
        "@px add.s32 %3, %3, %8;\n\t"
        "@py add.s32 %3, %3, %9;\n\t"
        "@pz add.s32 %3, %3, %10;\n\t}"
        : "+f"(tmx), "+f"(tmy), "+f"(tmz), "+r"(idx)
        : "f"(m), "f"(tdx), "f"(tdy), "f"(tdz), "r"(dix), "r"(diy), "r"(diz));
}

// a += b where p (the join of two rays of a lane that sit in the same voxel) as one predicated add
__device__ __forceinline__ void add_if(bool p, float &a, float b)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q add.rn.f32 %0, %0, %2;\n\t}" : "+f"(a) : "r"((unsigned)p), "f"(b));
}
__device__ __forceinline__ void add_if(bool p, unsigned long long &a, unsigned long long b)
{
    if (p) a += b;
}

// one DDA step (ray_voxel.cpp:375-391); returns the step's t (the minimum of the three t_max).  While it is below the ray's
// limit the ray certainly goes on; otherwise rays_ended decides exactly.
// PTX_ADVANCE picks the form of the axis choice + advance: the predicated PTX of dda_advance, or plain C++ (the compiler then
// uses selects).  Same arithmetic; which one ptxas schedules better depends on the rest of the loop -- measured 1080p -> 512^3:
// float32 786 G voxel-steps/s with the C++ form against 773, fixed point 758 with the PTX form against 742.
template <bool PTX_ADVANCE>
__device__ __forceinline__ float rays_step(RayS &s)
{
    const float m = fmin3(s.tmx, s.tmy, s.tmz);
    if (PTX_ADVANCE) {
        dda_advance(m, s.tmx, s.tmy, s.tmz, s.idx, s.tdx, s.tdy, s.tdz, s.dix, s.diy, s.diz);
    } else {
        const bool cz = (s.tmz == m);
        const bool cy = !cz && (s.tmy == m);
        const bool cx = !cz && (s.tmy != m);
        if (cx) s.tmx = fadd(s.tmx, s.tdx);
        if (cy) s.tmy = fadd(s.tmy, s.tdy);
        if (cz) s.tmz = fadd(s.tmz, s.tdz);
        s.idx += (unsigned)(cz ? s.diz : (cy ? s.diy : s.dix));
    }
    s.t_cur = m;  // :376-386 (dead in the kernels that do not attenuate)
    return m;
}

// rays_step with the voxel offset advanced by multiply-adds: offset += sign * stride, the signs (+-1) per ray in dix / diy / diz, the
// strides n^2 / n / 1 the same for every ray.  Same arithmetic as dda_advance; the point is the pipe: IMAD issues on the FMA pipe, and
// the lock-step loops saturate the ALU pipe (compares, min3, selects, integer adds: 76 % busy at 77 % of the issue slots in lean4,
// the FMA pipe at 23 % -- profiles/r2j_k2_lean4_dense_full_raw.csv), so the three offset adds of every ray step move over.
template <bool X_ONLY>
__device__ __forceinline__ float rays_step_mad(RayS &s, int nn, int n, int one)
{
    const float m = fmin3(s.tmx, s.tmy, s.tmz);
    if (X_ONLY) {   // diy holds the y stride itself (rays_setup<..., SIGNS = 2>)
        asm("{\n\t.reg .pred pz, py, px;\n\t"
            "setp.eq.f32 pz, %2, %4;\n\t"
            "setp.eq.and.f32 py, %1, %4, !pz;\n\t"
            "setp.neu.and.f32 px, %1, %4, !pz;\n\t"
            "@px add.rn.f32 %0, %0, %5;\n\t"
            "@py add.rn.f32 %1, %1, %6;\n\t"
            "@pz add.rn.f32 %2, %2, %7;\n\t"
            "@px mad.lo.s32 %3, %8, %11, %3;\n\t"
            "@py add.s32 %3, %3, %9;\n\t"
            "@pz add.s32 %3, %3, %10;\n\t}"
            : "+f"(s.tmx), "+f"(s.tmy), "+f"(s.tmz), "+r"(s.idx)
            : "f"(m), "f"(s.tdx), "f"(s.tdy), "f"(s.tdz), "r"(s.dix), "r"(s.diy), "r"(s.diz), "r"(nn));
        s.t_cur = m;
        return m;
    }
    asm("{\n\t.reg .pred pz, py, px;\n\t"
        "setp.eq.f32 pz, %2, %4;\n\t"
        "setp.eq.and.f32 py, %1, %4, !pz;\n\t"
        "setp.neu.and.f32 px, %1, %4, !pz;\n\t"
        "@px add.rn.f32 %0, %0, %5;\n\t"
        "@py add.rn.f32 %1, %1, %6;\n\t"
        "@pz add.rn.f32 %2, %2, %7;\n\t"
        "@px mad.lo.s32 %3, %8, %11, %3;\n\t"
        "@py mad.lo.s32 %3, %9, %12, %3;\n\t"
        "@pz mad.lo.s32 %3, %10, %13, %3;\n\t}"
        : "+f"(s.tmx), "+f"(s.tmy), "+f"(s.tmz), "+r"(s.idx)
        : "f"(m), "f"(s.tdx), "f"(s.tdy), "f"(s.tdz), "r"(s.dix), "r"(s.diy), "r"(s.diz), "r"(nn), "r"(n), "r"(one));
    s.t_cur = m;
    return m;
}

// rays_ended for rays whose dix / diy / diz hold signs (rays_setup<..., SIGNS = true>)
template <bool SLAB, bool Y_SIGN>
__device__ __forceinline__ bool rays_ended_signs(const RayS &s, unsigned cur, float m, const GridP &g)
{
    const int d = (int)(s.idx - cur);
    const bool cz = d == s.diz, cy = !cz && (Y_SIGN ? d == s.diy * g.n : d == s.diy);
    return step_leaves<SLAB>(cur, cz, cy, s.dix, s.diy, s.diz, g) || !(m <= s.t_end);
}

// exact end-of-ray test of the step (t = m) just taken from voxel `cur` (:388-391 and the loop condition :365).  The axis of the
// step is read back from the offset change (for N == 1 the strides coincide, and every step leaves whatever its axis).
template <bool SLAB>
__device__ __forceinline__ bool rays_ended(const RayS &s, unsigned cur, float m, const GridP &g)
{
    const int d = (int)(s.idx - cur);
    const bool cz = d == s.diz, cy = !cz && d == s.diy;
    return step_leaves<SLAB>(cur, cz, cy, s.dix, s.diy, s.diz, g) || !(m <= s.t_end);
}

template <int MODE, bool SLAB, bool NARROW, int MINB, bool ATT = false>
__global__ void __launch_bounds__(LEAN_THREADS, MINB)
dda_accumulate_lean2_kernel(Queue q, const pvp_pair *__restrict__ pairs, GridP g, typename Accum<MODE>::cell *__restrict__ grid,
                            int width, int height, Ctl *__restrict__ ctl)
{
    using V = LeanVal<MODE, NARROW>;
    using T = typename V::T;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const LaneMasks lm = lane_masks<PVP_RUN_CAP>(lane);
    const unsigned long long count = ctl->sel_count[1];  // the whole queue, or 0 when another lock-step kernel takes the batch (select_lockstep_kernel)
    const unsigned slab_cells = (unsigned)(g.slab_hi - g.slab_lo) * (unsigned)g.n * (unsigned)g.n;
    // per-lane counters in 32 bits (a lane walks count / 113 664 rays of at most 3 n steps each; the queue holds < 2^32 entries): three
    // registers less in a loop that sits at its register cap; widened for the warp reduction below
    unsigned steps32 = 0, written32 = 0, atomics32 = 0;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&ctl->sel_head[1], 64ull);
        base = __shfl_sync(FULL, base, 0);
        if (base >= count) break;
        RayS a, b;
        float da, db;  // the rays' pix_val (ray_voxel.cpp:500); the attenuating kernel derives every step's value from it
        T va = rays_setup<V, SLAB>(a, q, base + 2 * lane, count, pairs, g, width, height, da, ctl);
        T vb = rays_setup<V, SLAB>(b, q, base + 2 * lane + 1, count, pairs, g, width, height, db, ctl);
        float n_issued = 0.f, n_inslab = 0.f;
        float walked = 0.f;  // iterations of this chunk so far == voxels emitted by every ray that is still alive
        if (__all_sync(FULL, !rays_live(a) && !rays_live(b))) continue;

        unsigned heads = __ballot_sync(FULL, run_head(SLAB ? (a.idx < slab_cells ? a.idx : KEY_NONE) : a.idx));
        for (;;) {
            // ---- one DDA step of both rays first: the shuffle of A's new key overlaps the scan below ----------------
            unsigned ka = a.idx, kb = b.idx;
            const unsigned ka0 = ka, kb0 = kb;
            if (ATT) {  // value of THIS voxel step: pix_val / (1 + alpha * distance), before the step moves t_cur on
                va = V::make(attenuate(da, g.alpha, a.t_cur), g);
                vb = V::make(attenuate(db, g.alpha, b.t_cur), g);
            }
            const float ma = rays_step<MODE == PVP_ACCUM_FIXED>(a), mb = rays_step<MODE == PVP_ACCUM_FIXED>(b);
            const bool any_unsure = __any_sync(FULL, !(ma < a.lim) || !(mb < b.lim));
            walked += 1.f;
            T sa = va, sb = vb;
            if (SLAB) {  // writes outside the slab are suppressed; parked rays (KEY_NONE) are outside every slab
                const bool ina = ka < slab_cells, inb = kb < slab_cells;
                ka = ina ? ka : KEY_NONE; kb = inb ? kb : KEY_NONE;
                sa = ina ? va : (T)0; sb = inb ? vb : (T)0;
                n_inslab += (ina ? 1.f : 0.f) + (inb ? 1.f : 0.f);
            }
            const bool next_head = run_head(SLAB ? (a.idx < slab_cells ? a.idx : KEY_NONE) : a.idx);
            // ---- B joins A where both sit in the same voxel; otherwise B issues its own RED ------------------------
            const bool same = (kb == ka);
            const bool b_alone = !same && kb != KEY_NONE;
            T sum = same ? V::add(sa, sb) : sa, t;
            t = __shfl_up_sync(FULL, sum, 1);  if (!(heads & lm.w1))  sum = V::add(sum, t);
            t = __shfl_up_sync(FULL, sum, 2);  if (!(heads & lm.w2))  sum = V::add(sum, t);
            t = __shfl_up_sync(FULL, sum, 4);  if (!(heads & lm.w4))  sum = V::add(sum, t);
            if (PVP_RUN_CAP > 8)  { t = __shfl_up_sync(FULL, sum, 8);  if (!(heads & lm.w8))  sum = V::add(sum, t); }
            if (PVP_RUN_CAP > 16) { t = __shfl_up_sync(FULL, sum, 16); if (!(heads & lm.w16)) sum = V::add(sum, t); }
            const bool issue = (heads & lm.next) && ka != KEY_NONE;
            red_if(issue && PVP_DEV_CHECK(ctl, PVP_CHK_GRID, ka < slab_cells), grid + ka, V::cell_value(sum, g));
            red_if(b_alone && PVP_DEV_CHECK(ctl, PVP_CHK_GRID, kb < slab_cells), grid + kb, V::cell_value(sb, g));
#ifndef PVP_NO_ATOMIC_COUNT
            if (issue) n_issued += 1.f;
            if (b_alone) n_issued += 1.f;
#endif
            heads = __ballot_sync(FULL, next_head);
            // ---- rays that ended: park them; leave when every ray of the warp is parked ----------------------------
            if (any_unsure) {
                // (the original keys, not ka / kb: a slab kernel has replaced those outside its slab by KEY_NONE)
                if (!(ma < a.lim) && rays_live(a) && rays_ended<SLAB>(a, ka0, ma, g)) {
                    steps32 += (unsigned)walked;
                    rays_retire(a); va = 0; da = 0.f;
                }
                if (!(mb < b.lim) && rays_live(b) && rays_ended<SLAB>(b, kb0, mb, g)) {
                    steps32 += (unsigned)walked;
                    rays_retire(b); vb = 0; db = 0.f;
                }
                if (__all_sync(FULL, !rays_live(a) && !rays_live(b))) break;
                heads = __ballot_sync(FULL, run_head(SLAB ? (a.idx < slab_cells ? a.idx : KEY_NONE) : a.idx));
            }
        }
        atomics32 += (unsigned)n_issued;
        if (SLAB) written32 += (unsigned)n_inslab;
    }
    unsigned long long my_steps = steps32, my_written = SLAB ? written32 : steps32, my_atomics = atomics32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_steps += __shfl_xor_sync(FULL, my_steps, o);
        my_written += __shfl_xor_sync(FULL, my_written, o);
        my_atomics += __shfl_xor_sync(FULL, my_atomics, o);
    }
    if (lane == 0 && my_steps) {
        atomicAdd(&ctl->n_steps, my_steps);
        atomicAdd(&ctl->n_written, my_written);
        atomicAdd(&ctl->n_atomics, my_atomics);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->n_rays, count);
}

// ---- variant 6: "lean4", four rays per lane -------------------------------------------------------------------------
// A warp takes 128 consecutive queue entries; lane l walks entries 4l .. 4l+3, which K1's column-snake order makes one
// column of a 4-row block (or half a column of an 8-row block).  The per-iteration overhead of the lock-step loop -- the
// run scan, the run-head ballot and shuffle, the end-of-ray vote, the loop itself -- is paid once per 128 voxel
// steps instead of once per 64.  Inside a lane the four rays are chained bottom-up: ray j joins ray j-1 where both sit in
// the same voxel (a predicated add), otherwise it issues its own RED; ray 0 carries the lane's sum into the run scan.
// Runs are cut at every 8th lane (lane_masks<8>): 32 columns of the image span ~5 voxels, so a natural run is ~7 lanes long
// anyway; the cut costs ~0.02 REDs per voxel step and saves two of the five scan levels.
// COUNT: count the REDs (pvp_stats.n_atomics; context option count_atomics) -- a diagnostic that costs four instructions
// of the loop, so the default instantiation leaves it out and reports n_atomics = 0.
// Per ray the arithmetic, the step order and the counters are those of variants 4 and 5 (same bit-exact visited sets).
// Measured 1080p dense -> 512^3 float32 (G voxel-steps/s; lean2: 829): 256-thread CTAs x 2 (108 registers, 16 warps per SM) 898, 128 x 5 (96
// registers, 20 warps) 941, 192 x 3 at 104 registers 816, 128 x 5 at 88 registers (a 124-instruction loop) 790; run cap 4 / 8 / 16 / none
// 817 / 898 / 894 / 775 (256 x 2); counting the REDs 852 against 898; the loop unrolled twice 954, the chain's joins as predicated adds
// 966 (not unrolled; both together 960); three rays per lane 876-908, two rays per lane in this form 790-800 (profiles/r2_lean4_tuning.log).
#ifndef PVP_LEAN4_THREADS
#define PVP_LEAN4_THREADS 128
#endif
#ifndef PVP_LEAN4_MINB
#define PVP_LEAN4_MINB 5  // resident CTAs: 96 registers at 128 threads, 20 warps per SM
#endif
#ifndef PVP_LEAN4_CAP
#define PVP_LEAN4_CAP 8
#endif
#ifndef PVP_LEAN4_UNROLL
#define PVP_LEAN4_UNROLL 1
#endif
#ifndef PVP_LEAN4_MAD
#define PVP_LEAN4_MAD 1  // voxel offset advanced by IMAD (FMA pipe) instead of IADD (ALU pipe), see rays_step_mad
#endif
#ifndef PVP_LEAN4_NR
#define PVP_LEAN4_NR 4  // rays per lane
#endif
constexpr int LEAN4_THREADS = PVP_LEAN4_THREADS;
#ifdef PVP_LEAN4_MAXNREG
#define PVP_LEAN4_BOUNDS __maxnreg__(PVP_LEAN4_MAXNREG)
#else
#define PVP_LEAN4_BOUNDS __launch_bounds__(LEAN4_THREADS, PVP_LEAN4_MINB)
#endif
template <int MODE, bool SLAB, bool NARROW, bool COUNT>
__global__ void PVP_LEAN4_BOUNDS
dda_accumulate_lean4_kernel(Queue q, const pvp_pair *__restrict__ pairs, GridP g, typename Accum<MODE>::cell *__restrict__ grid,
                            int width, int height, Ctl *__restrict__ ctl)
{
    using V = LeanVal<MODE, NARROW>;
    using T = typename V::T;
    constexpr int NR = PVP_LEAN4_NR, CAP = PVP_LEAN4_CAP;
    constexpr bool L4_MAD = PVP_LEAN4_MAD != 0;
    // PVP_LEAN4_MAD: 1 = x and y offsets by IMAD, z by IADD (literal stride 1); 2 = all three by IMAD (stride g.one, a run-time 1: 149
    // instructions, ptxas runs out of predicate registers); 3 = x only
    const int stride_x = g.n * g.n, stride_y = (PVP_LEAN4_MAD == 3) ? 0 : g.n, stride_z = (PVP_LEAN4_MAD == 2) ? g.one : 1;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const LaneMasks lm = lane_masks<CAP>(lane);
    const unsigned long long count = ctl->sel_count[2];  // the whole queue, or 0 when another lock-step kernel takes the batch (select_lockstep_kernel)
    const unsigned slab_cells = (unsigned)(g.slab_hi - g.slab_lo) * (unsigned)g.n * (unsigned)g.n;
    unsigned steps32 = 0, written32 = 0, atomics32 = 0;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&ctl->sel_head[2], (unsigned long long)(32 * NR));
        base = __shfl_sync(FULL, base, 0);
        if (base >= count) break;
        RayS r[NR];
        T v[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            float raw;
            v[j] = rays_setup<V, SLAB, (PVP_LEAN4_MAD == 3 ? 2 : PVP_LEAN4_MAD != 0 ? 1 : 0)>(r[j], q, base + NR * lane + j, count, pairs, g, width, height, raw, ctl);
        }
        float n_issued = 0.f, n_inslab = 0.f;
        float walked = 0.f;
        bool none_live = true;
#pragma unroll
        for (int j = 0; j < NR; ++j) none_live = none_live && !rays_live(r[j]);
        if (__all_sync(FULL, none_live)) continue;

        unsigned heads = __ballot_sync(FULL, run_head(SLAB ? (r[0].idx < slab_cells ? r[0].idx : KEY_NONE) : r[0].idx));
        // one lock-step iteration; true when every ray of the warp has ended
        auto iteration = [&]() -> bool {
            unsigned k[NR], k0[NR];
            float m[NR];
            T s[NR];
            bool unsure = false;
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                k[j] = k0[j] = r[j].idx;
                // the predicated-PTX forms: with four rays the compiler's own form of the C++ turns into divergent branches
                m[j] = L4_MAD ? rays_step_mad<PVP_LEAN4_MAD == 3>(r[j], stride_x, stride_y, stride_z) : rays_step<true>(r[j]);
                unsure = unsure || !(m[j] < r[j].lim);
                s[j] = v[j];
            }
            const bool any_unsure = __any_sync(FULL, unsure);
            walked += 1.f;
            if (SLAB) {
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    const bool in = k[j] < slab_cells;
                    k[j] = in ? k[j] : KEY_NONE;
                    s[j] = in ? s[j] : (T)0;
                    n_inslab += in ? 1.f : 0.f;
                }
            }
            const bool next_head = run_head(SLAB ? (r[0].idx < slab_cells ? r[0].idx : KEY_NONE) : r[0].idx);
            // ---- the lane's chain: ray j joins ray j-1 where they share the voxel, else it issues its own RED ------------
#pragma unroll
            for (int j = NR - 1; j > 0; --j) {
                const bool same = (k[j] == k[j - 1]);
                const bool alone = !same && k[j] != KEY_NONE;
                red_if(alone && PVP_DEV_CHECK(ctl, PVP_CHK_GRID, k[j] < slab_cells), grid + k[j], V::cell_value(s[j], g));
                if (COUNT && alone) n_issued += 1.f;
                add_if(same, s[j - 1], s[j]);
            }
            T sum = s[0], t;
            t = __shfl_up_sync(FULL, sum, 1);  if (!(heads & lm.w1))  sum = V::add(sum, t);
            t = __shfl_up_sync(FULL, sum, 2);  if (!(heads & lm.w2))  sum = V::add(sum, t);
            if (CAP > 4)  { t = __shfl_up_sync(FULL, sum, 4);  if (!(heads & lm.w4))  sum = V::add(sum, t); }
            if (CAP > 8)  { t = __shfl_up_sync(FULL, sum, 8);  if (!(heads & lm.w8))  sum = V::add(sum, t); }
            if (CAP > 16) { t = __shfl_up_sync(FULL, sum, 16); if (!(heads & lm.w16)) sum = V::add(sum, t); }
            const bool issue = (heads & lm.next) && k[0] != KEY_NONE;
            red_if(issue && PVP_DEV_CHECK(ctl, PVP_CHK_GRID, k[0] < slab_cells), grid + k[0], V::cell_value(sum, g));
            if (COUNT && issue) n_issued += 1.f;
            heads = __ballot_sync(FULL, next_head);
            // ---- rays that ended: park them; leave when every ray of the warp is parked ----------------------------
            if (any_unsure) {
                bool live = false;
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    if (!(m[j] < r[j].lim) && rays_live(r[j]) && (L4_MAD ? rays_ended_signs<SLAB, PVP_LEAN4_MAD != 3>(r[j], k0[j], m[j], g) : rays_ended<SLAB>(r[j], k0[j], m[j], g))) {
                        steps32 += (unsigned)walked;
                        rays_retire(r[j]); v[j] = 0;
                    }
                    live = live || rays_live(r[j]);
                }
                if (__all_sync(FULL, !live)) return true;
                heads = __ballot_sync(FULL, run_head(SLAB ? (r[0].idx < slab_cells ? r[0].idx : KEY_NONE) : r[0].idx));
            }
            return false;
        };
        for (;;) {
            if (iteration()) break;
            if (PVP_LEAN4_UNROLL > 1 && iteration()) break;
        }
        if (COUNT) atomics32 += (unsigned)n_issued;
        if (SLAB) written32 += (unsigned)n_inslab;
    }
    unsigned long long my_steps = steps32, my_written = SLAB ? written32 : steps32, my_atomics = atomics32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_steps += __shfl_xor_sync(FULL, my_steps, o);
        my_written += __shfl_xor_sync(FULL, my_written, o);
        if (COUNT) my_atomics += __shfl_xor_sync(FULL, my_atomics, o);
    }
    if (lane == 0 && my_steps) {
        atomicAdd(&ctl->n_steps, my_steps);
        atomicAdd(&ctl->n_written, my_written);
        if (COUNT) atomicAdd(&ctl->n_atomics, my_atomics);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->n_rays, count);
}

// Ray set-up of ray_voxel.cpp:506-515 as a stage.
__global__ void pixel_rays_kernel(const uint32_t *__restrict__ pix, long long n, int width, int height, pvp_pair pose,
                                  float *__restrict__ dirs)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t p = pix[i];
        const int v = (int)(p / (uint32_t)width), u = (int)(p - (uint32_t)v * (uint32_t)width);
        float dx, dy, dz;
        pixel_ray(u, v, width, height, pose.focal, pose.R, dx, dy, dz);
        dirs[3 * i] = dx; dirs[3 * i + 1] = dy; dirs[3 * i + 2] = dz;
    }
}

// cast_ray_into_grid (ray_voxel.cpp:274-395) as a stage: counts, and optionally the step lists.
__global__ void cast_rays_kernel(const float *__restrict__ origins, const float *__restrict__ dirs, long long n, GridP g,
                                 int32_t *__restrict__ counts, const long long *__restrict__ offsets,
                                 int32_t *__restrict__ vox, float *__restrict__ tt)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Ray r;
        int c = 0;
        if (ray_setup(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2], dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], g, r)) {
            long long o = offsets ? offsets[i] : 0;
            while (r.t_cur <= r.t_end) {
                if (offsets) {
                    if (vox) { vox[3 * (o + c)] = r.ix; vox[3 * (o + c) + 1] = r.iy; vox[3 * (o + c) + 2] = r.iz; }
                    if (tt) tt[o + c] = r.t_cur;
                }
                ++c;
                if (!ray_advance(r, g.n)) break;
            }
        }
        counts[i] = c;
    }
}

__global__ void fixed_to_f32_kernel(const long long *__restrict__ src, float *__restrict__ dst, long long n, double inv)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = (float)((double)src[i] * inv);
}

__global__ void fixed_to_f64_kernel(const long long *__restrict__ src, double *__restrict__ dst, long long n, double inv)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = (double)src[i] * inv;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------

// Host evaluation of ray_voxel.cpp:284-290 in fp32 (volatile blocks any re-association).
static int make_gridp(const pvp_grid_desc *d, GridP *g)
{
    PVP_REQUIRE(d != nullptr, "grid descriptor is NULL");
    // up to 2048 the lock-step kernels' per-ray limits are proven (axis_limit: rem < 2^11); larger grids -- the ones that need slab
    // sharding because they exceed one GPU's memory, e.g. 4096^3 float32 = 256 GiB -- run the reference-structured kernel (variant 3)
    PVP_REQUIRE(d->n > 0 && d->n <= 4096, "grid n=%d out of range (1..4096)", d->n);
    PVP_REQUIRE(d->slab_lo >= 0 && d->slab_lo <= d->slab_hi && d->slab_hi <= d->n, "bad slab [%d,%d) for n=%d",
                d->slab_lo, d->slab_hi, d->n);
    PVP_REQUIRE(d->accum_mode == PVP_ACCUM_F32 || d->accum_mode == PVP_ACCUM_FIXED, "bad accum_mode %d", d->accum_mode);
    PVP_REQUIRE(d->accum_mode == PVP_ACCUM_F32 || (d->frac_bits >= 0 && d->frac_bits <= 40), "frac_bits %d out of range",
                d->frac_bits);
    g->n = d->n; g->slab_lo = d->slab_lo; g->slab_hi = d->slab_hi; g->vs = d->voxel_size;
    volatile float extent = (float)d->n * d->voxel_size;
    volatile float half = 0.5f * extent;
    for (int a = 0; a < 3; ++a) {
        volatile float lo = d->center[a] - half, hi = d->center[a] + half;
        g->gmin[a] = lo; g->gmax[a] = hi;
    }
    g->fixed_scale = (float)((long long)1 << (d->accum_mode == PVP_ACCUM_FIXED ? d->frac_bits : 0));
    g->frac_shift = d->accum_mode == PVP_ACCUM_FIXED ? d->frac_bits : 0;
    PVP_REQUIRE(d->attenuation_alpha >= 0.f && d->attenuation_alpha < 1e30f, "attenuation_alpha must be finite and >= 0");
    g->alpha = d->attenuation_alpha;
    g->mul_n = (unsigned)(0x100000000ull / (unsigned long long)d->n);  // n >= 2 here; n == 1 divides by shifting nothing: see udiv_by
    g->mul_nn = (unsigned)(0x100000000ull / ((unsigned long long)d->n * d->n));
    g->one = 1;
    return PVP_OK;
}

static int k2_grid_size(pvp_ctx *ctx, const void *kernel, int threads = K2_THREADS)
{
    int per_sm = 0;
    if (ctx->opt_ctas_per_sm > 0) {
        per_sm = (int)ctx->opt_ctas_per_sm;
    } else {
        auto hit = ctx->occupancy.find(kernel);  // the query costs microseconds per launch: keep the answer per kernel
        if (hit != ctx->occupancy.end()) {
            per_sm = hit->second;
        } else {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
            ctx->occupancy[kernel] = per_sm;
        }
    }
    return ctx->sm_count * per_sm;  // persistent: every CTA resident, a multiple of the SM count
}

// the queue as K1 sees it; the checked build can be told to under-state its capacity (option inject_fault = 1) so that a test can see the
// bounds checks fire: the writes past the stated capacity are skipped and counted, PVP_ERR_CHECK comes back
static Queue k1_queue(const pvp_ctx *ctx)
{
    Queue q = ctx->queue;
#ifdef PVP_CHECKED
    if (ctx->opt_inject_fault == 1) q.capacity /= 4;
#endif
    return q;
}

template <typename T>
static int launch_k1(pvp_ctx *ctx, const void *frames_dev, long long frame_stride, long long n_pix, int n_pairs, float threshold,
                     int width = 0, int height = 0, int row_group = 1, bool snake = false)
{
    constexpr int PX = FrameVec<T>::PX;
    if (row_group < 0) {  // Z-order queue
        const size_t vec_bytes = sizeof(T) * K1T_COLS;
        const bool vec = ((uintptr_t)frames_dev % vec_bytes == 0) && (frame_stride % K1T_COLS == 0) && (width % K1T_COLS == 0);
        const long long foots = (long long)((width + K1M_FOOT - 1) / K1M_FOOT) * ((height + K1M_FOOT - 1) / K1M_FOOT);
        dim3 grid((unsigned)foots, (unsigned)n_pairs);
        LaunchScope scope(ctx, 1);
        if (row_group == -4)
            motion_compact_morton_kernel<T, 4><<<grid, K1Morton<4>::THREADS, 0, ctx->stream>>>((const T *)frames_dev, frame_stride, ctx->pairs_dev, width, height, threshold, vec,
                                                                                                k1_queue(ctx), ctx->ctl_dev);
        else
            motion_compact_morton_kernel<T, 8><<<grid, K1Morton<8>::THREADS, 0, ctx->stream>>>((const T *)frames_dev, frame_stride, ctx->pairs_dev, width, height, threshold, vec,
                                                                                                k1_queue(ctx), ctx->ctl_dev);
        PVP_CUDA(cudaGetLastError());
        return PVP_OK;
    }
    if (row_group > 1) {
        const size_t vec_bytes = sizeof(T) * K1T_COLS;  // one row segment of a thread
        const bool vec = ((uintptr_t)frames_dev % vec_bytes == 0) && (frame_stride % K1T_COLS == 0) && (width % K1T_COLS == 0);
        const int R = row_group >= 8 ? 8 : row_group >= 4 ? 4 : 2;
        const int threads = R == 8 ? K1Tile<8>::THREADS : R == 4 ? K1Tile<4>::THREADS : K1Tile<2>::THREADS;
        const int per_cta = threads * (R == 8 ? K1Tile<8>::ITER : R == 4 ? K1Tile<4>::ITER : K1Tile<2>::ITER);
        const long long segs = (long long)((width + K1T_COLS - 1) / K1T_COLS) * ((height + R - 1) / R);
        dim3 grid((unsigned)((segs + per_cta - 1) / per_cta), (unsigned)n_pairs);
        LaunchScope scope(ctx, 1);
        auto k = R == 8 ? motion_compact_rowgroup_kernel<T, 8> : R == 4 ? motion_compact_rowgroup_kernel<T, 4> : motion_compact_rowgroup_kernel<T, 2>;
        k<<<grid, threads, 0, ctx->stream>>>((const T *)frames_dev, frame_stride, ctx->pairs_dev, width, height, threshold, vec, snake, k1_queue(ctx), ctx->ctl_dev);
        PVP_CUDA(cudaGetLastError());
        return PVP_OK;
    }
    const bool vec_ok = ((uintptr_t)frames_dev % sizeof(typename FrameVec<T>::V) == 0) && (frame_stride % PX == 0);
    dim3 grid((unsigned)((n_pix + (long long)K1_THREADS * PX - 1) / ((long long)K1_THREADS * PX)), (unsigned)n_pairs);
    LaunchScope scope(ctx, 1);
    motion_compact_kernel<T><<<grid, K1_THREADS, 0, ctx->stream>>>((const T *)frames_dev, frame_stride, ctx->pairs_dev, n_pix,
                                                                   threshold, vec_ok, k1_queue(ctx), ctx->ctl_dev);
    PVP_CUDA(cudaGetLastError());
    return PVP_OK;
}

template <int MODE, int VARIANT, bool SLAB>
static void launch_agg(pvp_ctx *ctx, const GridP &g, void *grid_dev, int width, int height)
{
    auto k = dda_accumulate_agg_kernel<MODE, VARIANT, SLAB>;
    k<<<k2_grid_size(ctx, (const void *)k), K2_THREADS, 0, ctx->stream>>>(ctx->queue, ctx->pairs_dev, g, (typename Accum<MODE>::cell *)grid_dev,
                                                                           width, height, ctx->ctl_dev);
}

// lean_regs: register budget = resident warps per SM: 0/72 -> 28 warps at 128-thread CTAs (24 at 256), 80 -> 24 warps, 64 -> 32 warps
template <int MODE, bool SLAB, bool NARROW>
static void launch_lean(pvp_ctx *ctx, const GridP &g, void *grid_dev, int width, int height)
{
    auto k = ctx->opt_lean_regs == 64 ? dda_accumulate_lean_kernel<MODE, SLAB, NARROW, LEAN_CTAS_C>
           : ctx->opt_lean_regs == 80 ? dda_accumulate_lean_kernel<MODE, SLAB, NARROW, LEAN_CTAS_B> : dda_accumulate_lean_kernel<MODE, SLAB, NARROW, LEAN_CTAS_A>;
    k<<<k2_grid_size(ctx, (const void *)k, LEAN_THREADS), LEAN_THREADS, 0, ctx->stream>>>(ctx->queue, ctx->pairs_dev, g, (typename Accum<MODE>::cell *)grid_dev,
                                                                                           width, height, ctx->ctl_dev);
}

template <int MODE, bool SLAB, bool NARROW>
static void launch_lean2(pvp_ctx *ctx, const GridP &g, void *grid_dev, int width, int height)
{
    if (g.alpha > 0.f) {  // opt-in distance attenuation: values change every step, so no integral shortcut (NARROW is false here)
        auto ka = dda_accumulate_lean2_kernel<MODE, SLAB, false, 3, true>;
        ka<<<k2_grid_size(ctx, (const void *)ka, LEAN_THREADS), LEAN_THREADS, 0, ctx->stream>>>(ctx->queue, ctx->pairs_dev, g, (typename Accum<MODE>::cell *)grid_dev,
                                                                                                 width, height, ctx->ctl_dev);
        return;
    }
    // lean_regs: 0/80 -> 3 resident CTAs (24 warps; measured best: 780 G voxel-steps/s at 512^3 on 2-row tiles), 128 -> 2 CTAs (686 G), 64 -> 4 CTAs (722 G)
    auto k = ctx->opt_lean_regs == 128 ? dda_accumulate_lean2_kernel<MODE, SLAB, NARROW, 2>
           : ctx->opt_lean_regs == 64  ? dda_accumulate_lean2_kernel<MODE, SLAB, NARROW, 4> : dda_accumulate_lean2_kernel<MODE, SLAB, NARROW, PVP_LEAN2_MINB>;
    k<<<k2_grid_size(ctx, (const void *)k, LEAN_THREADS), LEAN_THREADS, 0, ctx->stream>>>(ctx->queue, ctx->pairs_dev, g, (typename Accum<MODE>::cell *)grid_dev,
                                                                                           width, height, ctx->ctl_dev);
}

template <int MODE, bool SLAB, bool NARROW>
static void launch_lean4(pvp_ctx *ctx, const GridP &g, void *grid_dev, int width, int height)
{
    auto k = ctx->opt_count_atomics ? dda_accumulate_lean4_kernel<MODE, SLAB, NARROW, true> : dda_accumulate_lean4_kernel<MODE, SLAB, NARROW, false>;
    k<<<k2_grid_size(ctx, (const void *)k, LEAN4_THREADS), LEAN4_THREADS, 0, ctx->stream>>>(ctx->queue, ctx->pairs_dev, g, (typename Accum<MODE>::cell *)grid_dev,
                                                                                             width, height, ctx->ctl_dev);
}

// See the lean kernel's header: no t value of the batch can overflow to -inf.
static bool lean_geometry_ok(const GridP &g, const pvp_pair *pairs, int n_pairs)
{
    const float lim = 1e25f;
    if (!(g.vs > 0.f && g.vs < lim)) return false;
    for (int a = 0; a < 3; ++a)
        if (!(fabsf(g.gmin[a]) < lim && fabsf(g.gmax[a]) < lim)) return false;
    for (int i = 0; i < n_pairs; ++i)
        for (int a = 0; a < 3; ++a)
            if (!(fabsf(pairs[i].cam[a]) < lim)) return false;  // NaN fails too
    return true;
}

// 0 = auto: the lock-step kernels are all launched and the batch's ray count and paired share pick one on the device (select_lockstep_kernel:
// lean2 from LEAN2_MIN_RAYS_PER_WARP rays per resident warp up -- with fewer rays its 64-ray chunks leave SMs idle --, lean4 from
// LEAN4_MIN_RAYS_PER_WARP)
static int effective_variant(const pvp_ctx *ctx)
{
    const int v = (int)ctx->opt_dda_variant;
    return (v < 0 || v > 6) ? 0 : v;
}
constexpr unsigned long long LEAN2_MIN_RAYS_PER_WARP = 256;
constexpr unsigned long long LEAN4_MIN_RAYS_PER_WARP = 2048;  // 6 M rays on a B200: a moving object of 2.4 M rays per batch runs better on lean2, 38 M rays of 60 % iid motion on lean4

// Bytes of one x-plane of the grid (n * n cells).  Neighbouring image columns land in neighbouring x-planes, i.e. this
// far apart in memory; once a plane outgrows a 2 MiB page the RED misses of a wide, flat ray tile dominate K2 (measured at
// 768^3 / 1024^3 float32 and 512^3 int64: profiles/r1_summary.md, "large grids").  The auto settings follow it:
//   <= 2 MiB: lean2, queue tiles of 4 image rows (16 x 4 pixel chunks), snake order inside a tile
//             (1080p -> 512^3 float32: 823 G voxel-steps/s against 786 with 2-row tiles: with the per-ray limits the rays of a
//             compact tile also END in fewer distinct iterations, i.e. fewer trips through the slow path; 512^3 int64: 570 -> 626 G)
//   >  2 MiB: lean (32 rays per warp), tiles of 8 rows = 4 x 8 pixel chunks, snake order   (1024^3 float32: 294 -> 576 -> 600 G)
static size_t plane_bytes(const GridP &g, int accum_mode)
{
    return (size_t)g.n * g.n * (accum_mode == PVP_ACCUM_F32 ? sizeof(float) : sizeof(unsigned long long));
}
constexpr size_t PLANE_LARGE = 2u << 20;
// dda_variant: 0 = auto (lean2 / lean by ray count where their 32-bit voxel keys allow, else 3), 1 = segmented-scan aggregation (round-1
// baseline), 2 = match.any aggregation, 3 = one RED per lane and step (no aggregation, any grid size),
// 4 = lean lock-step kernel (one ray per lane), 5 = lean2 (two rays per lane, vertical neighbours of K1's tile order),
// 6 = lean4 (four rays per lane: a column of K1's 4-row blocks).  narrow_ok: uint8 frames, i.e. integer addends (LeanVal<FIXED, true>).
static int launch_k2(pvp_ctx *ctx, const GridP &g, int accum_mode, bool narrow_ok, bool lean_ok, void *grid_dev, int width, int height,
                     unsigned long long batch_pixels)  // NOLINT
{
    LaunchScope scope(ctx, 2);
    int variant = effective_variant(ctx);
    const bool slab = !(g.slab_lo == 0 && g.slab_hi == g.n);
    const unsigned long long cells = (unsigned long long)(g.slab_hi - g.slab_lo) * g.n * g.n;
    const unsigned long long full = (unsigned long long)g.n * g.n * g.n;
    if (cells >= 0xffffffffull || g.n > 2048) variant = 3;
    const bool lockstep = variant == 0 || variant >= 4;
    // 32-bit voxel keys: the whole grid for whole-grid kernels; under slabs only the slab's own cells -- since the exact fast-forward
    // (slab_enter) a walked ray is inside its slab from its first step to the step that passes the far face, so its slab-relative
    // offset never leaves [0, cells) (step_leaves decodes it as an int: cells <= 2^31).  A 2048^3 grid in 8 slabs thus runs on the
    // lock-step kernels (round 1 and the first half of round 2 fell back to the one-RED-per-step kernel above 1290^3)
    if (lockstep && (slab ? cells > (1ull << 31) : full >= 0xffffffffull)) variant = 3;
    if (!lockstep && variant != 3 && full >= 0xffffffffull) variant = 3;   // variants 1 / 2 walk every ray from its start with wrapped slab offsets: the whole grid must fit the key
    if (lockstep && !lean_ok) variant = 3;
    const bool f32 = accum_mode == PVP_ACCUM_F32;
    ctx->last_k2_lockstep = false;
    if (g.alpha > 0.f && variant != 3) { variant = 5; narrow_ok = false; }  // only lean2 and variant 3 implement the attenuation
    if (variant == 0 && plane_bytes(g, accum_mode) > PLANE_LARGE) variant = 4;
    if (variant == 0) {
        // all eligible lock-step kernels are launched; K1's counters pick one on the device (pick_lockstep): lean2 / lean4 from
        // LEAN2_MIN_RAYS_PER_WARP rays per resident warp up AND lean2_paired percent of the rays in vertical pairs (default 50), lean otherwise;
        // of the two, lean4 from LEAN4_MIN_RAYS_PER_WARP rays per resident warp up -- float32 cells without attenuation only (with int64
        // cells lean4 is no faster than lean2: 751 against 754 G voxel-steps/s).  The kernels that are not chosen exit at once.
        const unsigned long long T = (unsigned long long)ctx->sm_count * 3 * (LEAN_THREADS / 32) * LEAN2_MIN_RAYS_PER_WARP;
        const unsigned long long pct = (unsigned long long)(ctx->opt_lean2_density > 0 ? ctx->opt_lean2_density : 50);
        const bool lean4_ok = f32 && g.alpha == 0.f && ctx->opt_lean4 >= 0;
        const unsigned long long T4 = lean4_ok ? (unsigned long long)ctx->sm_count * PVP_LEAN4_MINB * (LEAN4_THREADS / 32) * LEAN4_MIN_RAYS_PER_WARP : ~0ull;
        ctx->prof.launches_k2 += lean4_ok ? 3 : 2;  // the selector and two or three kernels (the scope above counted one)
        (void)batch_pixels;
        ctx->last_k2_lockstep = true;
        select_lockstep_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl_dev, T, pct, T4);
        if (f32) {
            if (slab) { launch_lean<PVP_ACCUM_F32, true, false>(ctx, g, grid_dev, width, height); launch_lean2<PVP_ACCUM_F32, true, false>(ctx, g, grid_dev, width, height);
                        if (lean4_ok) launch_lean4<PVP_ACCUM_F32, true, false>(ctx, g, grid_dev, width, height); }
            else { launch_lean<PVP_ACCUM_F32, false, false>(ctx, g, grid_dev, width, height); launch_lean2<PVP_ACCUM_F32, false, false>(ctx, g, grid_dev, width, height);
                   if (lean4_ok) launch_lean4<PVP_ACCUM_F32, false, false>(ctx, g, grid_dev, width, height); }
        } else if (narrow_ok) {
            if (slab) { launch_lean<PVP_ACCUM_FIXED, true, true>(ctx, g, grid_dev, width, height); launch_lean2<PVP_ACCUM_FIXED, true, true>(ctx, g, grid_dev, width, height); }
            else { launch_lean<PVP_ACCUM_FIXED, false, true>(ctx, g, grid_dev, width, height); launch_lean2<PVP_ACCUM_FIXED, false, true>(ctx, g, grid_dev, width, height); }
        } else {
            if (slab) { launch_lean<PVP_ACCUM_FIXED, true, false>(ctx, g, grid_dev, width, height); launch_lean2<PVP_ACCUM_FIXED, true, false>(ctx, g, grid_dev, width, height); }
            else { launch_lean<PVP_ACCUM_FIXED, false, false>(ctx, g, grid_dev, width, height); launch_lean2<PVP_ACCUM_FIXED, false, false>(ctx, g, grid_dev, width, height); }
        }
    } else if (variant == 6) {
        ctx->last_k2_lockstep = true; ++ctx->prof.launches_k2;
        select_lockstep_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl_dev, 0ull, 0ull, 0ull);  // forced: (min_rays, paired_pct, min_rays4) = (0, 0, 0) is always lean4
        if (f32) { if (slab) launch_lean4<PVP_ACCUM_F32, true, false>(ctx, g, grid_dev, width, height); else launch_lean4<PVP_ACCUM_F32, false, false>(ctx, g, grid_dev, width, height); }
        else if (narrow_ok) { if (slab) launch_lean4<PVP_ACCUM_FIXED, true, true>(ctx, g, grid_dev, width, height); else launch_lean4<PVP_ACCUM_FIXED, false, true>(ctx, g, grid_dev, width, height); }
        else { if (slab) launch_lean4<PVP_ACCUM_FIXED, true, false>(ctx, g, grid_dev, width, height); else launch_lean4<PVP_ACCUM_FIXED, false, false>(ctx, g, grid_dev, width, height); }
    } else if (variant == 5) {
        ctx->last_k2_lockstep = true; ++ctx->prof.launches_k2;
        select_lockstep_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl_dev, 0ull, 0ull, ~0ull);  // always lean2
        if (f32) { if (slab) launch_lean2<PVP_ACCUM_F32, true, false>(ctx, g, grid_dev, width, height); else launch_lean2<PVP_ACCUM_F32, false, false>(ctx, g, grid_dev, width, height); }
        else if (narrow_ok) { if (slab) launch_lean2<PVP_ACCUM_FIXED, true, true>(ctx, g, grid_dev, width, height); else launch_lean2<PVP_ACCUM_FIXED, false, true>(ctx, g, grid_dev, width, height); }
        else { if (slab) launch_lean2<PVP_ACCUM_FIXED, true, false>(ctx, g, grid_dev, width, height); else launch_lean2<PVP_ACCUM_FIXED, false, false>(ctx, g, grid_dev, width, height); }
    } else if (variant == 4) {
        ctx->last_k2_lockstep = true; ++ctx->prof.launches_k2;
        select_lockstep_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl_dev, ~0ull, 0ull, ~0ull);  // always lean
        if (f32) { if (slab) launch_lean<PVP_ACCUM_F32, true, false>(ctx, g, grid_dev, width, height); else launch_lean<PVP_ACCUM_F32, false, false>(ctx, g, grid_dev, width, height); }
        else if (narrow_ok) { if (slab) launch_lean<PVP_ACCUM_FIXED, true, true>(ctx, g, grid_dev, width, height); else launch_lean<PVP_ACCUM_FIXED, false, true>(ctx, g, grid_dev, width, height); }
        else { if (slab) launch_lean<PVP_ACCUM_FIXED, true, false>(ctx, g, grid_dev, width, height); else launch_lean<PVP_ACCUM_FIXED, false, false>(ctx, g, grid_dev, width, height); }
    } else if (variant == 3) {
        if (f32) {
            auto k = dda_accumulate_kernel<PVP_ACCUM_F32>;
            k<<<k2_grid_size(ctx, (const void *)k), K2_THREADS, 0, ctx->stream>>>(ctx->queue, ctx->pairs_dev, g, (float *)grid_dev, width,
                                                                                   height, ctx->ctl_dev);
        } else {
            auto k = dda_accumulate_kernel<PVP_ACCUM_FIXED>;
            k<<<k2_grid_size(ctx, (const void *)k), K2_THREADS, 0, ctx->stream>>>(ctx->queue, ctx->pairs_dev, g,
                                                                                   (unsigned long long *)grid_dev, width, height, ctx->ctl_dev);
        }
    } else if (variant == 1) {
        if (f32) { if (slab) launch_agg<PVP_ACCUM_F32, 1, true>(ctx, g, grid_dev, width, height); else launch_agg<PVP_ACCUM_F32, 1, false>(ctx, g, grid_dev, width, height); }
        else { if (slab) launch_agg<PVP_ACCUM_FIXED, 1, true>(ctx, g, grid_dev, width, height); else launch_agg<PVP_ACCUM_FIXED, 1, false>(ctx, g, grid_dev, width, height); }
    } else {
        if (f32) { if (slab) launch_agg<PVP_ACCUM_F32, 2, true>(ctx, g, grid_dev, width, height); else launch_agg<PVP_ACCUM_F32, 2, false>(ctx, g, grid_dev, width, height); }
        else { if (slab) launch_agg<PVP_ACCUM_FIXED, 2, true>(ctx, g, grid_dev, width, height); else launch_agg<PVP_ACCUM_FIXED, 2, false>(ctx, g, grid_dev, width, height); }
    }
    PVP_CUDA(cudaGetLastError());
    return PVP_OK;
}

static int validate_pairs(const pvp_pair *pairs, int n_pairs, long long n_frames)
{
    for (int i = 0; i < n_pairs; ++i)
        PVP_REQUIRE(pairs[i].prev >= 0 && pairs[i].curr >= 0 && (n_frames < 0 || (pairs[i].prev < n_frames && pairs[i].curr < n_frames)),
                    "pair %d references frames (%d,%d) outside [0,%lld)", i, pairs[i].prev, pairs[i].curr, n_frames);
    return PVP_OK;
}

static int fetch_stats(pvp_ctx *ctx, pvp_stats *stats)
{
    if (!stats) return PVP_OK;
    Ctl h;
    PVP_CUDA(cudaMemcpyAsync(&h, ctx->ctl_dev, sizeof(Ctl), cudaMemcpyDeviceToHost, ctx->stream));
    PVP_CUDA(cudaMemsetAsync(&ctx->ctl_dev->n_rays, 0, 4 * sizeof(unsigned long long), ctx->stream));
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    stats->n_rays += h.n_rays; stats->n_steps += h.n_steps; stats->n_written += h.n_written; stats->n_atomics += h.n_atomics;
    return PVP_OK;
}

// One device-resident batch: K1 over all pairs, then K2 over the queue.  Batches are split so
// the worst-case queue (every pixel of every pair moves) fits the configured capacity.
static int run_batches(pvp_ctx *ctx, const void *frames_dev, int frame_dtype, long long frame_stride, int width, int height,
                       const pvp_pair *pairs, int n_pairs, float threshold, void *grid_dev, const GridP &g, int accum_mode)
{
    const long long n_pix = (long long)width * height;
    long long per_batch = std::max<long long>(1, ctx->opt_max_queue / n_pix);
    per_batch = std::min<long long>(per_batch, 65535);  // pair id is 16 bit in the queue
    for (int p0 = 0; p0 < n_pairs; p0 += (int)per_batch) {
        const int np = (int)std::min<long long>(per_batch, n_pairs - p0);
        int rc = ensure_queue(ctx, (unsigned long long)np * (unsigned long long)n_pix);
        if (rc) return rc;
        rc = upload_pairs(ctx, pairs + p0, np);
        if (rc) return rc;
        PVP_CUDA(cudaMemsetAsync(ctx->ctl_dev, 0, 3 * sizeof(unsigned long long), ctx->stream));  // count, head, n_paired
        // queue order (k1_row_group): 0 = auto: Z-order for the lock-step kernels (motion_compact_morton_kernel; blocks of 4 rows while an
        // x-plane of the grid stays within 2 MiB, of 8 rows above), row-major for variants 1-3; -4 / -8 = Z-order with blocks of that many
        // rows, 1 = plain row-major, 2 / 4 / 8 = the round-1 tiles of that many image rows (kept for A/B runs)
        const int ev = effective_variant(ctx);
        const int row_group = ctx->opt_k1_row_group != 0 ? (int)ctx->opt_k1_row_group
                              : (ev == 0 || ev >= 4) ? (plane_bytes(g, accum_mode) <= PLANE_LARGE ? -4 : -8) : 1;
        // snake order of the tile columns: default for tiles of 4 rows and more (k1_snake: -1 auto, 0 off, 1 on)
        const bool snake = ctx->opt_k1_snake >= 0 ? ctx->opt_k1_snake != 0 : row_group >= 4;
        rc = frame_dtype == PVP_FRAME_U8 ? launch_k1<uint8_t>(ctx, frames_dev, frame_stride, n_pix, np, threshold, width, height, row_group, snake)
                                         : launch_k1<float>(ctx, frames_dev, frame_stride, n_pix, np, threshold, width, height, row_group, snake);
        if (rc) return rc;
        // "integral" fixed-point sums: uint8 frames have integer diffs <= 255 (see LeanVal<FIXED, true>)
        const bool narrow_ok = frame_dtype == PVP_FRAME_U8;
        rc = launch_k2(ctx, g, accum_mode, narrow_ok, lean_geometry_ok(g, pairs + p0, np), grid_dev, width, height, (unsigned long long)np * (unsigned long long)n_pix);
        if (rc) return rc;
    }
    return PVP_OK;
}

static int check_common(pvp_ctx *ctx, const void *frames, int frame_dtype, long long frame_stride, int width, int height,
                        const pvp_pair *pairs, int n_pairs, void *grid_dev)
{
    PVP_REQUIRE(ctx != nullptr, "ctx is NULL");
    PVP_REQUIRE(n_pairs >= 0, "n_pairs < 0");
    PVP_REQUIRE(n_pairs == 0 || (frames && pairs), "frames / pairs is NULL");
    PVP_REQUIRE(frame_dtype == PVP_FRAME_U8 || frame_dtype == PVP_FRAME_F32, "bad frame_dtype %d", frame_dtype);
    PVP_REQUIRE(width > 0 && height > 0 && (long long)width * height < (1ll << 32), "bad frame size %dx%d", width, height);
    PVP_REQUIRE(frame_stride >= (long long)width * height, "frame_stride smaller than a frame");
    return PVP_OK;
}

}  // namespace pvp

using namespace pvp;

extern "C" {

int pvp_detect_motion(pvp_ctx *ctx, const void *prev_dev, const void *next_dev, int frame_dtype, int64_t n_pixels,
                      float threshold, uint8_t *changed_dev, float *diff_dev)
{
    PVP_REQUIRE(ctx && prev_dev && next_dev && changed_dev && diff_dev, "pvp_detect_motion: NULL argument");
    PVP_REQUIRE(frame_dtype == PVP_FRAME_U8 || frame_dtype == PVP_FRAME_F32, "pvp_detect_motion: bad frame_dtype %d", frame_dtype);
    if (n_pixels <= 0) return PVP_OK;
    DeviceGuard guard(ctx->device);
    int blocks = (int)std::min<long long>((n_pixels + 255) / 256, (long long)ctx->sm_count * 8);
    LaunchScope scope(ctx, 0);
    if (frame_dtype == PVP_FRAME_U8)
        detect_motion_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>((const uint8_t *)prev_dev, (const uint8_t *)next_dev, n_pixels,
                                                                       threshold, changed_dev, diff_dev);
    else
        detect_motion_kernel<float><<<blocks, 256, 0, ctx->stream>>>((const float *)prev_dev, (const float *)next_dev, n_pixels,
                                                                     threshold, changed_dev, diff_dev);
    PVP_CUDA(cudaGetLastError());
    return PVP_OK;
}

int pvp_pixel_rays(pvp_ctx *ctx, const uint32_t *pix_dev, int64_t n, int width, int height, const pvp_pair *pose, float *dirs_dev)
{
    PVP_REQUIRE(ctx && pix_dev && pose && dirs_dev, "pvp_pixel_rays: NULL argument");
    PVP_REQUIRE(width > 0 && height > 0, "pvp_pixel_rays: bad frame size");
    if (n <= 0) return PVP_OK;
    DeviceGuard guard(ctx->device);
    int blocks = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 8);
    LaunchScope scope(ctx, 0);
    pixel_rays_kernel<<<blocks, 256, 0, ctx->stream>>>(pix_dev, n, width, height, *pose, dirs_dev);
    PVP_CUDA(cudaGetLastError());
    return PVP_OK;
}

int pvp_cast_rays(pvp_ctx *ctx, const float *origins_dev, const float *dirs_dev, int64_t n, const pvp_grid_desc *grid,
                  int32_t *counts_dev, const int64_t *offsets_dev, int32_t *vox_dev, float *t_dev)
{
    PVP_REQUIRE(ctx && origins_dev && dirs_dev && counts_dev, "pvp_cast_rays: NULL argument");
    GridP g;
    int rc = make_gridp(grid, &g);
    if (rc) return rc;
    if (n <= 0) return PVP_OK;
    DeviceGuard guard(ctx->device);
    int blocks = (int)std::min<long long>((n + 127) / 128, (long long)ctx->sm_count * 16);
    LaunchScope scope(ctx, 0);
    cast_rays_kernel<<<blocks, 128, 0, ctx->stream>>>(origins_dev, dirs_dev, n, g, counts_dev, (const long long *)offsets_dev, vox_dev, t_dev);
    PVP_CUDA(cudaGetLastError());
    return PVP_OK;
}

int pvp_motion_accumulate(pvp_ctx *ctx, const void *frames_dev, int frame_dtype, int64_t frame_stride, int width, int height,
                          int n_frames, const pvp_pair *pairs, int n_pairs, float threshold, void *grid_dev, const pvp_grid_desc *grid,
                          pvp_stats *stats)
{
    int rc = check_common(ctx, frames_dev, frame_dtype, frame_stride, width, height, pairs, n_pairs, grid_dev);
    if (rc) return rc;
    GridP g;
    if ((rc = make_gridp(grid, &g))) return rc;
    PVP_REQUIRE(grid_dev != nullptr || g.slab_lo == g.slab_hi, "grid_dev is NULL");  // an empty slab stores nothing
    if ((rc = validate_pairs(pairs, n_pairs, n_frames))) return rc;
    DeviceGuard guard(ctx->device);
    if (n_pairs > 0) {
        rc = run_batches(ctx, frames_dev, frame_dtype, frame_stride, width, height, pairs, n_pairs, threshold, grid_dev, g, grid->accum_mode);
        if (rc) return rc;
        if ((rc = checked_verify(ctx, "pvp_motion_accumulate"))) return rc;
    }
    return fetch_stats(ctx, stats);
}

int pvp_motion_accumulate_host(pvp_ctx *ctx, const void *frames_host, int frame_dtype, int64_t frame_stride, int width, int height,
                               int n_frames, const pvp_pair *pairs, int n_pairs, float threshold, void *grid_dev,
                               const pvp_grid_desc *grid, pvp_stats *stats)
{
    int rc = check_common(ctx, frames_host, frame_dtype, frame_stride, width, height, pairs, n_pairs, grid_dev);
    if (rc) return rc;
    GridP g;
    if ((rc = make_gridp(grid, &g))) return rc;
    PVP_REQUIRE(grid_dev != nullptr || g.slab_lo == g.slab_hi, "grid_dev is NULL");
    if ((rc = validate_pairs(pairs, n_pairs, n_frames))) return rc;
    if (n_pairs == 0) return fetch_stats(ctx, stats);
    DeviceGuard guard(ctx->device);

    // Chunk the pair list; each chunk needs the frames [fmin, fmax] of its pairs on the device.
    // Two staging buffers: the copy stream fills buffer k+1 while the kernels of chunk k run.
    const size_t elem = frame_dtype == PVP_FRAME_U8 ? 1 : 4;
    const long long n_pix = (long long)width * height;
    const long long queue_pairs = std::max<long long>(1, ctx->opt_max_queue / n_pix);
    const int chunk_pairs = (int)std::min<long long>(std::min<long long>(queue_pairs, 65535), 64);

    struct Chunk { int p0, np, fmin, fmax; };
    std::vector<Chunk> chunks;
    size_t max_frames = 0;
    // chunk sizes ramp up (8, 16, 32, ... pairs): only the first, small copy is exposed; every later copy runs under the kernels of
    // the chunk before it
    int ramp = std::min(chunk_pairs, 8);
    for (int p0 = 0; p0 < n_pairs;) {
        Chunk c{p0, std::min(ramp, n_pairs - p0), INT32_MAX, -1};
        p0 += c.np;
        ramp = std::min(chunk_pairs, ramp * 2);
        for (int i = 0; i < c.np; ++i) {
            c.fmin = std::min(c.fmin, std::min(pairs[c.p0 + i].prev, pairs[c.p0 + i].curr));
            c.fmax = std::max(c.fmax, std::max(pairs[c.p0 + i].prev, pairs[c.p0 + i].curr));
        }
        max_frames = std::max(max_frames, (size_t)(c.fmax - c.fmin + 1));
        chunks.push_back(c);
    }
    const size_t need = max_frames * (size_t)frame_stride * elem;
    if (ctx->stage_bytes < need) {
        PVP_CUDA(cudaStreamSynchronize(ctx->stream));
        PVP_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        for (int i = 0; i < 2; ++i) {
            if (ctx->stage_dev[i]) cudaFree(ctx->stage_dev[i]);
            ctx->stage_dev[i] = nullptr;
        }
        ctx->stage_bytes = 0;
        for (int i = 0; i < 2; ++i) PVP_CUDA(cudaMalloc(&ctx->stage_dev[i], need));
        ctx->stage_bytes = need;
    }

    std::vector<pvp_pair> local;
    // order the copy stream behind whatever the caller already enqueued that touches the buffers
    PVP_CUDA(cudaEventRecord(ctx->ev_done[0], ctx->stream));
    PVP_CUDA(cudaEventRecord(ctx->ev_done[1], ctx->stream));
    for (size_t c = 0; c < chunks.size(); ++c) {
        const Chunk &ch = chunks[c];
        const int buf = (int)(c & 1);
        const size_t bytes = (size_t)(ch.fmax - ch.fmin + 1) * (size_t)frame_stride * elem;
        PVP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[buf], 0));  // buffer free again?
        PVP_CUDA(cudaMemcpyAsync(ctx->stage_dev[buf], (const char *)frames_host + (size_t)ch.fmin * (size_t)frame_stride * elem, bytes,
                                 cudaMemcpyHostToDevice, ctx->copy_stream));
        PVP_CUDA(cudaEventRecord(ctx->ev_copy[buf], ctx->copy_stream));
        PVP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[buf], 0));
        local.assign(pairs + ch.p0, pairs + ch.p0 + ch.np);
        for (auto &p : local) { p.prev -= ch.fmin; p.curr -= ch.fmin; }
        rc = run_batches(ctx, ctx->stage_dev[buf], frame_dtype, frame_stride, width, height, local.data(), ch.np, threshold, grid_dev, g,
                         grid->accum_mode);
        if (rc) return rc;
        PVP_CUDA(cudaEventRecord(ctx->ev_done[buf], ctx->stream));
    }
    if ((rc = checked_verify(ctx, "pvp_motion_accumulate_host"))) return rc;
    return fetch_stats(ctx, stats);
}

int pvp_fixed_to_f32(pvp_ctx *ctx, const int64_t *src_dev, float *dst_dev, int64_t n, int frac_bits)
{
    PVP_REQUIRE(ctx && src_dev && dst_dev && frac_bits >= 0 && frac_bits <= 62, "pvp_fixed_to_f32: bad argument");
    if (n <= 0) return PVP_OK;
    DeviceGuard guard(ctx->device);
    int blocks = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    LaunchScope scope(ctx, 0);
    fixed_to_f32_kernel<<<blocks, 256, 0, ctx->stream>>>((const long long *)src_dev, dst_dev, n, 1.0 / (double)(1ll << frac_bits));
    PVP_CUDA(cudaGetLastError());
    return PVP_OK;
}

int pvp_fixed_to_f64(pvp_ctx *ctx, const int64_t *src_dev, double *dst_dev, int64_t n, int frac_bits)
{
    PVP_REQUIRE(ctx && src_dev && dst_dev && frac_bits >= 0 && frac_bits <= 62, "pvp_fixed_to_f64: bad argument");
    if (n <= 0) return PVP_OK;
    DeviceGuard guard(ctx->device);
    int blocks = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    LaunchScope scope(ctx, 0);
    fixed_to_f64_kernel<<<blocks, 256, 0, ctx->stream>>>((const long long *)src_dev, dst_dev, n, 1.0 / (double)(1ll << frac_bits));
    PVP_CUDA(cudaGetLastError());
    return PVP_OK;
}

}  // extern "C"
