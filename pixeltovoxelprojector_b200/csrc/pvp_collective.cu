// pvp_collective.cu -- the one exchange step of the path (SURVEY.md 8e mode 1): the sum of the per-GPU private voxel
// grids at the end of a frame-sharded run, done by our own kernel over NVLink 5 / NVSwitch peer memory instead of a
// library collective.
//
// Every rank owns 1/world of the cells: it reads that slice of ALL ranks' grids -- through the NVSwitch multicast
// address with multimem.ld_reduce (the switch adds the copies in flight, NVLS: one 16-byte load returns the sum of
// all GPUs) or, without multicast, with plain peer loads in rank order -- and stores the sum into the destination
// rank's grid (or into every rank's grid through multimem.st / peer stores).  Each byte of the result crosses NVLink
// once; the slices of all ranks move concurrently, so the 512 MiB fp32 grid of the headline configuration is merged in
// roughly (512 MiB / world) / link bandwidth.  No rank waits on another inside the kernel: the caller brackets the
// launch with barriers (pixeltovoxelprojector_b200/distributed.py).
//
// fp32: the multicast reduction order is the switch's; for integer-valued addends (uint8 frames) and for the int64
// fixed-point grids the sum is exact in any order.  The peer-load path adds in rank order 0..world-1 (deterministic).
#include <algorithm>

#include "pvp_internal.h"

namespace pvp {

constexpr int MAX_PEERS = 16;
struct PeerPtrs { void *p[MAX_PEERS]; };

__device__ __forceinline__ float4 mc_ld_reduce(const float4 *mc)
{
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float4 *mc, float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long mc_ld_reduce(const unsigned long long *mc)
{
    unsigned long long v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(unsigned long long *mc, unsigned long long v)
{
    asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(mc), "l"(v) : "memory");
}
__device__ __forceinline__ float mc_ld_reduce(const float *mc)
{
    float v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float *mc, float v)
{
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}

__device__ __forceinline__ float4 vadd(float4 a, float4 b) { return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w)); }
__device__ __forceinline__ float vadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ unsigned long long vadd(unsigned long long a, unsigned long long b) { return a + b; }

// V = float4 | float | unsigned long long; elements [lo, hi) of V; dst < 0: every rank receives the sum.
// UNROLL independent loads per thread are in flight before the first add: NVLink round trips are ~2 us.
template <typename V, bool MC, int UNROLL>
__global__ void __launch_bounds__(256)
grid_reduce_peers_kernel(PeerPtrs peers, int world, int dst, V *__restrict__ mc, long long lo, long long hi)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; base < hi; base += stride * UNROLL) {
        V acc[UNROLL];
        if (MC) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (base + u * stride < hi) acc[u] = mc_ld_reduce(mc + base + u * stride);
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (base + u * stride < hi) acc[u] = reinterpret_cast<const V *>(peers.p[0])[base + u * stride];
            for (int r = 1; r < world; ++r) {
                V t[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u)
                    if (base + u * stride < hi) t[u] = reinterpret_cast<const V *>(peers.p[r])[base + u * stride];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u)
                    if (base + u * stride < hi) acc[u] = vadd(acc[u], t[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long i = base + u * stride;
            if (i >= hi) continue;
            if (dst >= 0) {
                reinterpret_cast<V *>(peers.p[dst])[i] = acc[u];
            } else if (MC) {
                mc_st(mc + i, acc[u]);
            } else {
                for (int r = 0; r < world; ++r) reinterpret_cast<V *>(peers.p[r])[i] = acc[u];
            }
        }
    }
}

template <typename V, bool MC>
static void launch_reduce(pvp_ctx *ctx, const PeerPtrs &pp, int world, int dst, void *mc, long long lo, long long hi)
{
    if (hi <= lo) return;
    // loads in flight = blocks x 256 threads x UNROLL x sizeof(V); options reduce_blocks_per_sm and reduce_unroll (4, 8).  Measured at 8
    // GPUs (tools/reduce_bench.py, profiles/r2_n8_reduce_*.log): the all-ranks variant does not care (512 MiB 1.20-1.23 ms, 4 GiB 9.0-9.3 ms
    // for every setting); the one-destination variant, whose results leave as peer stores into ONE GPU, is best with 8 loads per thread and
    // a grid of 16 CTAs per SM (512 MiB 1.44 -> 1.28 ms, 4 GiB 10.4 -> 9.1 ms)
    const bool to_one = dst >= 0;
    const int per_sm = ctx->opt_reduce_blocks_per_sm > 0 ? (int)ctx->opt_reduce_blocks_per_sm : to_one ? 16 : 8;
    const int unroll = ctx->opt_reduce_unroll > 0 ? (int)ctx->opt_reduce_unroll : to_one ? 8 : 4;
    const int blocks = (int)std::min<long long>((hi - lo + 1023) / 1024, (long long)ctx->sm_count * per_sm);
    ++ctx->prof.launches_other;
    if (unroll == 8) grid_reduce_peers_kernel<V, MC, 8><<<blocks, 256, 0, ctx->stream>>>(pp, world, dst, (V *)mc, lo, hi);
    else grid_reduce_peers_kernel<V, MC, 4><<<blocks, 256, 0, ctx->stream>>>(pp, world, dst, (V *)mc, lo, hi);
}

}  // namespace pvp

using namespace pvp;

extern "C" int pvp_grid_reduce_peers(pvp_ctx *ctx, void *const *peers, int world, int rank, int dst_rank, void *multicast, int64_t n_cells,
                                     int accum_mode)
{
    const bool pull_mode = ctx && ctx->opt_reduce_pull != 0;
    PVP_REQUIRE(ctx && peers, "pvp_grid_reduce_peers: NULL argument");
    PVP_REQUIRE(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world && dst_rank < world, "pvp_grid_reduce_peers: bad rank/world (%d/%d -> %d)",
                rank, world, dst_rank);
    PVP_REQUIRE(accum_mode == PVP_ACCUM_F32 || accum_mode == PVP_ACCUM_FIXED, "pvp_grid_reduce_peers: bad accum_mode %d", accum_mode);
    PVP_REQUIRE(n_cells >= 0, "pvp_grid_reduce_peers: n_cells < 0");
    PeerPtrs pp = {};
    bool aligned16 = ((uintptr_t)multicast % 16) == 0;
    for (int r = 0; r < world; ++r) {
        PVP_REQUIRE(peers[r] != nullptr, "pvp_grid_reduce_peers: peers[%d] is NULL", r);
        pp.p[r] = peers[r];
        aligned16 = aligned16 && ((uintptr_t)peers[r] % 16) == 0;
    }
    if (world == 1 || n_cells == 0) return PVP_OK;
    DeviceGuard guard(ctx->device);
    // Slice of this rank.  With the multicast address and ONE destination the destination pulls the whole grid through the
    // switch by itself (slice = everything on dst_rank, nothing elsewhere): the reduced cells cross NVLink once, into the
    // rank that wants them, instead of once into the reducing rank and once more into the destination.
    const bool dst_pulls = multicast != nullptr && dst_rank >= 0 && pull_mode;
    auto slice = [&](long long n, long long *lo, long long *hi) {
        if (dst_pulls) { *lo = 0; *hi = rank == dst_rank ? n : 0; }
        else { *lo = n * rank / world; *hi = n * (rank + 1) / world; }
    };
    long long lo, hi;
    if (accum_mode == PVP_ACCUM_F32) {
        if (aligned16 && n_cells % 4 == 0) {
            slice(n_cells / 4, &lo, &hi);
            if (multicast) launch_reduce<float4, true>(ctx, pp, world, dst_rank, multicast, lo, hi);
            else launch_reduce<float4, false>(ctx, pp, world, dst_rank, nullptr, lo, hi);
        } else {
            slice(n_cells, &lo, &hi);
            if (multicast) launch_reduce<float, true>(ctx, pp, world, dst_rank, multicast, lo, hi);
            else launch_reduce<float, false>(ctx, pp, world, dst_rank, nullptr, lo, hi);
        }
    } else {
        slice(n_cells, &lo, &hi);
        if (multicast) launch_reduce<unsigned long long, true>(ctx, pp, world, dst_rank, multicast, lo, hi);
        else launch_reduce<unsigned long long, false>(ctx, pp, world, dst_rank, nullptr, lo, hi);
    }
    PVP_CUDA(cudaGetLastError());
    return PVP_OK;
}
