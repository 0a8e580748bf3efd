// pvp_microbench.cu -- the roofline denominator for K2/K3: measured device throughput of
// independent no-return atomic adds (RED) into a buffer, for several address patterns
// ("the builder must ship this micro-benchmark").
#include <algorithm>

#include "pvp_internal.h"

namespace pvp {

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <typename T> __device__ __forceinline__ T one();
template <> __device__ __forceinline__ float one<float>() { return 1.0f; }
template <> __device__ __forceinline__ double one<double>() { return 1.0; }
template <> __device__ __forceinline__ unsigned long long one<unsigned long long>() { return 1ull; }

// Each lane issues `iters` REDs; address stream per pattern (see include/pvp_b200.h).
template <typename T>
__global__ void __launch_bounds__(256)
atomic_bench_kernel(T *__restrict__ buf, unsigned long long n_elems, int pattern, int iters)
{
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31, warp = tid >> 5;
    constexpr unsigned per_sector = 32 / sizeof(T), per_line = 128 / sizeof(T);
    uint32_t s = mix32(tid * 2654435761u + 12345u);
    const T v = one<T>();
    for (int it = 0; it < iters; ++it) {
        unsigned long long idx;
        if (pattern == 0) {
            s = mix32(s + 0x9e3779b9u);
            idx = ((unsigned long long)s * n_elems) >> 32;
        } else if (pattern == 1) {
            uint32_t h = mix32((warp * 4u + lane / per_sector) * 0x9e3779b9u + (uint32_t)it * 0x85ebca6bu);
            idx = ((((unsigned long long)h * (n_elems / per_sector)) >> 32) * per_sector) + (lane % per_sector);
        } else if (pattern == 2) {
            uint32_t h = mix32(warp * 0x9e3779b9u + (uint32_t)it * 0x85ebca6bu);
            idx = ((unsigned long long)h * n_elems) >> 32;
        } else {
            uint32_t h = mix32(warp * 0x9e3779b9u + (uint32_t)it * 0x85ebca6bu);
            idx = ((((unsigned long long)h * (n_elems / per_line)) >> 32) * per_line) + (lane % per_line);
        }
        atomicAdd(buf + idx, v);
    }
}

}  // namespace pvp

using namespace pvp;

extern "C" int pvp_atomic_microbench(pvp_ctx *ctx, void *buf_dev, size_t buf_bytes, int elem, int pattern, int64_t n_ops,
                                     float *elapsed_ms)
{
    PVP_REQUIRE(ctx && buf_dev && elapsed_ms, "pvp_atomic_microbench: NULL argument");
    PVP_REQUIRE(elem == 4 || elem == 8 || elem == 9, "pvp_atomic_microbench: elem must be 4, 8 or 9");
    PVP_REQUIRE(pattern >= 0 && pattern <= 3, "pvp_atomic_microbench: bad pattern");
    PVP_REQUIRE(buf_bytes >= 4096 && n_ops > 0, "pvp_atomic_microbench: bad sizes");
    DeviceGuard guard(ctx->device);
    const int blocks = ctx->sm_count * 8, threads = 256;
    const long long lanes = (long long)blocks * threads;
    const int iters = (int)std::max<long long>(1, n_ops / lanes);
    PVP_CUDA(cudaEventRecord(ctx->ev_a, ctx->stream));
    ++ctx->prof.launches_other;
    if (elem == 4)
        atomic_bench_kernel<float><<<blocks, threads, 0, ctx->stream>>>((float *)buf_dev, buf_bytes / 4, pattern, iters);
    else if (elem == 8)
        atomic_bench_kernel<unsigned long long><<<blocks, threads, 0, ctx->stream>>>((unsigned long long *)buf_dev, buf_bytes / 8, pattern, iters);
    else
        atomic_bench_kernel<double><<<blocks, threads, 0, ctx->stream>>>((double *)buf_dev, buf_bytes / 8, pattern, iters);
    PVP_CUDA(cudaGetLastError());
    PVP_CUDA(cudaEventRecord(ctx->ev_b, ctx->stream));
    PVP_CUDA(cudaEventSynchronize(ctx->ev_b));
    PVP_CUDA(cudaEventElapsedTime(elapsed_ms, ctx->ev_a, ctx->ev_b));
    return PVP_OK;
}

extern "C" int64_t pvp_atomic_microbench_ops(pvp_ctx *ctx, int64_t n_ops)
{
    if (!ctx) return 0;
    const long long lanes = (long long)ctx->sm_count * 8 * 256;
    return lanes * std::max<long long>(1, n_ops / lanes);
}
