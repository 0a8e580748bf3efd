// Micro-benchmark: what does redux.sync.add cost when the lanes of a warp form g disjoint groups (each lane passes its own group's
// member mask), against the 5-level segmented shuffle scan the lock-step kernels use?  (round 2, K2 instruction diet: could run sums
// come from one REDUX instead of 5 SHFL + 5 FADD + 5 LOP3?)     nvcc -arch=sm_100a -O3 -o redux_bench redux_bench.cu && ./redux_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(unsigned *out, int iters, int group)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned first = lane / group * group;
    const unsigned mask = group == 32 ? 0xffffffffu : (((1u << group) - 1u) << first);
    unsigned heads = 0;
    for (unsigned l = 0; l < 32; l += group) heads |= 1u << l;
    unsigned v = lane + 1, acc = 0;
    float fv = (float)(lane + 1), facc = 0.f;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {          // redux with per-group masks
            acc += __reduce_add_sync(mask, v);
            v = (v * 3u + acc) & 255u;
        } else if (MODE == 1) {   // segmented scan, 5 levels (float), as in the lean kernels
            float s = fv, t;
            t = __shfl_up_sync(0xffffffffu, s, 1);  if (lane >= first + 1)  s += t;
            t = __shfl_up_sync(0xffffffffu, s, 2);  if (lane >= first + 2)  s += t;
            t = __shfl_up_sync(0xffffffffu, s, 4);  if (lane >= first + 4)  s += t;
            t = __shfl_up_sync(0xffffffffu, s, 8);  if (lane >= first + 8)  s += t;
            t = __shfl_up_sync(0xffffffffu, s, 16); if (lane >= first + 16) s += t;
            facc += s;
            fv = fv * 0.5f + (float)(i & 3);
        } else {                  // empty loop body with the same dependent update
            acc += v;
            v = (v * 3u + acc) & 255u;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (unsigned)facc + heads;
}

template <int MODE>
float run(int group, int iters)
{
    unsigned *out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148 * 8, 256>>>(out, iters, group);
    cudaEventRecord(a);
    k<MODE><<<148 * 8, 256>>>(out, iters, group);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaFree(out);
    return ms;
}

int main()
{
    const int iters = 20000;
    const double warps = 148.0 * 8 * 8;
    printf("per warp-iteration, ns (148 x 8 CTAs x 256 threads, %d iterations); 1965 MHz: 1 cycle = 0.51 ns; per SM sub-partition 16 warps resident\n", iters);
    const float base = run<2>(32, iters);
    printf("empty loop: %.3f ms\n", base);
    for (int g : {32, 16, 8, 4, 2, 1}) {
        const float r = run<0>(g, iters), s = run<1>(g, iters);
        printf("groups of %2d lanes (%2d groups): redux %.3f ms (%.2f ns per warp-iter per SM-quarter slot), scan %.3f ms\n", g, 32 / g, r,
               (r - base) * 1e6 / iters / (warps / (148 * 4)), s);
    }
    return 0;
}
