// pvp_internal.h -- host-side context shared by the translation units of libpvp_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/pvp_b200.h"

namespace pvp {

// Device-resident control block of the ray queue and the run counters.
struct Ctl {
    unsigned long long count;      // entries appended by the compaction kernel
    unsigned long long head;       // next entry to be claimed by the traversal kernel
    unsigned long long n_rays;
    unsigned long long n_steps;
    unsigned long long n_written;
    unsigned long long n_atomics;
    unsigned long long img_rays;   // path B
    unsigned long long img_samples;
};

// SoA ray queue: one entry per motion pixel of the batch.
struct Queue {
    uint32_t *pix;    // v*width + u
    float *diff;      // |prev - curr| (the value accumulated, ray_voxel.cpp:500)
    uint16_t *pair;   // index into the batch's pair table
    unsigned long long capacity;
};

}  // namespace pvp

struct pvp_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;  // H2D staging for the *_host entry points
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr};

    pvp::Ctl *ctl_dev = nullptr;
    pvp::Queue queue = {nullptr, nullptr, nullptr, 0};
    pvp_pair *pairs_dev = nullptr;
    pvp_pair *pairs_pinned = nullptr;
    int pairs_cap = 0;
    cudaEvent_t ev_pairs = nullptr;  // guards reuse of pairs_pinned

    void *stage_dev[2] = {nullptr, nullptr};  // frame staging buffers for *_host
    size_t stage_bytes = 0;

    // per-kernel accounting (pvp_ctx_profile / pvp_ctx_get_profile)
    bool profiling = false;
    pvp_profile prof = {0, 0, 0, 0, 0.0, 0.0, 0.0};
    struct Timed { cudaEvent_t a, b; int kind; };
    std::vector<Timed> timed;          // event pairs recorded since the last get_profile
    std::vector<cudaEvent_t> ev_pool;  // recycled events

    std::map<const void *, int> occupancy;  // resident CTAs per SM of the persistent kernels (cached occupancy query)

    // options (pvp_ctx_set_option)
    int64_t opt_dda_variant = 0;       // 0 = default
    int64_t opt_max_queue = 64ll << 20;  // ray-queue capacity in entries per batch
    int64_t opt_ctas_per_sm = 0;       // 0 = occupancy query
    int64_t opt_k1_snake = -1;         // K1 queue order inside a tile: odd columns bottom-up (-1 = auto: on for tiles of 4+ rows, 0 off, 1 on)
    int64_t opt_k1_row_group = 0;      // K1 queue order: tiles of this many image rows (0 = default of the K2 variant, 1 = row major, 2, 4, 8)
    int64_t opt_reduce_pull = 0;       // grid merge to one rank with multicast: 0 = every rank reduces a slice (default), 1 = the destination pulls the whole grid
    int64_t opt_image_tile_h = 0;      // lock-step path-B kernel: rows of a warp's pixel tile (0/2 default, 1, 4, 8)
    int64_t opt_image_variant = 0;     // path B kernel: 0 = auto, 1 = generic per-pixel, 2 = lock-step
    int64_t opt_lean_regs = 0;         // register budget of the lean K2 kernel: 0/72, 80 or 64
};

namespace pvp {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define PVP_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return pvp::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define PVP_REQUIRE(cond, ...)        \
    do {                              \
        if (!(cond)) {                \
            pvp::set_error(__VA_ARGS__); \
            return PVP_ERR_INVALID;   \
        }                             \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// RAII bracket around one kernel launch: counts it and, when profiling, times it with events.
struct LaunchScope {
    pvp_ctx *ctx; int kind; cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(pvp_ctx *c, int k);
    ~LaunchScope();
};

// pvp_imageio.cpp: `path` decoded to 8-bit gray (the bytes stbi_load(...,1) returns, ray_voxel.cpp:195) in ONE pass; false = the
// reference's failed stbi_load.  Never throws.
bool decode_gray_file(const char *path, std::vector<uint8_t> &gray, int *width, int *height);

int ensure_queue(pvp_ctx *ctx, unsigned long long entries);
int ensure_pairs(pvp_ctx *ctx, int n_pairs);
int upload_pairs(pvp_ctx *ctx, const pvp_pair *pairs, int n_pairs);

}  // namespace pvp
