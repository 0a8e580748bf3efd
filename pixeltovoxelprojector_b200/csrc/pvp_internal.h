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
    unsigned long long n_paired;   // motion pixels of the batch whose vertical partner moves too (Z-order K1; picks lean2 / lean)
    unsigned long long n_rays;
    unsigned long long n_steps;
    unsigned long long n_written;
    unsigned long long n_atomics;
    unsigned long long img_rays;   // path B
    unsigned long long img_samples;
    unsigned long long n_pairs;    // checked build: pairs of the batch (written by K1)
    unsigned long long check[8];   // checked build: violations per PVP_CHK_* site class
    // the lock-step K2 kernels (lean / lean2 / lean4 = slots 0 / 1 / 2) each read their own copy of the queue length and claim chunks from
    // their own head: select_lockstep_kernel gives the whole queue to the one kernel that suits the batch and 0 to the others, which then
    // leave through their normal "queue is empty" exit.  (A pick tested at the top of a kernel, with an early return, costs the four-ray
    // kernel 6 instructions of its 106-instruction loop: ptxas lays the loop out differently.)
    unsigned long long sel_count[3];
    unsigned long long sel_head[3];
};

// ---- checked build (make checked -> libpvp_b200_checked.so, -DPVP_CHECKED) ---------------------------------------------------------
// compute-sanitizer is not available on the GPU pool this project runs on, so the library carries its own memcheck for the indices it
// computes: every queue slot K1 writes, every shared-memory staging slot, every queue entry K2 reads (and what it decodes from it), and
// every grid / texture cell address K2 and K3 hand to a RED is tested against the buffer it must lie in.  A violating access is
// SKIPPED and counted in Ctl::check; the host turns a non-zero count into PVP_ERR_CHECK after the call.  In the normal build the macro
// is the constant `true` and the kernels' SASS is unchanged (tests/test_profiles_fresh.py pins the md5).
enum { PVP_CHK_QUEUE_WRITE = 0, PVP_CHK_STAGE = 1, PVP_CHK_QUEUE_READ = 2, PVP_CHK_ENTRY = 3, PVP_CHK_GRID = 4, PVP_CHK_IMG_GRID = 5, PVP_CHK_IMG_TEX = 6, PVP_CHK_OTHER = 7 };
#ifdef PVP_CHECKED
#define PVP_DEV_CHECK(ctl, id, cond) ((cond) ? true : (atomicAdd(&(ctl)->check[id], 1ull), false))
#else
#define PVP_DEV_CHECK(ctl, id, cond) true
#endif

// SoA ray queue: one entry per motion pixel of the batch.
struct Queue {
    uint32_t *pix;    // v*width + u
    float *diff;      // |prev - curr| (the value accumulated, ray_voxel.cpp:500)
    uint16_t *pair;   // index into the batch's pair table
    unsigned long long capacity;
};

}  // namespace pvp

constexpr int PVP_PAIR_SLOTS = 4;

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
    cudaEvent_t ev_pairs[4] = {nullptr, nullptr, nullptr, nullptr};  // guard reuse of the pinned pair slots (PVP_PAIR_SLOTS)
    unsigned long long pairs_seq = 0;

    // buffers of the file pipeline (pvp_ray_voxel_accumulate) and of the .bin writer, kept across calls: pinning 75 MB costs ~25 ms
    uint8_t *fs_hring = nullptr; size_t fs_hring_bytes = 0;   // pinned frame ring
    uint8_t *fs_dring = nullptr; size_t fs_dring_bytes = 0;   // device frame ring
    uint8_t *bin_pin[2] = {nullptr, nullptr}; size_t bin_pin_bytes = 0;

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
    int64_t opt_reduce_unroll = 0;     // grid merge kernel: independent 16-byte loads per thread (0/4 default, 8)
    int64_t opt_reduce_blocks_per_sm = 0;  // grid merge kernel: CTAs per SM (0 = 8)
    int64_t opt_reduce_pull = 0;       // grid merge to one rank with multicast: 0 = every rank reduces a slice (default), 1 = the destination pulls the whole grid
    int64_t opt_image_tile_h = 0;      // lock-step path-B kernel: rows of a warp's pixel tile (0/2 default, 1, 4, 8)
    int64_t opt_image_variant = 0;     // path B kernel: 0 = auto, 1 = generic per-pixel, 2 = lock-step
    int64_t opt_lean2_density = 0;     // auto K2 ("lean2_paired"): lean2 from this percentage of the rays in vertical pairs up (0 = default 50), lean below
    int64_t opt_inject_fault = 0;      // checked build only: 1 = K1 is told a quarter of the queue's capacity, 2 = K3 half of the grid's span (tests of the checks themselves)
    int64_t opt_lean_regs = 0;         // register budget of the lean K2 kernel: 0/72, 80 or 64
    bool last_k2_lockstep = false;     // the last K2 batch went through select_lockstep_kernel (pvp_ctx_last_k2_pick)
    int64_t opt_lean4 = 0;             // auto K2: -1 = never pick lean4 (the round-2 pair lean / lean2 only)
    int64_t opt_count_atomics = 0;     // lean4 (four rays per lane): 1 = the instantiation that counts its REDs (pvp_stats.n_atomics); 0 = n_atomics is 0 for that kernel
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

// Same decode straight into `dst` (capacity cap bytes; a pinned staging slot): 1 = ok, 0 = failed, -1 = ok but larger than cap (pixels in spill).
int decode_gray_file_into(const char *path, uint8_t *dst, size_t cap, int *width, int *height, std::vector<uint8_t> &spill);

int ensure_queue(pvp_ctx *ctx, unsigned long long entries);
// checked build: synchronise, read Ctl::check and fail with PVP_ERR_CHECK when a kernel skipped an out-of-bounds access (no-op otherwise)
int checked_verify(pvp_ctx *ctx, const char *where);
int ensure_pairs(pvp_ctx *ctx, int n_pairs);
int upload_pairs(pvp_ctx *ctx, const pvp_pair *pairs, int n_pairs);

}  // namespace pvp
