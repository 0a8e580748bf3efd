// pvp_context.cu -- context, error reporting, pose helpers, .bin I/O and memory helpers of
// libpvp_b200.so (the parts of include/pvp_b200.h that launch no hot-path kernel).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>

#include <nvtx3/nvToolsExt.h>

#include "pvp_internal.h"

namespace pvp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    return e == cudaErrorMemoryAllocation ? PVP_ERR_NOMEM : PVP_ERR_CUDA;
}

static cudaEvent_t pool_get(pvp_ctx *ctx)
{
    if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

LaunchScope::LaunchScope(pvp_ctx *c, int k) : ctx(c), kind(k)
{
    uint64_t *cnt[4] = {&c->prof.launches_other, &c->prof.launches_k1, &c->prof.launches_k2, &c->prof.launches_k3};
    ++*cnt[k];
    static const char *const names[4] = {"pvp:stage kernel", "pvp:K1 diff/threshold/compact", "pvp:K2 DDA/accumulate", "pvp:K3 process_image"};
    nvtxRangePushA(names[k]);   // NVTX range per stage; a no-op unless a tool is attached
    if (c->profiling && k != 0) {
        a = pool_get(c); b = pool_get(c);
        cudaEventRecord(a, c->stream);
    }
}

LaunchScope::~LaunchScope()
{
    nvtxRangePop();
    if (a) {
        cudaEventRecord(b, ctx->stream);
        ctx->timed.push_back({a, b, kind});
    }
}

int checked_verify(pvp_ctx *ctx, const char *where)
{
#ifdef PVP_CHECKED
    Ctl h;
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    PVP_CUDA(cudaMemcpy(&h, ctx->ctl_dev, sizeof(Ctl), cudaMemcpyDeviceToHost));
    unsigned long long bad = 0;
    for (int i = 0; i < 8; ++i) bad += h.check[i];
    if (bad) {
        set_error("%s: checked build caught out-of-bounds indices: queue write %llu, staging %llu, queue read %llu, queue entry %llu, grid cell %llu, "
                  "image grid cell %llu, texture cell %llu, other %llu", where, h.check[0], h.check[1], h.check[2], h.check[3], h.check[4], h.check[5],
                  h.check[6], h.check[7]);
        PVP_CUDA(cudaMemset(&ctx->ctl_dev->check[0], 0, sizeof(h.check)));
        return PVP_ERR_CHECK;
    }
#else
    (void)ctx; (void)where;
#endif
    return PVP_OK;
}

int ensure_queue(pvp_ctx *ctx, unsigned long long entries)
{
    if (ctx->queue.capacity >= entries) return PVP_OK;
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->queue.pix) cudaFree(ctx->queue.pix);
    if (ctx->queue.diff) cudaFree(ctx->queue.diff);
    if (ctx->queue.pair) cudaFree(ctx->queue.pair);
    ctx->queue = {nullptr, nullptr, nullptr, 0};
    PVP_CUDA(cudaMalloc(&ctx->queue.pix, entries * sizeof(uint32_t)));
    PVP_CUDA(cudaMalloc(&ctx->queue.diff, entries * sizeof(float)));
    PVP_CUDA(cudaMalloc(&ctx->queue.pair, entries * sizeof(uint16_t)));
    ctx->queue.capacity = entries;
    return PVP_OK;
}

int ensure_pairs(pvp_ctx *ctx, int n_pairs)
{
    if (ctx->pairs_cap >= n_pairs) return PVP_OK;
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->pairs_dev) cudaFree(ctx->pairs_dev);
    if (ctx->pairs_pinned) cudaFreeHost(ctx->pairs_pinned);
    ctx->pairs_dev = nullptr; ctx->pairs_pinned = nullptr; ctx->pairs_cap = 0;
    int cap = n_pairs < 64 ? 64 : n_pairs;
    PVP_CUDA(cudaMalloc(&ctx->pairs_dev, sizeof(pvp_pair) * (size_t)cap));
    PVP_CUDA(cudaMallocHost(&ctx->pairs_pinned, sizeof(pvp_pair) * (size_t)cap * PVP_PAIR_SLOTS));
    ctx->pairs_cap = cap;
    return PVP_OK;
}

// Pair table -> device through a ring of pinned staging slots.  The copy is stream-ordered behind the kernels of the batch before
// it, so its event only fires when THAT batch has finished: with one slot the host would stall a whole batch of K2 time in every
// call (measured through the file pipeline on dense motion: 0.70 of 0.81 s); with the ring it runs PVP_PAIR_SLOTS - 1 batches ahead.
int upload_pairs(pvp_ctx *ctx, const pvp_pair *pairs, int n_pairs)
{
    int rc = ensure_pairs(ctx, n_pairs);
    if (rc) return rc;
    const int slot = (int)(ctx->pairs_seq++ % PVP_PAIR_SLOTS);
    pvp_pair *stage = ctx->pairs_pinned + (size_t)slot * (size_t)ctx->pairs_cap;
    PVP_CUDA(cudaEventSynchronize(ctx->ev_pairs[slot]));
    memcpy(stage, pairs, sizeof(pvp_pair) * (size_t)n_pairs);
    PVP_CUDA(cudaMemcpyAsync(ctx->pairs_dev, stage, sizeof(pvp_pair) * (size_t)n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    PVP_CUDA(cudaEventRecord(ctx->ev_pairs[slot], ctx->stream));
    return PVP_OK;
}

}  // namespace pvp

using namespace pvp;

extern "C" {

int pvp_version(void) { return PVP_VERSION; }

// bit 0: checked build (own bounds checks compiled into the kernels, see pvp_internal.h)
int pvp_build_flags(void)
{
#ifdef PVP_CHECKED
    return 1;
#else
    return 0;
#endif
}

const char *pvp_last_error(void) { return g_err; }

int pvp_ctx_destroy(pvp_ctx *ctx);

int pvp_ctx_create(int device, void *stream, pvp_ctx **out)
{
    PVP_REQUIRE(out != nullptr, "pvp_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("pvp_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return PVP_ERR_CUDA;
    }
    PVP_REQUIRE(device >= 0 && device < ndev, "pvp_ctx_create: device %d out of range [0,%d)", device, ndev);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    PVP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("pvp_ctx_create: device %d is sm_%d%d; libpvp_b200 is built for sm_100a only", device, prop.major,
                  prop.minor);
        return PVP_ERR_CUDA;
    }
    pvp_ctx *ctx = new (std::nothrow) pvp_ctx();
    if (!ctx) { set_error("pvp_ctx_create: out of host memory"); return PVP_ERR_NOMEM; }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    auto init = [&]() -> int {
        if (stream) {
            ctx->stream = (cudaStream_t)stream;
        } else {
            PVP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
            ctx->own_stream = true;
        }
        PVP_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        PVP_CUDA(cudaEventCreate(&ctx->ev_a));
        PVP_CUDA(cudaEventCreate(&ctx->ev_b));
        for (int i = 0; i < 2; ++i) {
            PVP_CUDA(cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
            PVP_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
        }
        for (auto &e : ctx->ev_pairs) PVP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        PVP_CUDA(cudaMalloc(&ctx->ctl_dev, sizeof(Ctl)));
        PVP_CUDA(cudaMemset(ctx->ctl_dev, 0, sizeof(Ctl)));
        return PVP_OK;
    };
    const int rc = init();
    if (rc != PVP_OK) {  // release whatever was created (the error text is already set)
        pvp_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return PVP_OK;
}

int pvp_ctx_destroy(pvp_ctx *ctx)
{
    if (!ctx) return PVP_OK;
    DeviceGuard guard(ctx->device);
    if (ctx->stream || ctx->own_stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->queue.pix) cudaFree(ctx->queue.pix);
    if (ctx->queue.diff) cudaFree(ctx->queue.diff);
    if (ctx->queue.pair) cudaFree(ctx->queue.pair);
    if (ctx->pairs_dev) cudaFree(ctx->pairs_dev);
    if (ctx->pairs_pinned) cudaFreeHost(ctx->pairs_pinned);
    if (ctx->fs_hring) cudaFreeHost(ctx->fs_hring);
    if (ctx->fs_dring) cudaFree(ctx->fs_dring);
    for (auto p : ctx->bin_pin) if (p) cudaFreeHost(p);
    for (int i = 0; i < 2; ++i) {
        if (ctx->stage_dev[i]) cudaFree(ctx->stage_dev[i]);
        if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    }
    if (ctx->ctl_dev) cudaFree(ctx->ctl_dev);
    for (auto &t : ctx->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    for (auto e : ctx->ev_pairs) if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PVP_OK;
}

int pvp_ctx_set_stream(pvp_ctx *ctx, void *stream)
{
    PVP_REQUIRE(ctx != nullptr, "pvp_ctx_set_stream: ctx is NULL");
    DeviceGuard guard(ctx->device);
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        PVP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return PVP_OK;
}

int pvp_ctx_sync(pvp_ctx *ctx)
{
    PVP_REQUIRE(ctx != nullptr, "pvp_ctx_sync: ctx is NULL");
    DeviceGuard guard(ctx->device);
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    return PVP_OK;
}

int pvp_ctx_profile(pvp_ctx *ctx, int enable)
{
    PVP_REQUIRE(ctx != nullptr, "pvp_ctx_profile: ctx is NULL");
    ctx->profiling = enable != 0;
    return PVP_OK;
}

int pvp_ctx_last_k2_pick(pvp_ctx *ctx, int *out_pick)
{
    PVP_REQUIRE(ctx && out_pick, "pvp_ctx_last_k2_pick: NULL argument");
    DeviceGuard guard(ctx->device);
    *out_pick = 0;
    if (!ctx->ctl_dev || !ctx->last_k2_lockstep) return PVP_OK;
    unsigned long long sel[3] = {0, 0, 0};
    PVP_CUDA(cudaMemcpyAsync(sel, ctx->ctl_dev->sel_count, sizeof(sel), cudaMemcpyDeviceToHost, ctx->stream));
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    *out_pick = sel[2] ? 4 : sel[1] ? 2 : sel[0] ? 1 : 0;
    return PVP_OK;
}

int pvp_ctx_get_profile(pvp_ctx *ctx, pvp_profile *out, int reset)
{
    PVP_REQUIRE(ctx && out, "pvp_ctx_get_profile: NULL argument");
    DeviceGuard guard(ctx->device);
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &t : ctx->timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
            if (t.kind == 1) ctx->prof.ms_k1 += ms; else if (t.kind == 2) ctx->prof.ms_k2 += ms; else ctx->prof.ms_k3 += ms;
        }
        ctx->ev_pool.push_back(t.a); ctx->ev_pool.push_back(t.b);
    }
    ctx->timed.clear();
    *out = ctx->prof;
    if (reset) ctx->prof = {0, 0, 0, 0, 0.0, 0.0, 0.0};
    return PVP_OK;
}

int pvp_ctx_set_option(pvp_ctx *ctx, const char *key, int64_t value)
{
    PVP_REQUIRE(ctx != nullptr && key != nullptr, "pvp_ctx_set_option: NULL argument");
    if (!strcmp(key, "dda_variant")) ctx->opt_dda_variant = value;
    else if (!strcmp(key, "max_queue")) { PVP_REQUIRE(value >= 1024, "max_queue must be >= 1024"); ctx->opt_max_queue = value; }
    else if (!strcmp(key, "ctas_per_sm")) ctx->opt_ctas_per_sm = value;
    else if (!strcmp(key, "lean_regs")) ctx->opt_lean_regs = value;
    else if (!strcmp(key, "image_variant")) ctx->opt_image_variant = value;
    else if (!strcmp(key, "reduce_pull")) ctx->opt_reduce_pull = value;
    else if (!strcmp(key, "reduce_unroll")) ctx->opt_reduce_unroll = value;
    else if (!strcmp(key, "reduce_blocks_per_sm")) ctx->opt_reduce_blocks_per_sm = value;
    else if (!strcmp(key, "image_tile_h")) ctx->opt_image_tile_h = value;
    else if (!strcmp(key, "k1_row_group")) ctx->opt_k1_row_group = value;
    else if (!strcmp(key, "k1_snake")) ctx->opt_k1_snake = value;
    else if (!strcmp(key, "lean2_paired")) ctx->opt_lean2_density = value;
    else if (!strcmp(key, "inject_fault")) ctx->opt_inject_fault = value;
    else if (!strcmp(key, "count_atomics")) ctx->opt_count_atomics = value;
    else if (!strcmp(key, "lean4")) ctx->opt_lean4 = value;
    else { set_error("pvp_ctx_set_option: unknown key '%s'", key); return PVP_ERR_INVALID; }
    return PVP_OK;
}

int pvp_dev_malloc(pvp_ctx *ctx, size_t bytes, void **out_dev)
{
    PVP_REQUIRE(ctx && out_dev, "pvp_dev_malloc: NULL argument");
    DeviceGuard guard(ctx->device);
    PVP_CUDA(cudaMalloc(out_dev, bytes));
    return PVP_OK;
}

int pvp_dev_free(pvp_ctx *ctx, void *dev)
{
    PVP_REQUIRE(ctx != nullptr, "pvp_dev_free: ctx is NULL");
    DeviceGuard guard(ctx->device);
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    PVP_CUDA(cudaFree(dev));
    return PVP_OK;
}

int pvp_dev_memset(pvp_ctx *ctx, void *dev, int value, size_t bytes)
{
    PVP_REQUIRE(ctx != nullptr, "pvp_dev_memset: ctx is NULL");
    DeviceGuard guard(ctx->device);
    PVP_CUDA(cudaMemsetAsync(dev, value, bytes, ctx->stream));
    return PVP_OK;
}

int pvp_memcpy_h2d(pvp_ctx *ctx, void *dst_dev, const void *src, size_t bytes)
{
    PVP_REQUIRE(ctx != nullptr, "pvp_memcpy_h2d: ctx is NULL");
    DeviceGuard guard(ctx->device);
    PVP_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    return PVP_OK;
}

int pvp_memcpy_d2h(pvp_ctx *ctx, void *dst, const void *src_dev, size_t bytes)
{
    PVP_REQUIRE(ctx != nullptr, "pvp_memcpy_d2h: ctx is NULL");
    DeviceGuard guard(ctx->device);
    PVP_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PVP_CUDA(cudaStreamSynchronize(ctx->stream));
    return PVP_OK;
}

int pvp_host_alloc_pinned(size_t bytes, void **out)
{
    PVP_REQUIRE(out != nullptr, "pvp_host_alloc_pinned: out is NULL");
    PVP_CUDA(cudaMallocHost(out, bytes));
    return PVP_OK;
}

int pvp_host_free_pinned(void *p)
{
    PVP_CUDA(cudaFreeHost(p));
    return PVP_OK;
}

// ---- pose: host code, glibc libm, fp32, no contraction (x86-64 baseline ISA has no FMA) ----

static inline float h_deg2rad(float deg) { return deg * 3.14159265358979323846f / 180.0f; }  // ray_voxel.cpp:60-62

// R = Rz(yaw) * Ry(roll) * Rx(pitch), each 3x3 product a left-to-right 3-term sum (ray_voxel.cpp:84-135).
void pvp_rotation_matrix(float yaw_deg, float pitch_deg, float roll_deg, float R[9])
{
    const float y = h_deg2rad(yaw_deg), p = h_deg2rad(pitch_deg), r = h_deg2rad(roll_deg);
    const float cy = cosf(y), sy = sinf(y), cr = cosf(r), sr = sinf(r), cp = cosf(p), sp = sinf(p);
    const float A[9] = {cy, -sy, 0.f, sy, cy, 0.f, 0.f, 0.f, 1.f};     // about z by yaw
    const float B[9] = {cr, 0.f, sr, 0.f, 1.f, 0.f, -sr, 0.f, cr};     // about y by ROLL
    const float Cm[9] = {1.f, 0.f, 0.f, 0.f, cp, -sp, 0.f, sp, cp};    // about x by PITCH
    float AB[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            volatile float s = A[i * 3] * B[j];
            s = s + A[i * 3 + 1] * B[3 + j];
            s = s + A[i * 3 + 2] * B[6 + j];
            AB[i * 3 + j] = s;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            volatile float s = AB[i * 3] * Cm[j];
            s = s + AB[i * 3 + 1] * Cm[3 + j];
            s = s + AB[i * 3 + 2] * Cm[6 + j];
            R[i * 3 + j] = s;
        }
}

float pvp_focal_length(float fov_degrees, int width)
{
    const float fov_rad = h_deg2rad(fov_degrees);
    return ((float)width * 0.5f) / tanf(fov_rad * 0.5f);  // ray_voxel.cpp:491
}

// ---- .bin ------------------------------------------------------------------

size_t pvp_grid_bytes(const pvp_grid_desc *g)
{
    if (!g || g->n <= 0 || g->slab_hi < g->slab_lo) return 0;
    size_t elem = g->accum_mode == PVP_ACCUM_FIXED ? 8 : 4;
    return (size_t)(g->slab_hi - g->slab_lo) * (size_t)g->n * (size_t)g->n * elem;
}

int pvp_write_bin(const char *path, int32_t n, float voxel_size, const float *grid_host)
{
    PVP_REQUIRE(path && grid_host && n > 0, "pvp_write_bin: bad argument");
    FILE *f = fopen(path, "wb");
    if (!f) { set_error("Cannot open output file: %s", path); return PVP_ERR_IO; }
    size_t count = (size_t)n * n * n;
    bool ok = fwrite(&n, sizeof(int32_t), 1, f) == 1 && fwrite(&voxel_size, sizeof(float), 1, f) == 1 &&
              fwrite(grid_host, sizeof(float), count, f) == count;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { set_error("pvp_write_bin: short write to %s", path); return PVP_ERR_IO; }
    return PVP_OK;
}

// The same writer fed from DEVICE memory: the N^3 float32 cells stream through two pinned buffers, the D2H copy of chunk k+1
// overlapping the fwrite of chunk k (a 512 MiB grid otherwise pays a zero-filled host vector and a pageable copy before the
// first byte reaches the file).  Synchronises the context stream.
int pvp_write_bin_dev(pvp_ctx *ctx, const char *path, int32_t n, float voxel_size, const float *grid_f32_dev)
{
    PVP_REQUIRE(ctx && path && grid_f32_dev && n > 0, "pvp_write_bin_dev: bad argument");
    DeviceGuard guard(ctx->device);
    FILE *f = fopen(path, "wb");
    if (!f) { set_error("Cannot open output file: %s", path); return PVP_ERR_IO; }
    nvtxRangePushA("pvp:.bin D2H + write");
    const size_t total = (size_t)n * n * n * sizeof(float);
    const size_t chunk = total < ((size_t)16 << 20) ? total : ((size_t)16 << 20);
    const size_t n_chunks = (total + chunk - 1) / chunk;
    uint8_t *pin[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    auto body = [&]() -> int {
        if (ctx->bin_pin_bytes < chunk) {   // the two pinned chunk buffers stay with the context
            for (auto &p : ctx->bin_pin) { if (p) cudaFreeHost(p); p = nullptr; }
            ctx->bin_pin_bytes = 0;
            for (auto &p : ctx->bin_pin) PVP_CUDA(cudaMallocHost((void **)&p, chunk));
            ctx->bin_pin_bytes = chunk;
        }
        for (int i = 0; i < 2; ++i) {
            pin[i] = ctx->bin_pin[i];
            PVP_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        auto issue = [&](size_t k) -> int {
            const size_t off = k * chunk, bytes = total - off < chunk ? total - off : chunk;
            PVP_CUDA(cudaMemcpyAsync(pin[k & 1], (const uint8_t *)grid_f32_dev + off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
            PVP_CUDA(cudaEventRecord(ev[k & 1], ctx->stream));
            return PVP_OK;
        };
        bool ok = fwrite(&n, sizeof(int32_t), 1, f) == 1 && fwrite(&voxel_size, sizeof(float), 1, f) == 1;
        int rc = issue(0);
        for (size_t k = 0; k < n_chunks && rc == PVP_OK; ++k) {
            if (k + 1 < n_chunks) rc = issue(k + 1);  // its buffer went to the file in the previous iteration
            if (rc != PVP_OK) break;
            PVP_CUDA(cudaEventSynchronize(ev[k & 1]));
            const size_t off = k * chunk, bytes = total - off < chunk ? total - off : chunk;
            ok = ok && fwrite(pin[k & 1], 1, bytes, f) == bytes;
        }
        if (rc != PVP_OK) return rc;
        if (!ok) { set_error("pvp_write_bin_dev: short write to %s", path); return PVP_ERR_IO; }
        return PVP_OK;
    };
    int rc = body();
    nvtxRangePop();
    cudaStreamSynchronize(ctx->stream);  // no copy may still target the pinned buffers when they are released
    for (int i = 0; i < 2; ++i)
        if (ev[i]) cudaEventDestroy(ev[i]);
    if (fclose(f) != 0 && rc == PVP_OK) { set_error("pvp_write_bin_dev: short write to %s", path); rc = PVP_ERR_IO; }
    return rc;
}

int pvp_read_bin_header(const char *path, int32_t *n, float *voxel_size)
{
    PVP_REQUIRE(path && n && voxel_size, "pvp_read_bin_header: NULL argument");
    FILE *f = fopen(path, "rb");
    if (!f) { set_error("pvp_read_bin_header: cannot open %s", path); return PVP_ERR_IO; }
    bool ok = fread(n, sizeof(int32_t), 1, f) == 1 && fread(voxel_size, sizeof(float), 1, f) == 1;
    fclose(f);
    if (!ok || *n <= 0) { set_error("pvp_read_bin_header: %s is not a voxel grid .bin", path); return PVP_ERR_FORMAT; }
    return PVP_OK;
}

int pvp_read_bin(const char *path, int32_t n, float *grid_host)
{
    PVP_REQUIRE(path && grid_host && n > 0, "pvp_read_bin: bad argument");
    FILE *f = fopen(path, "rb");
    if (!f) { set_error("pvp_read_bin: cannot open %s", path); return PVP_ERR_IO; }
    int32_t fn = 0; float vs = 0.f;
    size_t count = (size_t)n * n * n;
    bool ok = fread(&fn, sizeof(int32_t), 1, f) == 1 && fread(&vs, sizeof(float), 1, f) == 1 && fn == n &&
              fread(grid_host, sizeof(float), count, f) == count;
    fclose(f);
    if (!ok) { set_error("pvp_read_bin: %s truncated or N mismatch", path); return PVP_ERR_FORMAT; }
    return PVP_OK;
}

}  // extern "C"
