// Host-logic replay of path A's frame loop WITHOUT a GPU (test infrastructure, built by tests/test_host_replay.py).
//
// The product's own orchestration sources (csrc/pvp_pipeline.cu: metadata, camera grouping, frame order, skip-on-error, size
// mismatch, prev<-curr, frame sharding with look-back, chunking, the decode pool; csrc/pvp_imageio.cpp: the decoder) are compiled
// as they are; only what they call on the device is replaced by a recorder: every frame pair handed to
// pvp_motion_accumulate is printed (checksums of the two frames, size, camera, pose).  The test compares that list with a
// Python model of ray_voxel.cpp:450-537.  Nothing here computes voxels.
#include "pvp_pipeline.cu"
#include "pvp_imageio.cpp"

#include <stdarg.h>
#include <zlib.h>

#include <map>

static char g_err[512];
void pvp::set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
int pvp::cuda_fail(cudaError_t, const char *, const char *, int) { return PVP_ERR_CUDA; }

extern "C" {
// CUDA runtime entry points the orchestration uses for its pinned ring, its device ring ("device memory" is plain host memory here) and
// the events that recycle their slots
// allocations are tracked so that every copy the frame loop enqueues can be range-checked: the destination must lie inside ONE live
// "device" allocation, a source inside the pinned ring must lie inside ONE live pinned allocation (anything else is pageable memory)
static std::map<const char *, size_t> g_dev, g_pin;
static bool inside(const std::map<const char *, size_t> &m, const void *p, size_t n)
{
    auto it = m.upper_bound((const char *)p);
    if (it == m.begin()) return false;
    --it;
    return (const char *)p >= it->first && (const char *)p + n <= it->first + it->second;
}
static bool touches(const std::map<const char *, size_t> &m, const void *p)
{
    auto it = m.upper_bound((const char *)p);
    if (it == m.begin()) return false;
    --it;
    return (const char *)p < it->first + it->second;
}
cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n); if (*p) g_pin[(const char *)*p] = n; return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { if (p && !g_pin.erase((const char *)p)) printf("BAD cudaFreeHost of an unknown pointer\n"); free(p); return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n); if (*p) g_dev[(const char *)*p] = n; return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { if (p && !g_dev.erase((const char *)p)) printf("BAD cudaFree of an unknown pointer\n"); free(p); return cudaSuccess; }
static cudaError_t checked_copy(void *d, const void *s, size_t n, cudaMemcpyKind kind)
{
    if (!inside(g_dev, d, n)) { printf("BAD copy: destination [%p, +%zu) is not inside a live device allocation\n", d, n); return cudaErrorInvalidValue; }
    if (kind == cudaMemcpyDeviceToDevice ? !inside(g_dev, s, n) : (touches(g_pin, s) && !inside(g_pin, s, n)) || touches(g_dev, s)) {
        printf("BAD copy: source [%p, +%zu) of kind %d is out of range\n", s, n, (int)kind);
        return cudaErrorInvalidValue;
    }
    memcpy(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind kind, cudaStream_t) { return checked_copy(d, s, n, kind); }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind kind) { return checked_copy(d, s, n, kind); }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { memset(a, 0, sizeof *a); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)1; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }

// the recorder: every batch the frame loop submits (frames = the device ring, pairs index its slots)
int pvp_motion_accumulate(pvp_ctx *, const void *frames, int dtype, int64_t stride, int w, int h, int n_frames, const pvp_pair *pairs,
                          int n_pairs, float threshold, void *, const pvp_grid_desc *, pvp_stats *)
{
    const size_t elem = dtype == PVP_FRAME_F32 ? 4 : 1;
    if (n_pairs > 0 && !inside(g_dev, frames, (size_t)n_frames * stride * elem)) printf("BAD frames: the ring handed to K1 is not a live device allocation\n");
    for (int i = 0; i < n_pairs; ++i) {
        const pvp_pair &p = pairs[i];
        if (p.prev < 0 || p.curr < 0 || p.prev >= n_frames || p.curr >= n_frames || p.prev == p.curr) { printf("BAD pair index %d %d of %d\n", p.prev, p.curr, n_frames); continue; }
        const unsigned long a = crc32(0, (const Bytef *)frames + (size_t)p.prev * stride * elem, (uInt)((size_t)w * h * elem));
        const unsigned long b = crc32(0, (const Bytef *)frames + (size_t)p.curr * stride * elem, (uInt)((size_t)w * h * elem));
        printf("PAIR %08lx %08lx %d %d cam %a %a %a ypr %a %a %a focal %a thr %a dtype %d\n", a, b, w, h, p.cam[0], p.cam[1], p.cam[2], p.R[0], p.R[1],
               p.R[2], p.focal, threshold, dtype);
    }
    return PVP_OK;
}
int pvp_motion_accumulate_host(pvp_ctx *, const void *, int, int64_t, int, int, int, const pvp_pair *, int, float, void *, const pvp_grid_desc *, pvp_stats *) { return PVP_OK; }
// pose helpers: transparent stand-ins, so the record shows WHICH frame's pose went into a pair (ray_voxel.cpp:486-491: the current one)
void pvp_rotation_matrix(float yaw, float pitch, float roll, float R[9]) { for (int i = 0; i < 9; ++i) R[i] = 0.f; R[0] = yaw; R[1] = pitch; R[2] = roll; }
float pvp_focal_length(float fov_degrees, int width) { return fov_degrees * 65536.f + (float)width; }
// referenced by pvp_ray_voxel_main only (not run here)
int pvp_ctx_create(int, void *, pvp_ctx **) { return PVP_ERR_CUDA; }
int pvp_ctx_destroy(pvp_ctx *) { return PVP_OK; }
const char *pvp_last_error(void) { return g_err; }
int pvp_dev_malloc(pvp_ctx *, size_t, void **) { return PVP_ERR_CUDA; }
int pvp_dev_free(pvp_ctx *, void *) { return PVP_OK; }
int pvp_dev_memset(pvp_ctx *, void *, int, size_t) { return PVP_OK; }
size_t pvp_grid_bytes(const pvp_grid_desc *) { return 0; }
int pvp_fixed_to_f32(pvp_ctx *, const int64_t *, float *, int64_t, int) { return PVP_OK; }
int pvp_write_bin_dev(pvp_ctx *, const char *, int32_t, float, const float *) { return PVP_OK; }
}

// replay <metadata.json> <folder> <decode_threads> <chunk_frames> <shard_rank> <shard_count>
int main(int argc, char **argv)
{
    if (argc != 7) return 2;
    pvp_ctx ctx;
    pvp_run_options opt;
    pvp_run_options_default(&opt);
    opt.verbose = 0;
    opt.decode_threads = atoi(argv[3]);
    opt.chunk_frames = atoi(argv[4]);
    opt.shard_rank = atoi(argv[5]);
    opt.shard_count = atoi(argv[6]);
    int cell = 0;
    pvp_stats st = {0, 0, 0, 0};
    pvp_run_report rep;
    memset(&rep, 0, sizeof rep);
    int rc = pvp_ray_voxel_accumulate(&ctx, argv[1], argv[2], &cell, &opt, &st, &rep);
    if (rc != PVP_OK) { printf("ERROR %d %s\n", rc, g_err); return 1; }
    // a second run on the same context reuses the rings the first one left behind; its record must be the same (printed after AGAIN)
    printf("AGAIN\n");
    memset(&rep, 0, sizeof rep);
    rc = pvp_ray_voxel_accumulate(&ctx, argv[1], argv[2], &cell, &opt, &st, &rep);
    if (rc != PVP_OK) { printf("ERROR %d %s\n", rc, g_err); return 1; }
    printf("REPORT listed %llu loaded %llu skipped %llu pairs %llu mismatch %llu\n", (unsigned long long)rep.n_frames_listed,
           (unsigned long long)rep.n_frames_loaded, (unsigned long long)rep.n_frames_skipped, (unsigned long long)rep.n_pairs,
           (unsigned long long)rep.n_size_mismatch);
    return 0;
}
