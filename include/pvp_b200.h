/*
 * pvp_b200.h -- C ABI of libpvp_b200.so, the B200 (sm_100a) implementation of the
 * Pixeltovoxelprojector hot path.  This header is the drop-in boundary: plain
 * pointers and sizes, no torch / pybind types.  Every entry point names the reference
 * interface it replaces (file:line under the reference repo).
 *
 * Conventions
 *   - every function returns PVP_OK (0) or a negative pvp_status; pvp_last_error()
 *     gives the message of the calling thread's last failure;
 *   - `*_dev` pointers are CUDA device pointers on the context's device; everything
 *     else is host memory;
 *   - all device work is enqueued on the context's stream; entry points that return
 *     host-visible results (stats, host copies) synchronise that stream before returning,
 *     the others are asynchronous;
 *   - there is NO CPU fallback: without a usable CUDA device pvp_ctx_create fails;
 *   - threading: a pvp_ctx is not thread-safe (one context per host thread, or external locking; the reference
 *     extension likewise holds the GIL for a whole call, process_image.cpp:369-390); distinct contexts may run
 *     concurrently; the host-only helpers (pose, metadata, image decode, .bin) are re-entrant;
 *   - the entry points that parse files or run the frame loop (metadata, image decode, pvp_ray_voxel_accumulate) catch every
 *     C++ exception and report it as a status (allocation failures: PVP_ERR_NOMEM); the others allocate only small bookkeeping.
 */
#ifndef PVP_B200_H
#define PVP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVP_VERSION 100

typedef enum pvp_status {
    PVP_OK = 0,
    PVP_ERR_INVALID = -1,   /* bad argument */
    PVP_ERR_CUDA = -2,      /* CUDA runtime / launch failure */
    PVP_ERR_NOMEM = -3,     /* host or device allocation failed */
    PVP_ERR_IO = -4,        /* file could not be opened / read / written */
    PVP_ERR_FORMAT = -5,    /* malformed .bin / image / metadata */
    PVP_ERR_CHECK = -6      /* checked build only (libpvp_b200_checked.so): a kernel computed an out-of-bounds index (skipped and counted) */
} pvp_status;

/* frame element types accepted by the motion pipeline */
#define PVP_FRAME_U8 0   /* 8-bit gray, converted exactly like ray_voxel.cpp:213 (noise-free) */
#define PVP_FRAME_F32 1  /* float gray, i.e. the contents of ImageGray::pixels (ray_voxel.cpp:183-187) */

/* accumulation modes of the voxel grid */
#define PVP_ACCUM_F32 0   /* float32 grid, red.global.add.f32 -- reference arithmetic (ray_voxel.cpp:528) */
#define PVP_ACCUM_FIXED 1 /* int64 grid, value * 2^frac_bits rounded to nearest-even: order independent */
#define PVP_ACCUM_F64 0   /* path B: float64 grid / texture (process_image.cpp:357) */

typedef struct pvp_ctx pvp_ctx; /* device + stream + workspace (ray queue, counters) */

/*
 * The voxel grid of ray_voxel.cpp:434-440, with the compile-time constants of :434-437
 * turned into fields.  Storage is [x][y][z], z fastest (ray_voxel.cpp:527).  A grid
 * pointer may address only the x-planes [slab_lo, slab_hi) ("slab shard", SURVEY 8e mode 2):
 * rays are walked from their start with full-grid arithmetic and nothing is written outside the slab; the default kernels
 * clip exactly (a ray that cannot reach the slab is not walked, a walk stops once x has passed it).
 */
typedef struct pvp_grid_desc {
    int32_t n;            /* N (ray_voxel.cpp:434) */
    float voxel_size;     /* ray_voxel.cpp:435 */
    float center[3];      /* grid_center (ray_voxel.cpp:437) */
    int32_t accum_mode;   /* PVP_ACCUM_F32 | PVP_ACCUM_FIXED */
    int32_t frac_bits;    /* fixed-point fractional bits (0..40), ignored for F32 */
    int32_t slab_lo;      /* first x-plane stored at grid_dev */
    int32_t slab_hi;      /* one past the last x-plane stored (n for a whole grid) */
    float attenuation_alpha; /* 0 (default) = reference behaviour: the distance attenuation of ray_voxel.cpp:524-525 is computed and
                                discarded (:526).  > 0 = opt-in extension: every voxel step adds
                                pix_val * 1/(1 + alpha * distance), distance = RayStep::distance, all in IEEE fp32 */
} pvp_grid_desc;

/*
 * One consecutive-frame pair of one camera: what the loop body ray_voxel.cpp:484-531
 * consumes.  Pose is the CURRENT frame's (ray_voxel.cpp:486-491).  prev / curr index
 * the frame array handed to the same call.
 */
typedef struct pvp_pair {
    float cam[3];   /* FrameInfo::camera_position (ray_voxel.cpp:50) */
    float R[9];     /* rotation_matrix_yaw_pitch_roll, row major (ray_voxel.cpp:84-135) */
    float focal;    /* focal_len (ray_voxel.cpp:491) */
    int32_t prev;   /* frame index of the previous image */
    int32_t curr;   /* frame index of the current image */
} pvp_pair;

typedef struct pvp_stats {
    uint64_t n_rays;    /* pixels that passed the motion test and cast a ray (ray_voxel.cpp:496-503) */
    uint64_t n_steps;   /* RayStep count = voxel-steps of the reference (ray_voxel.cpp:373,523).  Slab grids: the steps THIS call walked --
                         * rays that cannot reach the slab are not walked and a walk stops once x has passed the slab -- so <= the whole-grid
                         * count and >= n_written (the default kernels; dda_variant 1-3 walk every ray in full) */
    uint64_t n_written; /* grid updates = voxel steps inside the grid's x-planes (== n_steps for a whole grid; over the slabs of a grid
                         * they add up to the whole-grid n_steps) */
    uint64_t n_atomics; /* global atomic operations actually issued (<= n_written after aggregation); lean4 counts them only with option "count_atomics" */
} pvp_stats;

/* ---- library / context ---------------------------------------------------- */
int pvp_version(void);
/* bit 0 set: this is the checked build (libpvp_b200_checked.so: the kernels test every index they compute, PVP_ERR_CHECK) */
int pvp_build_flags(void);
const char *pvp_last_error(void);
/* stream: a cudaStream_t to enqueue on (e.g. torch's current stream) or NULL for a private non-blocking one;
 * pass cudaStreamLegacy ((cudaStream_t)0x1) to name the legacy default stream, whose own handle is NULL */
int pvp_ctx_create(int device, void *stream, pvp_ctx **out);
int pvp_ctx_destroy(pvp_ctx *ctx);
int pvp_ctx_set_stream(pvp_ctx *ctx, void *stream);
int pvp_ctx_sync(pvp_ctx *ctx);
/* Tuning / experiment switches (all default to "auto"; none changes results, only which kernel or queue order is used -- DESIGN.md 4):
 *   "dda_variant"   0 auto | 1 scan aggregation | 2 match.any | 3 one RED per step (any grid size) | 4 lean | 5 lean2 | 6 lean4
 *   "lean4"         0 auto | -1 the auto mode never picks lean4                              "lean2_paired" percent of rays in vertical pairs from which lean2 / lean4 run (50)
 *   "count_atomics" 1: lean4 counts its REDs (pvp_stats.n_atomics; 4 instructions of its loop, ~5 %); 0 (default): lean4 reports n_atomics = 0
 *   "k1_row_group"  0 auto | 1 row-major queue | 2, 4, 8 image rows per queue tile        "k1_snake"  -1 auto | 0 | 1
 *   "lean_regs"     0 auto | 64, 80, 128 register budget of the lock-step kernels         "ctas_per_sm" 0 = by occupancy
 *   "max_queue"     ray queue capacity in entries (>= 1024)
 *   "image_variant" 0 auto | 1 generic per-pixel kernel | 2 lock-step, one pixel per lane | 3 two pixels per lane | 4 three pixels per lane
 *   "image_tile_h"  tile height of image_variant 2 (1, 2, 4, 8)                           "reduce_pull" pvp_grid_reduce_peers strategy
 * unknown key -> PVP_ERR_INVALID */
int pvp_ctx_set_option(pvp_ctx *ctx, const char *key, int64_t value);

/*
 * Per-kernel accounting for bench.py: launch counts are always kept; with profiling enabled every
 * K1/K2/K3 launch is bracketed by CUDA events on the context stream (live timing, no profiler).
 */
typedef struct pvp_profile {
    uint64_t launches_k1, launches_k2, launches_k3, launches_other; /* kernels launched since the last reset */
    double ms_k1, ms_k2, ms_k3;                                     /* summed event time (profiling enabled only) */
} pvp_profile;
int pvp_ctx_profile(pvp_ctx *ctx, int enable);
/* Which lock-step kernel the device-side pick gave the LAST batch to: 1 lean (one ray per lane), 2 lean2, 4 lean4; 0 when the last
 * batch ran a non-lock-step variant (or nothing ran yet).  Diagnostics for bench.py / the tests; synchronises the stream. */
int pvp_ctx_last_k2_pick(pvp_ctx *ctx, int *out_pick);
int pvp_ctx_get_profile(pvp_ctx *ctx, pvp_profile *out, int reset); /* synchronises the stream */

/* plain device-memory helpers for hosts that do not bring their own allocator (the CLI) */
int pvp_dev_malloc(pvp_ctx *ctx, size_t bytes, void **out_dev);
int pvp_dev_free(pvp_ctx *ctx, void *dev);
int pvp_dev_memset(pvp_ctx *ctx, void *dev, int value, size_t bytes);
int pvp_memcpy_h2d(pvp_ctx *ctx, void *dst_dev, const void *src, size_t bytes);
int pvp_memcpy_d2h(pvp_ctx *ctx, void *dst, const void *src_dev, size_t bytes); /* synchronises */
int pvp_host_alloc_pinned(size_t bytes, void **out);
int pvp_host_free_pinned(void *p);

/* ---- pose (host, glibc libm => bit-identical to the reference) ------------ */
/* replaces rotation_matrix_yaw_pitch_roll, ray_voxel.cpp:84-135 */
void pvp_rotation_matrix(float yaw_deg, float pitch_deg, float roll_deg, float R[9]);
/* replaces the focal-length expression of ray_voxel.cpp:490-491 */
float pvp_focal_length(float fov_degrees, int width);

/* ---- path A, stage level (parity hooks; same kernels as the fused path) --- */
/* replaces detect_motion, ray_voxel.cpp:236-255: changed[i] = |prev-next| > threshold, diff[i] = |prev-next| */
int pvp_detect_motion(pvp_ctx *ctx, const void *prev_dev, const void *next_dev, int frame_dtype, int64_t n_pixels,
                      float threshold, uint8_t *changed_dev, float *diff_dev);
/* replaces the ray set-up of ray_voxel.cpp:506-515 for a list of pixels (pix = v*width + u) */
int pvp_pixel_rays(pvp_ctx *ctx, const uint32_t *pix_dev, int64_t n, int width, int height, const pvp_pair *pose,
                   float *dirs_dev /* [n][3] */);
/*
 * replaces cast_ray_into_grid, ray_voxel.cpp:274-395, for n rays.  counts_dev[i] receives the
 * number of RaySteps of ray i.  If offsets_dev != NULL (exclusive prefix sum of the counts,
 * n+1 entries) the visited voxels are also written, in visit order, to
 * vox_dev[(offsets[i]+k)*3 + {0,1,2}] = (ix,iy,iz) and t_dev[offsets[i]+k] = RayStep::distance.
 */
int pvp_cast_rays(pvp_ctx *ctx, const float *origins_dev, const float *dirs_dev, int64_t n, const pvp_grid_desc *grid,
                  int32_t *counts_dev, const int64_t *offsets_dev, int32_t *vox_dev, float *t_dev);

/* ---- path A, fused: diff/threshold -> ray queue -> DDA -> accumulate ------- */
/*
 * replaces the body of main's frame loop, ray_voxel.cpp:484-531, for a batch of frame pairs.
 * frames_dev: [n_frames][height][width] (frame_stride elements between frames), device memory; every pair's prev / curr must
 *             lie in [0, n_frames) (PVP_ERR_INVALID otherwise: K1 would read past the frames).
 * grid_dev:   float32 or int64 [slab_hi-slab_lo][n][n], accumulated in place (never zeroed here).
 * stats:      NULL => fully asynchronous; else counters are ADDED into *stats after a stream sync.
 */
int pvp_motion_accumulate(pvp_ctx *ctx, const void *frames_dev, int frame_dtype, int64_t frame_stride, int width,
                          int height, int n_frames, const pvp_pair *pairs, int n_pairs, float threshold, void *grid_dev,
                          const pvp_grid_desc *grid, pvp_stats *stats);
/*
 * Same, frames in HOST memory (pinned for full copy bandwidth): frames are staged to the
 * device in chunks on a copy stream that overlaps the kernels of the previous chunk.
 * This is the call the CLI and the end-to-end benchmark use.
 */
int pvp_motion_accumulate_host(pvp_ctx *ctx, const void *frames_host, int frame_dtype, int64_t frame_stride, int width,
                               int height, int n_frames, const pvp_pair *pairs, int n_pairs, float threshold,
                               void *grid_dev, const pvp_grid_desc *grid, pvp_stats *stats);

/* ---- path A, whole pipeline: metadata.json + image folder -> grid ---------- */
/* One record of metadata.json, the FrameInfo of ray_voxel.cpp:47-55. */
typedef struct pvp_frame_info {
    int32_t camera_index;      /* default 0   (ray_voxel.cpp:157) */
    int32_t frame_index;       /* default 0   (:158) */
    float camera_position[3];  /* (:166-173); zero when absent, see has_position */
    float yaw, pitch, roll;    /* degrees, default 0 (:159-161) */
    float fov_degrees;         /* default 60 (:162) */
    int32_t has_position;      /* 0 when camera_position was missing / shorter than 3 */
    char image_file[512];      /* default "" (:163) */
} pvp_frame_info;

typedef struct pvp_run_options {
    pvp_grid_desc grid;   /* defaults: N=500, voxel_size=6, centre (0,0,500) (ray_voxel.cpp:434-437) */
    float threshold;      /* default 2.0 (:447) */
    int32_t shard_rank;   /* frame sharding: this process handles block shard_rank of shard_count */
    int32_t shard_count;
    int32_t chunk_frames; /* frames per K1/K2 batch (0 = 33: 32 pairs); results do not depend on it */
    int32_t verbose;      /* print the reference's per-frame diagnostics to stderr */
    int32_t decode_threads; /* host threads decoding image files ahead of the GPU (0 = $PVP_DECODE_THREADS or min(32, cores / $LOCAL_WORLD_SIZE)) */
    int32_t loader_noise; /* 1: add the reference loader's per-pixel uniform noise in [-1,1) (ray_voxel.cpp:203-218): the same
                             std::mt19937 / uniform_real_distribution<float> stream over all frames in load order, clamped to
                             [0,255]; frames are then float32.  The reference seeds it from std::random_device (not
                             reproducible); here the seed is noise_seed.  Default 0 = exact 8-bit values.  Not shardable. */
    uint32_t noise_seed;
} pvp_run_options;

typedef struct pvp_run_report {
    uint64_t n_frames_listed, n_frames_loaded, n_frames_skipped, n_pairs, n_size_mismatch;
    /* where the host side of the frame loop spent its wall time (the same ranges carry NVTX names "pvp:decode wait", "pvp:H2D frame",
     * "pvp:K1+K2 batch"): waiting for the decode pool, enqueueing H2D copies (incl. the noise of loader_noise), launching batches */
    double seconds_total, seconds_decode_wait, seconds_upload, seconds_launch;
    double seconds_setup;    /* metadata.json, the first frame's size probe, ring allocation (first call of a context) */
    double seconds_drain;    /* the final wait for the GPU after the last batch was submitted */
    int32_t decode_threads;  /* threads the decode pool ran with */
    int32_t reserved;
} pvp_run_report;

/* replaces load_metadata, ray_voxel.cpp:140-178; free the array with pvp_free_metadata */
int pvp_load_metadata(const char *json_path, pvp_frame_info **frames_out, int *n_out);
void pvp_free_metadata(pvp_frame_info *frames);
/* replaces stbi_load(path,&w,&h,&c,1) of load_image_gray, ray_voxel.cpp:195, without the unseeded
 * noise of :206-215 (PNM P5/P6, PNG incl. Adam7, uncompressed BMP, JPEG; stb_image's integer luma / IDCT).  out may be NULL. */
int pvp_load_image_gray(const char *path, uint8_t *out, int64_t cap, int *width, int *height);
void pvp_run_options_default(pvp_run_options *opt);
/* replaces the frame loop of main, ray_voxel.cpp:450-537, accumulating into grid_dev */
int pvp_ray_voxel_accumulate(pvp_ctx *ctx, const char *metadata_json, const char *image_folder, void *grid_dev,
                             const pvp_run_options *opt, pvp_stats *stats, pvp_run_report *report);
/* replaces main, ray_voxel.cpp:400-558: argv = ray_voxel <metadata.json> <image_folder> <out.bin> [flags] */
int pvp_ray_voxel_main(int argc, char **argv);

/* ---- grid helpers ---------------------------------------------------------- */
size_t pvp_grid_bytes(const pvp_grid_desc *grid);
/* int64 fixed point -> float32 (value / 2^frac_bits, one rounding) */
int pvp_fixed_to_f32(pvp_ctx *ctx, const int64_t *src_dev, float *dst_dev, int64_t n, int frac_bits);
/* int64 fixed point -> float64 */
int pvp_fixed_to_f64(pvp_ctx *ctx, const int64_t *src_dev, double *dst_dev, int64_t n, int frac_bits);
/* replaces the writer of ray_voxel.cpp:542-555: int32 N, float32 voxel_size, N^3 float32 */
int pvp_write_bin(const char *path, int32_t n, float voxel_size, const float *grid_host);
/* the same writer fed from device memory (float32 cells; D2H chunks overlap the file writes); synchronises */
int pvp_write_bin_dev(pvp_ctx *ctx, const char *path, int32_t n, float voxel_size, const float *grid_f32_dev);
/* reader for the same layout (the format's only consumer is voxelmotionviewer.py:28-53) */
int pvp_read_bin_header(const char *path, int32_t *n, float *voxel_size);
int pvp_read_bin(const char *path, int32_t n, float *grid_host);

/* ---- multi-GPU: the run's single grid merge (SURVEY 8e mode 1) over NVLink peer memory ---- */
/*
 * This rank sums its 1/world slice of the cells over ALL ranks' private grids and stores the result into rank dst_rank's
 * grid (dst_rank < 0: into every rank's grid).  peers[r] = the address through which THIS process reaches rank r's grid
 * (symmetric / peer-mapped memory, peers[rank] = the local grid); multicast = the NVSwitch multicast address of the same
 * allocation (then one multimem.ld_reduce per 16 bytes returns the sum of all GPUs) or NULL (plain peer loads, added in
 * rank order).  The call only enqueues a kernel on the context stream and contains no inter-GPU wait: every rank must
 * have finished accumulating before any rank's kernel starts, and finished this kernel before any grid is touched again
 * (barriers in the caller).  accum_mode: PVP_ACCUM_F32 (float32 cells) or PVP_ACCUM_FIXED (int64 cells).
 */
int pvp_grid_reduce_peers(pvp_ctx *ctx, void *const *peers, int world, int rank, int dst_rank, void *multicast, int64_t n_cells,
                          int accum_mode);

/* ---- consumer side of the .bin: voxelmotionviewer.py:56-96 with the grid in HBM ---- */
/*
 * The two order statistics np.percentile(flat, q) interpolates between (voxelmotionviewer.py:72): *v_k = the k-th
 * smallest of the n float32 cells (k counted from 0), *v_k1 = the (k+1)-th (== *v_k when k = n-1).  Radix select, 4-5
 * passes over the grid.  The interpolation itself (numpy's "linear" method) is host arithmetic on these two values.
 */
int pvp_grid_order_stats(pvp_ctx *ctx, const float *grid_dev, int64_t n, int64_t k, float *v_k, float *v_k1);
/*
 * replaces np.argwhere(voxel_grid > thresh) and the index -> world conversion of voxelmotionviewer.py:77-95, in the
 * same (C) order.  Any of the outputs may be NULL: coords_dev [cap][3] int32 = (i0,i1,i2) of voxel_grid[i0,i1,i2];
 * points_dev [cap][3] float64 = (x,y,z) with x from i2, y from i1, z from i0 (the viewer's "z up" reading, :85-93):
 * grid_min[a] + (idx + 0.5) * voxel_size; intens_dev [cap] float32.  *count = number of cells above thresh (may
 * exceed cap: nothing past cap is written; call with cap 0 to size the outputs).  The comparison is done in float64,
 * which is exact for a float32 or a float64 threshold.
 */
int pvp_grid_extract_above(pvp_ctx *ctx, const float *grid_dev, int32_t n_side, double thresh, double voxel_size, const double grid_min[3],
                           int64_t cap, double *points_dev, float *intens_dev, int32_t *coords_dev, int64_t *count);

/* ---- path B: process_image_cpp --------------------------------------------- */
/*
 * replaces process_image_cpp, process_image.cpp:125-366.  Arrays are addressed by element
 * strides exactly like the reference's numpy accessors; grid_dev may be NULL or have a zero
 * extent (voxel stage skipped, :152).  accum_mode PVP_ACCUM_F64: float64 grid/texture with
 * red.global.add.f64; PVP_ACCUM_FIXED: int64 grid/texture (value * 2^frac_bits).
 */
typedef struct pvp_image_args {
    double earth_position[3];      /* process_image.cpp:127 */
    double pointing_direction[3];  /* :128 (normalised inside, :194-199) */
    double fov;                    /* radians, :129 */
    int64_t image_width;           /* :130 */
    int64_t image_height;          /* :131 */
    int64_t image_stride[2];       /* element strides of `image` */
    int64_t grid_shape[3];         /* voxel_grid.shape, :171-175 */
    int64_t grid_stride[3];        /* element strides of voxel_grid */
    double extent[6];              /* voxel_grid_extent (xmin,xmax,ymin,ymax,zmin,zmax), :177-182 */
    double max_distance;           /* :134 */
    int32_t num_steps;             /* :135 */
    int64_t tex_shape[2];          /* celestial_sphere_texture.shape, :148-149 */
    int64_t tex_stride[2];
    double center_ra_rad, center_dec_rad, angular_width_rad, angular_height_rad; /* :137-140 */
    int32_t update_celestial_sphere;        /* :141 */
    int32_t perform_background_subtraction; /* :142 */
    int32_t accum_mode;
    int32_t frac_bits;
    /* rows [row_begin, row_end) of the image are processed -- a row shard of a multi-GPU run (the pixel loop of
     * process_image.cpp:232-235 restricted in i; pixel coordinates stay those of the whole image).  0, 0 = all rows;
     * row_begin == row_end > 0 = an empty shard. */
    int64_t row_begin, row_end;
} pvp_image_args;

typedef struct pvp_image_stats {
    uint64_t n_rays;    /* pixels that reached the ray stage */
    uint64_t n_samples; /* grid updates = voxel-samples (process_image.cpp:357) */
} pvp_image_stats;

int pvp_process_image(pvp_ctx *ctx, const double *image_dev, void *grid_dev, void *tex_dev, const pvp_image_args *args,
                      pvp_image_stats *stats);
/* host arrays (the numpy call of spacevoxelviewer.py:233-251): copy in, run, copy back in place */
int pvp_process_image_host(pvp_ctx *ctx, const double *image, double *grid, double *tex, const pvp_image_args *args,
                           pvp_image_stats *stats);

/*
 * Self-test of the lock-step path-B kernel's division by the grid extent (a reciprocal-based sequence that must equal
 * the reference's IEEE division, process_image.cpp:94-96, bit for bit): compares both on n_random numerators drawn
 * uniformly in [0, extent] and on +-64 ulps around every k*extent/n_grid voxel boundary (quotient bit-identical for
 * numerators >= 2^-960, voxel index identical always).  *eligible = 0 means the
 * library would use the generic kernel (plain division) for this extent and nothing is compared.
 */
int pvp_selftest_div_by_extent(pvp_ctx *ctx, double extent, int64_t n_random, int32_t n_grid, uint64_t seed, uint64_t *mismatches,
                               int32_t *eligible);

/* ---- roofline micro-benchmark (measured L2-atomic peak) ---------- */
/*
 * Issues n_ops independent atomic adds into a buffer of buf_bytes with the given address
 * pattern and returns the elapsed milliseconds of the kernel (CUDA events on the ctx stream).
 * elem: 4 = red.add.f32, 8 = red.add.u64, 9 = red.add.f64.
 * pattern: 0 = uniformly random address per lane, 1 = random 32 B sector per 8 lanes
 * (lanes hit consecutive words), 2 = one random address per warp (all 32 lanes collide),
 * 3 = random 128 B line per warp (lanes consecutive).
 */
int pvp_atomic_microbench(pvp_ctx *ctx, void *buf_dev, size_t buf_bytes, int elem, int pattern, int64_t n_ops,
                          float *elapsed_ms);
/* the number of atomic operations a pvp_atomic_microbench call with this n_ops really issues */
int64_t pvp_atomic_microbench_ops(pvp_ctx *ctx, int64_t n_ops);

#ifdef __cplusplus
}
#endif
#endif /* PVP_B200_H */
