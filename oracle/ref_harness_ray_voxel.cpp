/*
 * ref_harness_ray_voxel.cpp -- TEST INFRASTRUCTURE ONLY (builds into oracle/_ref/, git-ignored).
 *
 * Compiles the UNMODIFIED reference translation unit where it lies
 * (REF_RAY_VOXEL_CPP = "/root/reference/ray_voxel.cpp", passed by oracle/Makefile) and
 * exposes its own functions through a C ABI so tests can pin oracle/pvp_oracle.c and the
 * CUDA kernels against the real thing.  No reference source is copied into this repo:
 * the reference file is #included by absolute path at build time.
 *
 * Two build-time knobs, neither edits the reference:
 *   - main is renamed to ref_main so the TU can live in a shared object;
 *   - with -DPVP_REF_NOISE_OFF the token `uniform_real_distribution` is re-pointed at a
 *     distribution that always returns 0, which removes the unseeded per-pixel noise of
 *     load_image_gray (ray_voxel.cpp:206-215) and makes the reference CLI reproducible.
 *     All standard headers are included BEFORE the macro so only the reference's use is hit.
 *   - with -DPVP_REF_NOISE_SEEDED the token `random_device` is re-pointed at a device that returns the value of
 *     $PVP_REF_NOISE_SEED, so the reference's own noise code (static mt19937 seeded once, one
 *     uniform_real_distribution<float>(-1,1) draw per pixel in load order) runs unchanged but reproducibly: the golden
 *     for the product's --loader-noise mode.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <random>
#include <string>
#include <vector>
#include "nlohmann/json.hpp"

#ifdef PVP_REF_NOISE_OFF
namespace pvp_ref_shim {
template <class T> struct zero_distribution {
    zero_distribution(T, T) {}
    template <class G> T operator()(G &) { return T(0); }
};
}  // namespace pvp_ref_shim
namespace std { using pvp_ref_shim::zero_distribution; }
#define uniform_real_distribution zero_distribution
#endif

#ifdef PVP_REF_NOISE_SEEDED
namespace pvp_ref_shim {
struct env_seed_device {
    unsigned operator()() const { const char *e = std::getenv("PVP_REF_NOISE_SEED"); return e ? (unsigned)std::strtoul(e, nullptr, 10) : 0u; }
};
}  // namespace pvp_ref_shim
namespace std { using pvp_ref_shim::env_seed_device; }
#define random_device env_seed_device
#endif

#define main ref_main
#include REF_RAY_VOXEL_CPP
#undef main

extern "C" {

/* ray_voxel.cpp:84-135, called as is */
void ref_rotation_matrix(float yaw, float pitch, float roll, float R[9])
{
    Mat3 m = rotation_matrix_yaw_pitch_roll(yaw, pitch, roll);
    for (int i = 0; i < 9; ++i) R[i] = m.m[i];
}

/* ray_voxel.cpp:490-491 (expression lives inside main; restated with the reference's deg2rad) */
float ref_focal_length(float fov_degrees, int width)
{
    float fov_rad = deg2rad(fov_degrees);
    return (width * 0.5f) / std::tan(fov_rad * 0.5f);
}

/* ray_voxel.cpp:236-255, called as is */
int ref_detect_motion(const float *prev, const float *next, int width, int height, float threshold,
                      uint8_t *changed, float *diff)
{
    ImageGray a, b;
    a.width = b.width = width;
    a.height = b.height = height;
    a.pixels.assign(prev, prev + (size_t)width * height);
    b.pixels.assign(next, next + (size_t)width * height);
    MotionMask mm = detect_motion(a, b, threshold);
    for (size_t i = 0; i < mm.diff.size(); ++i) {
        changed[i] = mm.changed[i] ? 1 : 0;
        diff[i] = mm.diff[i];
    }
    return (int)mm.diff.size();
}

/* ray_voxel.cpp:506-515 (inside main) using the reference's normalize / mat3_mul_vec3 */
void ref_pixel_ray(int u, int v, int width, int height, float focal_len, const float R[9], float dir[3])
{
    Mat3 cam_rot;
    for (int i = 0; i < 9; ++i) cam_rot.m[i] = R[i];
    float x = (float(u) - 0.5f * width);
    float y = -(float(v) - 0.5f * height);
    float z = -focal_len;
    Vec3 ray_cam = { x, y, z };
    ray_cam = normalize(ray_cam);
    Vec3 ray_world = mat3_mul_vec3(cam_rot, ray_cam);
    ray_world = normalize(ray_world);
    dir[0] = ray_world.x; dir[1] = ray_world.y; dir[2] = ray_world.z;
}

/* ray_voxel.cpp:274-395, called as is */
int64_t ref_cast_ray(const float cam[3], const float dir[3], int N, float voxel_size, const float center[3],
                     int32_t *vox, float *t, int64_t cap)
{
    Vec3 c = { cam[0], cam[1], cam[2] }, d = { dir[0], dir[1], dir[2] }, g = { center[0], center[1], center[2] };
    std::vector<RayStep> steps = cast_ray_into_grid(c, d, N, voxel_size, g);
    int64_t n = (int64_t)steps.size();
    for (int64_t i = 0; i < n && i < cap; ++i) {
        if (vox) { vox[3 * i] = steps[i].ix; vox[3 * i + 1] = steps[i].iy; vox[3 * i + 2] = steps[i].iz; }
        if (t) t[i] = steps[i].distance;
    }
    return n;
}

/*
 * ray_voxel.cpp:484-531 for one frame pair: the loop body of main with N / voxel_size /
 * grid_center / motion_threshold as parameters (they are constants at :434-448) and every
 * computation delegated to the reference's own functions.  Serial fp32 += in the
 * reference's order.
 */
int ref_accumulate_pair(const float *prev, const float *curr, int width, int height, const float cam[3],
                        const float R[9], float focal_len, int N, float voxel_size, const float center[3],
                        float motion_threshold, float *voxel_grid, uint64_t *n_rays, uint64_t *n_steps)
{
    ImageGray a, b;
    a.width = b.width = width;
    a.height = b.height = height;
    a.pixels.assign(prev, prev + (size_t)width * height);
    b.pixels.assign(curr, curr + (size_t)width * height);
    MotionMask mm = detect_motion(a, b, motion_threshold);
    Vec3 cam_pos = { cam[0], cam[1], cam[2] };
    Vec3 grid_center = { center[0], center[1], center[2] };
    Mat3 cam_rot;
    for (int i = 0; i < 9; ++i) cam_rot.m[i] = R[i];
    uint64_t rays = 0, nsteps = 0;
    for (int v = 0; v < mm.height; v++) {
        for (int u = 0; u < mm.width; u++) {
            if (!mm.changed[v * mm.width + u]) continue;
            float pix_val = mm.diff[v * mm.width + u];
            if (pix_val < 1e-3f) continue;
            float x = (float(u) - 0.5f * mm.width);
            float y = -(float(v) - 0.5f * mm.height);
            float z = -focal_len;
            Vec3 ray_cam = { x, y, z };
            ray_cam = normalize(ray_cam);
            Vec3 ray_world = mat3_mul_vec3(cam_rot, ray_cam);
            ray_world = normalize(ray_world);
            std::vector<RayStep> steps = cast_ray_into_grid(cam_pos, ray_world, N, voxel_size, grid_center);
            for (const auto &rs : steps) {
                float val = pix_val * 1.f;
                size_t idx = (size_t)rs.ix * N * N + (size_t)rs.iy * N + (size_t)rs.iz;
                voxel_grid[idx] += val;
            }
            rays++;
            nsteps += steps.size();
        }
    }
    if (n_rays) *n_rays += rays;
    if (n_steps) *n_steps += nsteps;
    return 0;
}

/* The reference CLI, ray_voxel.cpp:400-558, unmodified (constants N=500 etc. included). */
int ref_cli(const char *metadata_json, const char *image_folder, const char *output_bin)
{
    char a0[] = "ray_voxel";
    std::string s1(metadata_json), s2(image_folder), s3(output_bin);
    char *argv[4] = { a0, s1.data(), s2.data(), s3.data() };
    return ref_main(4, argv);
}

}  // extern "C"

#ifdef PVP_REF_BUILD_CLI
int main(int argc, char **argv) { return ref_main(argc, argv); }
#endif
