// pvp_pipeline.cu -- host orchestration of path A: everything main() of ray_voxel.cpp does
// around the hot loop (:400-558): metadata.json parsing (:140-178), per-camera grouping and
// frame_index ordering (:419-429), first-frame priming / skip-on-load-error / prev<-curr
// (:464-481, :534-535), the hard-coded grid constants as defaults (:434-448) and the .bin
// writer (:542-555).  The per-pair work goes to pvp_motion_accumulate_host (K1 + K2).
//
// Host-only C++; compiled by nvcc only because it records CUDA events to recycle its pinned
// staging buffers.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <map>
#include <mutex>
#include <new>
#include <stdexcept>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "pvp_internal.h"

namespace {

// ---- a small JSON reader (the reference uses nlohmann::json, un-vendored) ------------------
struct JValue {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;
    const JValue *find(const char *key) const
    {
        const JValue *hit = nullptr;  // nlohmann keeps the last duplicate key
        for (auto &kv : obj) if (kv.first == key) hit = &kv.second;
        return hit;
    }
};

struct JParser {
    const char *p, *end;
    std::string err;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool fail(const char *m) { if (err.empty()) err = m; return false; }
    bool parse_string(std::string &out)
    {
        if (p >= end || *p != '"') return fail("expected string");
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail("bad escape");
                switch (*p) {
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    case '/': out += '/'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'n': out += '\n'; break;
                    case 'r': out += '\r'; break;
                    case 't': out += '\t'; break;
                    case 'u': {
                        if (end - p < 5) return fail("bad \\u escape");
                        unsigned cp = 0;
                        for (int i = 1; i <= 4; ++i) {
                            char c = p[i];
                            cp <<= 4;
                            if (c >= '0' && c <= '9') cp |= c - '0';
                            else if (c >= 'a' && c <= 'f') cp |= c - 'a' + 10;
                            else if (c >= 'A' && c <= 'F') cp |= c - 'A' + 10;
                            else return fail("bad \\u escape");
                        }
                        p += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: return fail("bad escape");
                }
                ++p;
            } else {
                out += *p++;
            }
        }
        if (p >= end) return fail("unterminated string");
        ++p;
        return true;
    }
    bool parse(JValue &v, int depth = 0)
    {
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end of input");
        if (*p == '{') {
            v.kind = JValue::Obj; ++p; ws();
            if (p < end && *p == '}') { ++p; return true; }
            for (;;) {
                ws();
                std::string k;
                if (!parse_string(k)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':'");
                ++p;
                JValue c;
                if (!parse(c, depth + 1)) return false;
                v.obj.emplace_back(std::move(k), std::move(c));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v.kind = JValue::Arr; ++p; ws();
            if (p < end && *p == ']') { ++p; return true; }
            for (;;) {
                JValue c;
                if (!parse(c, depth + 1)) return false;
                v.arr.push_back(std::move(c));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (*p == '"') { v.kind = JValue::Str; return parse_string(v.str); }
        if (end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = JValue::Bool; v.b = true; p += 4; return true; }
        if (end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = JValue::Bool; v.b = false; p += 5; return true; }
        if (end - p >= 4 && !strncmp(p, "null", 4)) { v.kind = JValue::Null; p += 4; return true; }
        {
            std::string tok;
            // (strchr would also match the terminating NUL of its own pattern: an embedded 0x00 byte must end the token, not join it)
            while (p < end && *p != '\0' && (strchr("+-.eE", *p) || (*p >= '0' && *p <= '9'))) tok += *p++;
            if (tok.empty()) return fail("unexpected character");
            char *e = nullptr;
            v.num = strtod(tok.c_str(), &e);
            if (!e || e == tok.c_str() || *e) return fail("bad number");
            v.kind = JValue::Num;
            return true;
        }
    }
};

bool jnum(const JValue *v, double *out)
{
    if (!v) return false;
    if (v->kind == JValue::Num) { *out = v->num; return true; }
    if (v->kind == JValue::Bool) { *out = v->b ? 1.0 : 0.0; return true; }
    return false;
}

}  // namespace

using namespace pvp;

extern "C" {

// replaces load_metadata, ray_voxel.cpp:140-178.  *frames_out is malloc'ed; release with
// pvp_free_metadata.  Defaults follow :157-163 (camera_index 0, frame_index 0, yaw/pitch/roll 0,
// fov_degrees 60, image_file "").  A missing / short camera_position leaves the reference's
// Vec3 uninitialised (:166-173); here it is (0,0,0) with has_position = 0.
static int load_metadata_impl(const char *json_path, pvp_frame_info **frames_out, int *n_out)
{
    PVP_REQUIRE(json_path && frames_out && n_out, "pvp_load_metadata: NULL argument");
    *frames_out = nullptr; *n_out = 0;
    FILE *f = fopen(json_path, "rb");
    if (!f) { set_error("ERROR: Cannot open %s", json_path); return PVP_ERR_IO; }
    std::string text;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
    fclose(f);
    JParser ps{text.data(), text.data() + text.size(), {}};
    JValue root;
    if (!ps.parse(root)) { set_error("ERROR: JSON parse error in %s: %s", json_path, ps.err.c_str()); return PVP_ERR_FORMAT; }
    if (root.kind != JValue::Arr) { set_error("ERROR: JSON top level is not an array."); return PVP_ERR_FORMAT; }
    const size_t cnt = root.arr.size();
    pvp_frame_info *out = (pvp_frame_info *)calloc(cnt ? cnt : 1, sizeof(pvp_frame_info));
    if (!out) { set_error("pvp_load_metadata: out of memory"); return PVP_ERR_NOMEM; }
    for (size_t i = 0; i < cnt; ++i) {
        const JValue &e = root.arr[i];
        pvp_frame_info &fi = out[i];
        fi.fov_degrees = 60.f;
        if (e.kind != JValue::Obj) { free(out); set_error("ERROR: metadata entry %zu is not an object", i); return PVP_ERR_FORMAT; }
        struct { const char *key; int *ip; float *fp; } fields[] = {
            {"camera_index", &fi.camera_index, nullptr}, {"frame_index", &fi.frame_index, nullptr},
            {"yaw", nullptr, &fi.yaw}, {"pitch", nullptr, &fi.pitch}, {"roll", nullptr, &fi.roll},
            {"fov_degrees", nullptr, &fi.fov_degrees}};
        for (auto &fd : fields) {
            const JValue *v = e.find(fd.key);
            if (!v) continue;
            double x;
            if (!jnum(v, &x)) { free(out); set_error("ERROR: metadata entry %zu: '%s' is not a number", i, fd.key); return PVP_ERR_FORMAT; }
            if (fd.ip) {  // a double outside int's range (or NaN) has no defined conversion: reject it like nlohmann's get<int> range error
                if (!(x >= -2147483648.0 && x <= 2147483647.0)) { free(out); set_error("ERROR: metadata entry %zu: '%s' is out of the int range", i, fd.key); return PVP_ERR_FORMAT; }
                *fd.ip = (int)x;
            } else {
                *fd.fp = (float)x;
            }
        }
        if (const JValue *v = e.find("image_file")) {
            if (v->kind != JValue::Str) { free(out); set_error("ERROR: metadata entry %zu: 'image_file' is not a string", i); return PVP_ERR_FORMAT; }
            if (v->str.size() >= sizeof(fi.image_file)) { free(out); set_error("ERROR: metadata entry %zu: image_file too long", i); return PVP_ERR_FORMAT; }
            memcpy(fi.image_file, v->str.data(), v->str.size());
        }
        const JValue *cp = e.find("camera_position");
        if (cp && cp->kind == JValue::Arr && cp->arr.size() >= 3) {
            for (int k = 0; k < 3; ++k) {
                double x;
                if (!jnum(&cp->arr[k], &x)) { free(out); set_error("ERROR: metadata entry %zu: camera_position[%d] is not a number", i, k); return PVP_ERR_FORMAT; }
                fi.camera_position[k] = (float)x;
            }
            fi.has_position = 1;
        }
    }
    *frames_out = out; *n_out = (int)cnt;
    return PVP_OK;
}

// No C++ exception crosses the C ABI: the parser and the frame loop allocate (strings, vectors, threads).
#define PVP_GUARDED(call)                                                                      \
    try {                                                                                      \
        return call;                                                                           \
    } catch (const std::bad_alloc &) {                                                         \
        set_error("%s: out of host memory", __func__);                                         \
        return PVP_ERR_NOMEM;                                                                  \
    } catch (const std::exception &e) {                                                        \
        set_error("%s: %s", __func__, e.what());                                               \
        return PVP_ERR_INVALID;                                                                \
    } catch (...) {                                                                            \
        set_error("%s: unexpected exception", __func__);                                       \
        return PVP_ERR_INVALID;                                                                \
    }

int pvp_load_metadata(const char *json_path, pvp_frame_info **frames_out, int *n_out)
{
    PVP_GUARDED(load_metadata_impl(json_path, frames_out, n_out))
}

void pvp_free_metadata(pvp_frame_info *frames) { free(frames); }

void pvp_run_options_default(pvp_run_options *o)
{
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->grid.n = 500;                 // ray_voxel.cpp:434
    o->grid.voxel_size = 6.f;        // :435
    o->grid.center[0] = -0.f; o->grid.center[1] = 0.f; o->grid.center[2] = 500.f;  // :437
    o->grid.accum_mode = PVP_ACCUM_F32;
    o->grid.frac_bits = 16;
    o->grid.slab_lo = 0; o->grid.slab_hi = 500;
    o->grid.attenuation_alpha = 0.f;  // ray_voxel.cpp:526 multiplies by 1.f (alpha = 0.1 of :448 is dead in the reference)
    o->threshold = 2.0f;             // :447
    o->shard_rank = 0; o->shard_count = 1;
    o->chunk_frames = 0;
    o->verbose = 1;
    o->decode_threads = 0;
    o->loader_noise = 0;
    o->noise_seed = 0;
}

}  // extern "C"

namespace {

struct Slot { int cam_pos; };  // index into the camera's sorted frame list

struct PinnedChunk {
    uint8_t *buf = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
};

// Decode prefetcher: once K2 is fast, decoding 1080p/4K files is what limits frames/s through the CLI.
// The frames a run will visit are known up front (metadata order), so a pool of host threads decodes them ahead of
// the consumer into a bounded ring; the consumer sees them strictly in order, so skip-on-error / size-mismatch /
// prev<-curr semantics (ray_voxel.cpp:464-481, :534) are untouched.
class DecodePool {
public:
    struct Item { bool ok = false; int w = 0, h = 0; std::vector<uint8_t> px; };

    DecodePool(std::vector<std::string> paths, int threads) : paths_(std::move(paths))
    {
        const int n = (int)paths_.size();
        threads = std::max(1, std::min(threads, std::max(1, n)));
        ring_.resize((size_t)std::min(n, 2 * threads + 2));
        state_.assign(ring_.size(), -1);
        for (int t = 0; t < threads && n > 0; ++t) workers_.emplace_back([this] { work(); });
    }
    ~DecodePool()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    // frame `idx` (consumed in increasing order); valid until release(idx)
    Item &get(int idx)
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return state_[(size_t)idx % ring_.size()] == idx; });
        return ring_[(size_t)idx % ring_.size()];
    }
    void release(int idx)
    {
        {
            std::lock_guard<std::mutex> g(m_);
            state_[(size_t)idx % ring_.size()] = -1;
            consumed_ = idx + 1;
        }
        cv_.notify_all();
    }

private:
    void work()
    {
        for (;;) {
            const int idx = next_.fetch_add(1);
            if (idx >= (int)paths_.size()) return;
            const size_t slot = (size_t)idx % ring_.size();
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return stop_ || (idx < consumed_ + (int)ring_.size() && state_[slot] == -1); });
                if (stop_) return;
                state_[slot] = -2;  // being filled
            }
            Item &it = ring_[slot];
            it.ok = pvp::decode_gray_file(paths_[(size_t)idx].c_str(), it.px, &it.w, &it.h);
            {
                std::lock_guard<std::mutex> g(m_);
                state_[slot] = idx;
            }
            cv_.notify_all();
        }
    }
    std::vector<std::string> paths_;
    std::vector<Item> ring_;
    std::vector<int> state_;  // -1 free, -2 being filled, else the frame index it holds
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::atomic<int> next_{0};
    int consumed_ = 0;
    bool stop_ = false;
};

struct Runner {
    pvp_ctx *ctx;
    const pvp_run_options *opt;
    void *grid_dev;
    std::string folder;
    PinnedChunk chunk[2];
    int cur = 0;
    // open chunk
    int W = 0, H = 0, n_frames = 0;
    std::vector<pvp_pair> pairs;
    int max_frames = 32;
    uint64_t frames_loaded = 0, pairs_done = 0;
    // loader noise (ray_voxel.cpp:203-218): frames become float32, one generator for the whole run, seeded once
    bool noise = false;
    size_t elem = 1;
    std::mt19937 gen;

    int ensure(PinnedChunk &c, size_t bytes)
    {
        if (c.cap >= bytes) return PVP_OK;
        if (c.buf) { PVP_CUDA(cudaEventSynchronize(c.done)); cudaFreeHost(c.buf); c.buf = nullptr; c.cap = 0; }
        PVP_CUDA(cudaMallocHost((void **)&c.buf, bytes));
        c.cap = bytes;
        return PVP_OK;
    }
    int init()
    {
        for (auto &c : chunk) PVP_CUDA(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
        return PVP_OK;
    }
    // the pinned chunk buffers and their events go with the Runner on every path out of the frame loop, error returns and
    // exceptions (caught at the C ABI) included
    ~Runner()
    {
        for (auto &c : chunk) {
            if (c.done) { cudaEventSynchronize(c.done); cudaEventDestroy(c.done); c.done = nullptr; }
            if (c.buf) { cudaFreeHost(c.buf); c.buf = nullptr; c.cap = 0; }
        }
    }
    // Submit the open chunk; keep its last frame as frame 0 of the next chunk when `carry`.
    int flush(bool carry)
    {
        PinnedChunk &c = chunk[cur];
        const size_t fbytes = (size_t)W * H * elem;
        if (!pairs.empty()) {
            int rc = pvp_motion_accumulate_host(ctx, c.buf, noise ? PVP_FRAME_F32 : PVP_FRAME_U8, (int64_t)W * H, W, H, n_frames, pairs.data(),
                                                (int)pairs.size(), opt->threshold, grid_dev, &opt->grid, nullptr);
            if (rc) return rc;
            PVP_CUDA(cudaEventRecord(c.done, ctx->stream));
            pairs_done += pairs.size();
        }
        PinnedChunk &nx = chunk[cur ^ 1];
        if (carry && n_frames > 0) {
            int rc = ensure(nx, (size_t)max_frames * fbytes);
            if (rc) return rc;
            PVP_CUDA(cudaEventSynchronize(nx.done));  // the other buffer's previous submission has been copied
            memcpy(nx.buf, c.buf + (size_t)(n_frames - 1) * fbytes, fbytes);
            n_frames = 1;
        } else {
            n_frames = 0;
        }
        pairs.clear();
        cur ^= 1;
        return PVP_OK;
    }
};

}  // namespace

extern "C" {

// replaces main's frame loop, ray_voxel.cpp:450-537, accumulating into grid_dev (not zeroed,
// not written to disk).  With shard_count > 1 this rank handles a contiguous block of the
// (camera, frame) slots; pairing across skipped frames is resolved exactly as the reference
// would (the previous frame of a block is the last loadable frame before it).
static int ray_voxel_accumulate_impl(pvp_ctx *ctx, const char *metadata_json, const char *image_folder, void *grid_dev,
                                     const pvp_run_options *opt, pvp_stats *stats, pvp_run_report *report)
{
    PVP_REQUIRE(ctx && metadata_json && image_folder && grid_dev && opt, "pvp_ray_voxel_accumulate: NULL argument");
    PVP_REQUIRE(opt->shard_count >= 1 && opt->shard_rank >= 0 && opt->shard_rank < opt->shard_count, "bad shard %d/%d",
                opt->shard_rank, opt->shard_count);
    PVP_REQUIRE(!(opt->loader_noise && opt->shard_count > 1),
                "loader_noise draws one pseudo-random stream over all frames in load order (ray_voxel.cpp:203-209); it cannot be frame-sharded");
    pvp_frame_info *frames = nullptr;
    int n = 0;
    int rc = pvp_load_metadata(metadata_json, &frames, &n);
    if (rc) return rc;
    if (n == 0) { pvp_free_metadata(frames); set_error("No frames loaded."); return PVP_ERR_FORMAT; }

    std::map<int, std::vector<pvp_frame_info>> by_cam;  // :419-422 (ascending camera id)
    for (int i = 0; i < n; ++i) by_cam[frames[i].camera_index].push_back(frames[i]);
    pvp_free_metadata(frames);
    for (auto &kv : by_cam)  // :424-429
        std::sort(kv.second.begin(), kv.second.end(), [](const pvp_frame_info &a, const pvp_frame_info &b) { return a.frame_index < b.frame_index; });

    // global slot list -> this rank's contiguous block
    long long total_slots = 0;
    for (auto &kv : by_cam) if (kv.second.size() >= 2) total_slots += (long long)kv.second.size();  // :454
    const long long lo = total_slots * opt->shard_rank / opt->shard_count, hi = total_slots * (opt->shard_rank + 1) / opt->shard_count;

    DeviceGuard guard(ctx->device);
    Runner R{ctx, opt, grid_dev, image_folder};
    R.max_frames = opt->chunk_frames > 1 ? opt->chunk_frames : 32;
    if (opt->loader_noise) {
        R.noise = true; R.elem = sizeof(float);
        R.gen.seed(opt->noise_seed);  // the reference: static std::mt19937 gen(rd()) -- here rd() is the caller's seed
    }
    if ((rc = R.init())) return rc;
    std::vector<uint8_t> tmp;
    uint64_t skipped = 0, mismatched = 0;

    auto load = [&](const pvp_frame_info &fi, int *w, int *h) -> bool {  // serial load (shard look-back only)
        std::string path = R.folder + "/" + fi.image_file;  // :467
        return decode_gray_file(path.c_str(), tmp, w, h);
    };
    // every frame of this rank's block, in visiting order, goes through the decode pool
    std::vector<std::string> visit;
    {
        long long s0 = 0;
        for (auto &kv : by_cam) {
            auto &cf = kv.second;
            if (cf.size() < 2) continue;
            const long long cam_lo = s0, cam_hi = s0 + (long long)cf.size();
            s0 = cam_hi;
            const long long a = std::max(lo, cam_lo) - cam_lo, b = std::min(hi, cam_hi) - cam_lo;
            for (long long k = a; k < b; ++k) visit.push_back(R.folder + "/" + cf[(size_t)k].image_file);
        }
    }
    int decode_threads = opt->decode_threads;
    if (decode_threads <= 0) {
        const char *env = getenv("PVP_DECODE_THREADS");
        decode_threads = env ? atoi(env) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    }
    DecodePool pool(std::move(visit), decode_threads);
    int visit_idx = 0;
    auto load_next = [&](int *w, int *h) -> bool {  // next frame of the visiting order -> tmp
        const int idx = visit_idx++;
        DecodePool::Item &it = pool.get(idx);
        const bool ok = it.ok;
        if (ok) { *w = it.w; *h = it.h; tmp.swap(it.px); }  // the slot keeps the old buffer for its next decode
        pool.release(idx);
        return ok;
    };
    auto push_frame = [&](int w, int h) -> int {  // append tmp as a new frame of the open chunk
        const size_t fbytes = (size_t)w * h * R.elem;
        PinnedChunk &c = R.chunk[R.cur];
        int rc2 = R.ensure(c, (size_t)R.max_frames * fbytes);
        if (rc2) return rc2;
        if (R.n_frames == 0) PVP_CUDA(cudaEventSynchronize(c.done));
        if (R.noise) {  // load_image_gray, ray_voxel.cpp:203-218: same generator, same distribution, same draw order
            std::uniform_real_distribution<float> noise_dist(-1.0f, 1.0f);
            float *dst = reinterpret_cast<float *>(c.buf + (size_t)R.n_frames * fbytes);
            const size_t n_px = (size_t)w * h;
            for (size_t i = 0; i < n_px; ++i) {
                float val = static_cast<float>(tmp[i]);
                val += noise_dist(R.gen);
                if (val < 0.0f) val = 0.0f;
                if (val > 255.0f) val = 255.0f;
                dst[i] = val;
            }
        } else {
            memcpy(c.buf + (size_t)R.n_frames * fbytes, tmp.data(), fbytes);
        }
        R.W = w; R.H = h; R.n_frames++;
        R.frames_loaded++;
        return PVP_OK;
    };

    long long slot = 0;
    for (auto &kv : by_cam) {
        auto &cf = kv.second;
        if (cf.size() < 2) continue;
        const long long cam_lo = slot, cam_hi = slot + (long long)cf.size();
        slot = cam_hi;
        const long long a = std::max(lo, cam_lo) - cam_lo, b = std::min(hi, cam_hi) - cam_lo;
        if (a >= b) continue;
        bool prev_valid = false;
        // the frame a pair starting this block compares against: last loadable frame before it
        for (long long k = a - 1; k >= 0 && !prev_valid; --k) {
            int w, h;
            if (load(cf[(size_t)k], &w, &h)) {
                if ((rc = push_frame(w, h))) return rc;
                prev_valid = true;
            }
        }
        for (long long k = a; k < b; ++k) {
            const pvp_frame_info &ci = cf[(size_t)k];
            int w, h;
            if (!load_next(&w, &h)) {  // :470-473
                if (opt->verbose) fprintf(stderr, "Failed to load image: %s/%s\nSkipping frame due to load error.\n", R.folder.c_str(), ci.image_file);
                ++skipped;
                continue;
            }
            if (!prev_valid) {  // :475-481
                if ((rc = push_frame(w, h))) return rc;
                prev_valid = true;
                continue;
            }
            if (w != R.W || h != R.H) {  // detect_motion's size check, :238-243: no rays, prev still advances (:534)
                if (opt->verbose) fprintf(stderr, "Images differ in size. Can't do motion detection!\n");
                ++mismatched;
                if ((rc = R.flush(false))) return rc;
                if ((rc = push_frame(w, h))) return rc;
                continue;
            }
            if (R.n_frames >= R.max_frames) {
                std::vector<uint8_t> keep;
                keep.swap(tmp);  // flush() carries the previous frame over; tmp holds the current one
                rc = R.flush(true);
                tmp.swap(keep);
                if (rc) return rc;
            }
            if ((rc = push_frame(w, h))) return rc;
            pvp_pair p;  // pose of the CURRENT frame, :486-491
            memcpy(p.cam, ci.camera_position, sizeof(p.cam));
            pvp_rotation_matrix(ci.yaw, ci.pitch, ci.roll, p.R);
            p.focal = pvp_focal_length(ci.fov_degrees, w);
            p.prev = R.n_frames - 2; p.curr = R.n_frames - 1;
            R.pairs.push_back(p);
        }
        if ((rc = R.flush(false))) return rc;  // camera boundary: no pair across cameras
    }
    pvp_stats st = {0, 0, 0, 0};
    rc = pvp_motion_accumulate(ctx, nullptr, PVP_FRAME_U8, 1, 1, 1, 0, nullptr, 0, 0.f, grid_dev, &opt->grid, &st);  // sync + counters
    if (rc) return rc;
    if (stats) { stats->n_rays += st.n_rays; stats->n_steps += st.n_steps; stats->n_written += st.n_written; stats->n_atomics += st.n_atomics; }
    if (report) {
        report->n_frames_listed = (uint64_t)n; report->n_frames_loaded = R.frames_loaded; report->n_frames_skipped = skipped;
        report->n_pairs = R.pairs_done; report->n_size_mismatch = mismatched;
    }
    return PVP_OK;
}

int pvp_ray_voxel_accumulate(pvp_ctx *ctx, const char *metadata_json, const char *image_folder, void *grid_dev,
                             const pvp_run_options *opt, pvp_stats *stats, pvp_run_report *report)
{
    PVP_GUARDED(ray_voxel_accumulate_impl(ctx, metadata_json, image_folder, grid_dev, opt, stats, report))
}

// The CLI of ray_voxel.cpp:400-558: `ray_voxel <metadata.json> <image_folder> <output_voxel_bin>`
// with the reference's constants as defaults and its messages / exit codes, plus optional flags
// after the three positionals:
//   --n N --voxel-size S --center X Y Z --threshold T --fixed-point [BITS] --device D --decode-threads T --loader-noise SEED --attenuation ALPHA --quiet --stats
int pvp_ray_voxel_main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "Usage: %s <metadata.json> <image_folder> <output_voxel_bin>\n", argc > 0 ? argv[0] : "ray_voxel");
        return 1;
    }
    pvp_run_options opt;
    pvp_run_options_default(&opt);
    int device = 0, want_stats = 0;
    for (int i = 4; i < argc; ++i) {
        auto need = [&](int k) { return i + k < argc; };
        if (!strcmp(argv[i], "--n") && need(1)) { opt.grid.n = atoi(argv[++i]); opt.grid.slab_hi = opt.grid.n; }
        else if (!strcmp(argv[i], "--voxel-size") && need(1)) opt.grid.voxel_size = strtof(argv[++i], nullptr);
        else if (!strcmp(argv[i], "--center") && need(3)) { for (int k = 0; k < 3; ++k) opt.grid.center[k] = strtof(argv[++i], nullptr); }
        else if (!strcmp(argv[i], "--threshold") && need(1)) opt.threshold = strtof(argv[++i], nullptr);
        else if (!strcmp(argv[i], "--fixed-point")) { opt.grid.accum_mode = PVP_ACCUM_FIXED; if (need(1) && argv[i + 1][0] != '-') opt.grid.frac_bits = atoi(argv[++i]); }
        else if (!strcmp(argv[i], "--device") && need(1)) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--decode-threads") && need(1)) opt.decode_threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--attenuation") && need(1)) opt.grid.attenuation_alpha = strtof(argv[++i], nullptr);
        else if (!strcmp(argv[i], "--loader-noise") && need(1)) { opt.loader_noise = 1; opt.noise_seed = (uint32_t)strtoul(argv[++i], nullptr, 10); }
        else if (!strcmp(argv[i], "--quiet")) opt.verbose = 0;
        else if (!strcmp(argv[i], "--stats")) want_stats = 1;
        else { fprintf(stderr, "ray_voxel: unknown or incomplete option '%s'\n", argv[i]); return 1; }
    }
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t_start = now();
    pvp_ctx *ctx = nullptr;
    if (pvp_ctx_create(device, nullptr, &ctx) != PVP_OK) { fprintf(stderr, "%s\n", pvp_last_error()); return 1; }
    const auto t_ctx = now();
    const size_t cells = (size_t)opt.grid.n * opt.grid.n * opt.grid.n;
    void *grid_dev = nullptr;
    float *grid_f32_dev = nullptr;
    int rc = pvp_dev_malloc(ctx, pvp_grid_bytes(&opt.grid), &grid_dev);
    if (!rc) rc = pvp_dev_memset(ctx, grid_dev, 0, pvp_grid_bytes(&opt.grid));
    pvp_stats st = {0, 0, 0, 0};
    pvp_run_report rep;
    memset(&rep, 0, sizeof(rep));
    if (!rc) rc = pvp_ray_voxel_accumulate(ctx, argv[1], argv[2], grid_dev, &opt, &st, &rep);
    if (rc) {
        fprintf(stderr, "%s\n", pvp_last_error());
        if (grid_dev) pvp_dev_free(ctx, grid_dev);
        pvp_ctx_destroy(ctx);
        return 1;
    }
    const auto t_run = now();
    // .bin payload is always float32 (ray_voxel.cpp:552); it streams from the device to the file (:542-555)
    const float *cells_dev = (const float *)grid_dev;
    if (opt.grid.accum_mode == PVP_ACCUM_FIXED) {
        rc = pvp_dev_malloc(ctx, cells * sizeof(float), (void **)&grid_f32_dev);
        if (!rc) rc = pvp_fixed_to_f32(ctx, (const int64_t *)grid_dev, grid_f32_dev, (int64_t)cells, opt.grid.frac_bits);
        cells_dev = grid_f32_dev;
    }
    if (!rc) rc = pvp_write_bin_dev(ctx, argv[3], opt.grid.n, opt.grid.voxel_size, cells_dev);  // "Cannot open output file: ..." (:545)
    if (grid_f32_dev) pvp_dev_free(ctx, grid_f32_dev);
    const auto t_write = now();
    pvp_dev_free(ctx, grid_dev);
    pvp_ctx_destroy(ctx);
    if (rc) { fprintf(stderr, "%s\n", pvp_last_error()); return 1; }
    printf("Saved voxel grid to %s\n", argv[3]);  // :554
    if (want_stats)
        fprintf(stderr, "frames=%llu pairs=%llu rays=%llu voxel_steps=%llu  seconds: context %.2f, frames %.2f, .bin %.2f\n",
                (unsigned long long)rep.n_frames_loaded, (unsigned long long)rep.n_pairs, (unsigned long long)st.n_rays,
                (unsigned long long)st.n_steps, secs(t_start, t_ctx), secs(t_ctx, t_run), secs(t_run, t_write));
    return 0;
}

}  // extern "C"
