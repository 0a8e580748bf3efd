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

#include <nvtx3/nvToolsExt.h>

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

// ---- frame stream: files -> pinned ring -> device ring -> K1/K2 batches ----------------------------------------
// Once K2 is fast, getting frames to it is what limits frames/s through the CLI.  The frames a run will visit are known up front
// (metadata order), so:
//   * a pool of host threads decodes them ahead of the consumer STRAIGHT INTO a ring of pinned slots (a raw PGM frame is one copy,
//     page cache -> pinned memory; no heap traffic, no consumer-side memcpy);
//   * the consumer sees the frames strictly in order -- skip-on-error / size-mismatch / prev<-curr semantics (ray_voxel.cpp:464-481,
//     :534) are decided there, untouched -- and sends each loadable frame ONCE over PCIe into a ring of device slots (the previous
//     frame of a pair is already there: it was the current frame of the pair before);
//   * every `batch` pairs one K1 + K2 batch is launched on the device ring; the copy stream runs ahead of the kernels by up to two
//     batches (a device slot is rewritten only after the batch that last read it has finished: event wait on the copy stream), the
//     decoders run ahead of the copies by the pinned ring.
// Nothing here blocks the consumer on the GPU except ring pressure.
class DecodePool {
public:
    struct Item { int status = 0; int w = 0, h = 0; std::vector<uint8_t> spill; };   // status: 1 pixels in the slot, -1 in spill (larger than a slot), 0 failed

    // ring: n_slots pinned slots of slot_stride bytes; a worker decodes into slot + dst_off, capacity dst_cap bytes.  copied[slot]: the
    // event the consumer records after enqueueing the slot's H2D copy; a worker waits for it before it overwrites the slot.
    DecodePool(std::vector<std::string> paths, int threads, uint8_t *ring, size_t slot_stride, size_t dst_off, size_t dst_cap, int n_slots,
               cudaEvent_t *copied, int device)
        : paths_(std::move(paths)), ring_(ring), stride_(slot_stride), off_(dst_off), cap_(dst_cap), copied_(copied), device_(device)
    {
        const int n = (int)paths_.size();
        threads = std::max(1, std::min(threads, std::max(1, n)));
        items_.resize((size_t)std::max(1, n_slots));
        state_.assign(items_.size(), -1);
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
    int slot_of(int idx) const { return (int)((size_t)idx % items_.size()); }
    uint8_t *pixels(int idx) { return ring_ + (size_t)slot_of(idx) * stride_; }
    // frame `idx` (consumed in increasing order); valid until release(idx)
    Item &get(int idx)
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return state_[(size_t)slot_of(idx)] == idx; });
        return items_[(size_t)slot_of(idx)];
    }
    void release(int idx)
    {
        {
            std::lock_guard<std::mutex> g(m_);
            state_[(size_t)slot_of(idx)] = -1;
            consumed_ = idx + 1;
        }
        cv_.notify_all();
    }

private:
    void work()
    {
        cudaSetDevice(device_);   // the slots' events belong to the context's device
        for (;;) {
            const int idx = next_.fetch_add(1);
            if (idx >= (int)paths_.size()) return;
            const size_t slot = (size_t)slot_of(idx);
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return stop_ || (idx < consumed_ + (int)items_.size() && state_[slot] == -1); });
                if (stop_) return;
                state_[slot] = -2;  // being filled
            }
            if (copied_[slot]) cudaEventSynchronize(copied_[slot]);   // the slot's previous frame has left for the device
            Item &it = items_[slot];
            it.status = pvp::decode_gray_file_into(paths_[(size_t)idx].c_str(), ring_ + slot * stride_ + off_, cap_, &it.w, &it.h, it.spill);
            {
                std::lock_guard<std::mutex> g(m_);
                state_[slot] = idx;
            }
            cv_.notify_all();
        }
    }
    std::vector<std::string> paths_;
    uint8_t *ring_;
    size_t stride_, off_, cap_;
    cudaEvent_t *copied_;
    int device_;
    std::vector<Item> items_;
    std::vector<int> state_;  // -1 free, -2 being filled, else the frame index it holds
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::atomic<int> next_{0};
    int consumed_ = 0;
    bool stop_ = false;
};

double wall_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct FrameStream {
    pvp_ctx *ctx;
    const pvp_run_options *opt;
    void *grid_dev;
    // pinned ring (host side of PCIe)
    uint8_t *hring = nullptr;
    size_t hslot = 0;               // bytes per pinned slot
    int n_hslots = 0;
    std::vector<cudaEvent_t> copied;  // per pinned slot: its H2D copy has finished
    // device ring
    uint8_t *dring = nullptr;
    size_t dslot = 0;               // bytes per device slot (a multiple of 16)
    int n_dslots = 0, dnext = 0;
    std::vector<long long> dslot_batch;  // id of the last batch that reads the slot (-1: none)
    static constexpr int NB = 64;
    cudaEvent_t batch_done[NB] = {};
    cudaEvent_t ev_up = nullptr;
    long long open_batch = 0;
    // open batch
    int W = 0, H = 0, prev_ds = -1, batch_pairs = 32;
    std::vector<pvp_pair> pairs;
    uint64_t frames_loaded = 0, pairs_done = 0;
    double s_wait = 0, s_upload = 0, s_launch = 0;
    // loader noise (ray_voxel.cpp:203-218): frames become float32, one generator for the whole run, seeded once
    bool noise = false;
    size_t elem = 1;
    std::mt19937 gen;
    std::vector<float> noise_tmp, spill_f32;

    ~FrameStream()
    {
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        for (auto e : copied) if (e) cudaEventDestroy(e);
        for (auto e : batch_done) if (e) cudaEventDestroy(e);
        if (ev_up) cudaEventDestroy(ev_up);
        // the rings stay with the context for the next call (pvp_ctx_destroy frees them)
    }
    int init(size_t frame_px, int host_slots)
    {
        const size_t px = std::max<size_t>(frame_px, 16);
        hslot = (px * elem + 4095) / 4096 * 4096;
        n_hslots = std::max(2, host_slots);
        if (ctx->fs_hring_bytes < hslot * (size_t)n_hslots) {
            if (ctx->fs_hring) cudaFreeHost(ctx->fs_hring);
            ctx->fs_hring = nullptr; ctx->fs_hring_bytes = 0;
            PVP_CUDA(cudaMallocHost((void **)&ctx->fs_hring, hslot * (size_t)n_hslots));
            ctx->fs_hring_bytes = hslot * (size_t)n_hslots;
        }
        hring = ctx->fs_hring;
        copied.assign((size_t)n_hslots, nullptr);
        for (auto &e : copied) PVP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : batch_done) PVP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        PVP_CUDA(cudaEventCreateWithFlags(&ev_up, cudaEventDisableTiming));
        n_dslots = 3 * batch_pairs + 3;   // three batches of frames in flight + the carried previous frame
        return grow_device(px * elem);
    }
    // (re)allocate the device ring with slots of at least `bytes`; the previous frame, if any, moves to slot 0 of the new ring
    int grow_device(size_t bytes)
    {
        const size_t want = (bytes + 255) / 256 * 256;
        if (dring && want <= dslot) return PVP_OK;
        int rc = flush();
        if (rc) return rc;
        PVP_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        PVP_CUDA(cudaStreamSynchronize(ctx->stream));
        uint8_t *nr = nullptr;
        if (!dring && ctx->fs_dring_bytes >= want * (size_t)n_dslots) {
            nr = ctx->fs_dring;            // the ring of an earlier call is large enough
        } else {
            PVP_CUDA(cudaMalloc((void **)&nr, want * (size_t)n_dslots));
            if (dring && prev_ds >= 0) PVP_CUDA(cudaMemcpy(nr, dring + (size_t)prev_ds * dslot, (size_t)W * H * elem, cudaMemcpyDeviceToDevice));
            if (ctx->fs_dring) cudaFree(ctx->fs_dring);
            ctx->fs_dring = nr; ctx->fs_dring_bytes = want * (size_t)n_dslots;
        }
        dring = nr; dslot = want;
        dslot_batch.assign((size_t)n_dslots, -1);
        if (prev_ds >= 0) { prev_ds = 0; dnext = 1; } else { dnext = 0; }
        return PVP_OK;
    }
    // launch K1 + K2 over the open batch's pairs (device ring slots); the previous frame stays where it is
    int flush()
    {
        if (pairs.empty()) return PVP_OK;
        const double t0 = wall_s();
        nvtxRangePushA("pvp:K1+K2 batch");
        PVP_CUDA(cudaEventRecord(ev_up, ctx->copy_stream));
        PVP_CUDA(cudaStreamWaitEvent(ctx->stream, ev_up, 0));   // the batch's frames are on the device
        int rc = pvp_motion_accumulate(ctx, dring, noise ? PVP_FRAME_F32 : PVP_FRAME_U8, (int64_t)(dslot / elem), W, H, n_dslots, pairs.data(),
                                       (int)pairs.size(), opt->threshold, grid_dev, &opt->grid, nullptr);
        nvtxRangePop();
        if (rc) return rc;
        PVP_CUDA(cudaEventRecord(batch_done[open_batch % NB], ctx->stream));
        pairs_done += pairs.size();
        pairs.clear();
        ++open_batch;
        s_launch += wall_s() - t0;
        return PVP_OK;
    }
    // One decoded frame (w x h, 8-bit gray at src: a pinned slot, or pageable memory with pinned_slot < 0) -> a free device slot.
    // Returns the device slot through *ds_out.  Pinned source: asynchronous, `copied[pinned_slot]` marks the end of the copy.
    int upload(uint8_t *src, int pinned_slot, int w, int h, int *ds_out)
    {
        const double t0 = wall_s();
        const size_t npx = (size_t)w * h, bytes = npx * elem;
        int rc = grow_device(bytes);
        if (rc) return rc;
        const void *from = src;
        if (noise) {  // load_image_gray, ray_voxel.cpp:203-218: same generator, same distribution, same draw order (pixel order, frame after frame)
            std::uniform_real_distribution<float> noise_dist(-1.0f, 1.0f);
            noise_tmp.resize(npx);
            for (size_t i = 0; i < npx; ++i) noise_tmp[i] = noise_dist(gen);
            float *dst;
            const uint8_t *px = src;
            if (pinned_slot >= 0) {
                dst = reinterpret_cast<float *>(hring + (size_t)pinned_slot * hslot);   // the 8-bit pixels sit in the last quarter of the slot: expand in place, forwards
            } else {
                spill_f32.resize(npx);
                dst = spill_f32.data();
            }
            for (size_t i = 0; i < npx; ++i) {
                float val = static_cast<float>(px[i]);
                val += noise_tmp[i];
                if (val < 0.0f) val = 0.0f;
                if (val > 255.0f) val = 255.0f;
                dst[i] = val;
            }
            from = dst;
        }
        int ds = dnext;
        if (ds == prev_ds) ds = (ds + 1) % n_dslots;   // the previous frame is still needed by the next pair
        dnext = (ds + 1) % n_dslots;
        if (dslot_batch[(size_t)ds] == open_batch) {   // read by the batch that is still open: submit it first
            if ((rc = flush())) return rc;
        }
        if (dslot_batch[(size_t)ds] >= 0)              // read by a submitted batch: the copy waits for that batch's kernels
            PVP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, batch_done[dslot_batch[(size_t)ds] % NB], 0));
        nvtxRangePushA("pvp:H2D frame");
        {
            const cudaError_t e = cudaMemcpyAsync(dring + (size_t)ds * dslot, from, bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
            if (e != cudaSuccess) {   // say what was asked for: the ring arithmetic is the only thing that can be wrong here
                cudaPointerAttributes pa = {}, pb = {};
                cudaPointerGetAttributes(&pa, dring + (size_t)ds * dslot);
                cudaPointerGetAttributes(&pb, from);
                cuda_fail(e, "cudaMemcpyAsync(frame H2D)", __FILE__, __LINE__);
                std::string why = pvp_last_error();
                set_error("%s [device ring %p + slot %d x %zu B of %d slots (memory type %d), source %p (type %d, pinned ring %p, %d slots x %zu B, slot %d), %zu bytes, frame %dx%d]",
                          why.c_str(), (void *)dring, ds, dslot, n_dslots, (int)pa.type, from, (int)pb.type, (void *)hring, n_hslots, hslot, pinned_slot, bytes, w, h);
                nvtxRangePop();
                return PVP_ERR_CUDA;
            }
        }
        if (pinned_slot >= 0) PVP_CUDA(cudaEventRecord(copied[(size_t)pinned_slot], ctx->copy_stream));
        nvtxRangePop();
        dslot_batch[(size_t)ds] = -1;
        ++frames_loaded;
        *ds_out = ds;
        s_upload += wall_s() - t0;
        return PVP_OK;
    }
    // the frame in device slot ds becomes the current frame of a pair with the previous one (pose of the CURRENT frame, :486-491)
    int add_pair(int ds, const pvp_frame_info &ci, int w)
    {
        pvp_pair p;
        memcpy(p.cam, ci.camera_position, sizeof(p.cam));
        pvp_rotation_matrix(ci.yaw, ci.pitch, ci.roll, p.R);
        p.focal = pvp_focal_length(ci.fov_degrees, w);
        p.prev = prev_ds; p.curr = ds;
        pairs.push_back(p);
        dslot_batch[(size_t)prev_ds] = open_batch;
        dslot_batch[(size_t)ds] = open_batch;
        prev_ds = ds;
        if ((int)pairs.size() >= batch_pairs) return flush();
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
    const double t_begin = wall_s();
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
    const std::string folder = image_folder;
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
            for (long long k = a; k < b; ++k) visit.push_back(folder + "/" + cf[(size_t)k].image_file);  // :467
        }
    }
    int decode_threads = opt->decode_threads;
    if (decode_threads <= 0) {
        const char *env = getenv("PVP_DECODE_THREADS");
        if (env) {
            decode_threads = atoi(env);
        } else {  // up to 32 threads, sharing the host's cores with the other ranks of the node (torchrun exports LOCAL_WORLD_SIZE)
            const char *lws = getenv("LOCAL_WORLD_SIZE");
            const unsigned ranks = lws && atoi(lws) > 0 ? (unsigned)atoi(lws) : 1u;
            decode_threads = (int)std::min(32u, std::max(2u, std::thread::hardware_concurrency() / ranks));
        }
    }
    decode_threads = std::max(1, decode_threads);

    FrameStream R{ctx, opt, grid_dev};
    R.batch_pairs = opt->chunk_frames > 1 ? opt->chunk_frames - 1 : 32;
    if (opt->loader_noise) {
        R.noise = true; R.elem = sizeof(float);
        R.gen.seed(opt->noise_seed);  // the reference: static std::mt19937 gen(rd()) -- here rd() is the caller's seed
    }
    // ring slots are sized by the first frame that decodes (a sequence normally has one resolution); larger frames later on take
    // the spill path of the pool and grow the device ring
    std::vector<uint8_t> tmp;
    size_t first_px = 0;
    for (size_t i = 0; i < visit.size() && !first_px; ++i) {
        int w, h;
        if (decode_gray_file(visit[i].c_str(), tmp, &w, &h)) first_px = (size_t)w * h;
    }
    if ((rc = R.init(first_px, 2 * decode_threads + 4))) return rc;
    uint64_t skipped = 0, mismatched = 0;

    // noise mode: the 8-bit pixels go into the last quarter of the (4 bytes per pixel) slot and are expanded in place
    const size_t cap_px = R.hslot / R.elem;
    DecodePool pool(std::move(visit), decode_threads, R.hring, R.hslot, R.noise ? 3 * cap_px : 0, cap_px, R.n_hslots, R.copied.data(), ctx->device);
    int visit_idx = 0;
    const double t_loop = wall_s();

    long long slot = 0;
    for (auto &kv : by_cam) {
        auto &cf = kv.second;
        if (cf.size() < 2) continue;
        const long long cam_lo = slot, cam_hi = slot + (long long)cf.size();
        slot = cam_hi;
        const long long a = std::max(lo, cam_lo) - cam_lo, b = std::min(hi, cam_hi) - cam_lo;
        if (a >= b) continue;
        R.prev_ds = -1;   // no pair across cameras
        // the frame a pair starting this block compares against: last loadable frame before it (serial load, pageable source)
        for (long long k = a - 1; k >= 0 && R.prev_ds < 0; --k) {
            int w, h;
            const std::string path = folder + "/" + cf[(size_t)k].image_file;
            if (decode_gray_file(path.c_str(), tmp, &w, &h)) {
                int ds;
                if ((rc = R.upload(tmp.data(), -1, w, h, &ds))) return rc;
                R.W = w; R.H = h; R.prev_ds = ds;
            }
        }
        for (long long k = a; k < b; ++k) {
            const pvp_frame_info &ci = cf[(size_t)k];
            const int idx = visit_idx++;
            const double t0 = wall_s();
            nvtxRangePushA("pvp:decode wait");
            DecodePool::Item &it = pool.get(idx);
            nvtxRangePop();
            R.s_wait += wall_s() - t0;
            if (it.status == 0) {  // :470-473
                pool.release(idx);
                if (opt->verbose) fprintf(stderr, "Failed to load image: %s/%s\nSkipping frame due to load error.\n", folder.c_str(), ci.image_file);
                ++skipped;
                continue;
            }
            const int w = it.w, h = it.h;
            const bool in_slot = it.status == 1;
            uint8_t *src = in_slot ? pool.pixels(idx) + (R.noise ? 3 * cap_px : 0) : it.spill.data();
            const bool first = R.prev_ds < 0;                      // :475-481
            const bool mismatch = !first && (w != R.W || h != R.H);  // detect_motion's size check, :238-243: no rays, prev still advances (:534)
            if (mismatch) {
                if (opt->verbose) fprintf(stderr, "Images differ in size. Can't do motion detection!\n");
                ++mismatched;
                if ((rc = R.flush())) return rc;   // a batch has one frame size
                R.prev_ds = -1;                   // the old previous frame is not needed any more
            }
            int ds;
            rc = R.upload(src, in_slot ? pool.slot_of(idx) : -1, w, h, &ds);
            pool.release(idx);   // a pinned slot is rewritten only after its `copied` event
            if (rc) return rc;
            if (first || mismatch) {
                R.W = w; R.H = h; R.prev_ds = ds;
                continue;
            }
            if ((rc = R.add_pair(ds, ci, w))) return rc;
        }
        if ((rc = R.flush())) return rc;  // camera boundary
    }
    const double t_drain = wall_s();
    pvp_stats st = {0, 0, 0, 0};
    rc = pvp_motion_accumulate(ctx, nullptr, PVP_FRAME_U8, 1, 1, 1, 0, nullptr, 0, 0.f, grid_dev, &opt->grid, &st);  // sync + counters
    if (rc) return rc;
    const double t_end = wall_s();
    if (stats) { stats->n_rays += st.n_rays; stats->n_steps += st.n_steps; stats->n_written += st.n_written; stats->n_atomics += st.n_atomics; }
    if (report) {
        report->n_frames_listed = (uint64_t)n; report->n_frames_loaded = R.frames_loaded; report->n_frames_skipped = skipped;
        report->n_pairs = R.pairs_done; report->n_size_mismatch = mismatched;
        report->seconds_total = wall_s() - t_begin; report->seconds_decode_wait = R.s_wait; report->seconds_upload = R.s_upload;
        report->seconds_launch = R.s_launch; report->decode_threads = decode_threads;
        report->seconds_setup = t_loop - t_begin; report->seconds_drain = t_end - t_drain;
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
    if (want_stats) pvp_ctx_profile(ctx, 1);   // CUDA events around every K1 / K2 launch: the GPU side of the stage split below
    if (!rc) rc = pvp_ray_voxel_accumulate(ctx, argv[1], argv[2], grid_dev, &opt, &st, &rep);
    if (rc) {
        fprintf(stderr, "%s\n", pvp_last_error());
        if (grid_dev) pvp_dev_free(ctx, grid_dev);
        pvp_ctx_destroy(ctx);
        return 1;
    }
    const auto t_run = now();
    pvp_profile prof;
    memset(&prof, 0, sizeof(prof));
    if (want_stats) pvp_ctx_get_profile(ctx, &prof, 1);
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
    if (want_stats)   // the host stages carry the same names as NVTX ranges (pvp:decode wait, pvp:H2D frame, pvp:K1+K2 batch, pvp:.bin)
        fprintf(stderr, "frames=%llu pairs=%llu rays=%llu voxel_steps=%llu  seconds: context %.2f, frames %.2f, .bin %.2f | frame loop: decode wait %.3f, "
                "H2D enqueue %.3f, batch launch %.3f, set-up %.3f, final drain %.3f (host, %d decode threads); GPU K1 %.3f, K2 %.3f\n",
                (unsigned long long)rep.n_frames_loaded, (unsigned long long)rep.n_pairs, (unsigned long long)st.n_rays,
                (unsigned long long)st.n_steps, secs(t_start, t_ctx), secs(t_ctx, t_run), secs(t_run, t_write), rep.seconds_decode_wait,
                rep.seconds_upload, rep.seconds_launch, rep.seconds_setup, rep.seconds_drain, rep.decode_threads, prof.ms_k1 * 1e-3, prof.ms_k2 * 1e-3);
    return 0;
}

}  // extern "C"
