// tools/e2e_plugin.cpp -- end-to-end driver through the reference's real boundary (bench.py's `e2e` leg; also a standalone tool).
//
// It does what the cloud-phone caller does and nothing else: dlopen("libVideoCodec.so"), CreateVideoEncoder -> InitEncoder ->
// StartEncoder -> EncodeOneFrame x N -> StopEncoder -> DestroyEncoder -> DestroyVideoEncoder (reference video_codec/VideoCodecApi.h:22-96),
// configured only through the Android properties the wrapper reads (VideoEncoderOpenH264.cpp:62-122; selector 3 = the B200 sibling),
// with ONE CALLER THREAD PER SESSION, each blocked in its own EncodeOneFrame (the reference runs one single-threaded encoder per
// session, :294), and frames in plain malloc memory -- caller-owned, pageable, exactly what VideoCodecApi.h:57-58 hands over.
// Nothing of libb200enc's own API (batches, pinned allocators) is used here.
//
// As a library (tools/libe2e_plugin.so, loaded by bench.py through ctypes): e2e_open / e2e_run / e2e_close.
// As a program (tools/e2e_plugin.bin): e2e_plugin <libVideoCodec.so> <sessions> <steps> [width height fps bitrate profile paced(0|1) device]
#include "../include/VideoCodecApi.h"
#include <dlfcn.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {
using clk = std::chrono::steady_clock;
typedef EncoderRetCode (*CreateFn)(VideoEncoder **);
typedef EncoderRetCode (*DestroyFn)(VideoEncoder *);
typedef int (*PropSetFn)(const char *, const char *);

struct E2E {
    void *lib = nullptr; CreateFn create = nullptr; DestroyFn destroy = nullptr; PropSetFn prop_set = nullptr;
    std::vector<VideoEncoder *> enc;
    std::vector<uint8_t *> pool; size_t frame_bytes = 0;
    void (*pinned_free)(void *) = nullptr;          // E2E_POOL_PINNED=1: the frames live in memory from b200enc_host_alloc (an integrator that owns its capture buffers)
    int fps = 30;
    std::string error;
};

int pool_index(long step, int sess, int n)      // ping-pong walk so consecutive frames of a session stay temporally adjacent (bench.py)
{
    if (n < 2) return 0;
    const long t = (step + 3L * sess) % (2 * n - 2);
    return (int)(t < n ? t : 2 * n - 2 - t);
}
} // namespace

extern "C" {

struct e2e_result {
    double seconds;            // wall time from the common start to the last thread's last frame
    uint64_t frames, bytes;    // access units delivered, their total size
    uint64_t errors, late;     // EncodeOneFrame failures; paced mode: frames that finished after the next capture time
    double lat_p50_ms, lat_p99_ms, lat_max_ms;
};

const char *e2e_last_error(void *h) { return h ? static_cast<E2E *>(h)->error.c_str() : "null handle"; }

// pool: pool_frames tightly packed frames of frame_bytes each (copied into per-frame malloc blocks: pageable caller memory)
void *e2e_open(const char *codec_lib, int sessions, int width, int height, int fps, int bitrate, int gop, const char *profile,
               const char *input_format, int device, const uint8_t *pool, int pool_frames, size_t frame_bytes)
{
    E2E *e = new E2E();
    e->lib = dlopen(codec_lib, RTLD_NOW | RTLD_LOCAL);
    if (!e->lib) { e->error = std::string("dlopen: ") + dlerror(); return e; }
    e->create = reinterpret_cast<CreateFn>(dlsym(e->lib, "CreateVideoEncoder"));
    e->destroy = reinterpret_cast<DestroyFn>(dlsym(e->lib, "DestroyVideoEncoder"));
    // on Android these are bionic's; the Linux build of libVideoCodec.so carries an in-memory property store with the same two symbols
    e->prop_set = reinterpret_cast<PropSetFn>(dlsym(e->lib, "__system_property_set"));
    if (!e->create || !e->destroy || !e->prop_set) { e->error = "libVideoCodec.so lacks CreateVideoEncoder / DestroyVideoEncoder / __system_property_set"; return e; }
    auto set = [&](const char *k, const std::string &v) { e->prop_set(k, v.c_str()); };
    set("ro.vmi.demo.video.encode.format", "3");
    set("ro.sys.vmi.cloudphone", "video");
    set("ro.hardware.width", std::to_string(width)); set("ro.hardware.height", std::to_string(height)); set("ro.hardware.fps", std::to_string(fps));
    set("persist.vmi.video.encode.bitrate", std::to_string(bitrate)); set("persist.vmi.video.encode.gopsize", std::to_string(gop));
    set("persist.vmi.video.encode.profile", profile && *profile ? profile : "baseline");
    set("persist.vmi.video.encode.param_adjusting", "0"); set("persist.vmi.video.encode.keyframe", "0");
    set("persist.vmi.b200.encode.input_format", input_format && *input_format ? input_format : "i420");
    set("persist.vmi.b200.encode.device", device >= 0 ? std::to_string(device) : std::string(""));
    e->fps = fps; e->frame_bytes = frame_bytes;
    // the secondary leg of bench.py: frames in pinned memory handed out by the encoder library instead of malloc memory -- what an
    // integrator who owns the capture buffers can do; the symbols are found through libVideoCodec.so's dependency, nothing else of that API is used
    void *(*pinned_alloc)(size_t) = nullptr;
    if (const char *pe = getenv("E2E_POOL_PINNED")) if (atoi(pe)) {
        pinned_alloc = reinterpret_cast<void *(*)(size_t)>(dlsym(e->lib, "b200enc_host_alloc"));
        e->pinned_free = reinterpret_cast<void (*)(void *)>(dlsym(e->lib, "b200enc_host_free"));
        if (!pinned_alloc || !e->pinned_free) { e->error = "E2E_POOL_PINNED: b200enc_host_alloc / b200enc_host_free not found"; return e; }
    }
    for (int t = 0; t < pool_frames; t++) {
        uint8_t *p = static_cast<uint8_t *>(pinned_alloc ? pinned_alloc(frame_bytes) : malloc(frame_bytes));
        if (!p) { e->error = "malloc"; return e; }
        memcpy(p, pool + (size_t)t * frame_bytes, frame_bytes);
        e->pool.push_back(p);
    }
    for (int i = 0; i < sessions; i++) {
        VideoEncoder *v = nullptr;
        if (e->create(&v) != VIDEO_ENCODER_SUCCESS || !v) { e->error = "CreateVideoEncoder failed"; return e; }
        e->enc.push_back(v);
        if (v->InitEncoder() != VIDEO_ENCODER_SUCCESS) { e->error = "InitEncoder failed"; return e; }
        if (v->StartEncoder() != VIDEO_ENCODER_SUCCESS) { e->error = "StartEncoder failed"; return e; }
    }
    return e;
}

// Every session encodes `steps` frames (frames first_step .. first_step + steps - 1 of its walk through the pool) on its own thread.
// paced = 0: back to back (throughput); paced = 1: one frame per 1/fps, session start times staggered over one period (real time).
// `limit` > 0 runs only the first `limit` sessions (the paced leg may use fewer sessions than the throughput leg).
int e2e_run(void *h, long first_step, int steps, int paced, int limit, e2e_result *out)
{
    E2E *e = static_cast<E2E *>(h);
    if (!e || !e->error.empty() || !out || e->enc.empty() || e->pool.empty()) return -1;
    const int N = limit > 0 ? std::min(limit, (int)e->enc.size()) : (int)e->enc.size(), P = (int)e->pool.size();
    std::vector<std::vector<float>> lat(N);
    std::atomic<uint64_t> bytes{ 0 }, errors{ 0 }, late{ 0 };
    std::vector<clk::time_point> done(N);
    const auto period = std::chrono::nanoseconds(1000000000LL / e->fps);
    const auto t_start = clk::now() + std::chrono::milliseconds(N > 64 ? 30 : 10);
    std::vector<std::thread> th;
    th.reserve(N);
    for (int i = 0; i < N; i++) th.emplace_back([&, i] {
        lat[i].reserve(steps);
        auto next = t_start + (paced ? std::chrono::nanoseconds((long long)(period.count() * (double)i / N)) : std::chrono::nanoseconds(0));
        std::this_thread::sleep_until(next);
        for (int k = 0; k < steps; k++) {
            if (paced) std::this_thread::sleep_until(next);
            const auto t0 = clk::now();
            uint8_t *bs = nullptr; uint32_t n = 0;
            const EncoderRetCode rc = e->enc[i]->EncodeOneFrame(e->pool[pool_index(first_step + k, i, P)], (uint32_t)e->frame_bytes, &bs, &n);
            const auto t1 = clk::now();
            if (rc != VIDEO_ENCODER_SUCCESS || !bs || !n) errors++; else bytes += n;
            lat[i].push_back(std::chrono::duration<float, std::milli>(t1 - t0).count());
            if (paced) { next += period; if (t1 > next) { late++; while (next < t1) next += period; } }
        }
        done[i] = clk::now();
    });
    for (auto &t : th) t.join();
    const auto t_end = *std::max_element(done.begin(), done.end());
    std::vector<float> all;
    for (auto &v : lat) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    auto pct = [&](double q) { return all.empty() ? 0.0 : (double)all[std::min(all.size() - 1, (size_t)(q * all.size()))]; };
    out->seconds = std::chrono::duration<double>(t_end - t_start).count();
    out->frames = (uint64_t)N * steps - errors.load(); out->bytes = bytes.load(); out->errors = errors.load(); out->late = late.load();
    out->lat_p50_ms = pct(0.5); out->lat_p99_ms = pct(0.99); out->lat_max_ms = all.empty() ? 0.0 : all.back();
    return 0;
}

void e2e_close(void *h)
{
    E2E *e = static_cast<E2E *>(h);
    if (!e) return;
    for (VideoEncoder *v : e->enc) { v->StopEncoder(); v->DestroyEncoder(); e->destroy(v); }
    for (uint8_t *p : e->pool) { if (e->pinned_free) e->pinned_free(p); else free(p); }
    // the library stays loaded: its scheduler threads and CUDA context live until process exit
    delete e;
}

} // extern "C"

#ifdef E2E_MAIN
int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s <libVideoCodec.so> <sessions> <steps> [width height fps bitrate profile paced device]\n", argv[0]); return 2; }
    const int N = atoi(argv[2]), steps = atoi(argv[3]);
    const int W = argc > 4 ? atoi(argv[4]) : 1920, H = argc > 5 ? atoi(argv[5]) : 1080, fps = argc > 6 ? atoi(argv[6]) : 30, br = argc > 7 ? atoi(argv[7]) : 4000000;
    const char *profile = argc > 8 ? argv[8] : "baseline"; const int paced = argc > 9 ? atoi(argv[9]) : 0, device = argc > 10 ? atoi(argv[10]) : -1;
    // frame pool: a translating blurred-noise texture, I420 (deterministic)
    const int POOL = 8; const size_t fb = (size_t)W * H * 3 / 2;
    std::vector<uint8_t> tex((size_t)(W + 64) * (H + 64)), pool(fb * POOL);
    { uint32_t s = 12345; std::vector<int> n(tex.size()); for (auto &v : n) { s = s * 1664525u + 1013904223u; v = s >> 24; }
      const int TW = W + 64;
      for (size_t i = 0; i < tex.size(); i++) { long a = 0; int c = 0;
          for (int d = -2; d <= 2; d++) for (int g = -2; g <= 2; g++) { long j = (long)i + d + (long)g * TW; if (j >= 0 && j < (long)tex.size()) { a += n[j]; c++; } }
          tex[i] = (uint8_t)std::min(235L, std::max(16L, 128 + (a / c - 128) * 3)); } }
    for (int t = 0; t < POOL; t++) {
        uint8_t *f = pool.data() + fb * t;
        for (int y = 0; y < H; y++) memcpy(f + (size_t)y * W, &tex[(size_t)(y + 2 * t) * (W + 64) + 3 * t], W);
        memset(f + (size_t)W * H, 128, (size_t)W * H / 2);
    }
    void *h = e2e_open(argv[1], N, W, H, fps, br, 300, profile, "i420", device, pool.data(), POOL, fb);
    if (*e2e_last_error(h)) { printf("{\"error\": \"%s\"}\n", e2e_last_error(h)); return 1; }
    e2e_result r;
    e2e_run(h, 0, 3, 0, 0, &r);                                  // warm-up: the IDR and two P pictures of every session
    if (e2e_run(h, 3, steps, paced, 0, &r) != 0) { printf("{\"error\": \"run failed\"}\n"); return 1; }
    printf("{\"via\": \"VideoEncoder::EncodeOneFrame, one caller thread per session, pageable input\", \"sessions\": %d, \"steps\": %d, \"paced\": %d, \"frames_per_s\": %.1f, "
           "\"bytes_per_frame\": %.0f, \"errors\": %llu, \"late\": %llu, \"latency_ms\": {\"p50\": %.2f, \"p99\": %.2f, \"max\": %.2f}}\n",
           N, steps, paced, r.frames / r.seconds, r.frames ? (double)r.bytes / r.frames : 0.0, (unsigned long long)r.errors, (unsigned long long)r.late,
           r.lat_p50_ms, r.lat_p99_ms, r.lat_max_ms);
    e2e_close(h);
    return 0;
}
#endif
