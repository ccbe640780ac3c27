// tools/rt_sessions.cpp -- real-time session sweep (BASELINE.json config 5) through the reference's own threading model:
// N caller threads, one encoder session each, every thread paced at `fps` and blocked in b200enc_encode for its frame (what the
// cloud-phone caller does with EncodeOneFrame, reference video_codec/VideoEncoderOpenH264.cpp:304-352, iMultipleThreadIdc = 1 at :294).
// The per-GPU auto_batch scheduler of libb200enc coalesces the concurrent calls. Prints one JSON line: latency percentiles of the
// encode call, frames that finished after the next frame's capture time ("late"), and the sessions' achieved frame rate.
// usage: rt_sessions <sessions> <seconds> [width height fps bitrate input_format(0 i420 | 2 rgba) devices(0 = all) profile(0 baseline | 1 main | 2 high) slices]
#include "../include/b200enc.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

using clk = std::chrono::steady_clock;

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 8; const double seconds = argc > 2 ? atof(argv[2]) : 3.0;
    const int W = argc > 3 ? atoi(argv[3]) : 1920, H = argc > 4 ? atoi(argv[4]) : 1080, fps = argc > 5 ? atoi(argv[5]) : 30;
    const int bitrate = argc > 6 ? atoi(argv[6]) : 4000000, fmt = argc > 7 ? atoi(argv[7]) : 0;
    int ndev = b200enc_device_count(); if (argc > 8 && atoi(argv[8]) > 0) ndev = std::min(ndev, atoi(argv[8]));
    const int profile = argc > 9 ? atoi(argv[9]) : 0, slices = argc > 10 ? atoi(argv[10]) : 1;      // 0 Baseline / CAVLC, 1 Main / CABAC, 2 High / CABAC
    if (ndev <= 0) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    // frame pool in pinned memory: a translating band-limited texture (deterministic), shared by all sessions
    const int POOL = 8; const size_t fb = fmt == 2 ? (size_t)W * H * 4 : (size_t)W * H * 3 / 2;
    std::vector<uint8_t *> pool(POOL);
    std::vector<uint8_t> tex((size_t)(W + 64) * (H + 64));
    { uint32_t s = 12345; std::vector<int> n(tex.size());
      for (auto &v : n) { s = s * 1664525u + 1013904223u; v = (s >> 24); }
      const int TW = W + 64;
      for (size_t i = 0; i < tex.size(); i++) {     // 5-tap box blur in x and (approximately) y for some spatial correlation
          long a = 0; int c = 0;
          for (int d = -2; d <= 2; d++) for (int e = -2; e <= 2; e++) { long j = (long)i + d + (long)e * TW; if (j >= 0 && j < (long)tex.size()) { a += n[j]; c++; } }
          tex[i] = (uint8_t)std::min(235L, std::max(16L, 128 + (a / c - 128) * 3));
      } }
    for (int t = 0; t < POOL; t++) {
        pool[t] = static_cast<uint8_t *>(b200enc_host_alloc(fb));
        if (!pool[t]) { printf("{\"error\": \"pinned allocation failed\"}\n"); return 1; }
        const int ox = 3 * t, oy = 2 * t, TW = W + 64;
        if (fmt == 2) { for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) { uint8_t v = tex[(size_t)(y + oy) * TW + x + ox]; uint8_t *p = pool[t] + ((size_t)y * W + x) * 4; p[0] = p[1] = p[2] = v; p[3] = 255; } }
        else { for (int y = 0; y < H; y++) memcpy(pool[t] + (size_t)y * W, &tex[(size_t)(y + oy) * TW + ox], W); memset(pool[t] + (size_t)W * H, 128, (size_t)W * H / 2); }
    }
    std::vector<b200enc_session *> sess(N, nullptr);
    for (int i = 0; i < N; i++) {
        b200enc_config c; b200enc_default_config(&c);
        c.width = W; c.height = H; c.fps = fps; c.bitrate = bitrate; c.gop = 300; c.input_format = fmt; c.auto_batch = 1; c.device = i % ndev; c.profile = profile; c.num_slices = slices;
        const int rc = b200enc_create(&c, &sess[i]);
        if (rc) { printf("{\"error\": \"create %d failed: %s\"}\n", i, b200enc_strerror(rc)); return 1; }
    }
    // untimed first frame of every session (its IDR, plus one-time CUDA module loading on the very first call)
    { std::vector<std::thread> w;
      for (int i = 0; i < N; i++) w.emplace_back([&, i] { const uint8_t *bs; uint32_t n; if (b200enc_encode(sess[i], pool[0], (uint32_t)fb, &bs, &n, nullptr) != 0) exit(2); });
      for (auto &t : w) t.join(); }
    const auto period = std::chrono::nanoseconds(1000000000LL / fps);
    const auto t_start = clk::now() + std::chrono::milliseconds(50);
    std::atomic<long> idr_frames{ 0 };
    std::vector<std::vector<float>> lat(N);
    std::atomic<long> late{ 0 }, frames{ 0 }, errors{ 0 };
    std::vector<std::thread> th;
    for (int i = 0; i < N; i++) th.emplace_back([&, i] {
        // sessions start staggered over the first frame period (capture clocks of different phones are not aligned)
        auto next = t_start + std::chrono::nanoseconds((long long)(period.count() * (double)i / N));
        const auto t_end = t_start + std::chrono::nanoseconds((long long)(seconds * 1e9));
        int k = 1;          // the untimed first frame was pool[0]: continue with its temporal neighbour (a jump would be a scene cut, not steady state)
        while (next < t_end) {
            std::this_thread::sleep_until(next);
            const auto t0 = clk::now();
            const uint8_t *bs; uint32_t n;
            const int p = k % (2 * POOL - 2), idx = p < POOL ? p : 2 * POOL - 2 - p; k++;
            // IDR phases are spread over the GOP like sessions that started at different times: session i refreshes at frame
            // (37 i) mod 300 of the run and every 300 frames after that
            if ((k - 1) == (37 * i) % 300) { b200enc_force_idr(sess[i]); idr_frames++; }
            if (b200enc_encode(sess[i], pool[idx], (uint32_t)fb, &bs, &n, nullptr) != 0) errors++;
            const auto t1 = clk::now();
            lat[i].push_back(std::chrono::duration<float, std::milli>(t1 - t0).count());
            frames++;
            next += period;
            if (t1 > next) { late++; while (next < t1) next += period; }      // missed the next capture: drop to the following one
        }
    });
    for (auto &t : th) t.join();
    const double wall = std::chrono::duration<double>(clk::now() - t_start).count();
    std::vector<float> all; for (auto &v : lat) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    auto pct = [&](double q) { return all.empty() ? 0.f : all[std::min(all.size() - 1, (size_t)(q * all.size()))]; };
    uint64_t nb = 0, nf = 0; double avg_batch = 0; int nd = 0;
    for (int d = 0; d < ndev; d++) if (b200enc_scheduler_stats(d, &nb, &nf) == 0 && nb) { avg_batch += (double)nf / nb; nd++; }
    printf("{\"sessions\": %d, \"devices\": %d, \"width\": %d, \"height\": %d, \"fps\": %d, \"seconds\": %.1f, \"frames\": %ld, \"achieved_fps_per_session\": %.2f, "
           "\"latency_ms\": {\"p50\": %.2f, \"p95\": %.2f, \"p99\": %.2f, \"max\": %.2f}, \"late_frames\": %ld, \"idr_frames\": %ld, \"errors\": %ld, \"avg_batch\": %.1f, \"realtime\": %s}\n",
           N, ndev, W, H, fps, seconds, frames.load(), frames.load() / wall / N, pct(0.5), pct(0.95), pct(0.99), all.empty() ? 0.f : all.back(), late.load(), idr_frames.load(), errors.load(),
           nd ? avg_batch / nd : 0.0, (late.load() == 0 && errors.load() == 0 && pct(0.99) <= 1000.0 / fps) ? "true" : "false");
    { unsigned long r = 0; for (auto s : sess) r += b200enc_rc_retries(s); fprintf(stderr, "rate control: %lu pictures coded twice (hard cap)\n", r); }
    for (auto s : sess) b200enc_destroy(s);
    for (auto p : pool) b200enc_host_free(p);
    return 0;
}
