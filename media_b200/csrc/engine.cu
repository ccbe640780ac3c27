// media_b200/csrc/engine.cu -- host engine and C ABI (include/b200enc.h) of the B200 H.264 encoder.
//
// What it stands in for: everything behind ISVCEncoder for the VideoEncoderOpenH264 wrapper
// (reference: video_codec/VideoEncoderOpenH264.cpp:131-157 init, :228-296 policy, :304-352 per-frame call).
// One session = one encoder object of the reference (own reference picture, rate-control state, bitstream buffer);
// a batch advances N sessions of one GPU by one frame with a single chain of kernel launches.
#include "../../include/b200enc.h"
#include "h264_dev.cuh"
#include "k_pre.cuh"
#include "k_me.cuh"
#include "k_t8.cuh"
#include "k_intra.cuh"
#include "k_deblock.cuh"
#include "k_cavlc.cuh"
#include "k_cabac.cuh"
#include "k_test.cuh"
#include "rate_control.h"
#include <cuda.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

using namespace b200;
#define B200_CABAC_SMEM_KB 100   /* default slab for the coder CTAs (see k_cabac_code launch); B200ENC_CABAC_SMEM_KB overrides */

namespace {

// last CUDA error seen by any thread of the library (callers are served by scheduler worker threads, so a thread-local would hide it)
std::atomic<int> g_last_cuda_error{ 0 };
#define CU_TRY(expr, fail) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { g_last_cuda_error.store((int)e_); fail; } } while (0)

// ---- session -> GPU placement: least load by pixel rate (analogue of the Netint path's EN_ALLOC_LEAST_LOAD,
// reference video_codec/VideoEncoderNetint.cpp:300-302,552-554) ----
std::mutex g_sched_mu;
std::vector<double> g_dev_load;
int sched_acquire(int want, double load)
{
    std::lock_guard<std::mutex> lk(g_sched_mu);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return -1;
    if ((int)g_dev_load.size() < n) g_dev_load.resize(n, 0.0);
    int dev = want;
    if (dev < 0) { dev = 0; for (int i = 1; i < n; i++) if (g_dev_load[i] < g_dev_load[dev]) dev = i; }
    if (dev >= n) return -1;
    g_dev_load[dev] += load;
    return dev;
}
void sched_release(int dev, double load)
{
    std::lock_guard<std::mutex> lk(g_sched_mu);
    if (dev >= 0 && dev < (int)g_dev_load.size()) g_dev_load[dev] -= load;
}

// ---- host bit writer for the parameter sets (7.3.2.1, 7.3.2.2) ----
struct HostBits {
    std::vector<uint8_t> buf; uint32_t acc = 0; int nacc = 0;
    void put(int n, uint32_t v) { for (int i = n - 1; i >= 0; i--) { acc = (acc << 1) | ((v >> i) & 1); if (++nacc == 8) { buf.push_back((uint8_t)acc); acc = 0; nacc = 0; } } }
    void ue(uint32_t v) { uint32_t x = v + 1; int l = 0; while ((x >> l) > 1) l++; put(l, 0); put(l + 1, x); }
    void se(int v) { ue(v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }
    void trailing() { put(1, 1); while (nacc) put(1, 0); }
};
void append_nal(std::vector<uint8_t> &out, int hdr, const std::vector<uint8_t> &rbsp)
{
    out.insert(out.end(), { 0, 0, 0, 1, (uint8_t)hdr });
    int zeros = 0;
    for (uint8_t b : rbsp) {
        if (zeros >= 2 && b <= 3) { out.push_back(3); zeros = 0; }
        out.push_back(b); zeros = b == 0 ? zeros + 1 : 0;
    }
}
int level_for(int w, int h, int fps)   // Table A-1: smallest level whose MaxFS and MaxMBPS cover the stream
{
    static const struct { int idc, mbps, fs; } L[] = {
        { 10, 1485, 99 }, { 11, 3000, 396 }, { 12, 6000, 396 }, { 13, 11880, 396 }, { 20, 11880, 396 }, { 21, 19800, 792 },
        { 22, 20250, 1620 }, { 30, 40500, 1620 }, { 31, 108000, 3600 }, { 32, 216000, 5120 }, { 40, 245760, 8192 },
        { 42, 522240, 8704 }, { 50, 589824, 22080 }, { 51, 983040, 36864 }, { 52, 2073600, 36864 } };
    const int fs = ((w + 15) / 16) * ((h + 15) / 16); const long mbps = (long)fs * fps;
    for (auto &l : L) if (fs <= l.fs && mbps <= l.mbps) return l.idc;
    return 52;
}
// profile: 0 Constrained Baseline / CAVLC, 1 Main / CABAC, 2 High / CABAC with transform_8x8_mode_flag (7.3.2.1.1 adds the chroma
// format fields to the SPS, 7.3.2.2 the transform / scaling-matrix / second chroma offset tail to the PPS)
std::vector<uint8_t> make_parameter_sets(int w, int h, int level, int profile)
{
    std::vector<uint8_t> out;
    const int mbw = (w + 15) / 16, mbh = (h + 15) / 16;
    HostBits s;
    s.put(8, profile == 2 ? 100 : profile == 1 ? 77 : 66); s.put(8, profile == 2 ? 0 : profile == 1 ? 0x40 : 0xC0); s.put(8, (uint32_t)level);
    s.ue(0);
    if (profile == 2) { s.ue(1); s.ue(0); s.ue(0); s.put(1, 0); s.put(1, 0); }
    s.ue(4); s.ue(2); s.ue(1); s.put(1, 0);
    s.ue((uint32_t)(mbw - 1)); s.ue((uint32_t)(mbh - 1));
    s.put(1, 1); s.put(1, 1);
    const int cr = (mbw * 16 - w) / 2, cb = (mbh * 16 - h) / 2;
    if (cr || cb) { s.put(1, 1); s.ue(0); s.ue((uint32_t)cr); s.ue(0); s.ue((uint32_t)cb); } else s.put(1, 0);
    s.put(1, 0); s.trailing();
    append_nal(out, 0x67, s.buf);
    HostBits p;
    p.ue(0); p.ue(0); p.put(1, profile ? 1 : 0); p.put(1, 0); p.ue(0); p.ue(0); p.ue(0); p.put(1, 0); p.put(2, 0);
    p.se(0); p.se(0); p.se(0); p.put(1, 1); p.put(1, 0); p.put(1, 0);
    if (profile == 2) { p.put(1, 1); p.put(1, 0); p.se(0); }
    p.trailing();
    append_nal(out, 0x68, p.buf);
    return out;
}

struct KernelTime { const char *name; cudaEvent_t ev0, ev1; };

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
bool encode_tmap(CUtensorMap *out, void *base, int rank, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1, uint64_t stride2,
                 uint32_t b0, uint32_t b1, uint32_t b2)
{
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                           const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Fn fn = [] {
        void *p = nullptr; cudaDriverEntryPointQueryResult qr;
        return cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess ? reinterpret_cast<Fn>(p) : nullptr;
    }();
    if (!fn) return false;
    const cuuint64_t dims[3] = { d0, d1, d2 }, strides[2] = { stride1, stride2 };
    const cuuint32_t box[3] = { b0, b1, b2 }, es[3] = { 1, 1, 1 };
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace

struct b200enc_batch {
    int device = 0, cap = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;    // stream2: entropy coding runs beside the deblocking wavefront
    cudaStream_t stream_hi = nullptr;                   // high-priority twin of `stream`: batches made only of IDR frames (long wavefront)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_wave0 = nullptr, ev_wave1 = nullptr;  // hand-over to / from the high-priority stream the wavefront kernels run on
    Sess *h_sess = nullptr, *d_sess = nullptr;
    WaveCtl *d_ctl = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0; int last_launches = 0;
    bool profiling = false;
    std::vector<KernelTime> ktimes; int n_ktimes = 0;
    std::vector<int> last_rcs;                          // per-session outcome of the last b200enc_batch_encode
    double t_pre = 0, t_launch = 0, t_sync = 0, t_dev = 0; int t_steps = 0;   // B200ENC_TRACE=1: host-clock split of launch_step, printed when the batch closes
};

struct b200enc_session {
    b200enc_config cfg;
    Geom g, g_key;                               // slice layout of P pictures / of key pictures (b200enc_config.key_slices); everything else is equal
    uint32_t rbsp_words_per_slice_key = 0;
    int device = -1; double load = 0;
    uint8_t *d_pool = nullptr; size_t pool_bytes = 0;
    // device buffers (sub-allocated from d_pool)
    uint8_t *input = nullptr, *src[3], *srcB[3], *bufA[3], *bufB[3], *rec_pre[3];     // src / srcB: the source planes alternate, so the previous picture's stay (background detection)
    bool src_is_A = true, have_psrc = false;
    uint8_t *srcL1, *srcL2, *refL1, *refL2;      // padded pyramid planes (allocation bases)
    uint8_t *rpl, *rpc[2];                       // padded reference planes: G,b,h,j contiguous; Cb, Cr
    void *tmaps;                                 // CUtensorMap[3] in HBM
    MbInfo *mbi; MbCoef *coef; uint4 *dbk_bs; int16_t *me2, *me1, *me0; int32_t *inter_cost, *skip_run;
    uint32_t *mb_bits, *mb_off, *mb_slot, *rbsp, *slice_bits; uint8_t *hdr; int hdr_len = 0; int *row_prog;
    uint2 *dbk_ll; uint32_t dbk_seq = 0;          // deblocking hand-over messages between MB rows, launch counter
    MbSide *side; uint16_t *bins, *bins_mb, *bin_lane_cnt, *bins_hdr; uint32_t *slice_nbins;   // CABAC (profile main / high)
    uint32_t rbsp_words_per_slice = 0;
    std::vector<uint8_t> param_sets;
    // pinned, device-mapped output: [0..cap) bitstream, then one uint32 size
    uint8_t *h_out = nullptr, *d_out = nullptr; uint32_t out_cap = 0;
    bool cur_is_A = true;
    uint32_t frame_index = 0; int frames_since_idr = 0, frame_num = 0, idr_pic_id = 0; bool force_idr = true, have_ref = false;
    int last_qp = 26, last_type = 1;
    b200rc::RateCtl rc; b200rc::Decision rc_dec;
    b200enc_batch *own = nullptr;
    bool in_scheduler = false;
    // upload path of b200enc_encode: the caller's thread copies its frame (through the pinned staging buffer when the caller's
    // memory is pageable, as every reference caller's is) on the session's own stream; the batch step waits for ev_up
    uint8_t *h_stage = nullptr; cudaStream_t up_stream = nullptr; cudaEvent_t ev_up = nullptr;
    uint32_t retries = 0;                        // pictures coded twice by the rate control (hard cap)
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int cabac_slab_kb()
{
    static const int kb = [] { const char *e = getenv("B200ENC_CABAC_SMEM_KB"); return std::min(std::max(e ? atoi(e) : B200_CABAC_SMEM_KB, 0), 180); }();
    return kb;
}

bool wave_prio()
{
    static const bool on = [] { const char *e = getenv("B200ENC_WAVE_PRIO"); return e ? atoi(e) != 0 : false; }();
    return on;
}

int dbk_resident_pct()
{
    // measured on 96 x 1080p sessions in three batches (frames/s): 100 % 11 320, 50 % 11 640, 40 % 12 250, 30 % 12 640, 25 % 12 680, 20 % 12 120, 15 % 12 220
    static const int pct = [] { const char *e = getenv("B200ENC_DBK_RESIDENT"); return std::min(std::max(e ? atoi(e) : 30, 5), 100); }();
    return pct;
}

// CUDA loads a kernel's code on its first launch (lazy module loading): 38 ms for k_me_fine, measured as the first P picture of a process
// (tools/frame_kernel_times.py) -- a late frame for every session that happens to be in that step. Asking for the attributes loads the code.
__global__ void k_reset(const Sess *ss, Geom g, int nsess, WaveCtl *ctl);
void preload_kernels()
{
    cudaFuncAttributes a;
#define PRELOAD(k) cudaFuncGetAttributes(&a, k)
    PRELOAD(k_reset); PRELOAD(k_ingest_rgba); PRELOAD(k_ingest_planar); PRELOAD(k_refplanes); PRELOAD(k_refchroma); PRELOAD(k_downsample);
    PRELOAD((k_me_coarse<4, 8, true>)); PRELOAD((k_me_coarse<4, 8, false>)); PRELOAD((k_me_coarse<8, 8, true>)); PRELOAD((k_me_coarse<8, 8, false>));
    PRELOAD((k_me_coarse<16, 4, true>)); PRELOAD((k_me_coarse<16, 4, false>));
    PRELOAD(k_me_fine); PRELOAD(k_scene_change); PRELOAD(k_inter_t8); PRELOAD(k_intra_wave); PRELOAD(k_pskip_scan); PRELOAD(k_deblock_bs); PRELOAD(k_deblock_wave);
    PRELOAD(k_cabac_side); PRELOAD(k_cabac_hdr); PRELOAD(k_cabac_bins); PRELOAD(k_cabac_scan); PRELOAD(k_cabac_place_hdr); PRELOAD(k_cabac_compact); PRELOAD(k_cabac_code);
    PRELOAD(k_cavlc_hdr); PRELOAD(k_cavlc_mb); PRELOAD(k_slice_scan); PRELOAD(k_slice_copy); PRELOAD(k_nal_pack);
#undef PRELOAD
    cudaGetLastError();
}

int batch_init(b200enc_batch *b, int device, int cap)
{
    b->device = device; b->cap = cap;
    CU_TRY(cudaSetDevice(device), return B200ENC_ENODEV);
    preload_kernels();
    // function attributes are per device: every batch context sets the coder's dynamic shared-memory limit on its own device
    if (cabac_slab_kb() > 0) CU_TRY(cudaFuncSetAttribute(k_cabac_code, cudaFuncAttributeMaxDynamicSharedMemorySize, cabac_slab_kb() * 1024), return B200ENC_ENODEV);
    CU_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking), return B200ENC_ENODEV);
    CU_TRY(cudaStreamCreateWithFlags(&b->stream2, cudaStreamNonBlocking), return B200ENC_ENODEV);
    { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);
      CU_TRY(cudaStreamCreateWithPriority(&b->stream_hi, cudaStreamNonBlocking, hi), return B200ENC_ENODEV); }
    CU_TRY(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming), return B200ENC_ENODEV);
    CU_TRY(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming), return B200ENC_ENODEV);
    CU_TRY(cudaEventCreateWithFlags(&b->ev_wave0, cudaEventDisableTiming), return B200ENC_ENODEV);
    CU_TRY(cudaEventCreateWithFlags(&b->ev_wave1, cudaEventDisableTiming), return B200ENC_ENODEV);
    CU_TRY(cudaHostAlloc(&b->h_sess, sizeof(Sess) * cap, cudaHostAllocDefault), return B200ENC_ENOMEM);
    CU_TRY(cudaMalloc(&b->d_sess, sizeof(Sess) * cap), return B200ENC_ENOMEM);
    CU_TRY(cudaMalloc(&b->d_ctl, sizeof(WaveCtl)), return B200ENC_ENOMEM);
    CU_TRY(cudaEventCreate(&b->ev0), return B200ENC_ENODEV);
    CU_TRY(cudaEventCreate(&b->ev1), return B200ENC_ENODEV);
    return B200ENC_OK;
}
void batch_free(b200enc_batch *b)
{
    if (!b) return;
    if (b->t_steps) fprintf(stderr, "[b200enc trace] batch %p: %d steps, per step: descriptors+upload %.3f ms, launches %.3f ms, wait %.3f ms (host clock); device ev0..ev1 %.3f ms\n",
                            (void *)b, b->t_steps, b->t_pre / b->t_steps, b->t_launch / b->t_steps, b->t_sync / b->t_steps, b->t_dev / b->t_steps);
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (auto &k : b->ktimes) { cudaEventDestroy(k.ev0); cudaEventDestroy(k.ev1); }
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->d_ctl) cudaFree(b->d_ctl);
    if (b->d_sess) cudaFree(b->d_sess);
    if (b->h_sess) cudaFreeHost(b->h_sess);
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    if (b->ev_wave0) cudaEventDestroy(b->ev_wave0);
    if (b->ev_wave1) cudaEventDestroy(b->ev_wave1);
    if (b->stream2) cudaStreamDestroy(b->stream2);
    if (b->stream_hi) { cudaStreamSynchronize(b->stream_hi); cudaStreamDestroy(b->stream_hi); }
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

__global__ void k_reset(const Sess *ss, Geom g, int nsess, WaveCtl *ctl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { ctl->ticket_intra = 0; ctl->ticket_dbk = 0; ctl->error = 0; }
    if (i < nsess * g.mbh) { const Sess &s = ss[i / g.mbh]; s.row_prog_intra[i % g.mbh] = 0; s.row_prog_dbk[i % g.mbh] = 0; }
}

struct Prof {
    b200enc_batch *b;
    cudaStream_t mainst, cur;
    void begin(const char *name, cudaStream_t st)
    {
        if (!b->profiling) return;
        if (b->n_ktimes == (int)b->ktimes.size()) { KernelTime k; k.name = name; cudaEventCreate(&k.ev0); cudaEventCreate(&k.ev1); b->ktimes.push_back(k); }
        b->ktimes[b->n_ktimes].name = name; cur = st;
        cudaEventRecord(b->ktimes[b->n_ktimes].ev0, st);
    }
    void begin(const char *name) { begin(name, mainst); }
    void end() { if (!b->profiling) return; cudaEventRecord(b->ktimes[b->n_ktimes].ev1, cur); b->n_ktimes++; }
};

bool same_shape(const b200enc_session *a, const b200enc_session *c)
{
    return a->cfg.width == c->cfg.width && a->cfg.height == c->cfg.height && a->cfg.num_slices == c->cfg.num_slices && a->cfg.key_slices == c->cfg.key_slices &&
           a->cfg.search_range == c->cfg.search_range && a->cfg.input_format == c->cfg.input_format && a->device == c->device &&
           (a->cfg.profile != 0) == (c->cfg.profile != 0);
}

// where the frame of a step comes from
enum { IN_HOST = 0,       // frames[i] is host memory: copied here (through the session's pinned staging buffer when it is pageable)
       IN_DEVICE = 1,     // frames[i] is device memory of the batch's GPU (capture that already lives in HBM)
       IN_UPLOADED = 2,   // the caller's thread already queued the copy on the session's upload stream (b200enc_encode): wait for ev_up
       IN_RESIDENT = 3 }; // the frame is still in the session's input buffer (second attempt of the rate control)
#define SC_MIN_DISTANCE 10   /* a P picture is only promoted to a scene-change IDR when at least this many pictures passed since the last IDR */

// Staging copy pageable -> pinned with non-temporal stores: the destination is only ever read by the DMA engine, so it should neither be
// read for ownership nor displace the callers' working sets from the cache (3 MB per 1080p frame, a hundred sessions per GPU).
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2"))) void stream_copy_avx2(uint8_t *dst, const uint8_t *src, size_t n)
{
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; i++; }
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64)), d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
        _mm256_stream_si256((__m256i *)(dst + i), a); _mm256_stream_si256((__m256i *)(dst + i + 32), b);
        _mm256_stream_si256((__m256i *)(dst + i + 64), c); _mm256_stream_si256((__m256i *)(dst + i + 96), d);
    }
    for (; i < n; i++) dst[i] = src[i];
    _mm_sfence();
}
#endif
void stage_copy(uint8_t *dst, const uint8_t *src, size_t n)
{
#if defined(__x86_64__)
    static const bool nt = [] { const char *e = getenv("B200ENC_STAGE_NT"); return (e ? atoi(e) != 0 : true) && __builtin_cpu_supports("avx2"); }();
    if (nt) { stream_copy_avx2(dst, src, n); return; }
#endif
    memcpy(dst, src, n);
}

bool host_ptr_is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
// H2D copy of one frame on `st`. Caller-owned pageable memory (what the reference's callers hand to EncodeOneFrame,
// video_codec/VideoCodecApi.h:57-58; zero-copy into openh264 at VideoEncoderOpenH264.cpp:354-365) goes through the session's pinned
// staging buffer first -- the copy the Netint sibling also makes (VideoEncoderNetint.cpp:503-505) -- so the transfer itself is an
// asynchronous DMA and concurrent sessions do not serialise on the driver's internal bounce buffer.
int upload_frame(b200enc_session *s, const uint8_t *frame, cudaStream_t st)
{
    const size_t bytes = b200enc_frame_bytes(s);
    static const bool stage = [] { const char *e = getenv("B200ENC_STAGE_PAGEABLE"); return e ? atoi(e) != 0 : true; }();   // 0: experiments only
    static const size_t chunk = [] { const char *e = getenv("B200ENC_STAGE_CHUNK_KB"); return (size_t)std::max(64, e ? atoi(e) : 1 << 20) * 1024; }();
    if (stage && !host_ptr_is_pinned(frame)) {
        if (!s->h_stage) CU_TRY(cudaHostAlloc(&s->h_stage, bytes, cudaHostAllocDefault), return B200ENC_ENOMEM);
        // (B200ENC_STAGE_CHUNK_KB splits the copy so that the DMA of a chunk runs while the CPU stages the next one; measured on a 16-core host with
        // 128 sessions it loses to one copy per frame -- the extra driver calls of 128 threads cost more than the overlap gains)
        for (size_t off = 0; off < bytes; off += chunk) {
            const size_t n = std::min(chunk, bytes - off);
            stage_copy(s->h_stage + off, frame + off, n);
            CU_TRY(cudaMemcpyAsync(s->input + off, s->h_stage + off, n, cudaMemcpyHostToDevice, st), return B200ENC_ECUDA);
        }
        return B200ENC_OK;
    }
    CU_TRY(cudaMemcpyAsync(s->input, frame, bytes, cudaMemcpyHostToDevice, st), return B200ENC_ECUDA);
    return B200ENC_OK;
}
bool next_is_idr(const b200enc_session *s) { return s->force_idr || !s->have_ref || s->frames_since_idr >= s->cfg.gop; }

// One chain of kernel launches over n sessions: descriptors, launches, one stream synchronisation. kinds[i] 1 = IDR; qps[i] the picture QP.
// Does not touch the sessions' stream state (frame_num, reference roles, rate control): the caller commits or repeats the step.
int launch_step(b200enc_batch *b, b200enc_session *const *ss, int n, const uint8_t *const *frames, int mode, const int *kinds, const int *qps)
{
    bool all_key = true;
    for (int i = 0; i < n; i++) all_key &= kinds[i] != 0;
    const Geom g = all_key ? ss[0]->g_key : ss[0]->g;          // key pictures travel in steps of their own (encode_impl) and take their own slice layout
    static const bool trace = [] { const char *e = getenv("B200ENC_TRACE"); return e && atoi(e) != 0; }();
    const auto tr0 = std::chrono::steady_clock::now();
    bool any_p = false;
    for (int i = 0; i < n; i++) any_p |= !kinds[i];
    cudaStream_t st = any_p ? b->stream : b->stream_hi;      // a batch of key frames only gets the high-priority stream
    for (int i = 0; i < n; i++) {
        b200enc_session *s = ss[i];
        const bool idr = kinds[i] != 0;
        uint8_t **cur = s->cur_is_A ? s->bufA : s->bufB, **ref = s->cur_is_A ? s->bufB : s->bufA;
        Sess &d = b->h_sess[i];
        d.input = s->input;
        if (mode == IN_DEVICE) d.input = frames[i];
        else if (mode == IN_HOST) { const int rc = upload_frame(s, frames[i], st); if (rc != B200ENC_OK) return rc; }
        else if (mode == IN_UPLOADED) CU_TRY(cudaStreamWaitEvent(st, s->ev_up, 0), return B200ENC_ECUDA);
        for (int c = 0; c < 3; c++) {
            d.src[c] = s->src_is_A ? s->src[c] : s->srcB[c]; d.src_prev[c] = s->have_psrc ? (s->src_is_A ? s->srcB[c] : s->src[c]) : nullptr;
            d.rec[c] = cur[c]; d.ref[c] = ref[c];
        }
        d.src_tmap = s->src_is_A ? 0 : 3 * (int)sizeof(CUtensorMap);
        d.bgd = s->cfg.background_detection; d.no_p8x8 = s->cfg.complexity <= 1; d.no_i4x4 = s->cfg.complexity == 0;
        {   // padded planes: hand the kernels the address of the interior sample (0,0)
            const size_t o1 = (size_t)g.p1 * g.s1 + g.p1, o2 = (size_t)g.p2 * g.s2 + g.p2, ol = (size_t)g.lp * g.ls + g.lp, oc = (size_t)g.cp * g.cs + g.cp;
            const size_t plane = (size_t)g.ls * (g.hc + 2 * g.lp);
            d.srcL1 = s->srcL1 + o1; d.refL1 = s->refL1 + o1; d.srcL2 = s->srcL2 + o2; d.refL2 = s->refL2 + o2;
            for (int k = 0; k < 4; k++) d.rpl[k] = s->rpl + k * plane + ol;
            d.rpc[0] = s->rpc[0] + oc; d.rpc[1] = s->rpc[1] + oc; d.tmaps = s->tmaps;
        }
        d.mbi = s->mbi; d.coef = s->coef; d.dbk_bs = s->dbk_bs; d.me2 = s->me2; d.me1 = s->me1; d.me0 = s->me0; d.inter_cost = s->inter_cost;
        d.skip_run = s->skip_run; d.mb_bits = s->mb_bits; d.mb_off = s->mb_off; d.mb_slot = s->mb_slot; d.rbsp = s->rbsp; d.slice_bits = s->slice_bits;
        d.side = s->side; d.bins = s->bins; d.bins_mb = s->bins_mb; d.bin_lane_cnt = s->bin_lane_cnt; d.bins_hdr = s->bins_hdr; d.slice_nbins = s->slice_nbins;
        d.out = s->d_out; d.out_size = reinterpret_cast<uint32_t *>(s->d_out + s->out_cap); d.hdr = s->hdr; d.hdr_len = s->hdr_len;
        d.row_prog_intra = s->row_prog; d.row_prog_dbk = s->row_prog + g.mbh;
        d.dbk_ll = s->dbk_ll; d.dbk_seq = ++s->dbk_seq;
        if (!d.dbk_seq) d.dbk_seq = ++s->dbk_seq;          // 0 is what never-written messages hold
        d.qp = qps[i]; d.is_idr = idr; d.frame_num = idr ? 0 : s->frame_num; d.idr_pic_id = s->idr_pic_id; d.input_format = s->cfg.input_format;
        d.scene_change = s->cfg.scene_change && !idr && s->frames_since_idr >= SC_MIN_DISTANCE;
        d.t8x8 = s->cfg.profile == 2; d.dump = s->cfg.debug & 1;
        d.rbsp_words_per_slice = all_key ? s->rbsp_words_per_slice_key : s->rbsp_words_per_slice; d.out_cap = s->out_cap;
    }
    CU_TRY(cudaMemcpyAsync(b->d_sess, b->h_sess, sizeof(Sess) * n, cudaMemcpyHostToDevice, st), return B200ENC_ECUDA);
    const int nmb = g.mbw * g.mbh;
    int launches = 0; b->n_ktimes = 0; Prof pf{ b, st, st };
    const auto tr1 = std::chrono::steady_clock::now();
    cudaEventRecord(b->ev0, st);
    pf.begin("k_reset"); k_reset<<<(n * g.mbh + 255) / 256, 256, 0, st>>>(b->d_sess, g, n, b->d_ctl); pf.end(); launches++;
    if (ss[0]->cfg.input_format == B200ENC_FMT_RGBA) {
        pf.begin("k_ingest_rgba"); k_ingest_rgba<<<dim3(((g.wc / 8) * (g.hc / 2) + 255) / 256, 1, n), 256, 0, st>>>(b->d_sess, g); pf.end();
    } else {
        pf.begin("k_ingest_planar"); k_ingest_planar<<<dim3((INGEST_UNITS(g.wc, g.hc) + 255) / 256, 1, n), 256, 0, st>>>(b->d_sess, g); pf.end();
    }
    launches++;
    if (any_p) {
        pf.begin("k_refplanes"); k_refplanes<<<dim3((g.ls + RP_TW - 1) / RP_TW, (g.hc + 2 * g.lp + RP_TH - 1) / RP_TH, n), 256, 0, st>>>(b->d_sess, g); pf.end();
        pf.begin("k_refchroma"); k_refchroma<<<dim3(((g.cs / 16) * (g.hc / 2 + 2 * g.cp) + 255) / 256, 2, n), 256, 0, st>>>(b->d_sess, g); pf.end();
        pf.begin("k_downsample0"); k_downsample<<<dim3((DOWNSAMPLE_UNITS(g.wc / 2, g.hc / 2, g.p1) + 255) / 256, 2, n), 256, 0, st>>>(b->d_sess, g, 0); pf.end();
        pf.begin("k_downsample1"); k_downsample<<<dim3((DOWNSAMPLE_UNITS(g.wc / 4, g.hc / 4, g.p2) + 255) / 256, 2, n), 256, 0, st>>>(b->d_sess, g, 1); pf.end();
        pf.begin("k_me_coarse");
        {
            const dim3 g8((nmb + 7) / 8, 1, n), g4((nmb + 3) / 4, 1, n);
            if (g.search_range == 16) k_me_coarse<4, 8, true><<<g8, 256, 0, st>>>(b->d_sess, g);
            else if (g.search_range < 16) k_me_coarse<4, 8, false><<<g8, 256, 0, st>>>(b->d_sess, g);
            else if (g.search_range == 32) k_me_coarse<8, 8, true><<<g8, 256, 0, st>>>(b->d_sess, g);
            else if (g.search_range < 32) k_me_coarse<8, 8, false><<<g8, 256, 0, st>>>(b->d_sess, g);
            else if (g.search_range == 64) k_me_coarse<16, 4, true><<<g4, 128, 0, st>>>(b->d_sess, g);
            else k_me_coarse<16, 4, false><<<g4, 128, 0, st>>>(b->d_sess, g);
        }
        pf.end();
        pf.begin("k_me_fine"); k_me_fine<<<dim3((nmb + ME_WARPS - 1) / ME_WARPS, 1, n), ME_WARPS * 32, 0, st>>>(b->d_sess, g, b->d_ctl); pf.end();
        pf.begin("k_scene_change"); k_scene_change<<<n, 256, 0, st>>>(b->d_sess, g); pf.end();
        launches += 7;
        bool any_t8 = false;
        for (int i = 0; i < n; i++) any_t8 |= ss[i]->cfg.profile == 2;
        if (any_t8) { pf.begin("k_inter_t8"); k_inter_t8<<<dim3((nmb + T8_WARPS - 1) / T8_WARPS, 1, n), T8_WARPS * 32, 0, st>>>(b->d_sess, g); pf.end(); launches++; }
    }
    const int wave_ctas = (n * g.mbh + WAVE_WARPS - 1) / WAVE_WARPS;
    // The wavefront kernels are dependent chains per MB row (latency-bound, few warps). On the batch's high-priority stream their CTAs are
    // placed ahead of the queued CTAs of other batches' throughput kernels (k_me_fine alone is 32 640 CTAs per 32 sessions) instead of
    // behind them, so the chains of one batch run underneath the motion search of the next.
    cudaStream_t sw = wave_prio() ? b->stream_hi : st;
    if (sw != st) { cudaEventRecord(b->ev_wave0, st); cudaStreamWaitEvent(sw, b->ev_wave0, 0); }
    // Resident deblocking warps. A 2-MB-lag wavefront over mbh rows of mbw MBs keeps mbh * mbw / (mbw + 2 mbh) rows busy on average (47 % at
    // 1080p); warps beyond that only hold registers while they wait for their turn and push the other batches' motion search off the SMs
    // (one warp per row of 32 x 1080p sessions is 69 % of the GPU's register file for k_deblock_wave). Small batches keep every row
    // resident (floor: one CTA per SM), because there the wavefront's own latency is all that matters.
    auto resident_ctas = [&](int pct) { return std::max(1, std::min(wave_ctas, std::max(148, (n * ((g.mbh * pct + 99) / 100) + WAVE_WARPS - 1) / WAVE_WARPS))); };
    pf.begin("k_intra_wave", sw); k_intra_wave<<<wave_ctas, WAVE_WARPS * 32, 0, sw>>>(b->d_sess, g, n, b->d_ctl); pf.end(); launches++;
    for (int i = 0; i < n; i++) if (ss[i]->cfg.debug & 1) {
        uint8_t **cur = ss[i]->cur_is_A ? ss[i]->bufA : ss[i]->bufB;
        for (int c = 0; c < 3; c++) cudaMemcpyAsync(ss[i]->rec_pre[c], cur[c], (size_t)g.wc * g.hc / (c ? 4 : 1), cudaMemcpyDeviceToDevice, sw);
    }
    pf.begin("k_pskip_scan", sw); k_pskip_scan<<<dim3(g.num_slices, 1, n), 256, 0, sw>>>(b->d_sess, g); pf.end(); launches++;
    // fork: the entropy coder only needs MbInfo / MbCoef / skip runs, the deblocking wavefront only reconstruction + MbInfo;
    // the latency-bound wavefront and the issue-bound CAVLC chain overlap on two streams and join before the read-back
    cudaStream_t s2 = b->stream2;
    cudaEventRecord(b->ev_fork, sw); cudaStreamWaitEvent(s2, b->ev_fork, 0);
    pf.begin("k_deblock_bs", sw); k_deblock_bs<<<dim3((nmb + 127) / 128, 1, n), 128, 0, sw>>>(b->d_sess, g); pf.end(); launches++;
    const int dbk_ctas = resident_ctas(dbk_resident_pct());
    pf.begin("k_deblock_wave", sw); k_deblock_wave<<<dbk_ctas, WAVE_WARPS * 32, 0, sw>>>(b->d_sess, g, n, b->d_ctl); pf.end(); launches++;
    if (sw != st) { cudaEventRecord(b->ev_wave1, sw); cudaStreamWaitEvent(st, b->ev_wave1, 0); }
    if (ss[0]->cfg.profile) {
        // CABAC: side records, entry counts, offsets, bin lists (all parallel over MBs), then one warp per slice runs the coder
        const dim3 gb((nmb + CABAC_WARPS - 1) / CABAC_WARPS, 1, n);
        static const int xskip = [] { const char *e = getenv("B200ENC_X_SKIP"); return e ? atoi(e) : 0; }();    // timing experiments only (output invalid)
        pf.begin("k_cabac_side", s2); k_cabac_side<<<dim3((nmb + 255) / 256, 1, n), 256, 0, s2>>>(b->d_sess, g); pf.end();
        pf.begin("k_cabac_hdr", s2); k_cabac_hdr<<<dim3((nmb + 255) / 256, 1, n), 256, 0, s2>>>(b->d_sess, g); pf.end();
        pf.begin("k_cabac_bins", s2); if (!(xskip & 1)) k_cabac_bins<<<gb, CABAC_WARPS * 32, 0, s2>>>(b->d_sess, g); pf.end();
        pf.begin("k_cabac_scan", s2); k_cabac_scan<<<dim3(g.num_slices, 1, n), 256, 0, s2>>>(b->d_sess, g); pf.end();
        pf.begin("k_cabac_place_hdr", s2); if (!(xskip & 2)) k_cabac_place_hdr<<<dim3((nmb + 255) / 256, 1, n), 256, 0, s2>>>(b->d_sess, g); pf.end();
        pf.begin("k_cabac_compact", s2); if (!(xskip & 2)) k_cabac_compact<<<gb, CABAC_WARPS * 32, 0, s2>>>(b->d_sess, g); pf.end();
        // the coder threads are latency chains: every slot they lose to a co-resident throughput kernel's warps stretches the frame. Asking for
        // a slab of dynamic shared memory they do not use keeps the shared-memory-hungry kernels of the other batches off their SMs
        // (paced sessions: small batches, tail latency counts). Big batches are throughput work: there the slab only takes shared memory
        // from the other batches' motion search (96 x 1080p Main in batches of 32: 8 550 frames/s with it, 9 380 without)
        static const int slab_max_n = [] { const char *e = getenv("B200ENC_CABAC_SLAB_MAXN"); return e ? atoi(e) : 16; }();
        const int hog_kb = n <= slab_max_n ? cabac_slab_kb() : 0;
        pf.begin("k_cabac_code", s2); if (!(xskip & 4)) k_cabac_code<<<dim3(g.num_slices, 1, n), 96, (size_t)hog_kb * 1024, s2>>>(b->d_sess, g, b->d_ctl); pf.end();
        launches += 7;
    } else {
    pf.begin("k_cavlc_hdr", s2); k_cavlc_hdr<<<dim3((nmb + 127) / 128, 1, n), 128, 0, s2>>>(b->d_sess, g); pf.end(); launches++;
    pf.begin("k_cavlc_mb", s2); k_cavlc_mb<<<dim3((nmb + CAVLC_WARPS - 1) / CAVLC_WARPS, 1, n), CAVLC_WARPS * 32, 0, s2>>>(b->d_sess, g); pf.end(); launches++;
    pf.begin("k_slice_scan", s2); k_slice_scan<<<dim3(g.num_slices, 1, n), 256, 0, s2>>>(b->d_sess, g); pf.end(); launches++;
    pf.begin("k_slice_copy", s2); k_slice_copy<<<dim3((nmb + CAVLC_WARPS - 1) / CAVLC_WARPS, 1, n), CAVLC_WARPS * 32, 0, s2>>>(b->d_sess, g); pf.end(); launches++;
    }
    pf.begin("k_nal_pack", s2); k_nal_pack<<<n, 1024, 0, s2>>>(b->d_sess, g); pf.end(); launches++;
    cudaEventRecord(b->ev_join, s2); cudaStreamWaitEvent(st, b->ev_join, 0);
    cudaEventRecord(b->ev1, st);
    WaveCtl ctl;
    const auto tr2 = std::chrono::steady_clock::now();
    CU_TRY(cudaMemcpyAsync(&ctl, b->d_ctl, sizeof ctl, cudaMemcpyDeviceToHost, st), return B200ENC_ECUDA);
    CU_TRY(cudaStreamSynchronize(st), return B200ENC_ECUDA);
    CU_TRY(cudaGetLastError(), return B200ENC_ECUDA);
    float ms = 0; cudaEventElapsedTime(&ms, b->ev0, b->ev1);
    if (trace) {
        const auto tr3 = std::chrono::steady_clock::now();
        auto d = [](auto a, auto c) { return std::chrono::duration<double, std::milli>(c - a).count(); };
        b->t_pre += d(tr0, tr1); b->t_launch += d(tr1, tr2); b->t_sync += d(tr2, tr3); b->t_dev += ms; b->t_steps++;
    }
    b->last_ms += ms; b->last_launches += launches;
    return ctl.error ? B200ENC_EWAVE : B200ENC_OK;
}

// launch_step over pictures of both kinds: key pictures and P pictures differ in their slice layout (b200enc_config.key_slices), so a mixed set
// runs as two steps -- key pictures first (the longer chain), each subset with its own geometry.
int launch_steps_by_kind(b200enc_batch *b, b200enc_session *const *ss, int n, const uint8_t *const *frames, int mode, const int *kinds, const int *qps)
{
    bool any_key = false, any_p = false;
    for (int i = 0; i < n; i++) { if (kinds[i]) any_key = true; else any_p = true; }
    if (!(any_key && any_p) || ss[0]->g_key.num_slices == ss[0]->g.num_slices) return launch_step(b, ss, n, frames, mode, kinds, qps);
    for (int pass = 1; pass >= 0; pass--) {
        std::vector<b200enc_session *> sub; std::vector<const uint8_t *> fr; std::vector<int> k, q;
        for (int i = 0; i < n; i++) if ((kinds[i] != 0) == (pass == 1)) { sub.push_back(ss[i]); fr.push_back(frames[i]); k.push_back(kinds[i]); q.push_back(qps[i]); }
        const int rc = launch_step(b, sub.data(), (int)sub.size(), fr.data(), mode, k.data(), q.data());
        if (rc != B200ENC_OK) return rc;
    }
    return B200ENC_OK;
}

bool rc_retry_enabled()
{
    static const bool on = [] { const char *e = getenv("B200ENC_RC_RETRY"); return e ? atoi(e) != 0 : true; }();
    return on;
}

// Advance n sessions of one GPU by one picture. Returns a whole-step failure (bad arguments, CUDA error, device watchdog) or B200ENC_OK /
// B200ENC_EOVERFLOW when only individual sessions failed; rcs[i] (optional) is the outcome of session i. A session whose picture was not
// delivered keeps its stream state and starts over with an IDR, so its decoder never predicts from a picture it did not get.
int encode_impl(b200enc_batch *b, b200enc_session *const *ss, int n, const uint8_t *const *frames, int mode,
                const uint8_t **bs, uint32_t *bs_size, b200enc_frame_info *infos, int *rcs)
{
    if (!b || !ss || n <= 0 || n > b->cap || !frames) return B200ENC_EINVAL;
    for (int i = 0; i < n; i++) {
        if (!ss[i] || (!frames[i] && mode != IN_UPLOADED) || ss[i]->device != b->device || !same_shape(ss[0], ss[i])) return B200ENC_EINVAL;
        for (int j = 0; j < i; j++) if (ss[j] == ss[i]) return B200ENC_EINVAL;
    }
    CU_TRY(cudaSetDevice(b->device), return B200ENC_ECUDA);
    std::vector<int> kinds(n), qps(n);
    for (int i = 0; i < n; i++) {
        b200enc_session *s = ss[i];
        kinds[i] = next_is_idr(s) ? 1 : 0;
        if (s->cfg.const_qp >= 0) qps[i] = s->cfg.const_qp;
        else { s->rc_dec = s->rc.pick(kinds[i]); qps[i] = s->rc_dec.qp; }
    }
    b->last_ms = 0; b->last_launches = 0;
    int rc = launch_steps_by_kind(b, ss, n, frames, mode, kinds.data(), qps.data());
    auto out_size = [](const b200enc_session *s) { return *reinterpret_cast<volatile uint32_t *>(s->h_out + s->out_cap); };
    auto promoted = [](const b200enc_session *s) { return *reinterpret_cast<volatile uint32_t *>(s->h_out + s->out_cap + 4) != 0; };
    if (rc == B200ENC_OK && rc_retry_enabled()) {
        // rate control, second attempt: pictures that came out above their hard cap (an IDR many times the budget; a P picture after a cut)
        // are coded once more with a coarser QP before anything is delivered or committed. The frames are still in the input buffers.
        std::vector<b200enc_session *> again; std::vector<int> ak, aq, at; std::vector<const uint8_t *> af;
        for (int i = 0; i < n; i++) {
            b200enc_session *s = ss[i];
            if (s->cfg.const_qp >= 0 || out_size(s) >= s->out_cap) continue;
            const int coded = kinds[i] || (promoted(s) ? 1 : 0);
            const int q2 = s->rc.second_attempt_qp(kinds[i], coded, s->rc_dec, 8.0 * out_size(s));
            if (q2 < 0) continue;
            { static const bool tr = [] { const char *e = getenv("B200ENC_TRACE"); return e && atoi(e) != 0; }(); static std::atomic<int> shown{ 0 };
              if (tr && shown.fetch_add(1) < 40) fprintf(stderr, "[b200enc trace] coded twice: frame %u planned %d coded %d qp %d -> %d, %u bytes, budget %.0f cap %.0f bytes\n",
                                                         s->frame_index, kinds[i], coded, qps[i], q2, out_size(s), s->rc_dec.budget / 8, s->rc_dec.hard_cap / 8); }
            again.push_back(s); ak.push_back(coded); aq.push_back(q2); at.push_back(i); af.push_back(frames[i]);
        }
        if (!again.empty()) {
            rc = launch_steps_by_kind(b, again.data(), (int)again.size(), af.data(), mode == IN_DEVICE ? IN_DEVICE : IN_RESIDENT, ak.data(), aq.data());
            for (size_t k = 0; k < again.size(); k++) { kinds[at[k]] = ak[k]; qps[at[k]] = aq[k]; again[k]->retries++; }
        }
    }
    if (rc != B200ENC_OK) {
        for (int i = 0; i < n; i++) { ss[i]->force_idr = true; if (rcs) rcs[i] = rc; }
        return rc;
    }
    int worst = B200ENC_OK;
    for (int i = 0; i < n; i++) {
        b200enc_session *s = ss[i];
        const uint32_t size = out_size(s);
        const bool idr = kinds[i] || promoted(s);      // scene change: the device coded this P picture as an IDR (k_scene_change)
        if (infos) { infos[i].frame_type = idr ? 1 : 0; infos[i].qp = qps[i]; infos[i].size_bytes = size; infos[i].frame_index = s->frame_index; }
        if (size >= s->out_cap) {
            // k_nal_pack clamped the access unit: nothing is delivered, the stream state stays where it was, the next picture is an IDR
            s->force_idr = true; worst = B200ENC_EOVERFLOW;
            if (rcs) rcs[i] = B200ENC_EOVERFLOW;
            if (bs) bs[i] = nullptr;
            if (bs_size) bs_size[i] = 0;
            continue;
        }
        if (rcs) rcs[i] = B200ENC_OK;
        s->last_qp = qps[i]; s->last_type = idr ? 1 : 0;
        if (s->cfg.const_qp < 0) s->rc.update(s->last_type, s->last_qp, 8.0 * size);
        if (bs) bs[i] = s->h_out;
        if (bs_size) bs_size[i] = size;
        if (idr) { s->frame_num = 0; s->frames_since_idr = 0; s->idr_pic_id = (s->idr_pic_id + 1) & 1; }
        s->frame_num = (s->frame_num + 1) & 255; s->frames_since_idr++; s->frame_index++;
        s->force_idr = false; s->have_ref = true; s->cur_is_A = !s->cur_is_A;
        s->src_is_A = !s->src_is_A; s->have_psrc = true;        // this picture's source planes are the next picture's "previous source"
    }
    return worst;
}


// b200enc_encode's upload: runs on the CALLER's thread, so N sessions copy (and stage pageable frames) in parallel
int session_upload(b200enc_session *s, const uint8_t *frame)
{
    CU_TRY(cudaSetDevice(s->device), return B200ENC_ECUDA);
    const int rc = upload_frame(s, frame, s->up_stream);
    if (rc != B200ENC_OK) return rc;
    CU_TRY(cudaEventRecord(s->ev_up, s->up_stream), return B200ENC_ECUDA);
    return B200ENC_OK;
}

// ---- per-GPU session scheduler (auto_batch): the reference runs one caller thread per session, each blocked in its own
// EncodeOneFrame (video_codec/VideoEncoderOpenH264.cpp:304-352, iMultipleThreadIdc = 1 at :294). Here those concurrent calls
// rendezvous in a per-GPU worker that advances all waiting sessions with ONE batch step; while a step runs, the next
// callers queue up, so batches form by themselves under load and a lone caller only pays the short window. ----
struct SchedRequest {
    b200enc_session *s; const uint8_t *frame; const uint8_t *bs = nullptr; uint32_t size = 0; b200enc_frame_info info{};
    int rc = B200ENC_OK; bool done = false;
    std::condition_variable cv;      // one per request: a finished batch wakes exactly its own callers, not every blocked session thread
};
struct DeviceScheduler {
    static constexpr int MAX_WORKERS = 8;
    int workers = 8;                       // batch contexts (own streams): up to that many batches in flight, so the latency-bound wavefront kernels of one
                                           // batch run underneath the throughput kernels of the others while further sessions upload their frames
    int batch_div = 4, batch_max = 32;     // a batch aims at registered / batch_div sessions, at most batch_max (measured: 256 sessions through the plugin
                                           // boundary, 8 x 32: 13 370 frames/s; 4 x 64: 11 370)
    int device = -1, registered = 0, inflight = 0, active = 0;       // active: batches on the GPU right now
    std::mutex mu; std::condition_variable cv_submit;
    std::vector<SchedRequest *> pending;
    std::thread worker[MAX_WORKERS]; bool stop = false;
    b200enc_batch *ctx[MAX_WORKERS] = {}; int window_us = 300, fill_us = 4000;
    std::atomic<uint64_t> batches{ 0 }, frames{ 0 };

    void run(int w)
    {
        cudaSetDevice(device);
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_submit.wait(lk, [&] { return stop || !pending.empty(); });
            if (stop && pending.empty()) return;
            // Rendezvous. A batch should carry about a quarter of the GPU's sessions (four batches in flight fill the machine and hide each
            // other's dependent chains). While fewer than two batches are on the GPU, waiting is lost time: leave after the short window (a
            // lone caller pays at most that). With two or more in flight the GPU is busy anyway, so wait for the fuller batch -- until enough
            // callers arrived, a running batch finished, or fill_us passed. Leaves at once when every session not being encoded is waiting.
            const auto t0 = std::chrono::steady_clock::now();
            // a batch aims at a quarter of the GPU's sessions, never fewer than 8 (a handful of sessions travels together), never more than batch_max
            const int target = std::max(1, std::min(std::min(ctx[w]->cap, batch_max), std::max((registered + batch_div - 1) / batch_div, std::min(registered, 8))));
            while (!stop && !pending.empty()) {
                const int want = std::min(target, std::max(1, registered - inflight));
                if ((int)pending.size() >= want) break;
                const auto limit = t0 + std::chrono::microseconds(active >= 2 ? fill_us : window_us);
                if (std::chrono::steady_clock::now() >= limit) break;
                cv_submit.wait_until(lk, limit);
            }
            if (pending.empty()) continue;                 // another worker took them
            // one batch = the requests that share the first request's shape AND frame kind (others wait for the next round or
            // the other worker): an IDR costs several times a P frame on the intra wavefront, so key frames travel in their own
            // batch and do not hold back the P frames of the sessions that happen to arrive with them
            std::vector<SchedRequest *> take, rest;
            const bool kind0 = next_is_idr(pending[0]->s);
            for (SchedRequest *r : pending)
                (same_shape(pending[0]->s, r->s) && next_is_idr(r->s) == kind0 && (int)take.size() < std::min(ctx[w]->cap, 2 * batch_max) ? take : rest).push_back(r);
            pending.swap(rest);
            const int n = (int)take.size();
            inflight += n; active++;
            lk.unlock();
            std::vector<b200enc_session *> ss(n); std::vector<const uint8_t *> fr(n), bs(n, nullptr); std::vector<uint32_t> sz(n, 0); std::vector<b200enc_frame_info> inf(n);
            std::vector<int> rcs(n, B200ENC_OK);
            for (int i = 0; i < n; i++) { ss[i] = take[i]->s; fr[i] = take[i]->frame; }
            // the callers' threads have already queued their uploads (scheduler_encode); every request gets its own outcome
            const int rc = encode_impl(ctx[w], ss.data(), n, fr.data(), IN_UPLOADED, bs.data(), sz.data(), inf.data(), rcs.data());
            batches++; frames += n;
            lk.lock();
            inflight -= n; active--;
            for (int i = 0; i < n; i++) {
                take[i]->bs = bs[i]; take[i]->size = sz[i]; take[i]->info = inf[i]; take[i]->done = true;
                take[i]->rc = rc != B200ENC_OK && rc != B200ENC_EOVERFLOW ? rc : rcs[i];
                take[i]->cv.notify_one();
            }
            cv_submit.notify_all();                        // workers holding out for a fuller batch re-evaluate (one batch fewer on the GPU)
        }
    }
};
std::mutex g_scheds_mu;
std::vector<std::unique_ptr<DeviceScheduler>> g_scheds;

DeviceScheduler *scheduler_for(int device)
{
    std::lock_guard<std::mutex> lk(g_scheds_mu);
    if ((int)g_scheds.size() <= device) g_scheds.resize(device + 1);
    if (!g_scheds[device]) {
        std::unique_ptr<DeviceScheduler> d(new DeviceScheduler());
        d->device = device;
        if (const char *e = getenv("B200ENC_BATCH_WINDOW_US")) d->window_us = std::max(0, atoi(e));
        if (const char *e = getenv("B200ENC_BATCH_FILL_US")) d->fill_us = std::max(0, atoi(e));
        if (const char *e = getenv("B200ENC_BATCH_WORKERS")) d->workers = std::min(std::max(1, atoi(e)), (int)DeviceScheduler::MAX_WORKERS);
        if (const char *e = getenv("B200ENC_BATCH_DIV")) d->batch_div = std::max(1, atoi(e));
        if (const char *e = getenv("B200ENC_BATCH_MAX")) d->batch_max = std::max(1, atoi(e));
        for (int w = 0; w < d->workers; w++) {
            d->ctx[w] = new (std::nothrow) b200enc_batch();
            if (!d->ctx[w] || batch_init(d->ctx[w], device, 512) != B200ENC_OK) return nullptr;
        }
        DeviceScheduler *raw = d.get();
        for (int w = 0; w < d->workers; w++) d->worker[w] = std::thread([raw, w] { raw->run(w); });
        g_scheds[device] = std::move(d);
    }
    return g_scheds[device].get();
}
void scheduler_register(b200enc_session *s, int delta)
{
    DeviceScheduler *d = scheduler_for(s->device);
    if (!d) return;
    std::lock_guard<std::mutex> lk(d->mu);
    d->registered += delta;
    s->in_scheduler = delta > 0;
    d->cv_submit.notify_all();
}
int scheduler_encode(b200enc_session *s, const uint8_t *frame, const uint8_t **bs, uint32_t *bs_size, b200enc_frame_info *info)
{
    DeviceScheduler *d = scheduler_for(s->device);
    if (!d) return B200ENC_ENODEV;
    SchedRequest req; req.s = s; req.frame = frame;
    {   // this caller's H2D copy starts now, on the session's own stream, while the scheduler is still collecting the batch
        const int rc = session_upload(s, frame);
        if (rc != B200ENC_OK) return rc;
    }
    std::unique_lock<std::mutex> lk(d->mu);
    d->pending.push_back(&req);
    d->cv_submit.notify_all();
    req.cv.wait(lk, [&] { return req.done; });
    if (bs) *bs = req.bs;
    if (bs_size) *bs_size = req.size;
    if (info) *info = req.info;
    return req.rc;
}
struct SchedulerShutdown {       // join the workers before the CUDA runtime is torn down at process exit
    ~SchedulerShutdown()
    {
        std::lock_guard<std::mutex> lk(g_scheds_mu);
        for (auto &d : g_scheds) if (d) {
            { std::lock_guard<std::mutex> l2(d->mu); d->stop = true; d->cv_submit.notify_all(); }
            for (auto &t : d->worker) if (t.joinable()) t.join();
        }
    }
} g_scheduler_shutdown;

} // namespace

extern "C" {

void b200enc_default_config(b200enc_config *c)
{
    if (!c) return;
    memset(c, 0, sizeof *c);
    // defaults of the reference wrapper: 720x1280, 30 fps, 5 Mbps, gop 30 (video_codec/VideoEncoderOpenH264.h:13-24)
    c->width = 720; c->height = 1280; c->fps = 30; c->bitrate = 5000000; c->gop = 30; c->const_qp = -1;
    c->num_slices = 0; c->search_range = 16; c->input_format = B200ENC_FMT_I420; c->device = -1; c->scene_change = 1;
    // rate-control bounds of openh264's GetDefaultParams (what the wrapper keeps, :230), max bitrate = target (:239-240), HIGH_COMPLEXITY (:289)
    c->max_bitrate = 0; c->min_qp = 0; c->max_qp = 51; c->background_detection = 0; c->complexity = 2; c->key_slices = 0;
}

int b200enc_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
int b200enc_last_cuda_error(void) { return g_last_cuda_error.load(); }
int b200enc_scheduler_stats(int device, uint64_t *batches, uint64_t *frames)
{
    std::lock_guard<std::mutex> lk(g_scheds_mu);
    if (device < 0 || device >= (int)g_scheds.size() || !g_scheds[device]) return B200ENC_EINVAL;
    if (batches) *batches = g_scheds[device]->batches.load();
    if (frames) *frames = g_scheds[device]->frames.load();
    return B200ENC_OK;
}
const char *b200enc_strerror(int code)
{
    switch (code) {
    case B200ENC_OK: return "ok";
    case B200ENC_EINVAL: return "invalid argument or unsupported configuration";
    case B200ENC_ENODEV: return "no usable CUDA device (this encoder has no CPU fallback)";
    case B200ENC_ENOMEM: return "out of memory";
    case B200ENC_ECUDA: return "CUDA error during encode";
    case B200ENC_ESIZE: return "input buffer smaller than one frame";
    case B200ENC_EOVERFLOW: return "bitstream larger than the output buffer";
    case B200ENC_EWAVE: return "device watchdog fired (wavefront wait or TMA transaction)";
    default: return "unknown error";
    }
}

int b200enc_create(const b200enc_config *cfg, b200enc_session **out)
{
    if (!cfg || !out) return B200ENC_EINVAL;
    *out = nullptr;
    b200enc_config c = *cfg;
    if (c.search_range <= 0) c.search_range = 16;
    if (c.fps <= 0) c.fps = 30;
    if (c.gop <= 0) c.gop = 30;
    if (c.width < 16 || c.width > 4096 || c.height < 16 || c.height > 4096 || (c.width & 1) || (c.height & 1)) return B200ENC_EINVAL;
    if (c.search_range % 4 || c.search_range > 64 || c.num_slices > B200_MAX_SLICES || c.const_qp > 51) return B200ENC_EINVAL;
    if (c.input_format < 0 || c.input_format > 2) return B200ENC_EINVAL;
    if (c.const_qp < 0 && c.bitrate <= 0) return B200ENC_EINVAL;
    if (c.profile < 0 || c.profile > 2) return B200ENC_EINVAL;
    if (c.min_qp < 0 || c.min_qp > 51 || c.max_qp < 0 || c.max_qp > 51 || (c.max_qp && c.max_qp < c.min_qp) || c.max_bitrate < 0) return B200ENC_EINVAL;
    if (c.complexity < 0 || c.complexity > 2) return B200ENC_EINVAL;
    // num_slices <= 0 = automatic: one slice with CAVLC (the wrapper's SM_SINGLE_SLICE, VideoEncoderOpenH264.cpp:247); with CABAC one slice per
    // ~17 MB rows (1080p: 4, 720p: 2, 2160p: 7), because the arithmetic coder is a serial chain per slice and a frame's latency is its longest slice
    // (+0.6..0.9 % bits on P pictures at 1080p, measured with the oracle)
    const bool auto_slices = c.num_slices < 1;
    if (c.num_slices < 1) c.num_slices = c.profile ? std::min(std::max(((c.height + 15) / 16 + 8) / 17, 1), 8) : 1;
    // key pictures: their own slice count (b200enc_config.key_slices); automatic = one slice per 4 MB rows for CABAC sessions with automatic slices
    if (c.key_slices < 0 || c.key_slices > B200_MAX_SLICES) return B200ENC_EINVAL;
    if (c.key_slices < 1) c.key_slices = auto_slices && c.profile ? std::min(std::max(((c.height + 15) / 16 + 3) / 4, c.num_slices), (int)B200_MAX_SLICES) : c.num_slices;
    b200enc_session *s = new (std::nothrow) b200enc_session();
    if (!s) return B200ENC_ENOMEM;
    s->cfg = c;
    Geom &g = s->g;
    g.width = c.width; g.height = c.height; g.mbw = (c.width + 15) / 16; g.mbh = (c.height + 15) / 16; g.wc = g.mbw * 16; g.hc = g.mbh * 16;
    geom_set_magic(g);
    g.search_range = c.search_range;
    auto slice_layout = [](Geom &q, int count) {
        q.num_slices = std::min(count, q.mbh);
        const int base = q.mbh / q.num_slices, rem = q.mbh % q.num_slices; int r = 0;
        for (int i = 0; i < q.num_slices; i++) { q.slice_row0[i] = r; r += base + (i < rem); }
        for (int i = q.num_slices; i <= B200_MAX_SLICES; i++) q.slice_row0[i] = r;
        memset(q.slice_top, 0, sizeof q.slice_top);
        for (int i = 0; i < q.num_slices; i++) q.slice_top[q.slice_row0[i] >> 5] |= 1u << (q.slice_row0[i] & 31);
    };
    slice_layout(g, c.num_slices); s->cfg.num_slices = g.num_slices;
    {   // borders of the padded planes: the widest reach of a search window / interpolation tap past the picture edge
        auto up = [](int v, int a) { return (v + a - 1) / a * a; };
        g.lp = up(c.search_range + 10, 16); g.ls = g.wc + 2 * g.lp;
        g.cp = g.lp / 2; g.cs = up(g.wc / 2 + 2 * g.cp, 16);
        g.p1 = g.lp / 2; g.s1 = up(g.wc / 2 + 2 * g.p1, 16);
        g.p2 = up(c.search_range / 4 + 5, 4); g.s2 = up(g.wc / 4 + 2 * g.p2, 16);
    }
    s->load = (double)c.width * c.height * c.fps;
    s->device = sched_acquire(c.device, s->load);
    if (s->device < 0) { delete s; return B200ENC_ENODEV; }
    int rc = B200ENC_OK;
    do {
        CU_TRY(cudaSetDevice(s->device), rc = B200ENC_ENODEV; break);
        const size_t ny = (size_t)g.wc * g.hc, nc = ny / 4, nmb = (size_t)g.mbw * g.mbh;
        s->g_key = g; slice_layout(s->g_key, c.key_slices); s->cfg.key_slices = s->g_key.num_slices;
        const int max_slice_rows = g.mbh / g.num_slices + (g.mbh % g.num_slices ? 1 : 0);
        s->rbsp_words_per_slice = (uint32_t)((size_t)max_slice_rows * g.mbw * B200_MB_SLOT_WORDS + 64);
        const int max_key_rows = g.mbh / s->g_key.num_slices + (g.mbh % s->g_key.num_slices ? 1 : 0);
        s->rbsp_words_per_slice_key = (uint32_t)((size_t)max_key_rows * g.mbw * B200_MB_SLOT_WORDS + 64);
        const std::vector<uint8_t> ps = make_parameter_sets(c.width, c.height, c.level_idc ? c.level_idc : level_for(c.width, c.height, c.fps), c.profile);
        s->hdr_len = (int)ps.size(); s->param_sets = ps;
        struct Item { void **p; size_t bytes; };
        std::vector<Item> items;
        auto add = [&](auto &ptr, size_t bytes) { items.push_back({ reinterpret_cast<void **>(&ptr), bytes }); };
        add(s->input, b200enc_frame_bytes(s));
        for (int k = 0; k < 3; k++) { const size_t b = k ? nc : ny; add(s->src[k], b); add(s->srcB[k], b); add(s->bufA[k], b); add(s->bufB[k], b); add(s->rec_pre[k], (c.debug & 1) ? b : 16); }
        const size_t l1 = (size_t)g.s1 * (g.hc / 2 + 2 * g.p1) + 64, l2 = (size_t)g.s2 * (g.hc / 4 + 2 * g.p2) + 64;
        const size_t lplane = (size_t)g.ls * (g.hc + 2 * g.lp), cplane = (size_t)g.cs * (g.hc / 2 + 2 * g.cp) + 64;
        add(s->srcL1, l1); add(s->refL1, l1); add(s->srcL2, l2); add(s->refL2, l2);
        add(s->rpl, 4 * lplane + 256); add(s->rpc[0], cplane); add(s->rpc[1], cplane); add(s->tmaps, 4 * sizeof(CUtensorMap));
        add(s->mbi, nmb * sizeof(MbInfo)); add(s->coef, nmb * sizeof(MbCoef)); add(s->dbk_bs, nmb * sizeof(uint4));
        add(s->me2, nmb * 4); add(s->me1, nmb * 4); add(s->me0, nmb * 4); add(s->inter_cost, nmb * 4);
        add(s->skip_run, (nmb + B200_MAX_SLICES) * 4); add(s->mb_bits, nmb * 4); add(s->mb_off, nmb * 4);
        add(s->mb_slot, nmb * B200_MB_SLOT_WORDS * 4);
        add(s->rbsp, std::max((size_t)s->rbsp_words_per_slice * g.num_slices, (size_t)s->rbsp_words_per_slice_key * s->g_key.num_slices) * 4);
        add(s->side, c.profile ? nmb * sizeof(MbSide) : 16); add(s->slice_nbins, B200_MAX_SLICES * 4);
        add(s->bins, c.profile ? (nmb * B200_MB_BIN_SLOT + 64) * sizeof(uint16_t) : 16);
        add(s->bins_mb, c.profile ? (nmb * CABAC_MB_SLOT + 64) * sizeof(uint16_t) : 16);
        add(s->bin_lane_cnt, c.profile ? nmb * 32 * sizeof(uint16_t) : 16);
        add(s->bins_hdr, c.profile ? nmb * CABAC_HDR_SLOT * sizeof(uint16_t) : 16);
        add(s->slice_bits, B200_MAX_SLICES * 4); add(s->hdr, 256); add(s->row_prog, (size_t)g.mbh * 2 * 4);
        add(s->dbk_ll, (size_t)g.mbh * (g.mbw + 1) * DBK_LL_PER_MB * sizeof(uint2));
        size_t total = 0;
        for (auto &it : items) total += align_up(it.bytes, 256);
        CU_TRY(cudaMalloc(&s->d_pool, total), rc = B200ENC_ENOMEM; break);
        s->pool_bytes = total;
        CU_TRY(cudaMemset(s->d_pool, 0, total), rc = B200ENC_ECUDA; break);
        size_t off = 0;
        for (auto &it : items) { *it.p = s->d_pool + off; off += align_up(it.bytes, 256); }
        CU_TRY(cudaMemcpy(s->hdr, ps.data(), ps.size(), cudaMemcpyHostToDevice), rc = B200ENC_ECUDA; break);
        {   // TMA descriptors of the tiles k_me_fine fetches (u8 elements, no swizzle, zero fill outside the tensor)
            CUtensorMap tm[4];
            const uint64_t H = (uint64_t)g.hc + 2 * g.lp;
            if (!encode_tmap(&tm[0], s->src[0], 2, (uint64_t)g.wc, (uint64_t)g.hc, 1, (uint64_t)g.wc, 0, 16, 16, 1) ||
                !encode_tmap(&tm[1], s->rpl, 2, (uint64_t)g.ls, H, 1, (uint64_t)g.ls, 0, 48, 20, 1) ||
                !encode_tmap(&tm[2], s->rpl, 3, (uint64_t)g.ls, H, 4, (uint64_t)g.ls, (uint64_t)g.ls * H, 48, 18, 4) ||
                !encode_tmap(&tm[3], s->srcB[0], 2, (uint64_t)g.wc, (uint64_t)g.hc, 1, (uint64_t)g.wc, 0, 16, 16, 1)) { rc = B200ENC_ENODEV; break; }
            CU_TRY(cudaMemcpy(s->tmaps, tm, sizeof tm, cudaMemcpyHostToDevice), rc = B200ENC_ECUDA; break);
        }
        // output buffer: twice the raw frame + 64 KB (CAVLC without I_PCM can exceed the raw size on noise at very low QP:
        // 1.75x measured at QP 0); a frame that still does not fit is reported as B200ENC_EOVERFLOW, never truncated silently
        s->out_cap = (uint32_t)align_up(std::max<size_t>(ny * 3, 1 << 16) + (1 << 16), 256);
        if (c.debug & 2) s->out_cap = 8192;          // test hook: a buffer small enough to exercise the overflow path
        CU_TRY(cudaHostAlloc(&s->h_out, s->out_cap + 256, cudaHostAllocMapped), rc = B200ENC_ENOMEM; break);
        CU_TRY(cudaHostGetDevicePointer(reinterpret_cast<void **>(&s->d_out), s->h_out, 0), rc = B200ENC_ECUDA; break);
        memset(s->h_out, 0, s->out_cap + 256);
        s->own = new (std::nothrow) b200enc_batch();
        if (!s->own) { rc = B200ENC_ENOMEM; break; }
        rc = batch_init(s->own, s->device, 1);
        if (rc != B200ENC_OK) break;
        CU_TRY(cudaStreamCreateWithFlags(&s->up_stream, cudaStreamNonBlocking), rc = B200ENC_ENODEV; break);
        CU_TRY(cudaEventCreateWithFlags(&s->ev_up, cudaEventDisableTiming), rc = B200ENC_ENODEV; break);
        {   // the wrapper's rate-control settings: RC_BITRATE_MODE, iMaxBitrate = iTargetBitrate, iMinQp / iMaxQp (VideoEncoderOpenH264.cpp:230,239-240,274)
            b200rc::Config rcc; rcc.bitrate = c.bitrate; rcc.max_bitrate = c.max_bitrate; rcc.fps = c.fps; rcc.width = c.width; rcc.height = c.height;
            rcc.min_qp = c.min_qp; rcc.max_qp = c.max_qp > 0 ? c.max_qp : 51;
            s->rc.init(rcc);
        }
        if (rc == B200ENC_OK && c.auto_batch) scheduler_register(s, +1);
    } while (0);
    if (rc != B200ENC_OK) { b200enc_destroy(s); return rc; }
    *out = s;
    return B200ENC_OK;
}

void b200enc_destroy(b200enc_session *s)
{
    if (!s) return;
    if (s->device >= 0) {
        if (s->in_scheduler) scheduler_register(s, -1);
        cudaSetDevice(s->device);
        if (s->up_stream) { cudaStreamSynchronize(s->up_stream); cudaStreamDestroy(s->up_stream); }
        if (s->ev_up) cudaEventDestroy(s->ev_up);
        if (s->h_stage) cudaFreeHost(s->h_stage);
        if (s->own) batch_free(s->own);
        if (s->h_out) cudaFreeHost(s->h_out);
        if (s->d_pool) cudaFree(s->d_pool);
        sched_release(s->device, s->load);
    }
    delete s;
}

size_t b200enc_frame_bytes(const b200enc_session *s)
{
    if (!s) return 0;
    const size_t px = (size_t)s->cfg.width * s->cfg.height;
    return s->cfg.input_format == B200ENC_FMT_RGBA ? px * 4 : px * 3 / 2;
}
int b200enc_device_of(const b200enc_session *s) { return s ? s->device : -1; }
int b200enc_force_idr(b200enc_session *s) { if (!s) return B200ENC_EINVAL; s->force_idr = true; return B200ENC_OK; }
int b200enc_get_parameter_sets(b200enc_session *s, uint8_t *out, uint32_t cap, uint32_t *len)
{
    if (!s || !out || !len || cap < s->param_sets.size()) return B200ENC_EINVAL;
    memcpy(out, s->param_sets.data(), s->param_sets.size());
    *len = (uint32_t)s->param_sets.size();
    return B200ENC_OK;
}

int b200enc_encode(b200enc_session *s, const uint8_t *frame, uint32_t size, const uint8_t **bs, uint32_t *bs_size, b200enc_frame_info *info)
{
    if (!s || !frame) return B200ENC_EINVAL;
    if (size < b200enc_frame_bytes(s)) return B200ENC_ESIZE;
    if (s->in_scheduler) return scheduler_encode(s, frame, bs, bs_size, info);
    const int rc = session_upload(s, frame);
    if (rc != B200ENC_OK) return rc;
    return encode_impl(s->own, &s, 1, &frame, IN_UPLOADED, bs, bs_size, info, nullptr);
}
float b200enc_last_kernel_ms(const b200enc_session *s) { return s && s->own ? s->own->last_ms : 0.f; }

int b200enc_batch_create(int device, int max_sessions, b200enc_batch **out)
{
    if (!out || max_sessions <= 0 || device < 0 || device >= b200enc_device_count()) return out && b200enc_device_count() == 0 ? B200ENC_ENODEV : B200ENC_EINVAL;
    b200enc_batch *b = new (std::nothrow) b200enc_batch();
    if (!b) return B200ENC_ENOMEM;
    const int rc = batch_init(b, device, max_sessions);
    if (rc != B200ENC_OK) { batch_free(b); return rc; }
    *out = b;
    return B200ENC_OK;
}
void b200enc_batch_destroy(b200enc_batch *b) { batch_free(b); }
int b200enc_batch_encode(b200enc_batch *b, b200enc_session *const *sessions, int n, const uint8_t *const *frames, int device_input,
                         const uint8_t **bs, uint32_t *bs_size, b200enc_frame_info *infos)
{
    if (!b) return B200ENC_EINVAL;
    b->last_rcs.assign(n > 0 ? n : 0, B200ENC_EINVAL);
    return encode_impl(b, sessions, n, frames, device_input ? IN_DEVICE : IN_HOST, bs, bs_size, infos, b->last_rcs.data());
}
int b200enc_batch_last_status(const b200enc_batch *b, int *rcs, int cap)
{
    if (!b || !rcs) return 0;
    int n = 0;
    for (; n < (int)b->last_rcs.size() && n < cap; n++) rcs[n] = b->last_rcs[n];
    return n;
}
uint32_t b200enc_rc_retries(const b200enc_session *s) { return s ? s->retries : 0; }
float b200enc_batch_last_kernel_ms(const b200enc_batch *b) { return b ? b->last_ms : 0.f; }
int b200enc_batch_last_launches(const b200enc_batch *b) { return b ? b->last_launches : 0; }
int b200enc_batch_set_profiling(b200enc_batch *b, int on) { if (!b) return B200ENC_EINVAL; b->profiling = on != 0; return B200ENC_OK; }
int b200enc_batch_kernel_times(const b200enc_batch *b, const char **names, float *ms, int cap)
{
    if (!b) return 0;
    int n = 0;
    for (; n < b->n_ktimes && n < cap; n++) { names[n] = b->ktimes[n].name; cudaEventElapsedTime(&ms[n], b->ktimes[n].ev0, b->ktimes[n].ev1); }
    return n;
}

void *b200enc_host_alloc(size_t bytes) { void *p = nullptr; return cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess ? p : nullptr; }
void b200enc_host_free(void *p) { if (p) cudaFreeHost(p); }
void *b200enc_dev_alloc(int device, size_t bytes) { void *p = nullptr; if (cudaSetDevice(device) != cudaSuccess) return nullptr; return cudaMalloc(&p, bytes) == cudaSuccess ? p : nullptr; }
int b200enc_dev_upload(int device, void *dst, const void *src, size_t bytes)
{
    CU_TRY(cudaSetDevice(device), return B200ENC_ECUDA);
    CU_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice), return B200ENC_ECUDA);
    return B200ENC_OK;
}
void b200enc_dev_free(int device, void *p) { if (p && cudaSetDevice(device) == cudaSuccess) cudaFree(p); }

int b200enc_slice_counts(b200enc_session *s, int *p_slices, int *key_slices)
{
    if (!s) return B200ENC_EINVAL;
    if (p_slices) *p_slices = s->g.num_slices;
    if (key_slices) *key_slices = s->g_key.num_slices;
    return B200ENC_OK;
}

int b200enc_get_stage(b200enc_session *s, int stage, void *out, size_t cap, size_t *written)
{
    if (!s || !out) return B200ENC_EINVAL;
    CU_TRY(cudaSetDevice(s->device), return B200ENC_ECUDA);
    const size_t nmb = (size_t)s->g.mbw * s->g.mbh, ny = (size_t)s->g.wc * s->g.hc;
    uint8_t **last = s->cur_is_A ? s->bufB : s->bufA;     // the frame just encoded (roles were swapped after it)
    const void *src = nullptr; size_t bytes = 0; uint8_t **planes = nullptr;
    switch (stage) {
    case B200ENC_STAGE_MBINFO: src = s->mbi; bytes = nmb * sizeof(MbInfo); break;
    case B200ENC_STAGE_MBCOEF: src = s->coef; bytes = nmb * sizeof(MbCoef); break;
    case B200ENC_STAGE_ME2: src = s->me2; bytes = nmb * 4; break;
    case B200ENC_STAGE_ME1: src = s->me1; bytes = nmb * 4; break;
    case B200ENC_STAGE_ME0: src = s->me0; bytes = nmb * 4; break;
    case B200ENC_STAGE_INTER_COST: src = s->inter_cost; bytes = nmb * 4; break;
    case B200ENC_STAGE_MBSIDE: if (!s->cfg.profile) return B200ENC_EINVAL; src = s->side; bytes = nmb * sizeof(MbSide); break;
    case B200ENC_STAGE_BIN_COUNT: if (!s->cfg.profile) return B200ENC_EINVAL; src = s->mb_bits; bytes = nmb * 4; break;
    case B200ENC_STAGE_BIN_OFF: if (!s->cfg.profile) return B200ENC_EINVAL; src = s->mb_off; bytes = nmb * 4; break;
    case B200ENC_STAGE_BINS: if (!s->cfg.profile) return B200ENC_EINVAL; src = s->bins; bytes = nmb * B200_MB_BIN_SLOT * sizeof(uint16_t); break;
    case B200ENC_STAGE_SRC: planes = s->src_is_A ? s->srcB : s->src; break;       // the picture just encoded (roles were swapped after it)
    case B200ENC_STAGE_REC_PRE: if (!(s->cfg.debug & 1)) return B200ENC_EINVAL; planes = s->rec_pre; break;
    case B200ENC_STAGE_REC: planes = last; break;
    default: return B200ENC_EINVAL;
    }
    if (planes) {
        bytes = ny * 3 / 2;
        if (cap < bytes) return B200ENC_EINVAL;
        uint8_t *o = static_cast<uint8_t *>(out);
        CU_TRY(cudaMemcpy(o, planes[0], ny, cudaMemcpyDeviceToHost), return B200ENC_ECUDA);
        CU_TRY(cudaMemcpy(o + ny, planes[1], ny / 4, cudaMemcpyDeviceToHost), return B200ENC_ECUDA);
        CU_TRY(cudaMemcpy(o + ny + ny / 4, planes[2], ny / 4, cudaMemcpyDeviceToHost), return B200ENC_ECUDA);
    } else {
        if (cap < bytes) return B200ENC_EINVAL;
        CU_TRY(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost), return B200ENC_ECUDA);
    }
    if (written) *written = bytes;
    return B200ENC_OK;
}

int b200enc_get_recon(b200enc_session *s, uint8_t *i420, size_t cap)
{
    if (!s || !i420 || cap < (size_t)s->cfg.width * s->cfg.height * 3 / 2) return B200ENC_EINVAL;
    CU_TRY(cudaSetDevice(s->device), return B200ENC_ECUDA);
    uint8_t **last = s->cur_is_A ? s->bufB : s->bufA;
    const int w = s->cfg.width, h = s->cfg.height, wc = s->g.wc;
    CU_TRY(cudaMemcpy2D(i420, w, last[0], wc, w, h, cudaMemcpyDeviceToHost), return B200ENC_ECUDA);
    CU_TRY(cudaMemcpy2D(i420 + (size_t)w * h, w / 2, last[1], wc / 2, w / 2, h / 2, cudaMemcpyDeviceToHost), return B200ENC_ECUDA);
    CU_TRY(cudaMemcpy2D(i420 + (size_t)w * h * 5 / 4, w / 2, last[2], wc / 2, w / 2, h / 2, cudaMemcpyDeviceToHost), return B200ENC_ECUDA);
    return B200ENC_OK;
}

} // extern "C"

#include "k_test_abi.inl"
#include "rc_capi.inl"
