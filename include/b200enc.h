/*
 * include/b200enc.h -- thin C ABI of the B200 H.264 encoder (libb200enc.so).
 *
 * This is the boundary the host C++ sibling of VideoEncoderOpenH264 (include/VideoEncoderB200.h) calls, and the
 * only thing a binding (ctypes, cgo, JNI) needs. Each entry point states the reference interface it replaces;
 * reference paths are relative to kunpengcompute/media.
 *
 *   b200enc_create        <- WelsCreateSVCEncoder + ISVCEncoder::InitializeExt + SetOption(DATAFORMAT)
 *                            (video_codec/VideoEncoderOpenH264.cpp:142,257,262; vendor/openh264/codec_api.h:284,545)
 *   b200enc_encode        <- ISVCEncoder::EncodeFrame(SSourcePicture*, SFrameBSInfo*)
 *                            (video_codec/VideoEncoderOpenH264.cpp:344-350; vendor/openh264/codec_api.h:309)
 *   b200enc_force_idr     <- ISVCEncoder::ForceIntraFrame(true) (video_codec/VideoEncoderOpenH264.cpp:406-415)
 *   b200enc_destroy       <- ISVCEncoder::Uninitialize + WelsDestroySVCEncoder (video_codec/VideoEncoderOpenH264.cpp:379-386)
 *   auto_batch scheduler  <- the reference's one-thread-per-session model (:294): N caller threads, each in its own
 *                            EncodeOneFrame, are served by one batch step per GPU
 *   b200enc_batch_*       <- no reference counterpart: N sessions of one GPU advance one frame in one set of kernel
 *                            launches (the reference runs one single-threaded encoder per session, :294)
 *   b200k_*               <- per-kernel entry points for parity tests against oracle/ (openh264's C kernels,
 *                            SURVEY.md section 2b, are not in the tree)
 *
 * All functions return 0 on success or a negative B200ENC_E* code; no exceptions cross this boundary.
 * There is no CPU fallback: without a CUDA device every call that needs one fails with B200ENC_ENODEV.
 */
#ifndef B200ENC_H
#define B200ENC_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    B200ENC_OK = 0,
    B200ENC_EINVAL = -1,     /* bad argument / unsupported configuration */
    B200ENC_ENODEV = -2,     /* no CUDA device or CUDA runtime failure at create */
    B200ENC_ENOMEM = -3,
    B200ENC_ECUDA = -4,      /* CUDA error during encode (b200enc_last_cuda_error has the code) */
    B200ENC_ESIZE = -5,      /* input buffer smaller than one frame (reference: ENCODE_FAIL, VideoEncoderOpenH264.cpp:307-310) */
    B200ENC_EOVERFLOW = -6,  /* bitstream larger than the output buffer */
    B200ENC_EWAVE = -7       /* wavefront watchdog fired (internal error) */
};

enum { B200ENC_FMT_I420 = 0, B200ENC_FMT_NV12 = 1, B200ENC_FMT_RGBA = 2 };
enum { B200ENC_FRAME_P = 0, B200ENC_FRAME_IDR = 1 };

typedef struct b200enc_config {
    int width, height;     /* even, 16..4096 (VideoEncoderOpenH264.cpp:16-17,162) */
    int fps;               /* frames per second of the session (30 or 60 in the reference, :166) */
    int bitrate;           /* bits per second, CBR target used when const_qp < 0 (:239-240) */
    int gop;               /* IDR period in frames (uiIntraPeriod, :242) */
    int const_qp;          /* 0..51: fixed QP for every frame; < 0: rate control */
    int num_slices;        /* MB-row groups, 1..35; 0 = automatic: 1 with CAVLC (the reference's SM_SINGLE_SLICE, :247), one per ~17 MB rows with
                              CABAC, whose arithmetic coder is a serial chain per slice */
    int search_range;      /* full-pel, multiple of 4 in 4..64 */
    int input_format;      /* B200ENC_FMT_* */
    int device;            /* CUDA ordinal, or -1: least-loaded device by pixel rate */
    int level_idc;         /* 0: derive from size and fps (the wrapper's LEVEL_3_2 at :255 is too small for 1080p) */
    int debug;             /* bit 0: keep stage dumps (pre-deblock reconstruction) for b200enc_get_stage; bit 1 (test hook): 8 KB output buffer, so that
                              the overflow path can be exercised */
    int scene_change;      /* 1 (default): a P frame whose macroblocks come out >= 2/5 intra after motion estimation is coded as an
                              IDR instead (the wrapper asks openh264 for bEnableSceneChangeDetect, VideoEncoderOpenH264.cpp:283) */
    int auto_batch;        /* 1: concurrent b200enc_encode calls of sessions living on the same GPU are coalesced by a per-GPU
                              scheduler thread into one batch step (what gives many single-threaded callers GPU-wide throughput) */
    int profile;           /* 0: Constrained Baseline, CAVLC; 1: Main, CABAC; 2: High, CABAC, 8x8 transform on inter MBs (transform_8x8_mode_flag) -- the wrapper's profile property
                              baseline / main / high (VideoEncoderOpenH264.cpp:186-188,248-253) with iEntropyCodingModeFlag = 1 (:291) */
    int max_bitrate;       /* bits per second the one-second leaky bucket of the rate control drains at (iMaxBitrate; the wrapper sets it to the
                              target, :239-240); 0 = bitrate */
    int min_qp, max_qp;    /* QP bounds of the rate control (iMinQp / iMaxQp; the wrapper keeps GetDefaultParams' 0 / 51, :230); max_qp 0 = 51 */
    int background_detection; /* bEnableBackgroundDetection (:282): static macroblocks are skipped on a pre-analysis of the source pictures */
    int complexity;        /* iComplexityMode (:289; the wrapper asks for HIGH_COMPLEXITY): 0 LOW (no Intra_4x4 trial, no P_8x8), 1 MEDIUM (no P_8x8), 2 HIGH */
    int key_slices;        /* MB-row groups of KEY pictures (1..35). A key frame is the longest dependent chain of a session: the intra wavefront's critical
                              path is mbw + mbh-of-the-slice steps and the CABAC coder walks a slice's bins one after the other, so key pictures may take
                              more slices than P pictures (a picture's slice count is free per picture, 7.3.3). 0 = automatic: num_slices when that was
                              given explicitly or with CAVLC; with CABAC and automatic num_slices one slice per 4 MB rows (1080p: 17). A P picture the
                              scene-change detector promotes on the device keeps the P pictures' slices */
} b200enc_config;

typedef struct b200enc_frame_info {
    int frame_type;        /* B200ENC_FRAME_* */
    int qp;
    uint32_t size_bytes;
    uint32_t frame_index;
} b200enc_frame_info;

typedef struct b200enc_session b200enc_session;
typedef struct b200enc_batch b200enc_batch;

void b200enc_default_config(b200enc_config *cfg);
int b200enc_create(const b200enc_config *cfg, b200enc_session **out);
void b200enc_destroy(b200enc_session *s);
/* Encode one frame from HOST memory. *bs points into an encoder-owned pinned buffer that stays valid until the
 * next encode/destroy on the same session (ownership as at VideoEncoderOpenH264.cpp:349-350). */
int b200enc_encode(b200enc_session *s, const uint8_t *frame, uint32_t size, const uint8_t **bs, uint32_t *bs_size,
                   b200enc_frame_info *info);
int b200enc_force_idr(b200enc_session *s);
/* SPS + PPS NALs of the session (Annex-B), what openh264's EncodeParameterSets returns (vendor/openh264/codec_api.h:316) */
int b200enc_get_parameter_sets(b200enc_session *s, uint8_t *out, uint32_t cap, uint32_t *len);
int b200enc_device_of(const b200enc_session *s);
size_t b200enc_frame_bytes(const b200enc_session *s);
int b200enc_last_cuda_error(void);
const char *b200enc_strerror(int code);
int b200enc_device_count(void);
/* statistics of the auto_batch scheduler of `device`: batches run and frames encoded through it */
int b200enc_scheduler_stats(int device, uint64_t *batches, uint64_t *frames);

/* Batched stepping: all sessions must live on the batch's device and share width/height/slices/search range/format. */
int b200enc_batch_create(int device, int max_sessions, b200enc_batch **out);
void b200enc_batch_destroy(b200enc_batch *b);
/* frames[i]: HOST pointer (device_input = 0) or DEVICE pointer already resident in HBM (device_input = 1). */
int b200enc_batch_encode(b200enc_batch *b, b200enc_session *const *sessions, int n, const uint8_t *const *frames,
                         int device_input, const uint8_t **bs, uint32_t *bs_size, b200enc_frame_info *infos);
/* Outcome per session of the last b200enc_batch_encode (B200ENC_OK / B200ENC_EOVERFLOW ...): one session's failure does not
 * invalidate the others' access units. b200enc_batch_encode itself returns B200ENC_EOVERFLOW when any session overflowed. Returns the count. */
int b200enc_batch_last_status(const b200enc_batch *b, int *rcs, int cap);
/* pictures of this session the rate control coded twice (first attempt above the hard cap) */
uint32_t b200enc_rc_retries(const b200enc_session *s);
/* device time of the kernels of the last batch_encode / encode call, milliseconds (CUDA events on the encode stream) */
float b200enc_batch_last_kernel_ms(const b200enc_batch *b);
float b200enc_last_kernel_ms(const b200enc_session *s);
/* number of kernel launches issued by the last call */
int b200enc_batch_last_launches(const b200enc_batch *b);
/* per-kernel device times of the last call when profiling is on: fills names[i] (static strings) and ms[i]; returns count */
int b200enc_batch_set_profiling(b200enc_batch *b, int on);
int b200enc_batch_kernel_times(const b200enc_batch *b, const char **names, float *ms, int cap);

/* pinned host memory helpers for callers that want asynchronous H2D (cudaHostAlloc / cudaFreeHost) */
void *b200enc_host_alloc(size_t bytes);
void b200enc_host_free(void *p);
/* plain device memory helpers for the device_input = 1 path (cudaMalloc / cudaMemcpy / cudaFree on `device`) */
void *b200enc_dev_alloc(int device, size_t bytes);
int b200enc_dev_upload(int device, void *dst, const void *src, size_t bytes);
void b200enc_dev_free(int device, void *p);

/* the session's slice counts after the automatic rules: P pictures / key pictures (b200enc_config.num_slices / key_slices) */
int b200enc_slice_counts(b200enc_session *s, int *p_slices, int *key_slices);

/* ---- test / parity hooks ---- */
enum {
    B200ENC_STAGE_MBINFO = 0,      /* n_mb * 48 bytes */
    B200ENC_STAGE_MBCOEF = 1,      /* n_mb * 816 bytes */
    B200ENC_STAGE_ME2 = 2, B200ENC_STAGE_ME1 = 3, B200ENC_STAGE_ME0 = 4,   /* n_mb * 2 int16 */
    B200ENC_STAGE_INTER_COST = 5,  /* n_mb int32 */
    B200ENC_STAGE_SRC = 6,         /* coded-size I420 source planes */
    B200ENC_STAGE_REC_PRE = 7,     /* coded-size reconstruction before deblocking (debug = 1 only) */
    B200ENC_STAGE_REC = 8,         /* coded-size reconstruction after deblocking */
    B200ENC_STAGE_MBSIDE = 9,      /* CABAC sessions: n_mb * 20 bytes (mvd / Intra_4x4 mode syntax / DC coded_block_flags) */
    B200ENC_STAGE_BIN_COUNT = 10,  /* CABAC sessions: n_mb uint32, bin-list entries per MB */
    B200ENC_STAGE_BIN_OFF = 11,    /* CABAC sessions: n_mb uint32, offset of the MB's entries inside its slice's list */
    B200ENC_STAGE_BINS = 12        /* CABAC sessions: n_mb * 3136 uint16; the list of a slice starts at first_mb * 3136 */
};
int b200enc_get_stage(b200enc_session *s, int stage, void *out, size_t cap, size_t *written);
/* display-size I420 reconstruction of the last frame */
int b200enc_get_recon(b200enc_session *s, uint8_t *i420, size_t cap);

/* per-kernel entry points (host pointers in, host pointers out; each runs the production kernel on `device`) */
int b200k_convert_to_i420(int device, int input_format, const uint8_t *in, int width, int height, uint8_t *i420_coded,
                          int *coded_w, int *coded_h);
int b200k_downsample2(int device, const uint8_t *in, int width, int height, uint8_t *out);
int b200k_sad16x16(int device, const uint8_t *cur, const uint8_t *ref, int stride, int n_blocks, const int32_t *xy /* 4 ints per block: cx, cy, rx, ry */, int32_t *sad);
int b200k_satd16x16(int device, const uint8_t *cur, const uint8_t *ref, int stride, int n_blocks, const int32_t *xy, int32_t *satd);
int b200k_transform_block(int device, const int16_t *residual /* n*16 */, int n, int qp, int intra, int16_t *levels_zz /* n*16 */, int32_t *recon_residual /* n*16 */);
/* the 8x8 transform chain of the High profile (residual -> transform -> quant -> 8.5.13 scaling / inverse): n blocks of 64 */
int b200k_transform_block8(int device, const int16_t *residual /* n*64 raster */, int n, int qp, int intra, int16_t *levels_zz /* n*64, 8x8 zig-zag */, int32_t *recon_residual /* n*64 */);
int b200k_deblock(int device, uint8_t *i420_coded, int mbw, int mbh, const void *mbinfo, int qp);
/* the CABAC arithmetic coder (9.3.4.2) on a bin list ending with a terminate bin of value 1 (entry format: oracle/orc.h) */
int b200k_cabac_code(int device, const uint16_t *bins, int n, int qp, int is_p, uint8_t *out, int cap, int *out_len);
/* the coder kernel `reps` times on `copies` independent copies of one list (one CTA each): average device ms per launch; stats[8]
 * = phase cycle counts in a -DCABAC_TIMING build, else zeros */
int b200k_cabac_code_bench(int device, const uint16_t *bins, int n, int qp, int is_p, int reps, int copies, float *ms, long long *stats);
/* the rate control object of the sessions (media_b200/csrc/rate_control.h) driven from outside with picture sizes -- host logic, no GPU:
 * pick returns the QP for a picture of `type` (B200ENC_FRAME_*), retry_qp the QP of a second attempt or -1 (the picture was planned as
 * planned_type and came out as coded_type: a scene-change promotion turns a P picture into an IDR), update commits a picture */
void *b200k_rc_create(double bitrate, double max_bitrate, int fps, int min_qp, int max_qp, int width, int height);
int b200k_rc_pick(void *rc, int type, double *budget_bits, double *hard_cap_bits);
int b200k_rc_retry_qp(void *rc, int planned_type, int coded_type, double bits);
void b200k_rc_update(void *rc, int type, int qp, double bits);
double b200k_rc_vbv(void *rc, double *bucket_bits);
void b200k_rc_destroy(void *rc);
/* microbenchmark: register-resident VABSDIFF4.U8.ACC issue rate, giga lane-instructions per second, and the SM clock seen */
int b200k_vabsdiff4_peak(int device, double *ginstr_per_s, int *sm_clock_mhz);
/* libb200enc_checked.so (-DB200_CHECKED: device-side bound checks on every computed slot / ring / list / tile index): failures counted since
 * load, *first_site = id of the first failing check; returns -1 when the library is not the checked build */
int b200k_check_failures(int device, int *first_site);
/* host-only (no device needed): the division-free macroblock-index arithmetic of the warp-per-MB kernels (h264_dev.cuh: mb_xy) checked against / and % for
 * every mb in [0, n) at a picture width of mbw macroblocks; returns the number of mismatches (0 expected), -1 for invalid arguments */
int b200k_mb_xy_mismatches(int mbw, int n);
/* issue-rate microbenchmarks behind the INT roofline: kind 0 VABSDIFF4.U8.ACC, 1 IADD3, 2 LOP3, 3 IDP.4A, 4 IMAD, 5 VIMNMX, 6 IABS + IADD, 7 SHF,
 * 8 the 4x4 Hadamard SATD of the motion search counted as 64 lane-operations. out[4 * kind + 0..3] = G warp-instructions/s of the whole GPU,
 * warp-instructions per clock per SM, the SM clock of the run in MHz (clock64 / globaltimer measured inside the kernel), instructions per unit */
int b200k_int_peaks(int device, double *out, int kinds);

#ifdef __cplusplus
}
#endif
#endif
