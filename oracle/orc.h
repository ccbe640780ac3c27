/*
 * oracle/orc.h -- TEST INFRASTRUCTURE ONLY. CPU oracle ("orc") for the B200 H.264 encode path.
 *
 * What it restates: the work done behind ISVCEncoder::EncodeFrame at
 * /root/reference/video_codec/VideoEncoderOpenH264.cpp:344 (one I420 frame in -> one Annex-B
 * access unit out, Baseline / CAVLC / IPPP / 1 reference, policy table at :228-296).
 * The arithmetic of that call lives in cisco/openh264 (libopenh264.so, dlopen'ed by name at
 * VideoEncoderOpenH264.cpp:46,203; no version pinned, vendored headers correspond to API ~v2.0.0),
 * which is absent from /root/reference and from this image. The reference has no tests or golden
 * vectors, so PARITY WITH OPENH264 IS UNPINNED. What IS pinned:
 *   - every normative stage (CAVLC, intra/inter prediction, dequant/IDCT, deblocking) by decoding the
 *     oracle's streams with FFmpeg's independent h264 decoder and comparing with the oracle's own
 *     reconstruction bit-exactly (tests/test_oracle.py::test_oracle_stream_decodes_to_its_reconstruction and the other *_decode tests);
 *   - the CAVLC/deblock/cbp tables against the copies inside that decoder (tests/test_oracle.py::test_tables_match_the_independent_decoder,
 *     ::test_cabac_tables_match_the_independent_decoder).
 * Encoder-side free choices (search pattern, costs, quantiser rounding) follow the standard JM-style
 * definitions named in SURVEY.md section 2b and are the specification the CUDA path must reproduce
 * bit-for-bit (same bitstream bytes, same reconstruction).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library. The product (media_b200/csrc) never includes or links anything from oracle/.
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- macroblock types used in the per-MB side arrays (shared layout with the CUDA path) ---- */
enum { ORC_MB_P16x16 = 0, ORC_MB_I16x16 = 1, ORC_MB_I4x4 = 2, ORC_MB_PSKIP = 3, ORC_MB_P8x8 = 4,
       ORC_MB_I8x8 = 5 /* I_NxN with transform_size_8x8_flag = 1 (High profile) */ };
#define ORC_MB_IS_INTRA(m) ((m)->mb_type == ORC_MB_I16x16 || (m)->mb_type == ORC_MB_I4x4 || (m)->mb_type == ORC_MB_I8x8)
#define ORC_MB_IS_INXN(m) ((m)->mb_type == ORC_MB_I4x4 || (m)->mb_type == ORC_MB_I8x8)

/* Per-MB coefficient record: everything CAVLC needs, 816 bytes. Levels are in zig-zag scan order. */
typedef struct {
    int16_t luma[16][16];    /* [blkIdx][scan]; for I16x16 index 0 of each block is unused (AC only); with transform_size_8x8_flag
                                luma[4*b8] .. luma[4*b8+3] hold the 64 levels of 8x8 block b8 in 8x8 zig-zag order */
    int16_t luma_dc[16];     /* I16x16 only: Intra16x16DCLevel in scan order */
    int16_t chroma_dc[2][4]; /* [plane][c] raster order of the 2x2 block */
    int16_t chroma_ac[2][4][16]; /* [plane][blk][scan], index 0 unused */
} OrcMbCoef;

typedef struct {
    uint8_t  mb_type;        /* ORC_MB_* */
    uint8_t  i16_mode;       /* bits 0..1 Intra16x16PredMode; bit 2 transform_size_8x8_flag (inter MBs of High-profile streams) */
    uint8_t  chroma_mode;    /* intra_chroma_pred_mode 0..3 */
    uint8_t  cbp;            /* bits 0..3 luma 8x8, bits 4..5 chroma (0,1,2) */
    int16_t  mv[2];          /* quarter-pel; the 16x16 vector (partition 0 for P_8x8) */
    union {
        uint8_t i4_mode[16]; /* intra MBs: Intra4x4PredMode per blkIdx (Intra_8x8: the 8x8 block's mode in all four of its entries) */
        int16_t mv8[4][2];   /* inter MBs: vector of each 8x8 partition (all equal for P_L0_16x16 / P_Skip) */
    };
    uint8_t  nnz[24];        /* total_coeff: 0..15 luma blkIdx, 16..19 Cb, 20..23 Cr (AC count for I16x16/chroma) */
} OrcMbInfo;

/* Per-MB side record of the CABAC back end (shared layout with the CUDA path): what context selection needs from a neighbour
 * beyond OrcMbInfo. 20 bytes. */
typedef struct {
    union {
        int16_t mvd[4][2];   /* inter MBs: mvd_l0 of each 8x8 partition (replicated for P_L0_16x16; 0 for P_Skip) */
        uint8_t i4_syn[16];  /* Intra_4x4: 8 = prev_intra4x4_pred_mode_flag set, else rem_intra4x4_pred_mode 0..7 */
    };
    uint8_t dc_cbf;          /* coded_block_flag of the DC blocks: bit 0 Intra16x16DCLevel, bit 1 Cb DC, bit 2 Cr DC */
    uint8_t pad[3];
} OrcMbSide;
/* bin-list entry (16 bits): regular bins ctxIdx | bin << 10 | (repeat - 1) << 11 (repeat consecutive equal bins of one context,
 * used for the unary prefixes); ctxIdx 276 = terminate bin; ctxIdx ORC_CTX_BYPASS0 + n (n = 1..6) = n bypass bins held in
 * bits 10..10+n-1, the first one most significant */
#define ORC_CTX_BYPASS0 0x3F8
#define ORC_CABAC_NCTX 460

typedef struct OrcEncoder OrcEncoder;

typedef struct {
    int width, height;       /* display size (even, 16..4096) */
    int num_slices;          /* >=1, MB-row groups */
    int search_range;        /* full-pel, multiple of 4: 16/32/64 */
    int level_idc;           /* 0 = derive from size/fps */
    int fps;
    int no_i4x4;             /* 1: Intra_16x16 only (quality A/B runs in tests; the product has no such switch) */
    int no_p8x8;             /* 1: P_L0_16x16 only (same purpose) */
    int no_scene_change;     /* 1: never turn a P frame into an IDR (b200enc_config.scene_change = 0) */
    int profile;             /* 0 Constrained Baseline / CAVLC; 1 Main / CABAC; 2 High / CABAC with the 8x8 transform on inter MBs, the
                                wrapper's persist.vmi.video.encode.profile values (VideoEncoderOpenH264.cpp:248-253) */
    int no_t8x8;             /* 1: High profile without the 8x8 transform (transform_8x8_mode_flag = 0; quality A/B runs) */
    int background_detection;/* 1: static macroblocks (against the previous source picture) are skipped (b200enc_config.background_detection) */
    int complexity_set, complexity; /* complexity_set = 1: iComplexityMode 0 LOW (no Intra_4x4 trial, no P_8x8), 1 MEDIUM (no P_8x8), 2 HIGH */
    int intra8x8;            /* 1 (what the product does): High profile intra MBs may also take Intra_8x8 (8.3.2), tried before Intra_4x4 and chosen
                                against it by J = 64 SSD + 27 lambda^2 B; 0: Intra_4x4 / Intra_16x16 only (quality A/B runs) */
    int key_slices;          /* MB-row groups of pictures REQUESTED as key pictures (b200enc_config.key_slices); < 1: num_slices. A P picture the
                                scene-change rule promotes keeps num_slices */
} OrcConfig;
#define ORC_BGD_OU_SAD 128       /* background detection: largest SAD of an 8x8 unit against the previous source picture (mean |d| <= 2) */
#define ORC_BGD_MAXDIFF 12       /* ... and largest single sample difference */
#define ORC_SC_MIN_DISTANCE 10   /* scene-change promotion needs this many pictures since the last IDR (the IDR counts as the first) */
#define ORC_MB_T8(m) (((m)->i16_mode >> 2) & 1)

OrcEncoder *orc_create(const OrcConfig *cfg);
void orc_destroy(OrcEncoder *e);
/* Encode one frame. frame_type: 1 = IDR (SPS+PPS prepended), 0 = P. Returns bytes written or <0. */
int orc_encode(OrcEncoder *e, const uint8_t *i420, int frame_type, int qp, uint8_t *out, int out_cap);
/* The same without advancing the stream state (reference picture, frame_num, idr_pic_id): a picture the rate control codes again. */
int orc_encode_trial(OrcEncoder *e, const uint8_t *i420, int frame_type, int qp, uint8_t *out, int out_cap);
/* 1 when the last frame was coded as an IDR (requested, first frame, or scene change), else 0 */
int orc_last_frame_was_idr(const OrcEncoder *e);
/* Reconstruction of the last encoded frame (after deblocking), cropped to width x height I420. */
void orc_get_recon(const OrcEncoder *e, uint8_t *i420);
/* Stage dumps of the last frame for stage-by-stage parity with the CUDA path. */
const OrcMbInfo *orc_mb_info(const OrcEncoder *e);
const OrcMbCoef *orc_mb_coef(const OrcEncoder *e);
int orc_mb_count(const OrcEncoder *e);
/* coded (padded) planes of the last frame: which = 0 src, 1 recon before deblock, 2 recon after deblock */
const uint8_t *orc_plane(const OrcEncoder *e, int which, int comp, int *stride, int *w, int *h);
/* coarse ME results of last P frame: level 2 (1/4), 1 (1/2), 0 (full-pel) MVs in that level's pixel units */
const int16_t *orc_me_level(const OrcEncoder *e, int level);
/* best inter cost (SATD + lambda*mv bits) per MB of the last P frame */
const int32_t *orc_inter_cost(const OrcEncoder *e);
/* quarter-pel sample of the current reference through the encoder's half-pel planes (== orc_interp_luma) */
int orc_dbg_qpel(const OrcEncoder *e, int xq, int yq);
/* (re)build the half-pel planes from the current reference picture; orc_encode does this itself for P frames */
void orc_dbg_build_halfpel(OrcEncoder *e);

/* CABAC stage dumps of the last frame: side records, and the bin list of slice s (returns the count) */
const OrcMbSide *orc_mb_side(const OrcEncoder *e);
int orc_slice_bins(const OrcEncoder *e, int s, const uint16_t **bins);
/* codes a bin list (ending with a terminate bin of value 1) into bytes; returns the length */
int orc_cabac_code_bins(const uint16_t *bins, int n, int slice_qp, int is_p, uint8_t *out, int cap);

/* ---- headers ---- */
int orc_write_sps(uint8_t *out, int width, int height, int level_idc, int profile);
int orc_write_pps(uint8_t *out, int profile, int transform8x8);
int orc_level_for(int width, int height, int fps);

/* ---- kernel-level oracles (bit-exact targets of the per-kernel C-ABI entry points) ---- */
int  orc_sad(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h);
int  orc_satd4x4(const uint8_t *a, int sa, const uint8_t *b, int sb);
int  orc_satd16x16(const uint8_t *a, int sa, const uint8_t *b, int sb);
void orc_dct4x4(const int16_t *res /*16 raster*/, int16_t *coef /*16 raster*/);
void orc_idct4x4(const int32_t *d /*16 raster dequantised*/, int32_t *r /*16 raster, (x+32)>>6 applied*/);
/* quantise raster coef -> zigzag levels; intra selects the dead-zone; ac_only skips position 0. Returns nnz */
int  orc_quant4x4(const int16_t *coef, int16_t *level_zz, int qp, int intra, int ac_only);
void orc_dequant4x4(const int16_t *level_zz, int32_t *d, int qp, int ac_only);
/* 8x8 transform of the High profile: forward transform (rows, then columns), quantiser (64 levels in 8x8 zig-zag order; returns nnz),
 * 8.5.13 scaling and inverse transform ((x+32)>>6 applied) */
void orc_dct8x8(const int16_t *res /*64 raster*/, int32_t *coef /*64 raster*/);
int  orc_quant8x8(const int32_t *coef, int16_t *level_zz, int qp, int intra);
void orc_dequant8x8(const int16_t *level_zz, int32_t *d, int qp);
void orc_idct8x8(const int32_t *d, int32_t *r);
void orc_rgba_to_i420(const uint8_t *rgba, int width, int height, uint8_t *i420);
void orc_nv12_to_i420(const uint8_t *nv12, int width, int height, uint8_t *i420);
void orc_downsample2(const uint8_t *src, int sstride, int w, int h, uint8_t *dst, int dstride);
/* luma sample at quarter-pel position (xq,yq) with edge clamping (8.4.2.2.1) */
int  orc_interp_luma(const uint8_t *ref, int stride, int w, int h, int xq, int yq);
/* chroma sample at 1/8-pel position */
int  orc_interp_chroma(const uint8_t *ref, int stride, int w, int h, int x8, int y8);
/* in-place deblocking of a whole coded frame given per-MB info (8.7) */
void orc_deblock_frame(uint8_t *y, int ys, uint8_t *u, uint8_t *v, int cs, int mbw, int mbh,
                       const OrcMbInfo *mbi, int qp);
/* Intra_4x4 predictor of one block (8.3.1.2); avail bits: 1 top, 2 left, 4 corner, 8 top-right. Returns 0 if the mode is unavailable */
int  orc_pred_i4(const uint8_t *r, int stride, int mode, int avail, uint8_t *pred16);
/* Intra_8x8 predictor of one block (8.3.2.2, reference sample filter included); avail bits as orc_pred_i4. Returns 0 if the mode is unavailable */
int  orc_pred_i8(const uint8_t *r, int stride, int mode, int avail, uint8_t *pred64);
/* emulation prevention: returns output length */
int  orc_escape_rbsp(const uint8_t *in, int n, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
