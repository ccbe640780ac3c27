/*
 * oracle/orc_encoder.c -- TEST INFRASTRUCTURE ONLY (see orc.h for scope and pinning status).
 *
 * Scalar restatement of one ISVCEncoder::EncodeFrame call (reference call site
 * /root/reference/video_codec/VideoEncoderOpenH264.cpp:344) for the stream structure the wrapper
 * fixes at :228-296 (Baseline, CAVLC, IPPP, one reference, deblocking idc 0, SPS+PPS with every
 * IDR). Normative stages follow ITU-T H.264 (clauses quoted per function); encoder-side choices
 * (search pattern, costs, mode decision) are the ones written down in DESIGN.md section 3 and are
 * what media_b200/csrc must reproduce bit for bit. The phases mirror the CUDA pipeline:
 *   A  pyramid + hierarchical ME + intra/inter choice   (parallel over MBs)
 *   B  inter MBs: MC, transform, quant, reconstruction   (parallel over MBs)
 *   C  intra MBs: prediction from reconstructed neighbours, mode decision, transform, recon (wavefront)
 *   D  MV prediction, P_Skip detection                    (parallel over MBs)
 *   E  in-loop deblocking                                 (wavefront)
 *   F  CAVLC + NAL packaging                              (per slice)
 */
#include "orc_internal.h"
#include "h264_tables.h"
#include <stdlib.h>
#include <string.h>

#define ORC_MB_BINS_MAX 4096   /* bound of one MB's bin list: 384 levels of at most 2 + 14 + 27 + 1 bins (|level| <= 2064) is far above real use; checked */
#define HP_M 4   /* margin of the half-pel planes, see build_halfpel() */
#define T8X8_ON(e) ((e)->cfg.profile == 2 && !(e)->cfg.no_t8x8)   /* PPS transform_8x8_mode_flag */
#define I8X8_ON(e) (T8X8_ON(e) && (e)->cfg.intra8x8)                /* Intra_8x8 on trial (High profile) */

struct OrcEncoder {
    OrcConfig cfg;
    int mbw, mbh, wc, hc;
    uint8_t *src[3], *rec[3], *dbk[3], *ref[3];
    uint8_t *bg_skip;                /* per MB: 1 = skipped as static background (no residual is coded) */
    uint8_t *psrc[3]; int have_psrc, src_committed;   /* source planes of the last committed picture (background detection) */
    uint8_t *srcL1, *srcL2, *refL1, *refL2;
    uint8_t *hpb, *hph, *hpj;        /* half-pel planes of ref luma, (wc+2M) x (hc+2M) */
    OrcMbInfo *mbi; OrcMbCoef *coef;
    int16_t *me[3];                  /* [level][mb*2] */
    int32_t *inter_cost;
    uint8_t *pred_y, *pred_c;        /* per-MB inter prediction: 256 luma, 2*64 chroma */
    int *slice_row0;                 /* the layout of the picture being coded: cur_slices + 1 entries (one of row0_tab) */
    int *row0_tab[2]; int cur_slices;
    int frame_num, idr_pic_id, have_ref, last_idr;
    int since_idr;                   /* pictures since (and including) the last IDR */
    uint8_t *rbsp; int rbsp_cap;
    OrcMbSide *side;                 /* CABAC side records */
    uint16_t *bins; int bins_cap;    /* CABAC bin lists of the last frame, slice after slice */
    int *slice_bin0;                 /* num_slices+1 offsets into bins */
};

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int pxc(const uint8_t *p, int st, int w, int h, int x, int y) { return p[clip3(0, h - 1, y) * st + clip3(0, w - 1, x)]; }
static inline int median3(int a, int b, int c) { int mx = a > b ? a : b, mn = a < b ? a : b; return c > mx ? mx : (c < mn ? mn : c); }

OrcEncoder *orc_create(const OrcConfig *cfg)
{
    OrcEncoder *e = (OrcEncoder *)calloc(1, sizeof *e);
    e->cfg = *cfg;
    /* iComplexityMode (the wrapper asks for HIGH_COMPLEXITY = 2, VideoEncoderOpenH264.cpp:289): 0 LOW drops the Intra_4x4 trial and P_8x8, 1 MEDIUM drops P_8x8 */
    if (e->cfg.complexity_set) { if (e->cfg.complexity <= 1) e->cfg.no_p8x8 = 1; if (e->cfg.complexity == 0) e->cfg.no_i4x4 = 1; }
    if (e->cfg.num_slices < 1) e->cfg.num_slices = 1;
    if (e->cfg.search_range < 4) e->cfg.search_range = 16;
    if (e->cfg.fps <= 0) e->cfg.fps = 30;
    e->mbw = (cfg->width + 15) / 16; e->mbh = (cfg->height + 15) / 16;
    if (e->cfg.num_slices > e->mbh) e->cfg.num_slices = e->mbh;
    e->wc = e->mbw * 16; e->hc = e->mbh * 16;
    size_t ny = (size_t)e->wc * e->hc, nc = ny / 4;
    uint8_t **sets[5] = { e->src, e->rec, e->dbk, e->ref, e->psrc };
    for (int s = 0; s < 5; s++) { sets[s][0] = calloc(1, ny); sets[s][1] = calloc(1, nc); sets[s][2] = calloc(1, nc); }
    e->srcL1 = calloc(1, ny / 4); e->refL1 = calloc(1, ny / 4);
    e->srcL2 = calloc(1, ny / 16); e->refL2 = calloc(1, ny / 16);
    size_t nh = (size_t)(e->wc + 2 * HP_M) * (e->hc + 2 * HP_M);
    e->hpb = malloc(nh); e->hph = malloc(nh); e->hpj = malloc(nh);
    int n = e->mbw * e->mbh;
    e->mbi = calloc(n, sizeof(OrcMbInfo)); e->coef = calloc(n, sizeof(OrcMbCoef)); e->bg_skip = calloc(n, 1);
    for (int l = 0; l < 3; l++) e->me[l] = calloc(n * 2, sizeof(int16_t));
    e->inter_cost = calloc(n, sizeof(int32_t));
    e->pred_y = malloc((size_t)n * 256); e->pred_c = malloc((size_t)n * 128);
    if (e->cfg.key_slices < 1) e->cfg.key_slices = e->cfg.num_slices;
    if (e->cfg.key_slices > e->mbh) e->cfg.key_slices = e->mbh;
    for (int k = 0; k < 2; k++) {     /* slice layouts: [0] P pictures, [1] key pictures -- equal MB-row groups */
        int cnt = k ? e->cfg.key_slices : e->cfg.num_slices, base = e->mbh / cnt, rem = e->mbh % cnt, r = 0;
        e->row0_tab[k] = calloc(cnt + 1, sizeof(int));
        for (int s = 0; s < cnt; s++) { e->row0_tab[k][s] = r; r += base + (s < rem); }
        e->row0_tab[k][cnt] = r;
    }
    e->slice_row0 = e->row0_tab[0]; e->cur_slices = e->cfg.num_slices;
    e->rbsp_cap = n * 1024 + 4096; e->rbsp = malloc(e->rbsp_cap);
    e->side = calloc(n, sizeof(OrcMbSide));
    e->bins_cap = n * ORC_MB_BINS_MAX; e->bins = malloc((size_t)e->bins_cap * 2);
    e->slice_bin0 = calloc((e->cfg.num_slices > e->cfg.key_slices ? e->cfg.num_slices : e->cfg.key_slices) + 1, sizeof(int));
    return e;
}

void orc_destroy(OrcEncoder *e)
{
    if (!e) return;
    uint8_t **sets[5] = { e->src, e->rec, e->dbk, e->ref, e->psrc };
    for (int s = 0; s < 5; s++) for (int c = 0; c < 3; c++) free(sets[s][c]);
    free(e->srcL1); free(e->srcL2); free(e->refL1); free(e->refL2);
    free(e->hpb); free(e->hph); free(e->hpj);
    free(e->mbi); free(e->coef); free(e->bg_skip); for (int l = 0; l < 3; l++) free(e->me[l]);
    free(e->inter_cost); free(e->pred_y); free(e->pred_c); free(e->row0_tab[0]); free(e->row0_tab[1]); free(e->rbsp); free(e->side); free(e->bins); free(e->slice_bin0); free(e);
}

int orc_last_frame_was_idr(const OrcEncoder *e) { return e->last_idr; }
const OrcMbInfo *orc_mb_info(const OrcEncoder *e) { return e->mbi; }
const OrcMbCoef *orc_mb_coef(const OrcEncoder *e) { return e->coef; }
int orc_mb_count(const OrcEncoder *e) { return e->mbw * e->mbh; }
const int16_t *orc_me_level(const OrcEncoder *e, int level) { return e->me[level]; }
const int32_t *orc_inter_cost(const OrcEncoder *e) { return e->inter_cost; }
const uint8_t *orc_plane(const OrcEncoder *e, int which, int comp, int *stride, int *w, int *h)
{
    uint8_t *const *set = which == 0 ? e->src : which == 1 ? e->rec : which == 2 ? e->dbk : e->ref;
    int sh = comp ? 1 : 0;
    if (stride) *stride = e->wc >> sh;
    if (w) *w = e->wc >> sh;
    if (h) *h = e->hc >> sh;
    return set[comp];
}
void orc_get_recon(const OrcEncoder *e, uint8_t *out)
{
    int w = e->cfg.width, h = e->cfg.height;
    for (int y = 0; y < h; y++) memcpy(out + (size_t)y * w, e->dbk[0] + (size_t)y * e->wc, w);
    out += (size_t)w * h;
    for (int c = 1; c < 3; c++, out += (size_t)(w / 2) * (h / 2))
        for (int y = 0; y < h / 2; y++) memcpy(out + (size_t)y * (w / 2), e->dbk[c] + (size_t)y * (e->wc / 2), w / 2);
}

static int slice_of_row(const OrcEncoder *e, int row)
{
    int s = 0; while (row >= e->slice_row0[s + 1]) s++; return s;
}
/* is row `my` the first row of its slice (then no neighbour above is available, 6.4.9) */
static int row_is_slice_top(const OrcEncoder *e, int my) { return e->slice_row0[slice_of_row(e, my)] == my; }

/* ---- input: copy the display-size I420 frame into the coded-size planes, replicating the last
 * column/row into the padding (the wrapper hands tightly packed I420, VideoEncoderOpenH264.cpp:354-365) ---- */
static void load_source(OrcEncoder *e, const uint8_t *in)
{
    int w = e->cfg.width, h = e->cfg.height;
    /* the planes of the last COMMITTED picture become the previous source (a trial is overwritten, not remembered) */
    if (e->src_committed) { for (int c = 0; c < 3; c++) { uint8_t *t = e->src[c]; e->src[c] = e->psrc[c]; e->psrc[c] = t; } e->have_psrc = 1; e->src_committed = 0; }
    for (int c = 0; c < 3; c++) {
        int pw = c ? w / 2 : w, ph = c ? h / 2 : h, cw = c ? e->wc / 2 : e->wc, ch = c ? e->hc / 2 : e->hc;
        for (int y = 0; y < ch; y++) {
            const uint8_t *s = in + (size_t)(y < ph ? y : ph - 1) * pw;
            uint8_t *d = e->src[c] + (size_t)y * cw;
            memcpy(d, s, pw);
            for (int x = pw; x < cw; x++) d[x] = s[pw - 1];
        }
        in += (size_t)pw * ph;
    }
}

/* ---- half-pel planes of the reference luma (8.4.2.2.1), margin HP_M with clamped sources ---- */
static inline int tap6(int a, int b, int c, int d, int e_, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e_ + f; }
static void build_halfpel(OrcEncoder *e)
{
    const uint8_t *r = e->ref[0]; int w = e->wc, h = e->hc, st = e->wc, ps = w + 2 * HP_M;
    for (int y = -HP_M; y < h + HP_M; y++)
        for (int x = -HP_M; x < w + HP_M; x++) {
            int braw[6];
            for (int k = 0; k < 6; k++) {
                int yy = y - 2 + k;
                braw[k] = tap6(pxc(r, st, w, h, x - 2, yy), pxc(r, st, w, h, x - 1, yy), pxc(r, st, w, h, x, yy),
                               pxc(r, st, w, h, x + 1, yy), pxc(r, st, w, h, x + 2, yy), pxc(r, st, w, h, x + 3, yy));
            }
            size_t o = (size_t)(y + HP_M) * ps + x + HP_M;
            e->hpb[o] = (uint8_t)clip255((braw[2] + 16) >> 5);
            e->hph[o] = (uint8_t)clip255((tap6(pxc(r, st, w, h, x, y - 2), pxc(r, st, w, h, x, y - 1), pxc(r, st, w, h, x, y),
                                               pxc(r, st, w, h, x, y + 1), pxc(r, st, w, h, x, y + 2), pxc(r, st, w, h, x, y + 3)) + 16) >> 5);
            e->hpj[o] = (uint8_t)clip255((tap6(braw[0], braw[1], braw[2], braw[3], braw[4], braw[5]) + 512) >> 10);
        }
}
static inline int hp(const OrcEncoder *e, const uint8_t *pl, int x, int y)
{
    return pl[(size_t)(clip3(-HP_M, e->hc + HP_M - 1, y) + HP_M) * (e->wc + 2 * HP_M) + clip3(-HP_M, e->wc + HP_M - 1, x) + HP_M];
}
/* same value as orc_interp_luma(ref, ...) -- tests/test_oracle.py::test_qpel_planes_equal_the_normative_interpolation checks the identity */
static int qpel_sample(const OrcEncoder *e, int xq, int yq)
{
    int x = xq >> 2, y = yq >> 2, w = e->wc, h = e->hc; const uint8_t *r = e->ref[0];
#define G_ pxc(r, w, w, h, x, y)
#define H_ pxc(r, w, w, h, x + 1, y)
#define M_ pxc(r, w, w, h, x, y + 1)
#define b_ hp(e, e->hpb, x, y)
#define s_ hp(e, e->hpb, x, y + 1)
#define h_ hp(e, e->hph, x, y)
#define m_ hp(e, e->hph, x + 1, y)
#define j_ hp(e, e->hpj, x, y)
    switch ((yq & 3) * 4 + (xq & 3)) {
    case 0:  return G_;
    case 1:  return (G_ + b_ + 1) >> 1;
    case 2:  return b_;
    case 3:  return (H_ + b_ + 1) >> 1;
    case 4:  return (G_ + h_ + 1) >> 1;
    case 5:  return (b_ + h_ + 1) >> 1;
    case 6:  return (b_ + j_ + 1) >> 1;
    case 7:  return (b_ + m_ + 1) >> 1;
    case 8:  return h_;
    case 9:  return (h_ + j_ + 1) >> 1;
    case 10: return j_;
    case 11: return (j_ + m_ + 1) >> 1;
    case 12: return (M_ + h_ + 1) >> 1;
    case 13: return (h_ + s_ + 1) >> 1;
    case 14: return (j_ + s_ + 1) >> 1;
    default: return (m_ + s_ + 1) >> 1;
    }
}
int orc_dbg_qpel(const OrcEncoder *e, int xq, int yq) { return qpel_sample(e, xq, yq); }
void orc_dbg_build_halfpel(OrcEncoder *e) { build_halfpel(e); }

/* luma prediction of a bw x bh block whose top-left sample is (x0,y0), written into a 16-wide buffer */
static void mc_luma_blk(const OrcEncoder *e, int x0, int y0, int bw, int bh, int mvx, int mvy, uint8_t *dst /*stride 16*/)
{
    for (int y = 0; y < bh; y++)
        for (int x = 0; x < bw; x++) dst[y * 16 + x] = (uint8_t)qpel_sample(e, 4 * (x0 + x) + mvx, 4 * (y0 + y) + mvy);
}
static void mc_luma(const OrcEncoder *e, int x0, int y0, int mvx, int mvy, uint8_t *dst /*16x16*/) { mc_luma_blk(e, x0, y0, 16, 16, mvx, mvy, dst); }
static void mc_chroma_blk(const OrcEncoder *e, int comp, int cx0, int cy0, int bw, int bh, int mvx, int mvy, uint8_t *dst /*stride 8*/)
{
    for (int y = 0; y < bh; y++)
        for (int x = 0; x < bw; x++)
            dst[y * 8 + x] = (uint8_t)orc_interp_chroma(e->ref[comp], e->wc / 2, e->wc / 2, e->hc / 2,
                                                        8 * (cx0 + x) + mvx, 8 * (cy0 + y) + mvy);
}

#define ORC_P8X8_BIAS_BITS 8   /* extra header bits of P_8x8 over P_L0_16x16: mb_type ue(3) vs ue(0), four sub_mb_type ue(0) */
static int mv_bits(int mvx, int mvy) { return orc_se_len(mvx) + orc_se_len(mvy); }

/* SAD of a bw x bh block with clamped coordinates on both planes (same dimensions w x h) */
static int sad_clamped(const uint8_t *a, const uint8_t *b, int w, int h, int ax, int ay, int bx, int by, int bw, int bh)
{
    int s = 0;
    for (int y = 0; y < bh; y++)
        for (int x = 0; x < bw; x++) s += iabs(pxc(a, w, w, h, ax + x, ay + y) - pxc(b, w, w, h, bx + x, by + y));
    return s;
}

/* ---- Phase A: hierarchical motion search for one MB (role of WelsMotionEstimateSearch + MeRefineFracPixel).
 * Level 2 (1/4 res): 8x8 block centred on the MB, exhaustive +-R/4. Level 1 (1/2 res): 8x8, +-2 around 2*mv2.
 * Level 0: 16x16, +-2 around 2*mv1 plus the zero vector. Then 8 half-pel and 8 quarter-pel neighbours by SATD.
 * Every level picks argmin of key = (cost << k) | candidate_index, so ties resolve identically everywhere. ---- */
static int quant_dc_inter(int y, int qp);
/* Early skip test: would the macroblock, predicted from the reference at the ZERO vector, quantise to all-zero levels in
 * luma and chroma (same transform / inter dead zone as code_inter_mb and code_chroma)? */
static int zero_vector_residual_vanishes(const OrcEncoder *e, int mx, int my, int qp)
{
    int st = e->wc, cs = st / 2, qpc = CHROMA_QP[qp];
    const uint8_t *s = e->src[0] + (size_t)my * 16 * st + mx * 16, *r = e->ref[0] + (size_t)my * 16 * st + mx * 16;
    for (int b = 0; b < 16; b++) {
        int bx = BLK_X[b] * 4, by = BLK_Y[b] * 4; int16_t res[16], c[16], lz[16];
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) res[y * 4 + x] = (int16_t)(s[(by + y) * st + bx + x] - r[(by + y) * st + bx + x]);
        orc_dct4x4(res, c);
        if (orc_quant4x4(c, lz, qp, 0, 0)) return 0;
    }
    for (int pl = 0; pl < 2; pl++) {
        const uint8_t *sc = e->src[1 + pl] + (size_t)my * 8 * cs + mx * 8, *rc = e->ref[1 + pl] + (size_t)my * 8 * cs + mx * 8;
        int dc[4];
        for (int b = 0; b < 4; b++) {
            int bx = (b & 1) * 4, by = (b >> 1) * 4; int16_t res[16], c[16], lz[16];
            for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) res[y * 4 + x] = (int16_t)(sc[(by + y) * cs + bx + x] - rc[(by + y) * cs + bx + x]);
            orc_dct4x4(res, c); dc[b] = c[0];
            if (orc_quant4x4(c, lz, qpc, 0, 1)) return 0;
        }
        int h[4] = { dc[0] + dc[1] + dc[2] + dc[3], dc[0] - dc[1] + dc[2] - dc[3], dc[0] + dc[1] - dc[2] - dc[3], dc[0] - dc[1] - dc[2] + dc[3] };
        for (int i = 0; i < 4; i++) if (quant_dc_inter(h[i], qpc)) return 0;
    }
    return 1;
}

/* Background detection (the wrapper sets bEnableBackgroundDetection = 1, VideoEncoderOpenH264.cpp:282; openh264's pre-analysis compares the
 * current with the previous SOURCE picture per 8x8 unit and lets static macroblocks be skipped). OUR definition: a macroblock
 * is static background when, against the previous source picture at the same position, every 8x8 unit of luma and both 8x8 chroma blocks
 * have SAD <= ORC_BGD_OU_SAD and no sample differs by more than ORC_BGD_MAXDIFF, and its luma SAD against the REFERENCE at the zero
 * vector is at most 64 * (8 + lambda) -- what bounds the error a run of skips can accumulate. Such a macroblock is skipped like one
 * whose residual quantises to nothing. */
static int mb_is_static_background(const OrcEncoder *e, int mx, int my, int lambda)
{
    if (!e->cfg.background_detection || !e->have_psrc) return 0;
    int st = e->wc, cs = st / 2;
    const uint8_t *s = e->src[0] + (size_t)my * 16 * st + mx * 16, *p = e->psrc[0] + (size_t)my * 16 * st + mx * 16;
    const uint8_t *r = e->ref[0] + (size_t)my * 16 * st + mx * 16;
    int zsad = 0;
    for (int q = 0; q < 4; q++) {
        int sad = 0, o = (q >> 1) * 8 * st + (q & 1) * 8;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
            int d = iabs(s[o + y * st + x] - p[o + y * st + x]);
            if (d > ORC_BGD_MAXDIFF) return 0;
            sad += d; zsad += iabs(s[o + y * st + x] - r[o + y * st + x]);
        }
        if (sad > ORC_BGD_OU_SAD) return 0;
    }
    for (int pl = 0; pl < 2; pl++) {
        const uint8_t *sc = e->src[1 + pl] + (size_t)my * 8 * cs + mx * 8, *pc = e->psrc[1 + pl] + (size_t)my * 8 * cs + mx * 8;
        int sad = 0;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
            int d = iabs(sc[y * cs + x] - pc[y * cs + x]);
            if (d > ORC_BGD_MAXDIFF) return 0;
            sad += d;
        }
        if (sad > ORC_BGD_OU_SAD) return 0;
    }
    return zsad <= 64 * (8 + lambda);
}

static void motion_search(OrcEncoder *e, int mx, int my, int lambda, int qp)
{
    int mb = my * e->mbw + mx, R4 = e->cfg.search_range / 4, span = 2 * R4 + 1;
    int w2 = e->wc / 4, h2 = e->hc / 4, w1 = e->wc / 2, h1 = e->hc / 2;
    uint32_t best = 0xffffffffu; int bx = 0, by = 0;
    for (int dy = -R4; dy <= R4; dy++)
        for (int dx = -R4; dx <= R4; dx++) {
            int c = sad_clamped(e->srcL2, e->refL2, w2, h2, 4 * mx - 2, 4 * my - 2, 4 * mx - 2 + dx, 4 * my - 2 + dy, 8, 8) + iabs(dx) + iabs(dy);
            uint32_t key = ((uint32_t)c << 11) | (uint32_t)((dy + R4) * span + dx + R4);
            if (key < best) { best = key; bx = dx; by = dy; }
        }
    e->me[2][mb * 2] = (int16_t)bx; e->me[2][mb * 2 + 1] = (int16_t)by;
    int cx = 2 * bx, cy = 2 * by; best = 0xffffffffu;
    for (int dy = -2; dy <= 2; dy++)
        for (int dx = -2; dx <= 2; dx++) {
            int vx = cx + dx, vy = cy + dy;
            int c = sad_clamped(e->srcL1, e->refL1, w1, h1, 8 * mx, 8 * my, 8 * mx + vx, 8 * my + vy, 8, 8) + iabs(vx) + iabs(vy);
            uint32_t key = ((uint32_t)c << 5) | (uint32_t)((dy + 2) * 5 + dx + 2);
            if (key < best) { best = key; bx = vx; by = vy; }
        }
    e->me[1][mb * 2] = (int16_t)bx; e->me[1][mb * 2 + 1] = (int16_t)by;
    /* Estimate of this MB's MV predictor for the rate term of levels 0 and below: the 8.4.1.3 median taken over the LEVEL-1
     * vectors of the left, upper and upper-right (else upper-left) neighbours, scaled to quarter-pel (x8). It needs no final
     * vector of any neighbour, so the search stays independent per MB; missing neighbours count as zero vectors. */
    int ppx, ppy;
    {
        int top = !row_is_slice_top(e, my), ax = 0, ay = 0, tx = 0, ty = 0, rx = 0, ry = 0;
        const int16_t *m1 = e->me[1];
        if (mx > 0) { ax = m1[(mb - 1) * 2]; ay = m1[(mb - 1) * 2 + 1]; }
        if (top) { tx = m1[(mb - e->mbw) * 2]; ty = m1[(mb - e->mbw) * 2 + 1]; }
        if (top && mx + 1 < e->mbw) { rx = m1[(mb - e->mbw + 1) * 2]; ry = m1[(mb - e->mbw + 1) * 2 + 1]; }
        else if (top && mx > 0) { rx = m1[(mb - e->mbw - 1) * 2]; ry = m1[(mb - e->mbw - 1) * 2 + 1]; }
        ppx = 8 * median3(ax, tx, rx); ppy = 8 * median3(ay, ty, ry);
    }
    /* EARLY SKIP: with a zero predictor estimate, a macroblock whose zero-vector residual quantises to nothing is
     * final -- P_L0_16x16, vector (0,0), cbp 0 (P_Skip follows in phase D when the normative skip vector is zero too). Static
     * screen content never enters the search. */
    e->bg_skip[mb] = 0;
    if (ppx == 0 && ppy == 0 && (zero_vector_residual_vanishes(e, mx, my, qp) || (e->bg_skip[mb] = (uint8_t)mb_is_static_background(e, mx, my, lambda)))) {
        OrcMbInfo *mi0 = &e->mbi[mb];
        mi0->mb_type = ORC_MB_P16x16; e->me[0][mb * 2] = e->me[0][mb * 2 + 1] = 0; e->inter_cost[mb] = 0;
        return;                              /* mbi was zeroed by the caller: mv = mv8 = 0 */
    }
    cx = 2 * bx; cy = 2 * by; best = 0xffffffffu;
    for (int i = 0; i < 26; i++) {
        int vx = i < 25 ? cx + i % 5 - 2 : 0, vy = i < 25 ? cy + i / 5 - 2 : 0;
        int c = sad_clamped(e->src[0], e->ref[0], e->wc, e->hc, 16 * mx, 16 * my, 16 * mx + vx, 16 * my + vy, 16, 16)
              + lambda * mv_bits(4 * vx - ppx, 4 * vy - ppy);
        uint32_t key = ((uint32_t)c << 5) | (uint32_t)i;
        if (key < best) { best = key; bx = vx; by = vy; }
    }
    e->me[0][mb * 2] = (int16_t)bx; e->me[0][mb * 2 + 1] = (int16_t)by;
    /* sub-pel: centre + 8 neighbours at step 2 (half), then at step 1 (quarter). Every candidate's SATD is also kept per 8x8
     * quadrant: each quadrant remembers its own best candidate (key = (SATD8x8 + lambda*mv bits) << 5 | candidate sequence
     * number), which is the motion search of the four P_L0_8x8 partitions at no extra SATD work. */
    static const int8_t OX[9] = { 0, -1, 0, 1, -1, 1, -1, 0, 1 }, OY[9] = { 0, -1, -1, -1, 0, 0, 1, 1, 1 };
    int qx = 4 * bx, qy = 4 * by; uint8_t pred[256];
    const uint8_t *s = e->src[0] + (size_t)my * 16 * e->wc + mx * 16;
    uint32_t centre_key = 0, bestq[4] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu }; int mvq[4][2] = { { 0 } };
    for (int step = 2; step >= 1; step--) {
        best = 0xffffffffu; int nx = qx, ny = qy;
        for (int i = 0; i < 9; i++) {
            uint32_t key;
            if (i == 0 && step == 1) key = centre_key & ~15u;
            else {
                int vx = qx + step * OX[i], vy = qy + step * OY[i], bits = mv_bits(vx - ppx, vy - ppy), tot = 0;
                mc_luma(e, 16 * mx, 16 * my, vx, vy, pred);
                for (int q = 0; q < 4; q++) {
                    int o = (q >> 1) * 8 * e->wc + (q & 1) * 8, po = (q >> 1) * 8 * 16 + (q & 1) * 8;
                    int sq = orc_satd4x4(s + o, e->wc, pred + po, 16) + orc_satd4x4(s + o + 4, e->wc, pred + po + 4, 16)
                           + orc_satd4x4(s + o + 4 * e->wc, e->wc, pred + po + 64, 16) + orc_satd4x4(s + o + 4 * e->wc + 4, e->wc, pred + po + 68, 16);
                    uint32_t kq = ((uint32_t)(sq + lambda * bits) << 5) | (uint32_t)(step == 2 ? i : 8 + i);
                    if (kq < bestq[q]) { bestq[q] = kq; mvq[q][0] = vx; mvq[q][1] = vy; }
                    tot += sq;
                }
                key = ((uint32_t)(tot + lambda * bits) << 4) | (uint32_t)i;
            }
            if (key < best) { best = key; nx = qx + step * OX[i]; ny = qy + step * OY[i]; }
        }
        qx = nx; qy = ny; centre_key = best;
    }
    OrcMbInfo *mi = &e->mbi[mb];
    int cost16 = (int)(best >> 4), cost8 = lambda * ORC_P8X8_BIAS_BITS;
    for (int q = 0; q < 4; q++) cost8 += (int)(bestq[q] >> 5);
    mi->mv[0] = (int16_t)qx; mi->mv[1] = (int16_t)qy;
    if (cost8 < cost16 && !e->cfg.no_p8x8) {
        mi->mb_type = ORC_MB_P8x8; e->inter_cost[mb] = cost8;
        for (int q = 0; q < 4; q++) { mi->mv8[q][0] = (int16_t)mvq[q][0]; mi->mv8[q][1] = (int16_t)mvq[q][1]; }
        mi->mv[0] = mi->mv8[0][0]; mi->mv[1] = mi->mv8[0][1];
    } else {
        mi->mb_type = ORC_MB_P16x16; e->inter_cost[mb] = cost16;
        for (int q = 0; q < 4; q++) { mi->mv8[q][0] = (int16_t)qx; mi->mv8[q][1] = (int16_t)qy; }
    }
}

/* ---- Phase A: intra estimate from SOURCE neighbours (V/H/DC 16x16), used only to choose intra vs inter
 * in P frames without waiting for reconstructed neighbours. ---- */
static int intra_estimate(const OrcEncoder *e, int mx, int my)
{
    const uint8_t *s = e->src[0] + (size_t)my * 16 * e->wc + mx * 16; int st = e->wc;
    int top = !row_is_slice_top(e, my), left = mx > 0; uint8_t pred[256]; int best = 1 << 30;
    if (top) { for (int y = 0; y < 16; y++) memcpy(pred + y * 16, s - st, 16);
               int c = orc_satd16x16(s, st, pred, 16); if (c < best) best = c; }
    if (left) { for (int y = 0; y < 16; y++) memset(pred + y * 16, s[y * st - 1], 16);
                int c = orc_satd16x16(s, st, pred, 16); if (c < best) best = c; }
    int sum = 0, dc;
    if (top) for (int x = 0; x < 16; x++) sum += s[x - st];
    if (left) for (int y = 0; y < 16; y++) sum += s[y * st - 1];
    dc = top && left ? (sum + 16) >> 5 : (top || left) ? (sum + 8) >> 4 : 128;
    memset(pred, dc, 256);
    { int c = orc_satd16x16(s, st, pred, 16); if (c < best) best = c; }
    return best;
}
#define ORC_INTRA_BIAS_BITS 16

/* ---- transforms on DC terms ---- */
static void hadamard4x4(const int *in, int *out)   /* H X H, no scaling */
{
    int t[16];
    for (int y = 0; y < 4; y++) {
        int a0 = in[y * 4] + in[y * 4 + 3], a1 = in[y * 4 + 1] + in[y * 4 + 2];
        int a2 = in[y * 4 + 1] - in[y * 4 + 2], a3 = in[y * 4] - in[y * 4 + 3];
        t[y * 4] = a0 + a1; t[y * 4 + 1] = a3 + a2; t[y * 4 + 2] = a0 - a1; t[y * 4 + 3] = a3 - a2;
    }
    for (int x = 0; x < 4; x++) {
        int a0 = t[x] + t[12 + x], a1 = t[4 + x] + t[8 + x], a2 = t[4 + x] - t[8 + x], a3 = t[x] - t[12 + x];
        out[x] = a0 + a1; out[4 + x] = a3 + a2; out[8 + x] = a0 - a1; out[12 + x] = a3 - a2;
    }
}
static int quant_dc(int y, int qp)     /* (|y|*MF0 + 2f) >> (qbits+1), intra dead zone */
{
    int qbits = 15 + qp / 6, f = (1 << qbits) / 3;
    int l = (int)(((int64_t)iabs(y) * QUANT_MF[qp % 6][0] + 2 * f) >> (qbits + 1));
    if (l > 2063) l = 2063;
    return y < 0 ? -l : l;
}
static int quant_dc_inter(int y, int qp)
{
    int qbits = 15 + qp / 6, f = (1 << qbits) / 6;
    int l = (int)(((int64_t)iabs(y) * QUANT_MF[qp % 6][0] + 2 * f) >> (qbits + 1));
    if (l > 2063) l = 2063;
    return y < 0 ? -l : l;
}

/* chroma residual for one MB (both planes): fills coef, nnz, returns chroma cbp (8.5.11 on the decode side) */
static int code_chroma(OrcEncoder *e, int mx, int my, const uint8_t *pred /*2 x 64*/, int qp, int intra)
{
    int mb = my * e->mbw + mx, qpc = CHROMA_QP[qp], cs = e->wc / 2, any_dc = 0, any_ac = 0;
    OrcMbCoef *co = &e->coef[mb]; OrcMbInfo *mi = &e->mbi[mb];
    for (int pl = 0; pl < 2; pl++) {
        const uint8_t *s = e->src[1 + pl] + (size_t)my * 8 * cs + mx * 8;
        uint8_t *r = e->rec[1 + pl] + (size_t)my * 8 * cs + mx * 8;
        const uint8_t *p = pred + pl * 64;
        int16_t c[4][16]; int dc[4];
        for (int b = 0; b < 4; b++) {
            int bx = (b & 1) * 4, by = (b >> 1) * 4; int16_t res[16];
            for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
                res[y * 4 + x] = (int16_t)(s[(by + y) * cs + bx + x] - p[(by + y) * 8 + bx + x]);
            orc_dct4x4(res, c[b]); dc[b] = c[b][0];
            int n = orc_quant4x4(c[b], co->chroma_ac[pl][b], qpc, intra, 1);
            mi->nnz[16 + pl * 4 + b] = (uint8_t)n; any_ac |= n;
        }
        int h[4] = { dc[0] + dc[1] + dc[2] + dc[3], dc[0] - dc[1] + dc[2] - dc[3], dc[0] + dc[1] - dc[2] - dc[3], dc[0] - dc[1] - dc[2] + dc[3] };
        int l[4];
        for (int i = 0; i < 4; i++) { l[i] = intra ? quant_dc(h[i], qpc) : quant_dc_inter(h[i], qpc); co->chroma_dc[pl][i] = (int16_t)l[i]; any_dc |= l[i]; }
        /* 8.5.11.1/2: inverse 2x2 transform and scaling of chroma DC */
        int f[4] = { l[0] + l[1] + l[2] + l[3], l[0] - l[1] + l[2] - l[3], l[0] + l[1] - l[2] - l[3], l[0] - l[1] - l[2] + l[3] };
        for (int b = 0; b < 4; b++) {
            int bx = (b & 1) * 4, by = (b >> 1) * 4; int32_t d[16], rr[16];
            orc_dequant4x4(co->chroma_ac[pl][b], d, qpc, 1);
            d[0] = ((f[b] * 16 * DEQUANT_V[qpc % 6][0]) << (qpc / 6)) >> 5;
            orc_idct4x4(d, rr);
            for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
                r[(by + y) * cs + bx + x] = (uint8_t)clip255(p[(by + y) * 8 + bx + x] + rr[y * 4 + x]);
        }
    }
    return any_ac ? 2 : (any_dc ? 1 : 0);
}

/* rate estimate of one residual block in half bits: per nonzero level 2 * (3 + min(|level|, 16)) (significance, last, sign, unary
 * magnitude), 1 per zero before the last nonzero level (a significance flag of ~half a bit) */
static int level_cost2(const int16_t *lv, int n)
{
    int c = 0, last = -1, nz = 0;
    for (int i = 0; i < n; i++) { int a = iabs(lv[i]); if (a) { c += 2 * (3 + (a < 16 ? a : 16)); last = i; nz++; } }
    return c + (last + 1 - nz);
}
/* ---- Phase B: one inter MB ---- */
static void code_inter_mb(OrcEncoder *e, int mx, int my, int qp)
{
    int mb = my * e->mbw + mx, st = e->wc; OrcMbInfo *mi = &e->mbi[mb]; OrcMbCoef *co = &e->coef[mb];
    uint8_t *py = e->pred_y + (size_t)mb * 256, *pc = e->pred_c + (size_t)mb * 128;
    for (int q = 0; q < 4; q++) {      /* one vector per 8x8 partition (all equal for P_L0_16x16) */
        int ox = (q & 1) * 8, oy = (q >> 1) * 8, vx = mi->mv8[q][0], vy = mi->mv8[q][1];
        mc_luma_blk(e, 16 * mx + ox, 16 * my + oy, 8, 8, vx, vy, py + oy * 16 + ox);
        mc_chroma_blk(e, 1, 8 * mx + ox / 2, 8 * my + oy / 2, 4, 4, vx, vy, pc + (oy / 2) * 8 + ox / 2);
        mc_chroma_blk(e, 2, 8 * mx + ox / 2, 8 * my + oy / 2, 4, 4, vx, vy, pc + 64 + (oy / 2) * 8 + ox / 2);
    }
    const uint8_t *s = e->src[0] + (size_t)my * 16 * st + mx * 16; uint8_t *r = e->rec[0] + (size_t)my * 16 * st + mx * 16;
    int cbp = 0;
    memset(co, 0, sizeof *co);
    if (e->bg_skip[mb]) {              /* static background: the prediction (the reference at the zero vector) IS the reconstruction */
        int cs = st / 2;
        for (int y = 0; y < 16; y++) memcpy(r + y * st, py + y * 16, 16);
        for (int pl = 0; pl < 2; pl++) for (int y = 0; y < 8; y++) memcpy(e->rec[1 + pl] + (size_t)(my * 8 + y) * cs + mx * 8, pc + pl * 64 + y * 8, 8);
        memset(mi->nnz, 0, sizeof mi->nnz); mi->cbp = 0;
        return;
    }
    for (int b = 0; b < 16; b++) {
        int bx = BLK_X[b] * 4, by = BLK_Y[b] * 4; int16_t res[16], c[16]; int32_t d[16], rr[16];
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) res[y * 4 + x] = (int16_t)(s[(by + y) * st + bx + x] - py[(by + y) * 16 + bx + x]);
        orc_dct4x4(res, c);
        int n = orc_quant4x4(c, co->luma[b], qp, 0, 0);
        mi->nnz[b] = (uint8_t)n; if (n) cbp |= 1 << (b >> 2);
        orc_dequant4x4(co->luma[b], d, qp, 0); orc_idct4x4(d, rr);
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) r[(by + y) * st + bx + x] = (uint8_t)clip255(py[(by + y) * 16 + bx + x] + rr[y * 4 + x]);
    }
    /* High profile: the luma residual is coded again with the 8x8 transform (four 8x8 blocks, 8x8 zig-zag levels in luma[4*b8..]) and
     * the macroblock takes transform_size_8x8_flag = 1 when that wins the rate-distortion comparison J = 64 SSD + 27 lambda^2 B
     * (lambda_mode = 27/32 lambda^2, B in half bits), at least one 8x8 level being nonzero. */
    if (T8X8_ON(e) && cbp) {
        int16_t l8[4][64]; int n8[4], b4 = 0, b8_ = 0, any8 = 0; int64_t d4 = 0, d8 = 0; uint8_t r8[256];
        for (int b = 0; b < 16; b++) if (cbp & (1 << (b >> 2))) b4 += 1 + level_cost2(co->luma[b], 16);
        for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) { int d = s[y * st + x] - r[y * st + x]; d4 += d * d; }
        for (int b8 = 0; b8 < 4; b8++) {
            int ox = (b8 & 1) * 8, oy = (b8 >> 1) * 8; int16_t res[64]; int32_t c[64], d[64], rr[64];
            for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) res[y * 8 + x] = (int16_t)(s[(oy + y) * st + ox + x] - py[(oy + y) * 16 + ox + x]);
            orc_dct8x8(res, c);
            n8[b8] = orc_quant8x8(c, l8[b8], qp, 0); any8 |= n8[b8];
            b8_ += level_cost2(l8[b8], 64);
            orc_dequant8x8(l8[b8], d, qp); orc_idct8x8(d, rr);
            for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
                int v = clip255(py[(oy + y) * 16 + ox + x] + rr[y * 8 + x]), dd = s[(oy + y) * st + ox + x] - v;
                r8[(oy + y) * 16 + ox + x] = (uint8_t)v; d8 += dd * dd;
            }
        }
        int64_t l2 = 27 * (int64_t)LAMBDA_TAB[qp] * LAMBDA_TAB[qp];
        if (any8 && 64 * d8 + l2 * (b8_ + 4) < 64 * d4 + l2 * b4) {   /* + 2 bits: the flag's minority value */
            cbp = 0;
            for (int b8 = 0; b8 < 4; b8++) {
                memcpy(co->luma[4 * b8], l8[b8], sizeof l8[b8]);
                for (int k = 0; k < 4; k++) mi->nnz[4 * b8 + k] = (uint8_t)n8[b8];
                if (n8[b8]) cbp |= 1 << b8;
            }
            for (int y = 0; y < 16; y++) memcpy(r + y * st, r8 + y * 16, 16);
            mi->i16_mode = 4;      /* transform_size_8x8_flag */
        }
    }
    cbp |= code_chroma(e, mx, my, pc, qp, 0) << 4;
    mi->cbp = (uint8_t)cbp;
}

/* ---- intra predictors, 8.3.3 (Intra_16x16) and 8.3.4 (chroma), from the pre-deblock reconstruction ---- */
static void pred_i16(const uint8_t *r, int st, int mode, int top, int left, uint8_t *p)
{
    if (mode == 0) { for (int y = 0; y < 16; y++) memcpy(p + y * 16, r - st, 16); }
    else if (mode == 1) { for (int y = 0; y < 16; y++) memset(p + y * 16, r[y * st - 1], 16); }
    else if (mode == 2) {
        int sum = 0;
        if (top) for (int x = 0; x < 16; x++) sum += r[x - st];
        if (left) for (int y = 0; y < 16; y++) sum += r[y * st - 1];
        memset(p, top && left ? (sum + 16) >> 5 : (top || left) ? (sum + 8) >> 4 : 128, 256);
    } else {
        int H = 0, V = 0;
        for (int i = 0; i < 8; i++) { H += (i + 1) * (r[8 + i - st] - r[6 - i - st]); V += (i + 1) * (r[(8 + i) * st - 1] - r[(6 - i) * st - 1]); }
        int a = 16 * (r[15 * st - 1] + r[15 - st]), b = (5 * H + 32) >> 6, c = (5 * V + 32) >> 6;
        for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) p[y * 16 + x] = (uint8_t)clip255((a + b * (x - 7) + c * (y - 7) + 16) >> 5);
    }
}
static void pred_chroma(const uint8_t *r, int st, int mode, int top, int left, uint8_t *p /*8x8*/)
{
    if (mode == 0) {
        for (int b = 0; b < 4; b++) {
            int bx = (b & 1) * 4, by = (b >> 1) * 4, st_ = 0, sl = 0, v;
            if (top) for (int x = 0; x < 4; x++) st_ += r[bx + x - st];
            if (left) for (int y = 0; y < 4; y++) sl += r[(by + y) * st - 1];
            if (b == 0 || b == 3) v = top && left ? (st_ + sl + 4) >> 3 : top ? (st_ + 2) >> 2 : left ? (sl + 2) >> 2 : 128;
            else if (b == 1) v = top ? (st_ + 2) >> 2 : left ? (sl + 2) >> 2 : 128;
            else v = left ? (sl + 2) >> 2 : top ? (st_ + 2) >> 2 : 128;
            for (int y = 0; y < 4; y++) memset(p + (by + y) * 8 + bx, v, 4);
        }
    } else if (mode == 1) { for (int y = 0; y < 8; y++) memset(p + y * 8, r[y * st - 1], 8); }
    else if (mode == 2) { for (int y = 0; y < 8; y++) memcpy(p + y * 8, r - st, 8); }
    else {
        int H = 0, V = 0;
        for (int i = 0; i < 4; i++) { H += (i + 1) * (r[4 + i - st] - r[2 - i - st]); V += (i + 1) * (r[(4 + i) * st - 1] - r[(2 - i) * st - 1]); }
        int a = 16 * (r[7 * st - 1] + r[7 - st]), b = (34 * H + 32) >> 6, c = (34 * V + 32) >> 6;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) p[y * 8 + x] = (uint8_t)clip255((a + b * (x - 3) + c * (y - 3) + 16) >> 5);
    }
}
static int satd8x8(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
    return orc_satd4x4(a, sa, b, sb) + orc_satd4x4(a + 4, sa, b + 4, sb) + orc_satd4x4(a + 4 * sa, sa, b + 4 * sb, sb) + orc_satd4x4(a + 4 * sa + 4, sa, b + 4 * sb + 4, sb);
}

/* ---- Intra_4x4 prediction, 8.3.1.2.1-9. r points at the block's top-left sample inside the (pre-deblock) reconstruction;
 * avail: bit0 top row, bit1 left column, bit2 corner, bit3 top-right (when top is available but top-right is not,
 * p[4..7,-1] are substituted by p[3,-1], 8.3.1.2). Returns 0 when the mode's neighbours are missing. ---- */
#define I4_AV_T 1
#define I4_AV_L 2
#define I4_AV_X 4
#define I4_AV_TR 8
static int pred_i4(const uint8_t *r, int st, int mode, int avail, uint8_t *p /*16*/)
{
    int T[8], L[4], X = 128, top = avail & I4_AV_T, left = avail & I4_AV_L, corner = avail & I4_AV_X;
    for (int i = 0; i < 8; i++) T[i] = 128;
    for (int i = 0; i < 4; i++) L[i] = 128;
    if (top) { for (int i = 0; i < 4; i++) T[i] = r[i - st]; for (int i = 4; i < 8; i++) T[i] = (avail & I4_AV_TR) ? r[i - st] : T[3]; }
    if (left) for (int i = 0; i < 4; i++) L[i] = r[i * st - 1];
    if (corner) X = r[-st - 1];
#define PT(i) ((i) < 0 ? X : T[i])      /* p[i,-1] */
#define PL(i) ((i) < 0 ? X : L[i])      /* p[-1,i] */
    switch (mode) {
    case 0: if (!top) return 0; for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) p[y * 4 + x] = (uint8_t)T[x]; return 1;
    case 1: if (!left) return 0; for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) p[y * 4 + x] = (uint8_t)L[y]; return 1;
    case 2: {
        int sT = T[0] + T[1] + T[2] + T[3], sL = L[0] + L[1] + L[2] + L[3];
        int v = top && left ? (sT + sL + 4) >> 3 : top ? (sT + 2) >> 2 : left ? (sL + 2) >> 2 : 128;
        memset(p, v, 16); return 1; }
    case 3: if (!top) return 0;
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
            p[y * 4 + x] = (uint8_t)(x == 3 && y == 3 ? (T[6] + 3 * T[7] + 2) >> 2 : (T[x + y] + 2 * T[x + y + 1] + T[x + y + 2] + 2) >> 2);
        return 1;
    case 4: if (!(top && left && corner)) return 0;
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
            p[y * 4 + x] = (uint8_t)(x > y ? (PT(x - y - 2) + 2 * PT(x - y - 1) + PT(x - y) + 2) >> 2
                                   : x < y ? (PL(y - x - 2) + 2 * PL(y - x - 1) + PL(y - x) + 2) >> 2 : (T[0] + 2 * X + L[0] + 2) >> 2);
        return 1;
    case 5: if (!(top && left && corner)) return 0;
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
            int z = 2 * x - y, i = x - (y >> 1), v;
            if (z >= 0 && !(z & 1)) v = (PT(i - 1) + PT(i) + 1) >> 1;
            else if (z >= 0) v = (PT(i - 2) + 2 * PT(i - 1) + PT(i) + 2) >> 2;
            else if (z == -1) v = (L[0] + 2 * X + T[0] + 2) >> 2;
            else v = (PL(y - 1) + 2 * PL(y - 2) + PL(y - 3) + 2) >> 2;
            p[y * 4 + x] = (uint8_t)v;
        }
        return 1;
    case 6: if (!(top && left && corner)) return 0;
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
            int z = 2 * y - x, i = y - (x >> 1), v;
            if (z >= 0 && !(z & 1)) v = (PL(i - 1) + PL(i) + 1) >> 1;
            else if (z >= 0) v = (PL(i - 2) + 2 * PL(i - 1) + PL(i) + 2) >> 2;
            else if (z == -1) v = (L[0] + 2 * X + T[0] + 2) >> 2;
            else v = (PT(x - 1) + 2 * PT(x - 2) + PT(x - 3) + 2) >> 2;
            p[y * 4 + x] = (uint8_t)v;
        }
        return 1;
    case 7: if (!top) return 0;
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
            int i = x + (y >> 1);
            p[y * 4 + x] = (uint8_t)((y & 1) ? (T[i] + 2 * T[i + 1] + T[i + 2] + 2) >> 2 : (T[i] + T[i + 1] + 1) >> 1);
        }
        return 1;
    default: if (!left) return 0;
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
            int z = x + 2 * y, i = y + (x >> 1), v;
            if (z > 5) v = L[3];
            else if (z == 5) v = (L[2] + 3 * L[3] + 2) >> 2;
            else if (z & 1) v = (L[i] + 2 * L[i + 1] + L[i + 2] + 2) >> 2;
            else v = (L[i] + L[i + 1] + 1) >> 1;
            p[y * 4 + x] = (uint8_t)v;
        }
        return 1;
    }
#undef PT
#undef PL
}
int orc_pred_i4(const uint8_t *r, int st, int mode, int avail, uint8_t *p) { return pred_i4(r, st, mode, avail, p); }

/* predIntra4x4PredMode of block b (8.3.1.1): min of the left and upper blocks' modes; DC when a neighbouring MB is missing
 * (then for both) or is not Intra_4x4 (constrained_intra_pred_flag = 0) */
static int i4_pred_mode(const OrcEncoder *e, int mx, int my, int b)
{
    const OrcMbInfo *m = &e->mbi[my * e->mbw + mx];
    int bx = BLK_X[b], by = BLK_Y[b], ma, mb_;
    static const uint8_t XY2B[4][4] = { { 0, 1, 4, 5 }, { 2, 3, 6, 7 }, { 8, 9, 12, 13 }, { 10, 11, 14, 15 } };
    if (bx > 0) ma = m->i4_mode[XY2B[by][bx - 1]];
    else if (mx == 0) return 2;
    else ma = ORC_MB_IS_INXN(m - 1) ? (m - 1)->i4_mode[XY2B[by][3]] : 2;
    if (by > 0) mb_ = m->i4_mode[XY2B[by - 1][bx]];
    else if (row_is_slice_top(e, my)) return 2;
    else mb_ = ORC_MB_IS_INXN(m - e->mbw) ? (m - e->mbw)->i4_mode[XY2B[3][bx]] : 2;
    return ma < mb_ ? ma : mb_;
}
#define ORC_I4_BIAS_BITS 24   /* fixed cost of choosing Intra_4x4 (16 mode flags), in lambda units */

/* Try Intra_4x4 on the luma of one MB: blocks in decoding order, each picks argmin of key = ((SATD + lambda*modebits) << 4) | mode
 * (modebits 1 when the mode equals its prediction, else 4), is transformed/quantised/reconstructed at once because the next
 * block predicts from it. Returns the total cost; fills i4_mode, luma levels, nnz, cbp luma bits; reconstruction in e->rec. */
static int code_intra4x4_luma(OrcEncoder *e, int mx, int my, int qp, int lambda)
{
    int mb = my * e->mbw + mx, st = e->wc; OrcMbInfo *mi = &e->mbi[mb]; OrcMbCoef *co = &e->coef[mb];
    int top = !row_is_slice_top(e, my), left = mx > 0, topright = top && mx + 1 < e->mbw;
    const uint8_t *s = e->src[0] + (size_t)my * 16 * st + mx * 16; uint8_t *r = e->rec[0] + (size_t)my * 16 * st + mx * 16;
    int total = lambda * ORC_I4_BIAS_BITS, cbp = 0;
    mi->mb_type = ORC_MB_I4x4;    /* so that i4_pred_mode() of later blocks sees this MB's own modes */
    for (int b = 0; b < 16; b++) {
        int bx = BLK_X[b], by = BLK_Y[b], avail = 0;
        if (by > 0 || top) avail |= I4_AV_T;
        if (bx > 0 || left) avail |= I4_AV_L;
        if ((by > 0 || top) && (bx > 0 || left)) avail |= I4_AV_X;
        if (by == 0 ? (bx < 3 ? top : topright) : (bx < 3 && b != 3 && b != 11 && b != 7 && b != 13 && b != 15)) avail |= I4_AV_TR;
        if (b == 3 || b == 11) avail &= ~I4_AV_TR;
        const uint8_t *sb = s + by * 4 * st + bx * 4; uint8_t *rb = r + by * 4 * st + bx * 4;
        int pm = i4_pred_mode(e, mx, my, b); uint32_t best = 0xffffffffu; uint8_t pred[16], bp[16];
        for (int m = 0; m < 9; m++) {
            /* Block 5 (the MB's upper-right 4x4 block) is the only Intra_4x4 block that predicts from the UPPER-RIGHT macroblock, and only in
             * modes 3 (diagonal down-left) and 7 (vertical-left). Without them -- an encoder's choice, the stream stays conformant -- a
             * macroblock depends on its left, upper-left and upper neighbours only, and the wavefront of the GPU runs at a lag of one macroblock
             * per row instead of two (critical path mbw + mbh steps instead of mbw + 2 mbh). High-profile sessions keep both modes: there
             * Intra_8x8 needs the upper-right macroblock anyway (8.3.2.2.1 filters p[15,-1] with p[16,-1]). */
            if (b == 5 && topright && (m == 3 || m == 7) && !I8X8_ON(e)) continue;
            if (!pred_i4(rb, st, m, avail, pred)) continue;
            uint32_t key = ((uint32_t)(orc_satd4x4(sb, st, pred, 4) + lambda * (m == pm ? 1 : 4)) << 4) | (uint32_t)m;
            if (key < best) { best = key; memcpy(bp, pred, 16); }
        }
        mi->i4_mode[b] = (uint8_t)(best & 15); total += (int)(best >> 4);
        int16_t res[16], c[16]; int32_t d[16], rr[16];
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) res[y * 4 + x] = (int16_t)(sb[y * st + x] - bp[y * 4 + x]);
        orc_dct4x4(res, c);
        int n = orc_quant4x4(c, co->luma[b], qp, 1, 0);
        mi->nnz[b] = (uint8_t)n; if (n) cbp |= 1 << (b >> 2);
        orc_dequant4x4(co->luma[b], d, qp, 0); orc_idct4x4(d, rr);
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) rb[y * st + x] = (uint8_t)clip255(bp[y * 4 + x] + rr[y * 4 + x]);
    }
    mi->cbp = (uint8_t)cbp;
    return total;
}

/* ---- Intra_8x8 (High profile), 8.3.2; pinned by the decoder round trip. ----
 * Reference samples of an 8x8 block with the filtering of 8.3.2.2.1: T[0..15] = p'[x,-1], L[0..7] = p'[-1,y], X = p'[-1,-1].
 * avail: I4_AV_T / _L / _X / _TR as for Intra_4x4 (top-right missing: p[8..15,-1] = p[7,-1]). */
typedef struct { int T[16], L[8], X, top, left, corner; } I8Ref;
static void i8_reference(const uint8_t *r, int st, int avail, I8Ref *o)
{
    int t[16], l[8], x = 128;
    o->top = !!(avail & I4_AV_T); o->left = !!(avail & I4_AV_L); o->corner = !!(avail & I4_AV_X);
    for (int i = 0; i < 16; i++) t[i] = 128;
    for (int i = 0; i < 8; i++) l[i] = 128;
    if (o->top) { for (int i = 0; i < 8; i++) t[i] = r[i - st]; for (int i = 8; i < 16; i++) t[i] = (avail & I4_AV_TR) ? r[i - st] : t[7]; }
    if (o->left) for (int i = 0; i < 8; i++) l[i] = r[i * st - 1];
    if (o->corner) x = r[-st - 1];
    for (int i = 0; i < 16; i++) o->T[i] = t[i];
    for (int i = 0; i < 8; i++) o->L[i] = l[i];
    o->X = x;
    if (o->top) {
        o->T[0] = o->corner ? (x + 2 * t[0] + t[1] + 2) >> 2 : (3 * t[0] + t[1] + 2) >> 2;
        for (int i = 1; i < 15; i++) o->T[i] = (t[i - 1] + 2 * t[i] + t[i + 1] + 2) >> 2;
        o->T[15] = (t[14] + 3 * t[15] + 2) >> 2;
    }
    if (o->corner) {
        if (o->top && o->left) o->X = (t[0] + 2 * x + l[0] + 2) >> 2;
        else if (o->top) o->X = (3 * x + t[0] + 2) >> 2;
        else if (o->left) o->X = (3 * x + l[0] + 2) >> 2;
    }
    if (o->left) {
        o->L[0] = o->corner ? (x + 2 * l[0] + l[1] + 2) >> 2 : (3 * l[0] + l[1] + 2) >> 2;
        for (int i = 1; i < 7; i++) o->L[i] = (l[i - 1] + 2 * l[i] + l[i + 1] + 2) >> 2;
        o->L[7] = (l[6] + 3 * l[7] + 2) >> 2;
    }
}
/* the nine Intra_8x8 predictors, 8.3.2.2.2 .. 8.3.2.2.10; returns 0 when the mode's neighbours are missing */
static int pred_i8(const I8Ref *f, int mode, uint8_t *p /*64*/)
{
    const int *T = f->T, *L = f->L; int X = f->X, all = f->top && f->left && f->corner;
#define PT8(i) ((i) < 0 ? X : T[i])
#define PL8(i) ((i) < 0 ? X : L[i])
    switch (mode) {
    case 0: if (!f->top) return 0; for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) p[y * 8 + x] = (uint8_t)T[x]; return 1;
    case 1: if (!f->left) return 0; for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) p[y * 8 + x] = (uint8_t)L[y]; return 1;
    case 2: {
        int sT = 0, sL = 0; for (int i = 0; i < 8; i++) { sT += T[i]; sL += L[i]; }
        int v = f->top && f->left ? (sT + sL + 8) >> 4 : f->top ? (sT + 4) >> 3 : f->left ? (sL + 4) >> 3 : 128;
        memset(p, v, 64); return 1; }
    case 3: if (!f->top) return 0;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++)
            p[y * 8 + x] = (uint8_t)(x == 7 && y == 7 ? (T[14] + 3 * T[15] + 2) >> 2 : (T[x + y] + 2 * T[x + y + 1] + T[x + y + 2] + 2) >> 2);
        return 1;
    case 4: if (!all) return 0;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++)
            p[y * 8 + x] = (uint8_t)(x > y ? (PT8(x - y - 2) + 2 * PT8(x - y - 1) + PT8(x - y) + 2) >> 2
                                   : x < y ? (PL8(y - x - 2) + 2 * PL8(y - x - 1) + PL8(y - x) + 2) >> 2 : (T[0] + 2 * X + L[0] + 2) >> 2);
        return 1;
    case 5: if (!all) return 0;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
            int z = 2 * x - y, i = x - (y >> 1), v;
            if (z >= 0 && !(z & 1)) v = (PT8(i - 1) + PT8(i) + 1) >> 1;
            else if (z >= 0) v = (PT8(i - 2) + 2 * PT8(i - 1) + PT8(i) + 2) >> 2;
            else if (z == -1) v = (L[0] + 2 * X + T[0] + 2) >> 2;
            else v = (PL8(y - 2 * x - 1) + 2 * PL8(y - 2 * x - 2) + PL8(y - 2 * x - 3) + 2) >> 2;
            p[y * 8 + x] = (uint8_t)v;
        }
        return 1;
    case 6: if (!all) return 0;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
            int z = 2 * y - x, i = y - (x >> 1), v;
            if (z >= 0 && !(z & 1)) v = (PL8(i - 1) + PL8(i) + 1) >> 1;
            else if (z >= 0) v = (PL8(i - 2) + 2 * PL8(i - 1) + PL8(i) + 2) >> 2;
            else if (z == -1) v = (L[0] + 2 * X + T[0] + 2) >> 2;
            else v = (PT8(x - 2 * y - 1) + 2 * PT8(x - 2 * y - 2) + PT8(x - 2 * y - 3) + 2) >> 2;
            p[y * 8 + x] = (uint8_t)v;
        }
        return 1;
    case 7: if (!f->top) return 0;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
            int i = x + (y >> 1);
            p[y * 8 + x] = (uint8_t)((y & 1) ? (T[i] + 2 * T[i + 1] + T[i + 2] + 2) >> 2 : (T[i] + T[i + 1] + 1) >> 1);
        }
        return 1;
    default: if (!f->left) return 0;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
            int z = x + 2 * y, i = y + (x >> 1), v;
            if (z > 13) v = L[7];
            else if (z == 13) v = (L[6] + 3 * L[7] + 2) >> 2;
            else if (z & 1) v = (L[i] + 2 * L[i + 1] + L[i + 2] + 2) >> 2;
            else v = (L[i] + L[i + 1] + 1) >> 1;
            p[y * 8 + x] = (uint8_t)v;
        }
        return 1;
    }
#undef PT8
#undef PL8
}
/* kernel-level oracle of the Intra_8x8 predictors: r = the block's top-left sample inside a reconstruction with its neighbours, avail bits as orc_pred_i4 */
int orc_pred_i8(const uint8_t *r, int stride, int mode, int avail, uint8_t *pred64) { I8Ref ref; i8_reference(r, stride, avail, &ref); return pred_i8(&ref, mode, pred64); }
/* predIntra8x8PredMode of 8x8 block b (8.3.2.1): the neighbouring I_NxN MB's mode of the 4x4 block 4 * blk8 + 1 (left) / + 2 (above);
 * Intra_8x8 MBs keep their block modes in all four entries, so one lookup serves both kinds */
static int i8_pred_mode(const OrcEncoder *e, int mx, int my, int b)
{
    const OrcMbInfo *m = &e->mbi[my * e->mbw + mx]; int ma, mb_;
    if (b & 1) ma = m->i4_mode[4 * (b - 1) + 1];
    else if (mx == 0) return 2;
    else ma = ORC_MB_IS_INXN(m - 1) ? (m - 1)->i4_mode[4 * (b + 1) + 1] : 2;
    if (b & 2) mb_ = m->i4_mode[4 * (b - 2) + 2];
    else if (row_is_slice_top(e, my)) return 2;
    else mb_ = ORC_MB_IS_INXN(m - e->mbw) ? (m - e->mbw)->i4_mode[4 * (b + 2) + 2] : 2;
    return ma < mb_ ? ma : mb_;
}
#define ORC_I8_BIAS_BITS 12   /* fixed cost of choosing Intra_8x8 (four mode flags, the transform flag), in lambda units */
/* Try Intra_8x8 on the luma of one MB, the four blocks in decoding order, each reconstructed at once (the next one predicts from it).
 * Returns the cost; fills i4_mode (replicated), the 8x8 levels, nnz, cbp luma bits; reconstruction in e->rec. */
static int code_intra8x8_luma(OrcEncoder *e, int mx, int my, int qp, int lambda)
{
    int mb = my * e->mbw + mx, st = e->wc; OrcMbInfo *mi = &e->mbi[mb]; OrcMbCoef *co = &e->coef[mb];
    int top = !row_is_slice_top(e, my), left = mx > 0, topright = top && mx + 1 < e->mbw;
    const uint8_t *s = e->src[0] + (size_t)my * 16 * st + mx * 16; uint8_t *r = e->rec[0] + (size_t)my * 16 * st + mx * 16;
    int total = lambda * ORC_I8_BIAS_BITS, cbp = 0;
    mi->mb_type = ORC_MB_I8x8;
    for (int b = 0; b < 4; b++) {
        int bx = (b & 1) * 8, by = (b >> 1) * 8, avail = 0;
        int at = by ? 1 : top, al = bx ? 1 : left;
        int ax = b == 0 ? (top && left) : b == 1 ? top : b == 2 ? left : 1;
        int atr = b == 0 ? top : b == 1 ? topright : b == 2 ? 1 : 0;
        if (at) avail |= I4_AV_T;
        if (al) avail |= I4_AV_L;
        if (ax) avail |= I4_AV_X;
        if (at && atr) avail |= I4_AV_TR;
        const uint8_t *sb = s + by * st + bx; uint8_t *rb = r + by * st + bx;
        I8Ref ref; i8_reference(rb, st, avail, &ref);
        int pm = i8_pred_mode(e, mx, my, b); uint32_t best = 0xffffffffu; uint8_t pred[64], bp[64];
        for (int m = 0; m < 9; m++) {
            if (!pred_i8(&ref, m, pred)) continue;
            int satd = orc_satd4x4(sb, st, pred, 8) + orc_satd4x4(sb + 4, st, pred + 4, 8) + orc_satd4x4(sb + 4 * st, st, pred + 32, 8) + orc_satd4x4(sb + 4 * st + 4, st, pred + 36, 8);
            uint32_t key = ((uint32_t)(satd + lambda * (m == pm ? 1 : 4)) << 4) | (uint32_t)m;
            if (key < best) { best = key; memcpy(bp, pred, 64); }
        }
        for (int k = 0; k < 4; k++) mi->i4_mode[4 * b + k] = (uint8_t)(best & 15);
        total += (int)(best >> 4);
        int16_t res[64]; int32_t c[64], d[64], rr[64];
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) res[y * 8 + x] = (int16_t)(sb[y * st + x] - bp[y * 8 + x]);
        orc_dct8x8(res, c);
        int n = orc_quant8x8(c, co->luma[4 * b], qp, 1);
        for (int k = 0; k < 4; k++) mi->nnz[4 * b + k] = (uint8_t)n;
        if (n) cbp |= 1 << b;
        orc_dequant8x8(co->luma[4 * b], d, qp); orc_idct8x8(d, rr);
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) rb[y * st + x] = (uint8_t)clip255(bp[y * 8 + x] + rr[y * 8 + x]);
    }
    mi->cbp = (uint8_t)cbp;
    return total;
}

/* ---- Phase C: one intra MB (roles of WelsMdI16x16, WelsMdI4x4, WelsIChromaPred*, WelsHadamardT4Dc_c, WelsDequantIHadamard4x4_c).
 * Intra_16x16 mode by SATD; Intra_4x4 is coded on trial and kept when its cost is below the Intra_16x16 SATD. ---- */
static void code_intra_mb(OrcEncoder *e, int mx, int my, int qp)
{
    int mb = my * e->mbw + mx, st = e->wc, cs = st / 2; OrcMbInfo *mi = &e->mbi[mb]; OrcMbCoef *co = &e->coef[mb];
    int top = !row_is_slice_top(e, my), left = mx > 0;
    const uint8_t *s = e->src[0] + (size_t)my * 16 * st + mx * 16; uint8_t *r = e->rec[0] + (size_t)my * 16 * st + mx * 16;
    uint8_t pred[256], best_pred[256]; uint32_t best = 0xffffffffu; int mode = 2;
    memset(co, 0, sizeof *co);
    for (int m = 0; m < 4; m++) {
        if ((m == 0 && !top) || (m == 1 && !left) || (m == 3 && !(top && left))) continue;
        pred_i16(r, st, m, top, left, pred);
        uint32_t key = ((uint32_t)orc_satd16x16(s, st, pred, 16) << 2) | (uint32_t)m;
        if (key < best) { best = key; mode = m; memcpy(best_pred, pred, 256); }
    }
    mi->mv[0] = mi->mv[1] = 0; mi->i16_mode = 0;
    /* Intra_8x8 on trial first (High profile groundwork): its outcome is kept aside while Intra_4x4 is tried on the same samples. The two
     * I_NxN codings are compared by J = 64 SSD + 27 lambda^2 B like the inter transform choice (B in half bits: levels + 2 x mode bits);
     * the winner then meets Intra_16x16 with its SATD cost as before. */
    int cost8 = 1 << 30, use_i8 = 0; OrcMbInfo keep_mi; int16_t keep_luma[16][16]; uint8_t keep_rec[256]; int64_t j8 = 0;
    int64_t l2 = 27 * (int64_t)LAMBDA_TAB[qp] * LAMBDA_TAB[qp];
    if (I8X8_ON(e)) {
        cost8 = code_intra8x8_luma(e, mx, my, qp, LAMBDA_TAB[qp]);
        keep_mi = *mi; memcpy(keep_luma, co->luma, sizeof keep_luma);
        int rate = 2;
        for (int b = 0; b < 4; b++) {
            rate += 2 * (mi->i4_mode[4 * b] == i8_pred_mode(e, mx, my, b) ? 1 : 4);
            if (mi->cbp & (1 << b)) rate += level_cost2(co->luma[4 * b], 64);
        }
        int64_t ssd = 0;
        for (int y = 0; y < 16; y++) { memcpy(keep_rec + y * 16, r + y * st, 16); for (int x = 0; x < 16; x++) { int d = s[y * st + x] - r[y * st + x]; ssd += d * d; } }
        j8 = 64 * ssd + l2 * rate;
        memset(co, 0, sizeof *co); memset(mi->i4_mode, 0, 16); memset(mi->nnz, 0, 16);
    }
    int cost4 = e->cfg.no_i4x4 ? 1 << 30 : code_intra4x4_luma(e, mx, my, qp, LAMBDA_TAB[qp]);
    if (I8X8_ON(e) && cost4 < (1 << 30)) {
        int rate = 0; int64_t ssd = 0;
        for (int k = 0; k < 16; k++) {
            rate += 2 * (mi->i4_mode[k] == i4_pred_mode(e, mx, my, k) ? 1 : 4);
            if (mi->cbp & (1 << (k >> 2))) rate += 1 + level_cost2(co->luma[k], 16);
        }
        for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) { int d = s[y * st + x] - r[y * st + x]; ssd += d * d; }
        if (j8 < 64 * ssd + l2 * rate) use_i8 = 1;
    } else if (I8X8_ON(e)) use_i8 = 1;
    int use_i4 = (use_i8 ? cost8 : cost4) < (int)(best >> 2);
    if (use_i8 && use_i4) {                                      /* Intra_8x8 wins: put its outcome back */
        *mi = keep_mi; memcpy(co->luma, keep_luma, sizeof keep_luma);
        for (int y = 0; y < 16; y++) memcpy(r + y * st, keep_rec + y * 16, 16);
        mi->i16_mode = 4;                                        /* transform_size_8x8_flag */
    }
    (void)use_i8;
    int luma_cbp = mi->cbp & 15;
    if (!use_i4) {
    memset(co, 0, sizeof *co); memset(mi->i4_mode, 0, 16);
    mi->mb_type = ORC_MB_I16x16; mi->i16_mode = (uint8_t)mode;
    /* luma: 16 forward transforms, DC Hadamard, quant, and the normative inverse (8.5.2, 8.5.10, 8.5.12) */
    int16_t c[16][16]; int dcm[16], hd[16], any_ac = 0;
    for (int b = 0; b < 16; b++) {
        int bx = BLK_X[b] * 4, by = BLK_Y[b] * 4; int16_t res[16];
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) res[y * 4 + x] = (int16_t)(s[(by + y) * st + bx + x] - best_pred[(by + y) * 16 + bx + x]);
        orc_dct4x4(res, c[b]);
        dcm[BLK_Y[b] * 4 + BLK_X[b]] = c[b][0];
        int n = orc_quant4x4(c[b], co->luma[b], qp, 1, 1);
        mi->nnz[b] = (uint8_t)n; any_ac |= n;
    }
    hadamard4x4(dcm, hd);
    int lev[16];
    for (int i = 0; i < 16; i++) lev[i] = quant_dc((hd[i] + 1) >> 1, qp);
    for (int i = 0; i < 16; i++) co->luma_dc[i] = (int16_t)lev[ZIGZAG4x4[i]];
    int fdc[16]; hadamard4x4(lev, fdc);
    int LS = 16 * DEQUANT_V[qp % 6][0];
    for (int b = 0; b < 16; b++) {
        int bx = BLK_X[b] * 4, by = BLK_Y[b] * 4; int32_t d[16], rr[16];
        orc_dequant4x4(co->luma[b], d, qp, 1);
        int f = fdc[BLK_Y[b] * 4 + BLK_X[b]];
        d[0] = qp >= 36 ? (f * LS) << (qp / 6 - 6) : (f * LS + (1 << (5 - qp / 6))) >> (6 - qp / 6);
        orc_idct4x4(d, rr);
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) r[(by + y) * st + bx + x] = (uint8_t)clip255(best_pred[(by + y) * 16 + bx + x] + rr[y * 4 + x]);
    }
    luma_cbp = any_ac ? 15 : 0;
    }
    /* chroma mode: SATD over both planes, key = cost<<2 | mode */
    const uint8_t *su = e->src[1] + (size_t)my * 8 * cs + mx * 8, *sv = e->src[2] + (size_t)my * 8 * cs + mx * 8;
    const uint8_t *ru = e->rec[1] + (size_t)my * 8 * cs + mx * 8, *rv = e->rec[2] + (size_t)my * 8 * cs + mx * 8;
    uint8_t pc[128], best_pc[128]; int cmode = 0; best = 0xffffffffu;
    for (int m = 0; m < 4; m++) {
        if ((m == 1 && !left) || (m == 2 && !top) || (m == 3 && !(top && left))) continue;
        pred_chroma(ru, cs, m, top, left, pc); pred_chroma(rv, cs, m, top, left, pc + 64);
        uint32_t key = ((uint32_t)(satd8x8(su, cs, pc, 8) + satd8x8(sv, cs, pc + 64, 8)) << 2) | (uint32_t)m;
        if (key < best) { best = key; cmode = m; memcpy(best_pc, pc, 128); }
    }
    mi->chroma_mode = (uint8_t)cmode;
    int ccbp = code_chroma(e, mx, my, best_pc, qp, 1);
    mi->cbp = (uint8_t)(luma_cbp | (ccbp << 4));
}

/* ---- Phase D: luma MV prediction (8.4.1.3, neighbours per 6.4.11.7) for the 16x16 partition or one 8x8 partition, and
 * the P_Skip vector (8.4.1.1). One reference picture, so refIdx is 0 for inter neighbours and -1 for intra ones. ---- */
#define IS_INTER(p) ((p)->mb_type == ORC_MB_P16x16 || (p)->mb_type == ORC_MB_PSKIP || (p)->mb_type == ORC_MB_P8x8)
typedef struct { int avail, ref, x, y; } NbMv;
/* the partition covering luma location (x,y) relative to MB (mx,my); partitions of the current MB with index >= cur_part
 * are not decoded yet */
static NbMv nb_at(const OrcEncoder *e, int mx, int my, int x, int y, int cur_part)
{
    NbMv r = { 0, -1, 0, 0 };
    int nx = mx + (x < 0 ? -1 : x > 15 ? 1 : 0), ny = my - (y < 0);
    if (nx < 0 || nx >= e->mbw) return r;
    if (ny != my && row_is_slice_top(e, my)) return r;
    if (ny == my && nx > mx) return r;
    int part = (((y + 16) & 15) >> 3) * 2 + (((x + 16) & 15) >> 3);
    if (nx == mx && ny == my && part >= cur_part) return r;
    const OrcMbInfo *m = &e->mbi[ny * e->mbw + nx];
    r.avail = 1;
    if (IS_INTER(m)) { r.ref = 0; r.x = m->mv8[part][0]; r.y = m->mv8[part][1]; }
    return r;
}
/* part < 0: the 16x16 partition; else the 8x8 partition `part` */
static void predict_mv_part(const OrcEncoder *e, int mx, int my, int part, int *pmx, int *pmy, NbMv *outA, NbMv *outB)
{
    int px = part < 0 ? 0 : (part & 1) * 8, py = part < 0 ? 0 : (part >> 1) * 8, w = part < 0 ? 16 : 8, cur = part < 0 ? 0 : part;
    NbMv A = nb_at(e, mx, my, px - 1, py, cur), B = nb_at(e, mx, my, px, py - 1, cur), C = nb_at(e, mx, my, px + w, py - 1, cur);
    if (!C.avail) C = nb_at(e, mx, my, px - 1, py - 1, cur);
    if (outA) *outA = A;
    if (outB) *outB = B;
    if (!B.avail && !C.avail && A.avail) { B = A; C = A; }
    int n = (A.ref == 0) + (B.ref == 0) + (C.ref == 0);
    if (n == 1) { const NbMv *o = A.ref == 0 ? &A : B.ref == 0 ? &B : &C; *pmx = o->x; *pmy = o->y; }
    else { *pmx = median3(A.x, B.x, C.x); *pmy = median3(A.y, B.y, C.y); }
}
static void predict_mv(const OrcEncoder *e, int mx, int my, int *pmx, int *pmy, int *skx, int *sky)
{
    NbMv A, B;
    predict_mv_part(e, mx, my, -1, pmx, pmy, &A, &B);
    if (!A.avail || !B.avail || (A.ref == 0 && A.x == 0 && A.y == 0) || (B.ref == 0 && B.x == 0 && B.y == 0)) { *skx = 0; *sky = 0; }
    else { *skx = *pmx; *sky = *pmy; }
}

/* ---- Phase F: macroblock layer syntax, 7.3.5 (role of WelsSpatialWriteMbSyn + WelsWriteMbResidual) ---- */
static int nnz_ctx(const OrcEncoder *e, int mx, int my, int idx_cur, int idx_left_mb, int idx_top_mb, int has_left_in, int has_top_in)
{
    /* idx_cur unused; the caller passes neighbour indices: inside the MB (has_*_in) or in the adjacent MB */
    (void)idx_cur;
    const OrcMbInfo *m = &e->mbi[my * e->mbw + mx];
    int nA = -1, nB = -1;
    if (has_left_in) nA = m->nnz[idx_left_mb]; else if (mx > 0) nA = (m - 1)->nnz[idx_left_mb];
    if (has_top_in) nB = m->nnz[idx_top_mb]; else if (!row_is_slice_top(e, my)) nB = (m - e->mbw)->nnz[idx_top_mb];
    if (nA >= 0 && nB >= 0) return (nA + nB + 1) >> 1;
    return nA >= 0 ? nA : (nB >= 0 ? nB : 0);
}
static const uint8_t XY2BLK[4][4] = { { 0, 1, 4, 5 }, { 2, 3, 6, 7 }, { 8, 9, 12, 13 }, { 10, 11, 14, 15 } };
static int luma_nc(const OrcEncoder *e, int mx, int my, int b)
{
    int bx = BLK_X[b], by = BLK_Y[b];
    return nnz_ctx(e, mx, my, b, bx > 0 ? XY2BLK[by][bx - 1] : XY2BLK[by][3], by > 0 ? XY2BLK[by - 1][bx] : XY2BLK[3][bx], bx > 0, by > 0);
}
static int chroma_nc(const OrcEncoder *e, int mx, int my, int pl, int b)
{
    int bx = b & 1, by = b >> 1, base = 16 + pl * 4;
    return nnz_ctx(e, mx, my, b, base + by * 2 + (bx > 0 ? 0 : 1), base + (by > 0 ? 0 : 2) + bx, bx > 0, by > 0);
}

static void write_mb(OrcEncoder *e, BitWriter *b, int mx, int my, int is_p)
{
    int mb = my * e->mbw + mx; const OrcMbInfo *mi = &e->mbi[mb]; const OrcMbCoef *co = &e->coef[mb];
    int cl = mi->cbp & 15, cc = mi->cbp >> 4;
    if (mi->mb_type == ORC_MB_I16x16) {
        bw_ue(b, (uint32_t)((is_p ? 5 : 0) + 1 + mi->i16_mode + 4 * cc + (cl ? 12 : 0)));
        bw_ue(b, mi->chroma_mode);
        bw_se(b, 0);                                           /* mb_qp_delta */
        orc_write_residual_block(b, co->luma_dc, 16, luma_nc(e, mx, my, 0));
        if (cl) for (int k = 0; k < 16; k++) orc_write_residual_block(b, co->luma[k] + 1, 15, luma_nc(e, mx, my, k));
    } else if (mi->mb_type == ORC_MB_I4x4) {
        bw_ue(b, (uint32_t)(is_p ? 5 : 0));                    /* I_NxN; transform_8x8_mode_flag is 0 in the PPS */
        for (int k = 0; k < 16; k++) {                         /* prev_intra4x4_pred_mode_flag / rem_intra4x4_pred_mode, 7.3.5.1 */
            int pm = i4_pred_mode(e, mx, my, k), m = mi->i4_mode[k];
            if (m == pm) bw_put(b, 1, 1); else { bw_put(b, 1, 0); bw_put(b, 3, (uint32_t)(m < pm ? m : m - 1)); }
        }
        bw_ue(b, mi->chroma_mode);
        bw_ue(b, CBP_TO_CODENUM_INTRA[mi->cbp]);
        if (mi->cbp) bw_se(b, 0);
        for (int k = 0; k < 16; k++) if (cl & (1 << (k >> 2))) orc_write_residual_block(b, co->luma[k], 16, luma_nc(e, mx, my, k));
    } else if (mi->mb_type == ORC_MB_P8x8) {
        bw_ue(b, 3);                                           /* P_8x8 */
        for (int q = 0; q < 4; q++) bw_ue(b, 0);               /* sub_mb_type P_L0_8x8; ref_idx_l0 is not coded with one reference */
        for (int q = 0; q < 4; q++) {
            int pmx, pmy; predict_mv_part(e, mx, my, q, &pmx, &pmy, 0, 0);
            bw_se(b, mi->mv8[q][0] - pmx); bw_se(b, mi->mv8[q][1] - pmy);
        }
        bw_ue(b, CBP_TO_CODENUM_INTER[mi->cbp]);
        if (mi->cbp) bw_se(b, 0);
        for (int k = 0; k < 16; k++) if (cl & (1 << (k >> 2))) orc_write_residual_block(b, co->luma[k], 16, luma_nc(e, mx, my, k));
    } else {
        int pmx, pmy, sx, sy; predict_mv(e, mx, my, &pmx, &pmy, &sx, &sy);
        bw_ue(b, 0);                                           /* P_L0_16x16 */
        bw_se(b, mi->mv[0] - pmx); bw_se(b, mi->mv[1] - pmy);
        bw_ue(b, CBP_TO_CODENUM_INTER[mi->cbp]);
        if (mi->cbp) bw_se(b, 0);
        for (int k = 0; k < 16; k++) if (cl & (1 << (k >> 2))) orc_write_residual_block(b, co->luma[k], 16, luma_nc(e, mx, my, k));
    }
    if (cc) { orc_write_residual_block(b, co->chroma_dc[0], 4, -1); orc_write_residual_block(b, co->chroma_dc[1], 4, -1); }
    if (cc == 2) for (int pl = 0; pl < 2; pl++) for (int k = 0; k < 4; k++)
        orc_write_residual_block(b, co->chroma_ac[pl][k] + 1, 15, chroma_nc(e, mx, my, pl, k));
}

/* ---- CABAC side records (mvd per partition, Intra_4x4 mode syntax, DC coded_block_flags): parallel over MBs ---- */
static void cabac_side_records(OrcEncoder *e, int is_idr)
{
    (void)is_idr;
    for (int my = 0; my < e->mbh; my++)
        for (int mx = 0; mx < e->mbw; mx++) {
            int mb = my * e->mbw + mx; const OrcMbInfo *mi = &e->mbi[mb]; const OrcMbCoef *co = &e->coef[mb]; OrcMbSide *sd = &e->side[mb];
            memset(sd, 0, sizeof *sd);
            if (mi->mb_type == ORC_MB_P16x16) {
                int pmx, pmy, sx, sy; predict_mv(e, mx, my, &pmx, &pmy, &sx, &sy);
                for (int q = 0; q < 4; q++) { sd->mvd[q][0] = (int16_t)(mi->mv[0] - pmx); sd->mvd[q][1] = (int16_t)(mi->mv[1] - pmy); }
            } else if (mi->mb_type == ORC_MB_P8x8) {
                for (int q = 0; q < 4; q++) {
                    int pmx, pmy; predict_mv_part(e, mx, my, q, &pmx, &pmy, 0, 0);
                    sd->mvd[q][0] = (int16_t)(mi->mv8[q][0] - pmx); sd->mvd[q][1] = (int16_t)(mi->mv8[q][1] - pmy);
                }
            } else if (mi->mb_type == ORC_MB_I4x4) {
                for (int k = 0; k < 16; k++) { int pm = i4_pred_mode(e, mx, my, k), m = mi->i4_mode[k]; sd->i4_syn[k] = (uint8_t)(m == pm ? 8 : m < pm ? m : m - 1); }
            } else if (mi->mb_type == ORC_MB_I8x8) {
                for (int k = 0; k < 4; k++) { int pm = i8_pred_mode(e, mx, my, k), m = mi->i4_mode[4 * k]; sd->i4_syn[k] = (uint8_t)(m == pm ? 8 : m < pm ? m : m - 1); }
            }
            if (mi->mb_type == ORC_MB_PSKIP) continue;
            int dc = 0;
            if (mi->mb_type == ORC_MB_I16x16) for (int i = 0; i < 16; i++) dc |= co->luma_dc[i] != 0;
            if (mi->cbp >> 4) for (int p = 0; p < 2; p++) for (int i = 0; i < 4; i++) dc |= (co->chroma_dc[p][i] != 0) << (1 + p);
            sd->dc_cbf = (uint8_t)dc;
        }
}
const OrcMbSide *orc_mb_side(const OrcEncoder *e) { return e->side; }
int orc_slice_bins(const OrcEncoder *e, int s, const uint16_t **bins) { *bins = e->bins + e->slice_bin0[s]; return e->slice_bin0[s + 1] - e->slice_bin0[s]; }
int orc_cabac_code_bins(const uint16_t *bins, int n, int slice_qp, int is_p, uint8_t *out, int cap)
{
    BitWriter b; bw_init(&b, out, cap);
    orc_cabac_code(&b, bins, n, slice_qp, is_p);
    while (b.nacc) bw_put(&b, 1, 0);
    return b.overflow ? -1 : b.pos;
}

static int encode_frame(OrcEncoder *e, const uint8_t *i420, int frame_type, int qp, uint8_t *out, int out_cap, int commit)
{
    int is_idr = frame_type == 1 || !e->have_ref, n = e->mbw * e->mbh, lambda = LAMBDA_TAB[qp];
    const int frame_num_in = e->frame_num;
    /* requested key pictures take the key layout; a promoted P picture (scene change, below) keeps the P layout it was searched with */
    e->slice_row0 = e->row0_tab[is_idr]; e->cur_slices = is_idr ? e->cfg.key_slices : e->cfg.num_slices;
    load_source(e, i420);
    if (is_idr) { e->frame_num = 0; }
    if (!is_idr) {
        /* Phase A */
        orc_downsample2(e->src[0], e->wc, e->wc, e->hc, e->srcL1, e->wc / 2);
        orc_downsample2(e->srcL1, e->wc / 2, e->wc / 2, e->hc / 2, e->srcL2, e->wc / 4);
        orc_downsample2(e->ref[0], e->wc, e->wc, e->hc, e->refL1, e->wc / 2);
        orc_downsample2(e->refL1, e->wc / 2, e->wc / 2, e->hc / 2, e->refL2, e->wc / 4);
        build_halfpel(e);
        for (int my = 0; my < e->mbh; my++)
            for (int mx = 0; mx < e->mbw; mx++) {
                int mb = my * e->mbw + mx;
                memset(&e->mbi[mb], 0, sizeof(OrcMbInfo));
                motion_search(e, mx, my, lambda, qp);
                int ie = intra_estimate(e, mx, my);
                if (ie + lambda * ORC_INTRA_BIAS_BITS < e->inter_cost[mb]) { memset(&e->mbi[mb], 0, sizeof(OrcMbInfo)); e->mbi[mb].mb_type = ORC_MB_I16x16; }
            }
        /* Scene change (the wrapper enables openh264's detector, VideoEncoderOpenH264.cpp:283): when at least 2/5 of the MBs came out
         * intra from the motion search, the picture is coded as an IDR instead -- but not within ORC_SC_MIN_DISTANCE pictures of the last
         * IDR: noise or sustained violent motion must not turn every picture into a key frame */
        int n_intra = 0;
        for (int i = 0; i < n; i++) n_intra += e->mbi[i].mb_type == ORC_MB_I16x16;
        if (!e->cfg.no_scene_change && e->since_idr >= ORC_SC_MIN_DISTANCE && 5 * n_intra >= 2 * n) {
            is_idr = 1; e->frame_num = 0;
            memset(e->mbi, 0, (size_t)n * sizeof(OrcMbInfo));
            for (int i = 0; i < n; i++) e->mbi[i].mb_type = ORC_MB_I16x16;
        } else
        /* Phase B */
        for (int my = 0; my < e->mbh; my++)
            for (int mx = 0; mx < e->mbw; mx++)
                if (e->mbi[my * e->mbw + mx].mb_type == ORC_MB_P16x16 || e->mbi[my * e->mbw + mx].mb_type == ORC_MB_P8x8) code_inter_mb(e, mx, my, qp);
    } else {
        memset(e->mbi, 0, (size_t)n * sizeof(OrcMbInfo));
        for (int i = 0; i < n; i++) e->mbi[i].mb_type = ORC_MB_I16x16;
    }
    /* Phase C */
    for (int my = 0; my < e->mbh; my++)
        for (int mx = 0; mx < e->mbw; mx++)
            if (e->mbi[my * e->mbw + mx].mb_type == ORC_MB_I16x16) code_intra_mb(e, mx, my, qp);
    /* Phase D */
    if (!is_idr)
        for (int my = 0; my < e->mbh; my++)
            for (int mx = 0; mx < e->mbw; mx++) {
                OrcMbInfo *mi = &e->mbi[my * e->mbw + mx];
                if (mi->mb_type != ORC_MB_P16x16 || mi->cbp) continue;
                int pmx, pmy, sx, sy; predict_mv(e, mx, my, &pmx, &pmy, &sx, &sy);
                if (mi->mv[0] == sx && mi->mv[1] == sy) mi->mb_type = ORC_MB_PSKIP;
            }
    /* Phase E */
    for (int c = 0; c < 3; c++) memcpy(e->dbk[c], e->rec[c], (size_t)(c ? e->wc / 2 * e->hc / 2 : e->wc * e->hc));
    orc_deblock_frame(e->dbk[0], e->wc, e->dbk[1], e->dbk[2], e->wc / 2, e->mbw, e->mbh, e->mbi, qp);
    /* Phase F */
    int o = 0;
    if (is_idr) {
        if (out_cap < 64) return -1;
        int level = e->cfg.level_idc ? e->cfg.level_idc : orc_level_for(e->cfg.width, e->cfg.height, e->cfg.fps);
        o += orc_write_sps(out + o, e->cfg.width, e->cfg.height, level, e->cfg.profile);
        o += orc_write_pps(out + o, e->cfg.profile, T8X8_ON(e));
    }
    if (e->cfg.profile) cabac_side_records(e, is_idr);
    for (int s = 0; s < e->cur_slices; s++) {
        BitWriter b; bw_init(&b, e->rbsp, e->rbsp_cap);
        int r0 = e->slice_row0[s], r1 = e->slice_row0[s + 1], run = 0;
        orc_write_slice_header(&b, r0 * e->mbw, is_idr, e->frame_num, e->idr_pic_id, qp, e->cfg.profile != 0);
        if (e->cfg.profile) {
            /* slice_data() with entropy_coding_mode_flag = 1 (7.3.4): cabac_alignment_one_bit, then one arithmetic codeword */
            while (b.nacc) bw_put(&b, 1, 1);
            int nb = e->slice_bin0[s];
            for (int my = r0; my < r1; my++)
                for (int mx = 0; mx < e->mbw; mx++) {
                    int k = orc_cabac_mb_bins(e->mbi, e->coef, e->side, e->mbw, mx, my, my > r0, !is_idr, my == r1 - 1 && mx == e->mbw - 1,
                                              T8X8_ON(e), e->bins + nb, e->bins_cap - nb);
                    if (k > ORC_MB_BINS_MAX || nb + k > e->bins_cap) return -1;
                    nb += k;
                }
            e->slice_bin0[s + 1] = nb;
            orc_cabac_code(&b, e->bins + e->slice_bin0[s], nb - e->slice_bin0[s], qp, !is_idr);
            while (b.nacc) bw_put(&b, 1, 0);                   /* the flush wrote the stop bit; rbsp_alignment_zero_bit */
        } else {
        for (int my = r0; my < r1; my++)
            for (int mx = 0; mx < e->mbw; mx++) {
                if (!is_idr) {
                    if (e->mbi[my * e->mbw + mx].mb_type == ORC_MB_PSKIP) { run++; continue; }
                    bw_ue(&b, (uint32_t)run); run = 0;
                }
                write_mb(e, &b, mx, my, !is_idr);
            }
        if (run) bw_ue(&b, (uint32_t)run);
        bw_trailing(&b);
        }
        if (b.overflow || o + 5 + b.pos + b.pos / 2 + 16 > out_cap) return -1;
        out[o] = 0; out[o + 1] = 0; out[o + 2] = 0; out[o + 3] = 1; out[o + 4] = is_idr ? 0x65 : 0x61;
        o += 5 + orc_escape_rbsp(e->rbsp, b.pos, out + o + 5);
    }
    e->last_idr = is_idr;
    if (!commit) { e->frame_num = frame_num_in; return o; }
    e->src_committed = 1;      /* a trial leaves the stream state (reference, frame_num, idr_pic_id) untouched */
    /* the deblocked picture becomes the reference of the next frame */
    for (int c = 0; c < 3; c++) memcpy(e->ref[c], e->dbk[c], (size_t)(c ? e->wc / 2 * e->hc / 2 : e->wc * e->hc));
    e->have_ref = 1; e->frame_num = (e->frame_num + 1) & 255;
    e->since_idr = is_idr ? 1 : e->since_idr + 1;
    if (is_idr) e->idr_pic_id = (e->idr_pic_id + 1) & 1;
    return o;
}
int orc_encode(OrcEncoder *e, const uint8_t *i420, int frame_type, int qp, uint8_t *out, int out_cap) { return encode_frame(e, i420, frame_type, qp, out, out_cap, 1); }
/* the same picture coded without advancing the stream: what a rate-control retry at another QP discards (tools/rc_sim.py) */
int orc_encode_trial(OrcEncoder *e, const uint8_t *i420, int frame_type, int qp, uint8_t *out, int out_cap) { return encode_frame(e, i420, frame_type, qp, out, out_cap, 0); }
