/*
 * oracle/orc_cabac.c -- TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * CABAC entropy coding of one slice (ITU-T H.264 clauses 7.3.4, 7.3.5, 9.3), the back end the reference asks openh264 for with
 * iEntropyCodingModeFlag = 1 (/root/reference/video_codec/VideoEncoderOpenH264.cpp:291) and profile main / high (:248-253).
 * Role of openh264's WelsCabacEncodeDecision / WelsSpatialWriteMbSynCabac inside the absent libopenh264.so; CABAC is fully
 * normative, so this follows the standard and is pinned by FFmpeg's decoder reproducing the reconstruction bit for bit.
 *
 * Two stages, the same split as the CUDA path:
 *   1. binarisation + context selection (9.3.2, 9.3.3): every macroblock turns into a list of 16-bit entries
 *      (orc.h: ctxIdx | bin << 10 | (repeat - 1) << 11, ctxIdx 276 = terminate, ORC_CTX_BYPASS0 + n = n bypass bins). It only
 *      depends on the MB's own record and its left / upper neighbours' records, so it is parallel over macroblocks;
 *   2. the arithmetic coder proper (9.3.4.2), serial over the slice's bin list.
 */
#include "orc_internal.h"
#include "cabac_tables.h"
#include "h264_tables.h"
#include <stdlib.h>

typedef struct { uint16_t *p; int n, cap; } Bins;
static inline void put(Bins *s, int ctx, int bin) { if (s->n < s->cap) s->p[s->n] = (uint16_t)(ctx | (bin << 10)); s->n++; }
/* `rep` (1..32) consecutive bins of the same value in the same context: one entry */
static inline void put_run(Bins *s, int ctx, int bin, int rep) { if (s->n < s->cap) s->p[s->n] = (uint16_t)(ctx | (bin << 10) | ((rep - 1) << 11)); s->n++; }
static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* A string of bypass bins (first bin = most significant of the `len` bits): entries of six bins, the remainder last */
static void put_bypass(Bins *s, uint32_t bits, int len)
{
    while (len > 0) {
        int n = imin(len, 6);
        uint32_t v = (bits >> (len - n)) & ((1u << n) - 1);
        if (s->n < s->cap) s->p[s->n] = (uint16_t)((ORC_CTX_BYPASS0 + n) | (v << 10));
        s->n++; len -= n;
    }
}
/* k-th order Exp-Golomb suffix (9.3.2.3) followed by the sign: appended to (bits, len) */
static void egk_sign(int v, int k, int neg, uint32_t *bits, int *len)
{
    uint32_t b = 0; int n = 0;
    while (v >= (1 << k)) { b = (b << 1) | 1; n++; v -= 1 << k; k++; }
    b <<= 1; n++;
    b = (b << k) | (uint32_t)v; n += k;
    *bits = (b << 1) | (uint32_t)neg; *len = n + 1;
}

/* mvd_l0 component: UEG3, signedValFlag 1, uCoff 9 (9.3.2.3); ctxIdxInc of bin 0 from the neighbours' |mvd| sum (9.3.3.1.1.7) */
static void put_mvd(Bins *s, int base, int sum, int v)
{
    int a = iabs(v);
    put(s, base + (sum < 3 ? 0 : sum > 32 ? 2 : 1), a != 0);
    if (!a) return;
    for (int i = 1; i < imin(a, 4); i++) put(s, base + 2 + i, 1);                  /* binIdx 1..3: ctxIdxInc 3, 4, 5 */
    if (a < 4) put(s, base + 2 + a, 0);
    else {
        if (imin(a, 9) > 4) put_run(s, base + 6, 1, imin(a, 9) - 4);              /* binIdx 4..: ctxIdxInc 6 */
        if (a < 9) put(s, base + 6, 0);
    }
    uint32_t bits = (uint32_t)(v < 0); int len = 1;
    if (a >= 9) egk_sign(a - 9, 3, v < 0, &bits, &len);
    put_bypass(s, bits, len);
}

/* residual_block_cabac (7.3.5.3.3): c = levels in scan order, n = 16 / 15 / 4, cat = ctxBlockCat 0..4 */
static void put_residual(Bins *s, const int16_t *c, int n, int cat, int cbf_inc)
{
    static const uint8_t CBF_OFF[5] = { 0, 4, 8, 12, 16 }, SIG_OFF[5] = { 0, 15, 29, 44, 47 }, ABS_OFF[5] = { 0, 10, 20, 30, 39 };
    int last = -1;
    for (int i = 0; i < n; i++) if (c[i]) last = i;
    put(s, 85 + CBF_OFF[cat] + cbf_inc, last >= 0);
    if (last < 0) return;
    for (int i = 0; i < n - 1; i++) {
        int inc = cat == 3 ? imin(i, 2) : i;
        put(s, 105 + SIG_OFF[cat] + inc, c[i] != 0);
        if (c[i]) { put(s, 166 + SIG_OFF[cat] + inc, i == last); if (i == last) break; }
    }
    int eq1 = 0, gt1 = 0;
    for (int i = last; i >= 0; i--) {
        if (!c[i]) continue;
        int a = iabs(c[i]) - 1, base = 227 + ABS_OFF[cat];
        put(s, base + (gt1 ? 0 : imin(4, 1 + eq1)), a > 0);
        uint32_t bits = (uint32_t)(c[i] < 0); int len = 1;
        if (a > 0) {
            int inc = 5 + imin(4 - (cat == 3), gt1);
            if (imin(a, 14) > 1) put_run(s, base + inc, 1, imin(a, 14) - 1);
            if (a < 14) put(s, base + inc, 0); else egk_sign(a - 14, 0, c[i] < 0, &bits, &len);
            gt1++;
        } else eq1++;
        put_bypass(s, bits, len);
    }
}

/* residual_block_cabac for ctxBlockCat 5 (8x8 luma block, 64 levels): no coded_block_flag (inferred 1 with the cbp bit); the
 * significance map contexts come from Table 9-43 */
static void put_residual8(Bins *s, const int16_t *c)
{
    int last = -1;
    for (int i = 0; i < 64; i++) if (c[i]) last = i;
    for (int i = 0; i < 63; i++) {
        put(s, 402 + CABAC_SIG8[i], c[i] != 0);
        if (c[i]) { put(s, 417 + CABAC_LAST8[i], i == last); if (i == last) break; }
    }
    int eq1 = 0, gt1 = 0;
    for (int i = last; i >= 0; i--) {
        if (!c[i]) continue;
        int a = iabs(c[i]) - 1, base = 426;
        put(s, base + (gt1 ? 0 : imin(4, 1 + eq1)), a > 0);
        uint32_t bits = (uint32_t)(c[i] < 0); int len = 1;
        if (a > 0) {
            int inc = 5 + imin(4, gt1);
            if (imin(a, 14) > 1) put_run(s, base + inc, 1, imin(a, 14) - 1);
            if (a < 14) put(s, base + inc, 0); else egk_sign(a - 14, 0, c[i] < 0, &bits, &len);
            gt1++;
        } else eq1++;
        put_bypass(s, bits, len);
    }
}

#define IS_INTRA(m) ORC_MB_IS_INTRA(m)
static const uint8_t XY2BLK[4][4] = { { 0, 1, 4, 5 }, { 2, 3, 6, 7 }, { 8, 9, 12, 13 }, { 10, 11, 14, 15 } };
static const uint8_t BX[16] = { 0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3 }, BY[16] = { 0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3 };

/* coded_block_flag of the neighbouring MB's block for the ctxIdxInc of 9.3.3.1.1.9; nb == NULL: MB not available */
static int nb_cbf_luma(const OrcMbInfo *nb, int blk, int cur_intra) { return nb ? nb->nnz[blk] != 0 : cur_intra; }
static int nb_cbf_cac(const OrcMbInfo *nb, int idx, int cur_intra) { return nb ? ((nb->cbp >> 4) == 2 && nb->nnz[idx] != 0) : cur_intra; }

/* Bins of one macroblock; mb_skip_flag and end_of_slice_flag included. last = 1 for the last MB of the slice. */
int orc_cabac_mb_bins(const OrcMbInfo *mbi, const OrcMbCoef *coef, const OrcMbSide *side, int mbw, int mx, int my, int top_avail,
                      int is_p, int last, int transform8x8, uint16_t *out, int cap)
{
    Bins s = { out, 0, cap };
    int mb = my * mbw + mx;
    const OrcMbInfo *m = &mbi[mb], *L = mx > 0 ? m - 1 : 0, *T = top_avail ? m - mbw : 0;
    const OrcMbSide *sd = &side[mb], *sL = L ? sd - 1 : 0, *sT = T ? sd - mbw : 0;
    const OrcMbCoef *co = &coef[mb];
    int cl = m->cbp & 15, cc = m->cbp >> 4, intra = IS_INTRA(m);
    if (is_p) {
        put(&s, 11 + (L && L->mb_type != ORC_MB_PSKIP) + (T && T->mb_type != ORC_MB_PSKIP), m->mb_type == ORC_MB_PSKIP);
        if (m->mb_type == ORC_MB_PSKIP) { put(&s, 276, last); return s.n; }
    }
    /* mb_type (9.3.2.5, ctxIdx per Table 9-39) */
    if (!intra) {
        put(&s, 14, 0); put(&s, 15, 0); put(&s, 16, m->mb_type == ORC_MB_P8x8);
        if (m->mb_type == ORC_MB_P8x8) for (int q = 0; q < 4; q++) put(&s, 21, 1);         /* sub_mb_type P_L0_8x8 */
    } else {
        int b0, c_cl, c_cc, c_cc2, c_m1, c_m0;
        if (is_p) { put(&s, 14, 1); b0 = 17; c_cl = 18; c_cc = 19; c_cc2 = 19; c_m1 = 20; c_m0 = 20; }
        else { b0 = 3 + (L && !ORC_MB_IS_INXN(L)) + (T && !ORC_MB_IS_INXN(T)); c_cl = 6; c_cc = 7; c_cc2 = 8; c_m1 = 9; c_m0 = 10; }
        put(&s, b0, m->mb_type == ORC_MB_I16x16);
        if (m->mb_type == ORC_MB_I16x16) {
            put(&s, 276, 0);                                                               /* not I_PCM */
            put(&s, c_cl, cl != 0); put(&s, c_cc, cc != 0);
            if (cc) put(&s, c_cc2, cc == 2);
            put(&s, c_m1, m->i16_mode >> 1); put(&s, c_m0, m->i16_mode & 1);
        }
    }
    /* transform_size_8x8_flag (7.3.5, ctxIdx 399 + the neighbours' flags, 9.3.3.1.1.10): 0 for I_NxN (Intra_4x4 only) */
    int t8inc = (L && ORC_MB_T8(L)) + (T && ORC_MB_T8(T));
    if (ORC_MB_IS_INXN(m) && transform8x8) put(&s, 399 + t8inc, m->mb_type == ORC_MB_I8x8);
    if (m->mb_type == ORC_MB_I8x8)
        for (int k = 0; k < 4; k++) {                                                      /* prev_intra8x8_pred_mode_flag / rem_intra8x8_pred_mode: same contexts */
            int r = sd->i4_syn[k];
            put(&s, 68, r == 8);
            if (r != 8) { put(&s, 69, r & 1); put(&s, 69, (r >> 1) & 1); put(&s, 69, (r >> 2) & 1); }
        }
    if (m->mb_type == ORC_MB_I4x4)
        for (int k = 0; k < 16; k++) {                                                     /* prev_intra4x4_pred_mode_flag / rem_intra4x4_pred_mode */
            int r = sd->i4_syn[k];
            put(&s, 68, r == 8);
            if (r != 8) { put(&s, 69, r & 1); put(&s, 69, (r >> 1) & 1); put(&s, 69, (r >> 2) & 1); }
        }
    if (intra) {                                                                           /* intra_chroma_pred_mode, TU cMax 3 */
        int inc = (L && IS_INTRA(L) && L->chroma_mode != 0) + (T && IS_INTRA(T) && T->chroma_mode != 0), cm = m->chroma_mode;
        put(&s, 64 + inc, cm != 0);
        if (cm) { put(&s, 67, cm != 1); if (cm != 1) put(&s, 67, cm != 2); }
    }
    /* motion vector differences: mvd of the left / upper neighbouring partitions select the context (9.3.3.1.1.7) */
    if (!intra) {
        int np = m->mb_type == ORC_MB_P8x8 ? 4 : 1;
        for (int q = 0; q < np; q++)
            for (int c = 0; c < 2; c++) {
                int a, b;
                if (q & 1) a = iabs(sd->mvd[q - 1][c]); else a = L && !IS_INTRA(L) ? iabs(sL->mvd[q + 1][c]) : 0;   /* the union holds i4_syn for intra MBs */
                if (q & 2) b = iabs(sd->mvd[q - 2][c]); else b = T && !IS_INTRA(T) ? iabs(sT->mvd[q + 2][c]) : 0;
                put_mvd(&s, c ? 47 : 40, a + b, sd->mvd[q][c]);
            }
    }
    /* coded_block_pattern (9.3.2.6, 9.3.3.1.1.4) */
    if (m->mb_type != ORC_MB_I16x16) {
        for (int b8 = 0; b8 < 4; b8++) {
            int a, b;
            if (b8 & 1) a = !((cl >> (b8 - 1)) & 1); else a = L ? !((L->cbp >> (b8 + 1)) & 1) : 0;
            if (b8 & 2) b = !((cl >> (b8 - 2)) & 1); else b = T ? !((T->cbp >> (b8 + 2)) & 1) : 0;
            put(&s, 73 + a + 2 * b, (cl >> b8) & 1);
        }
        put(&s, 77 + (L && (L->cbp >> 4)) + 2 * (T && (T->cbp >> 4)), cc != 0);
        if (cc) put(&s, 81 + (L && (L->cbp >> 4) == 2) + 2 * (T && (T->cbp >> 4) == 2), cc == 2);
    }
    if (!intra && cl && transform8x8) put(&s, 399 + t8inc, ORC_MB_T8(m));                  /* every inter MB here has 8x8 partitions at least */
    if (m->mb_type == ORC_MB_I16x16 || m->cbp) put(&s, 60, 0);                             /* mb_qp_delta = 0; the previous MB's is 0 too */
    /* residual */
    if (m->mb_type == ORC_MB_I16x16) {
        int a = L ? (L->mb_type == ORC_MB_I16x16 && (sL->dc_cbf & 1)) : 1, b = T ? (T->mb_type == ORC_MB_I16x16 && (sT->dc_cbf & 1)) : 1;
        put_residual(&s, co->luma_dc, 16, 0, a + 2 * b);
    }
    if (ORC_MB_T8(m)) {
        for (int b8 = 0; b8 < 4; b8++) if (cl & (1 << b8)) put_residual8(&s, co->luma[4 * b8]);
    } else
    for (int k = 0; k < 16; k++) {
        if (!(cl & (1 << (k >> 2)))) continue;
        int bx = BX[k], by = BY[k];
        int a = bx ? m->nnz[XY2BLK[by][bx - 1]] != 0 : nb_cbf_luma(L, XY2BLK[by][3], intra);
        int b = by ? m->nnz[XY2BLK[by - 1][bx]] != 0 : nb_cbf_luma(T, XY2BLK[3][bx], intra);
        if (m->mb_type == ORC_MB_I16x16) put_residual(&s, co->luma[k] + 1, 15, 1, a + 2 * b);
        else put_residual(&s, co->luma[k], 16, 2, a + 2 * b);
    }
    if (cc) for (int p = 0; p < 2; p++) {
        int a = L ? ((L->cbp >> 4) && ((sL->dc_cbf >> (1 + p)) & 1)) : intra, b = T ? ((T->cbp >> 4) && ((sT->dc_cbf >> (1 + p)) & 1)) : intra;
        put_residual(&s, co->chroma_dc[p], 4, 3, a + 2 * b);
    }
    if (cc == 2) for (int p = 0; p < 2; p++) for (int k = 0; k < 4; k++) {
        int bx = k & 1, by = k >> 1, base = 16 + 4 * p;
        int a = bx ? m->nnz[base + k - 1] != 0 : nb_cbf_cac(L, base + by * 2 + 1, intra);
        int b = by ? m->nnz[base + k - 2] != 0 : nb_cbf_cac(T, base + 2 + bx, intra);
        put_residual(&s, co->chroma_ac[p][k] + 1, 15, 4, a + 2 * b);
    }
    put(&s, 276, last);                                                                    /* end_of_slice_flag */
    return s.n;
}

/* ---- arithmetic encoder, 9.3.4.2, written the way the standard's flow charts are ---- */
typedef struct { BitWriter *b; uint32_t low, range; int first, outstanding; uint8_t state[ORC_CABAC_NCTX], mps[ORC_CABAC_NCTX]; } Coder;
static void put_bit(Coder *c, int v)
{
    if (c->first) c->first = 0; else bw_put(c->b, 1, (uint32_t)v);
    while (c->outstanding > 0) { bw_put(c->b, 1, (uint32_t)(1 - v)); c->outstanding--; }
}
static void renorm(Coder *c)
{
    while (c->range < 256) {
        if (c->low < 256) put_bit(c, 0);
        else if (c->low >= 512) { c->low -= 512; put_bit(c, 1); }
        else { c->low -= 256; c->outstanding++; }
        c->range <<= 1; c->low <<= 1;
    }
}
/* Codes a bin list into the bit writer (which must be byte aligned: cabac_alignment_one_bit is the caller's). The list ends with a
 * terminate bin of value 1; its flush writes the rbsp_stop_one_bit (9.3.4.5), the caller pads to the byte. */
void orc_cabac_code(BitWriter *bw, const uint16_t *bins, int n, int slice_qp, int is_p)
{
    Coder c = { bw, 0, 510, 1, 0, { 0 }, { 0 } };
    const int8_t *init = is_p ? CABAC_INIT_P0 : CABAC_INIT_I;
    int q = slice_qp < 0 ? 0 : slice_qp > 51 ? 51 : slice_qp;
    for (int i = 0; i < ORC_CABAC_NCTX; i++) {                                             /* 9.3.1.1 */
        int pre = ((init[2 * i] * q) >> 4) + init[2 * i + 1];
        pre = pre < 1 ? 1 : pre > 126 ? 126 : pre;
        if (pre <= 63) { c.state[i] = (uint8_t)(63 - pre); c.mps[i] = 0; } else { c.state[i] = (uint8_t)(pre - 64); c.mps[i] = 1; }
    }
    for (int i = 0; i < n; i++) {
        int ctx = bins[i] & 1023, bin = (bins[i] >> 10) & 1;
        if (ctx > ORC_CTX_BYPASS0) {
            int nb = ctx - ORC_CTX_BYPASS0;
            for (int k = nb - 1; k >= 0; k--) {
                c.low <<= 1;
                if ((bins[i] >> (10 + k)) & 1) c.low += c.range;
                if (c.low >= 1024) { put_bit(&c, 1); c.low -= 1024; }
                else if (c.low < 512) put_bit(&c, 0);
                else { c.low -= 512; c.outstanding++; }
            }
        } else if (ctx == 276) {
            c.range -= 2;
            if (bin) {
                c.low += c.range; c.range = 2; renorm(&c);
                put_bit(&c, (c.low >> 9) & 1);
                bw_put(c.b, 2, ((c.low >> 7) & 3) | 1);
            } else renorm(&c);
        } else for (int rep = bins[i] >> 11; rep >= 0; rep--) {
            uint32_t rlps = CABAC_RANGE_LPS[c.state[ctx] * 4 + ((c.range >> 6) & 3)];
            c.range -= rlps;
            if (bin != c.mps[ctx]) {
                c.low += c.range; c.range = rlps;
                if (c.state[ctx] == 0) c.mps[ctx] ^= 1;
                c.state[ctx] = CABAC_NEXT_LPS[c.state[ctx]];
            } else c.state[ctx] = CABAC_NEXT_MPS[c.state[ctx]];
            renorm(&c);
        }
    }
}
