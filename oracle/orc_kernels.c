/*
 * oracle/orc_kernels.c -- TEST INFRASTRUCTURE ONLY (see orc.h for scope and pinning status).
 *
 * Scalar, obviously-correct restatements of the pixel kernels on the encode path. Each function
 * names the openh264 routine that plays the same role inside the absent libopenh264.so (names from
 * SURVEY.md section 2b, unverifiable here) and the H.264 clause that defines the normative ones.
 */
#include "orc.h"
#include "h264_tables.h"
#include <stdlib.h>
#include <string.h>

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* ---- SAD / SATD (role of WelsSampleSad*_c / WelsSampleSatd*_c) ---- */
int orc_sad(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h)
{
    int s = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            s += iabs((int)a[y * sa + x] - (int)b[y * sb + x]);
    return s;
}

/* 4x4 Hadamard SATD: sum |H d H^T| / 2 (the sum is always even). */
int orc_satd4x4(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
    int d[16], t[16], s = 0;
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
            d[y * 4 + x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    for (int y = 0; y < 4; y++) {
        int a0 = d[y * 4 + 0] + d[y * 4 + 1], a1 = d[y * 4 + 0] - d[y * 4 + 1];
        int a2 = d[y * 4 + 2] + d[y * 4 + 3], a3 = d[y * 4 + 2] - d[y * 4 + 3];
        t[y * 4 + 0] = a0 + a2; t[y * 4 + 1] = a1 + a3; t[y * 4 + 2] = a0 - a2; t[y * 4 + 3] = a1 - a3;
    }
    for (int x = 0; x < 4; x++) {
        int a0 = t[x] + t[4 + x], a1 = t[x] - t[4 + x];
        int a2 = t[8 + x] + t[12 + x], a3 = t[8 + x] - t[12 + x];
        s += iabs(a0 + a2) + iabs(a1 + a3) + iabs(a0 - a2) + iabs(a1 - a3);
    }
    return s >> 1;
}

int orc_satd16x16(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
    int s = 0;
    for (int y = 0; y < 16; y += 4)
        for (int x = 0; x < 16; x += 4)
            s += orc_satd4x4(a + y * sa + x, sa, b + y * sb + x, sb);
    return s;
}

/* ---- forward core transform (role of WelsDctT4_c); encoder side of 8.5.12 ---- */
void orc_dct4x4(const int16_t *r, int16_t *c)
{
    int t[16];
    for (int y = 0; y < 4; y++) {
        int a0 = r[y * 4 + 0] + r[y * 4 + 3], a1 = r[y * 4 + 1] + r[y * 4 + 2];
        int a2 = r[y * 4 + 1] - r[y * 4 + 2], a3 = r[y * 4 + 0] - r[y * 4 + 3];
        t[y * 4 + 0] = a0 + a1; t[y * 4 + 1] = 2 * a3 + a2; t[y * 4 + 2] = a0 - a1; t[y * 4 + 3] = a3 - 2 * a2;
    }
    for (int x = 0; x < 4; x++) {
        int a0 = t[x] + t[12 + x], a1 = t[4 + x] + t[8 + x];
        int a2 = t[4 + x] - t[8 + x], a3 = t[x] - t[12 + x];
        c[x] = (int16_t)(a0 + a1); c[4 + x] = (int16_t)(2 * a3 + a2);
        c[8 + x] = (int16_t)(a0 - a1); c[12 + x] = (int16_t)(a3 - 2 * a2);
    }
}

/* ---- inverse core transform, 8.5.12.2: rows then columns, (x+32)>>6 (role of WelsIDctT4Rec_c) ---- */
void orc_idct4x4(const int32_t *d, int32_t *r)
{
    int f[16];
    for (int y = 0; y < 4; y++) {
        int e0 = d[y * 4 + 0] + d[y * 4 + 2], e1 = d[y * 4 + 0] - d[y * 4 + 2];
        int e2 = (d[y * 4 + 1] >> 1) - d[y * 4 + 3], e3 = d[y * 4 + 1] + (d[y * 4 + 3] >> 1);
        f[y * 4 + 0] = e0 + e3; f[y * 4 + 1] = e1 + e2; f[y * 4 + 2] = e1 - e2; f[y * 4 + 3] = e0 - e3;
    }
    for (int x = 0; x < 4; x++) {
        int g0 = f[x] + f[8 + x], g1 = f[x] - f[8 + x];
        int g2 = (f[4 + x] >> 1) - f[12 + x], g3 = f[4 + x] + (f[12 + x] >> 1);
        r[x] = (g0 + g3 + 32) >> 6; r[4 + x] = (g1 + g2 + 32) >> 6;
        r[8 + x] = (g1 - g2 + 32) >> 6; r[12 + x] = (g0 - g3 + 32) >> 6;
    }
}

/* ---- quantiser (role of WelsQuant4x4_c): level = (|c|*MF + f) >> (15+qp/6), f = 2^qbits/3 intra, /6 inter.
 * Levels are clipped to +-2063 so that CAVLC level_prefix never exceeds 15 (Baseline). ---- */
#define ORC_MAX_LEVEL 2063
int orc_quant4x4(const int16_t *coef, int16_t *lz, int qp, int intra, int ac_only)
{
    int qbits = 15 + qp / 6, m = qp % 6, nnz = 0;
    int f = (1 << qbits) / (intra ? 3 : 6);
    for (int i = 0; i < 16; i++) {
        int pos = ZIGZAG4x4[i];
        if (ac_only && i == 0) { lz[0] = 0; continue; }
        int c = coef[pos];
        int l = (int)(((int64_t)iabs(c) * QUANT_MF[m][POS_CLASS[pos]] + f) >> qbits);
        if (l > ORC_MAX_LEVEL) l = ORC_MAX_LEVEL;
        lz[i] = (int16_t)(c < 0 ? -l : l);
        nnz += (l != 0);
    }
    return nnz;
}

/* ---- dequantiser, 8.5.12.1 with flat scaling lists: d = level * V << (qp/6) (role of WelsDequant4x4_c) ---- */
void orc_dequant4x4(const int16_t *lz, int32_t *d, int qp, int ac_only)
{
    int sh = qp / 6, m = qp % 6;
    for (int i = 0; i < 16; i++) {
        int pos = ZIGZAG4x4[i];
        if (ac_only && i == 0) { d[pos] = 0; continue; }
        d[pos] = (lz[i] * DEQUANT_V[m][POS_CLASS[pos]]) << sh;
    }
}

/* ---- 8x8 transform of the High profile (transform_size_8x8_flag = 1) ----
 * forward: the integer 8x8 transform matching 8.5.13's inverse (rows then columns; encoder side, the JM / x264 butterfly) */
static void fdct8_1d(const int *x, int *y)
{
    int a0 = x[0] + x[7], a1 = x[1] + x[6], a2 = x[2] + x[5], a3 = x[3] + x[4];
    int a4 = x[0] - x[7], a5 = x[1] - x[6], a6 = x[2] - x[5], a7 = x[3] - x[4];
    int b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2;
    int b4 = a5 + a6 + ((a4 >> 1) + a4), b5 = a4 - a7 - ((a6 >> 1) + a6), b6 = a4 + a7 - ((a5 >> 1) + a5), b7 = a5 - a6 + ((a7 >> 1) + a7);
    y[0] = b0 + b1; y[1] = b4 + (b7 >> 2); y[2] = b2 + (b3 >> 1); y[3] = b5 + (b6 >> 2);
    y[4] = b0 - b1; y[5] = b6 - (b5 >> 2); y[6] = (b2 >> 1) - b3; y[7] = (b4 >> 2) - b7;
}
void orc_dct8x8(const int16_t *r, int32_t *c)
{
    int t[64], in[8], out[8];
    for (int y = 0; y < 8; y++) { for (int x = 0; x < 8; x++) in[x] = r[y * 8 + x]; fdct8_1d(in, out); for (int x = 0; x < 8; x++) t[y * 8 + x] = out[x]; }
    for (int x = 0; x < 8; x++) { for (int y = 0; y < 8; y++) in[y] = t[y * 8 + x]; fdct8_1d(in, out); for (int y = 0; y < 8; y++) c[y * 8 + x] = out[y]; }
}
/* quantise raster coefficients -> 64 levels in 8x8 zig-zag order (JM: (|c| * MF + f) >> (16 + qp/6), dead zone 1/3 intra, 1/6 inter,
 * levels clipped to +-2063 like the 4x4 quantiser). Returns the number of nonzero levels. */
int orc_quant8x8(const int32_t *c, int16_t *lz, int qp, int intra)
{
    int qbits = 16 + qp / 6, m = qp % 6, n = 0; int64_t f = ((int64_t)1 << qbits) / (intra ? 3 : 6);
    for (int i = 0; i < 64; i++) {
        int pos = ZIGZAG8x8[i], v = c[pos];
        int l = (int)(((int64_t)iabs(v) * QUANT8_MF[m][pos_class8(pos >> 3, pos & 7)] + f) >> qbits);
        if (l > 2063) l = 2063;
        lz[i] = (int16_t)(v < 0 ? -l : l); n += l != 0;
    }
    return n;
}
/* 8.5.13 scaling (flat_8x8_16 weights): LevelScale8x8 = 16 * normAdjust8x8 */
void orc_dequant8x8(const int16_t *lz, int32_t *d, int qp)
{
    int sh = qp / 6, m = qp % 6;
    for (int i = 0; i < 64; i++) {
        int pos = ZIGZAG8x8[i], ls = 16 * DEQUANT8_V[m][pos_class8(pos >> 3, pos & 7)];
        d[pos] = qp >= 36 ? (lz[i] * ls) << (sh - 6) : (lz[i] * ls + (1 << (5 - sh))) >> (6 - sh);
    }
}
/* 8.5.13 inverse transform: each row, then each column, then (x + 32) >> 6 */
static void idct8_1d(const int *d, int *o)
{
    int a0 = d[0] + d[4], a1 = -d[3] + d[5] - d[7] - (d[7] >> 1), a2 = d[0] - d[4], a3 = d[1] + d[7] - d[3] - (d[3] >> 1);
    int a4 = (d[2] >> 1) - d[6], a5 = -d[1] + d[7] + d[5] + (d[5] >> 1), a6 = d[2] + (d[6] >> 1), a7 = d[3] + d[5] + d[1] + (d[1] >> 1);
    int b0 = a0 + a6, b1 = a1 + (a7 >> 2), b2 = a2 + a4, b3 = a3 + (a5 >> 2), b4 = a2 - a4, b5 = (a3 >> 2) - a5, b6 = a0 - a6, b7 = a7 - (a1 >> 2);
    o[0] = b0 + b7; o[1] = b2 + b5; o[2] = b4 + b3; o[3] = b6 + b1; o[4] = b6 - b1; o[5] = b4 - b3; o[6] = b2 - b5; o[7] = b0 - b7;
}
void orc_idct8x8(const int32_t *d, int32_t *r)
{
    int t[64], in[8], out[8];
    for (int y = 0; y < 8; y++) { for (int x = 0; x < 8; x++) in[x] = d[y * 8 + x]; idct8_1d(in, out); for (int x = 0; x < 8; x++) t[y * 8 + x] = out[x]; }
    for (int x = 0; x < 8; x++) { for (int y = 0; y < 8; y++) in[y] = t[y * 8 + x]; idct8_1d(in, out); for (int y = 0; y < 8; y++) r[y * 8 + x] = (out[y] + 32) >> 6; }
}

/* ---- colour conversion: BT.601 limited range, 8-bit fixed point; chroma from the 2x2 mean RGB.
 * No reference function computes this (the reference only accepts I420,
 * VideoEncoderOpenH264.cpp:256,262,358): definition is ours, see SURVEY.md 8c. ---- */
void orc_rgba_to_i420(const uint8_t *rgba, int w, int h, uint8_t *out)
{
    uint8_t *Y = out, *U = out + w * h, *V = U + (w / 2) * (h / 2);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t *p = rgba + 4 * (y * w + x);
            Y[y * w + x] = (uint8_t)(((66 * p[0] + 129 * p[1] + 25 * p[2] + 128) >> 8) + 16);
        }
    for (int y = 0; y < h / 2; y++)
        for (int x = 0; x < w / 2; x++) {
            int r = 0, g = 0, b = 0;
            for (int dy = 0; dy < 2; dy++)
                for (int dx = 0; dx < 2; dx++) {
                    const uint8_t *p = rgba + 4 * ((2 * y + dy) * w + 2 * x + dx);
                    r += p[0]; g += p[1]; b += p[2];
                }
            r = (r + 2) >> 2; g = (g + 2) >> 2; b = (b + 2) >> 2;
            U[y * (w / 2) + x] = (uint8_t)(((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128);
            V[y * (w / 2) + x] = (uint8_t)(((112 * r - 94 * g - 18 * b + 128) >> 8) + 128);
        }
}

void orc_nv12_to_i420(const uint8_t *nv12, int w, int h, uint8_t *out)
{
    memcpy(out, nv12, (size_t)w * h);
    const uint8_t *uv = nv12 + w * h;
    uint8_t *U = out + w * h, *V = U + (w / 2) * (h / 2);
    for (int i = 0; i < (w / 2) * (h / 2); i++) { U[i] = uv[2 * i]; V[i] = uv[2 * i + 1]; }
}

/* ---- 2x2 box downsample with rounding (role of DyadicBilinearDownsampler_c) ---- */
void orc_downsample2(const uint8_t *s, int ss, int w, int h, uint8_t *d, int ds)
{
    for (int y = 0; y < h / 2; y++)
        for (int x = 0; x < w / 2; x++)
            d[y * ds + x] = (uint8_t)((s[2 * y * ss + 2 * x] + s[2 * y * ss + 2 * x + 1] +
                                       s[(2 * y + 1) * ss + 2 * x] + s[(2 * y + 1) * ss + 2 * x + 1] + 2) >> 2);
}

/* ---- luma sub-pel interpolation, 8.4.2.2.1 (role of McHorVer20/02/22_c + PixelAvg_c) ---- */
static inline int refpx(const uint8_t *r, int st, int w, int h, int x, int y)
{
    return r[clip3(0, h - 1, y) * st + clip3(0, w - 1, x)];
}
static int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }
static int half_h_raw(const uint8_t *r, int st, int w, int h, int x, int y)   /* b1 at (x+1/2, y) */
{
    return tap6(refpx(r, st, w, h, x - 2, y), refpx(r, st, w, h, x - 1, y), refpx(r, st, w, h, x, y),
                refpx(r, st, w, h, x + 1, y), refpx(r, st, w, h, x + 2, y), refpx(r, st, w, h, x + 3, y));
}
static int half_v_raw(const uint8_t *r, int st, int w, int h, int x, int y)   /* h1 at (x, y+1/2) */
{
    return tap6(refpx(r, st, w, h, x, y - 2), refpx(r, st, w, h, x, y - 1), refpx(r, st, w, h, x, y),
                refpx(r, st, w, h, x, y + 1), refpx(r, st, w, h, x, y + 2), refpx(r, st, w, h, x, y + 3));
}
static int half_h(const uint8_t *r, int st, int w, int h, int x, int y) { return clip255((half_h_raw(r, st, w, h, x, y) + 16) >> 5); }
static int half_v(const uint8_t *r, int st, int w, int h, int x, int y) { return clip255((half_v_raw(r, st, w, h, x, y) + 16) >> 5); }
static int half_c(const uint8_t *r, int st, int w, int h, int x, int y)       /* j at (x+1/2, y+1/2) */
{
    int j1 = tap6(half_h_raw(r, st, w, h, x, y - 2), half_h_raw(r, st, w, h, x, y - 1), half_h_raw(r, st, w, h, x, y),
                  half_h_raw(r, st, w, h, x, y + 1), half_h_raw(r, st, w, h, x, y + 2), half_h_raw(r, st, w, h, x, y + 3));
    return clip255((j1 + 512) >> 10);
}

int orc_interp_luma(const uint8_t *r, int st, int w, int h, int xq, int yq)
{
    int x = xq >> 2, y = yq >> 2, fx = xq & 3, fy = yq & 3;
#define G_ refpx(r, st, w, h, x, y)
#define H_ refpx(r, st, w, h, x + 1, y)
#define M_ refpx(r, st, w, h, x, y + 1)
#define b_ half_h(r, st, w, h, x, y)
#define s_ half_h(r, st, w, h, x, y + 1)
#define h_ half_v(r, st, w, h, x, y)
#define m_ half_v(r, st, w, h, x + 1, y)
#define j_ half_c(r, st, w, h, x, y)
    switch (fy * 4 + fx) {
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

/* ---- chroma 1/8-pel bilinear, 8.4.2.2.2 (role of McChroma_c) ---- */
int orc_interp_chroma(const uint8_t *r, int st, int w, int h, int x8, int y8)
{
    int x = x8 >> 3, y = y8 >> 3, fx = x8 & 7, fy = y8 & 7;
    int A = refpx(r, st, w, h, x, y), B = refpx(r, st, w, h, x + 1, y);
    int C = refpx(r, st, w, h, x, y + 1), D = refpx(r, st, w, h, x + 1, y + 1);
    return ((8 - fx) * (8 - fy) * A + fx * (8 - fy) * B + (8 - fx) * fy * C + fx * fy * D + 32) >> 6;
}

/* ---- in-loop deblocking, 8.7 (roles: DeblockingBSCalcEnc_c, DeblockLumaLt4/Eq4, DeblockChromaLt4/Eq4) ---- */
static void filter_luma(uint8_t *p, int step /* across the edge */, int bs, int alpha, int beta, int tc0)
{
    int p0 = p[-step], p1 = p[-2 * step], p2 = p[-3 * step], p3 = p[-4 * step];
    int q0 = p[0], q1 = p[step], q2 = p[2 * step], q3 = p[3 * step];
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    int ap = iabs(p2 - p0), aq = iabs(q2 - q0);
    if (bs < 4) {
        int tc = tc0 + (ap < beta) + (aq < beta);
        int delta = clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        p[-step] = (uint8_t)clip255(p0 + delta);
        p[0] = (uint8_t)clip255(q0 - delta);
        if (ap < beta) p[-2 * step] = (uint8_t)(p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1));
        if (aq < beta) p[step] = (uint8_t)(q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1));
    } else {
        int small = iabs(p0 - q0) < ((alpha >> 2) + 2);
        if (ap < beta && small) {
            p[-step] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            p[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            p[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else {
            p[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        }
        if (aq < beta && small) {
            p[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            p[step] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
            p[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else {
            p[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
        }
    }
}

static void filter_chroma(uint8_t *p, int step, int bs, int alpha, int beta, int tc0)
{
    int p0 = p[-step], p1 = p[-2 * step], q0 = p[0], q1 = p[step];
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    if (bs < 4) {
        int tc = tc0 + 1;
        int delta = clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        p[-step] = (uint8_t)clip255(p0 + delta);
        p[0] = (uint8_t)clip255(q0 - delta);
    } else {
        p[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        p[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}

static int mb_is_intra(const OrcMbInfo *m) { return ORC_MB_IS_INTRA(m); }

/* bS between 4x4 block (bxq,byq) of MB q and the adjacent block (bxp,byp) of MB p; mb_edge = on a MB boundary */
static int boundary_strength(const OrcMbInfo *mp, int bxp, int byp, const OrcMbInfo *mq, int bxq, int byq, int mb_edge)
{
    static const uint8_t XY2BLK[4][4] = { { 0, 1, 4, 5 }, { 2, 3, 6, 7 }, { 8, 9, 12, 13 }, { 10, 11, 14, 15 } };
    if (mb_is_intra(mp) || mb_is_intra(mq)) return mb_edge ? 4 : 3;
    if (mp->nnz[XY2BLK[byp][bxp]] || mq->nnz[XY2BLK[byq][bxq]]) return 2;
    const int16_t *vp = mp->mv8[(byp >> 1) * 2 + (bxp >> 1)], *vq = mq->mv8[(byq >> 1) * 2 + (bxq >> 1)];
    if (iabs(vp[0] - vq[0]) >= 4 || iabs(vp[1] - vq[1]) >= 4) return 1;
    return 0;
}

void orc_deblock_frame(uint8_t *Y, int ys, uint8_t *U, uint8_t *V, int cs, int mbw, int mbh,
                       const OrcMbInfo *mbi, int qp)
{
    int qpc = CHROMA_QP[qp];
    int alphaY = DEBLOCK_ALPHA[qp], betaY = DEBLOCK_BETA[qp];
    int alphaC = DEBLOCK_ALPHA[qpc], betaC = DEBLOCK_BETA[qpc];
    for (int my = 0; my < mbh; my++)
        for (int mx = 0; mx < mbw; mx++) {
            const OrcMbInfo *q = &mbi[my * mbw + mx];
            uint8_t *py = Y + my * 16 * ys + mx * 16;
            uint8_t *pc[2] = { U + my * 8 * cs + mx * 8, V + my * 8 * cs + mx * 8 };
            int bsv[4][4], bsh[4][4];          /* [edge][segment] */
            for (int e = 0; e < 4; e++)
                for (int k = 0; k < 4; k++) {
                    if (e == 0) {
                        bsv[0][k] = mx > 0 ? boundary_strength(q - 1, 3, k, q, 0, k, 1) : 0;
                        bsh[0][k] = my > 0 ? boundary_strength(q - mbw, k, 3, q, k, 0, 1) : 0;
                    } else if ((e & 1) && ORC_MB_T8(q)) {
                        bsv[e][k] = bsh[e][k] = 0;     /* transform_size_8x8_flag: only the 8x8 transform block edges are filtered (8.7) */
                    } else {
                        bsv[e][k] = boundary_strength(q, e - 1, k, q, e, k, 0);
                        bsh[e][k] = boundary_strength(q, k, e - 1, q, k, e, 0);
                    }
                }
            /* luma: vertical edges left to right, then horizontal edges top to bottom */
            for (int e = 0; e < 4; e++)
                for (int r = 0; r < 16; r++) {
                    int bs = bsv[e][r >> 2];
                    if (bs) filter_luma(py + r * ys + 4 * e, 1, bs, alphaY, betaY, bs < 4 ? DEBLOCK_TC0[qp][bs - 1] : 0);
                }
            for (int e = 0; e < 4; e++)
                for (int c = 0; c < 16; c++) {
                    int bs = bsh[e][c >> 2];
                    if (bs) filter_luma(py + 4 * e * ys + c, ys, bs, alphaY, betaY, bs < 4 ? DEBLOCK_TC0[qp][bs - 1] : 0);
                }
            /* chroma: edges 0 and 2 of the luma grid map to chroma columns/rows 0 and 4 */
            for (int pl = 0; pl < 2; pl++) {
                for (int e = 0; e < 4; e += 2)
                    for (int r = 0; r < 8; r++) {
                        int bs = bsv[e][r >> 1];
                        if (bs) filter_chroma(pc[pl] + r * cs + 2 * e, 1, bs, alphaC, betaC, bs < 4 ? DEBLOCK_TC0[qpc][bs - 1] : 0);
                    }
                for (int e = 0; e < 4; e += 2)
                    for (int c = 0; c < 8; c++) {
                        int bs = bsh[e][c >> 1];
                        if (bs) filter_chroma(pc[pl] + 2 * e * cs + c, cs, bs, alphaC, betaC, bs < 4 ? DEBLOCK_TC0[qpc][bs - 1] : 0);
                    }
            }
        }
}

/* ---- emulation prevention (7.4.1): 00 00 0x -> 00 00 03 0x for x <= 3 ---- */
int orc_escape_rbsp(const uint8_t *in, int n, uint8_t *out)
{
    int o = 0, zeros = 0;
    for (int i = 0; i < n; i++) {
        if (zeros >= 2 && in[i] <= 3) { out[o++] = 3; zeros = 0; }
        out[o++] = in[i];
        zeros = in[i] == 0 ? zeros + 1 : 0;
    }
    return o;
}
