// media_b200/csrc/k_cabac.cuh -- CABAC entropy coding (ITU-T H.264 7.3.4, 7.3.5, 9.3) for Main / High profile streams.
//
// Role inside the reference: the wrapper asks openh264 for iEntropyCodingModeFlag = 1 (video_codec/VideoEncoderOpenH264.cpp:291)
// whenever the profile property is main or high (:248-253); openh264's WelsSpatialWriteMbSynCabac / WelsCabacEncodeDecision
// live in the absent libopenh264. CABAC is normative end to end, so the oracle (oracle/orc_cabac.c) follows the standard's
// flow charts and this file must produce the same bytes.
//
// The arithmetic coder is a serial chain over a slice's bins, but which bins there are and which context each one uses only
// depends on a macroblock and its left / upper neighbours. So:
//   k_cabac_side  : per-MB side record: mvd per 8x8 partition, Intra_4x4 mode syntax, DC coded_block_flags      [thread per MB]
//   k_cabac_bins  : binarisation + context selection of a whole MB into 16-bit entries; 29 lanes = 29 syntax groups
//                   (header, luma DC, 16 luma, 2 chroma DC, 8 chroma AC, end_of_slice). COUNT pass: entries per MB;
//                   WRITE pass: entries stored at the MB's offset in the slice's bin list                         [warp per MB]
//   k_cabac_scan  : prefix sum of the per-MB entry counts inside a slice                                           [CTA per slice]
//   k_cabac_code  : slice header + the arithmetic coder over the slice's bin list: a producer warp resolves the per-context
//                   probability states 32 entries at a time, one thread runs the range/low recurrence, one resolves carries and writes bytes [3 warps per slice]
// Entry format (shared with the oracle, oracle/orc.h): ctxIdx | bin << 10 | (repeat - 1) << 11; ctxIdx 276 = terminate;
// ctxIdx 0x3F8 + n = n bypass bins in bits 10.., first bin most significant.
#pragma once
#include "h264_dev.cuh"
#include "cabac_tables.cuh"
#include "k_cavlc.cuh"

namespace b200 {

#define CABAC_BYPASS0 0x3F8
#define CABAC_NCTX 460

template <int WRITE> struct BinSink {
    uint16_t *p; int n;
#ifdef B200_CHECKED
    int cap = 1 << 30;                                  // entries of room behind p
#define BINSINK_CHECK() B200_CHECK(!WRITE || n < cap, 2)
#else
#define BINSINK_CHECK() do { } while (0)
#endif
    __device__ __forceinline__ void put(int ctx, int bin) { BINSINK_CHECK(); if (WRITE) p[n] = (uint16_t)(ctx | (bin << 10)); n++; }
    __device__ __forceinline__ void run(int ctx, int bin, int rep) { BINSINK_CHECK(); if (WRITE) p[n] = (uint16_t)(ctx | (bin << 10) | ((rep - 1) << 11)); n++; }
    // a string of bypass bins, first bin = most significant of `len` bits: entries of six, the remainder last
    __device__ __forceinline__ void bypass(uint32_t bits, int len)
    {
        while (len > 0) {
            const int k = min(len, 6);
            BINSINK_CHECK();
            if (WRITE) p[n] = (uint16_t)((CABAC_BYPASS0 + k) | (((bits >> (len - k)) & ((1u << k) - 1u)) << 10));
            n++; len -= k;
        }
    }
};
// k-th order Exp-Golomb suffix (9.3.2.3) followed by the sign bit
__device__ __forceinline__ void egk_sign(int v, int k, int neg, uint32_t &bits, int &len)
{
    uint32_t b = 0; int n = 0;
    while (v >= (1 << k)) { b = (b << 1) | 1u; n++; v -= 1 << k; k++; }
    b <<= 1; n++;
    b = (b << k) | (uint32_t)v; n += k;
    bits = (b << 1) | (uint32_t)neg; len = n + 1;
}
// mvd_l0 component: UEG3 with uCoff 9; ctxIdxInc of bin 0 from the neighbouring partitions' |mvd| sum (9.3.3.1.1.7)
template <int W> __device__ __forceinline__ void bin_mvd(BinSink<W> &s, int base, int sum, int v)
{
    const int a = abs(v);
    s.put(base + (sum < 3 ? 0 : sum > 32 ? 2 : 1), a != 0);
    if (!a) return;
    for (int i = 1; i < min(a, 4); i++) s.put(base + 2 + i, 1);
    if (a < 4) s.put(base + 2 + a, 0);
    else {
        if (min(a, 9) > 4) s.run(base + 6, 1, min(a, 9) - 4);
        if (a < 9) s.put(base + 6, 0);
    }
    uint32_t bits = (uint32_t)(v < 0); int len = 1;
    if (a >= 9) egk_sign(a - 9, 3, v < 0, bits, len);
    s.bypass(bits, len);
}
// residual_block_cabac (7.3.5.3.3): lv = levels in scan order, n = 16 / 15 / 4, cat = ctxBlockCat 0..4
template <int W> __device__ void bin_residual(BinSink<W> &s, const int16_t *lv, int n, int cat, int cbf_inc)
{
    const int cbf_off = cat * 4, sig_off = cat == 0 ? 0 : cat == 1 ? 15 : cat == 2 ? 29 : cat == 3 ? 44 : 47,
              abs_off = cat == 4 ? 39 : cat * 10;
    int last = -1;
    for (int i = 0; i < n; i++) if (lv[i]) last = i;
    s.put(85 + cbf_off + cbf_inc, last >= 0);
    if (last < 0) return;
    for (int i = 0; i < n - 1; i++) {
        const int inc = cat == 3 ? min(i, 2) : i;
        s.put(105 + sig_off + inc, lv[i] != 0);
        if (lv[i]) { s.put(166 + sig_off + inc, i == last); if (i == last) break; }
    }
    int eq1 = 0, gt1 = 0;
    const int base = 227 + abs_off;
    for (int i = last; i >= 0; i--) {
        const int v = lv[i];
        if (!v) continue;
        const int a = abs(v) - 1;
        s.put(base + (gt1 ? 0 : min(4, 1 + eq1)), a > 0);
        uint32_t bits = (uint32_t)(v < 0); int len = 1;
        if (a > 0) {
            const int inc = 5 + min(4 - (cat == 3), gt1);
            if (min(a, 14) > 1) s.run(base + inc, 1, min(a, 14) - 1);
            if (a < 14) s.put(base + inc, 0); else egk_sign(a - 14, 0, v < 0, bits, len);
            gt1++;
        } else eq1++;
        s.bypass(bits, len);
    }
}

// Table 9-43 (frame coded blocks): ctxIdxInc of significant_coeff_flag / last_significant_coeff_flag of ctxBlockCat 5 by levelListIdx
static __device__ __constant__ uint8_t c_cabac_sig8[63] = {
    0, 1, 2, 3, 4, 5, 5, 4, 4, 3, 3, 4, 4, 4, 5, 5, 4, 4, 4, 4, 3, 3, 6, 7, 7, 7, 8, 9, 10, 9, 8, 7,
    7, 6, 11, 12, 13, 11, 6, 7, 8, 9, 14, 10, 9, 8, 6, 11, 12, 13, 11, 6, 9, 14, 10, 9, 11, 12, 13, 11, 14, 10, 12 };
static __device__ __constant__ uint8_t c_cabac_last8[63] = {
    0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2,
    3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8 };
// residual_block_cabac of an 8x8 luma block (ctxBlockCat 5, 64 levels in 8x8 zig-zag order read from the MB's coefficient record): no
// coded_block_flag (inferred 1 with the cbp bit); significance map contexts 402.. / 417.., magnitudes 426..
template <int W> __device__ void bin_residual8(BinSink<W> &s, const int16_t *lv)
{
    int last = -1;
    for (int i = 0; i < 64; i++) if (lv[i]) last = i;
    for (int i = 0; i < 63; i++) {
        const int v = lv[i];
        s.put(402 + c_cabac_sig8[i], v != 0);
        if (v) { s.put(417 + c_cabac_last8[i], i == last); if (i == last) break; }
    }
    int eq1 = 0, gt1 = 0;
    for (int i = last; i >= 0; i--) {
        const int v = lv[i];
        if (!v) continue;
        const int a = abs(v) - 1;
        s.put(426 + (gt1 ? 0 : min(4, 1 + eq1)), a > 0);
        uint32_t bits = (uint32_t)(v < 0); int len = 1;
        if (a > 0) {
            const int inc = 5 + min(4, gt1);
            if (min(a, 14) > 1) s.run(426 + inc, 1, min(a, 14) - 1);
            if (a < 14) s.put(426 + inc, 0); else egk_sign(a - 14, 0, v < 0, bits, len);
            gt1++;
        } else eq1++;
        s.bypass(bits, len);
    }
}

__device__ __forceinline__ bool is_intra_type(int t) { return t == MB_I16x16 || t == MB_I4x4 || t == MB_I8x8; }
__device__ __forceinline__ bool is_inxn_type(int t) { return t == MB_I4x4 || t == MB_I8x8; }

// grid: (ceil(n_mb / 256), 1, sessions)
__global__ void __launch_bounds__(256) k_cabac_side(const Sess *ss, Geom g)
{
    const int mb = blockIdx.x * 256 + threadIdx.x;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    const MbInfo *mi = s.mbi + mb; const MbCoef *co = s.coef + mb;
    int mx, my; mb_xy(g, mb, mx, my);
    MbSide sd; uint32_t *w = reinterpret_cast<uint32_t *>(&sd);
#pragma unroll
    for (int i = 0; i < 5; i++) w[i] = 0;
    const int t = mi->mb_type;
    if (t == MB_P16x16 || t == MB_P8x8) {
        int mvd[8]; mb_mvds(s, g, mx, my, mi, mvd);
#pragma unroll
        for (int q = 0; q < 4; q++) { const int k = t == MB_P8x8 ? q : 0; sd.mvd[q][0] = (int16_t)mvd[2 * k]; sd.mvd[q][1] = (int16_t)mvd[2 * k + 1]; }
    } else if (t == MB_I4x4) {
        const bool left = mx > 0, top = !row_is_slice_top(g, my);
        const MbInfo *ml = mi - 1, *mt = mi - g.mbw;
        const bool l4 = left && is_inxn_type(ml->mb_type), t4 = top && is_inxn_type(mt->mb_type);
        for (int k = 0; k < 16; k++) {       // predIntra4x4PredMode, 8.3.1.1
            const int bx = blk_x(k), by = blk_y(k);
            const int ma = bx > 0 ? mi->i4_mode[xy2blk(bx - 1, by)] : !left ? -1 : l4 ? ml->i4_mode[xy2blk(3, by)] : 2;
            const int mb_ = by > 0 ? mi->i4_mode[xy2blk(bx, by - 1)] : !top ? -1 : t4 ? mt->i4_mode[xy2blk(bx, 3)] : 2;
            const int pm = (ma < 0 || mb_ < 0) ? 2 : min(ma, mb_), m = mi->i4_mode[k];
            sd.i4_syn[k] = (uint8_t)(m == pm ? 8 : m < pm ? m : m - 1);
        }
    }
    else if (t == MB_I8x8) {                 // predIntra8x8PredMode, 8.3.2.1: the neighbouring I_NxN macroblock's mode at the block's first row / column
        const bool left = mx > 0, top = !row_is_slice_top(g, my);
        const MbInfo *ml = mi - 1, *mt = mi - g.mbw;
        const bool l4 = left && is_inxn_type(ml->mb_type), t4 = top && is_inxn_type(mt->mb_type);
        for (int k = 0; k < 4; k++) {
            const int ma = (k & 1) ? mi->i4_mode[4 * (k - 1) + 1] : !left ? -1 : l4 ? ml->i4_mode[4 * (k + 1) + 1] : 2;
            const int mb_ = (k & 2) ? mi->i4_mode[4 * (k - 2) + 2] : !top ? -1 : t4 ? mt->i4_mode[4 * (k + 2) + 2] : 2;
            const int pm = (ma < 0 || mb_ < 0) ? 2 : min(ma, mb_), m = mi->i4_mode[4 * k];
            sd.i4_syn[k] = (uint8_t)(m == pm ? 8 : m < pm ? m : m - 1);
        }
    }
    if (t != MB_PSKIP) {
        int dc = 0;
        if (t == MB_I16x16) { const uint4 *p = reinterpret_cast<const uint4 *>(co->luma_dc); const uint4 a = p[0], b = p[1]; dc |= (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) != 0u; }
        if (mi->cbp >> 4) {
            const uint2 *p = reinterpret_cast<const uint2 *>(co->chroma_dc[0]); const uint2 a = p[0], b = p[1];
            dc |= ((a.x | a.y) != 0u) << 1; dc |= ((b.x | b.y) != 0u) << 2;
        }
        sd.dc_cbf = (uint8_t)dc;
    }
    uint32_t *dst = reinterpret_cast<uint32_t *>(s.side + mb);
#pragma unroll
    for (int i = 0; i < 5; i++) dst[i] = w[i];
}

// the header bins of one MB (everything of macroblock_layer() before the residual), preceded by mb_skip_flag in P slices
template <int W> __device__ void bin_mb_header(BinSink<W> &s, const Sess &se, const Geom &g, int mx, int my, const MbInfo *m)
{
    const bool is_p = !se.is_idr;
    const MbInfo *L = mx > 0 ? m - 1 : nullptr, *T = row_is_slice_top(g, my) ? nullptr : m - g.mbw;
    const MbSide *sd = se.side + (my * g.mbw + mx), *sL = sd - 1, *sT = sd - g.mbw;
    const int t = m->mb_type, cl = m->cbp & 15, cc = m->cbp >> 4;
    const bool intra = is_intra_type(t);
    if (is_p) {
        s.put(11 + (L && L->mb_type != MB_PSKIP) + (T && T->mb_type != MB_PSKIP), t == MB_PSKIP);
        if (t == MB_PSKIP) return;
    }
    if (!intra) {                                                            // mb_type, Table 9-39
        s.put(14, 0); s.put(15, 0); s.put(16, t == MB_P8x8);
        if (t == MB_P8x8) for (int q = 0; q < 4; q++) s.put(21, 1);          // sub_mb_type P_L0_8x8
    } else {
        int b0, c_cl, c_cc, c_cc2, c_m1, c_m0;
        if (is_p) { s.put(14, 1); b0 = 17; c_cl = 18; c_cc = 19; c_cc2 = 19; c_m1 = 20; c_m0 = 20; }
        else { b0 = 3 + (L && !is_inxn_type(L->mb_type)) + (T && !is_inxn_type(T->mb_type)); c_cl = 6; c_cc = 7; c_cc2 = 8; c_m1 = 9; c_m0 = 10; }
        s.put(b0, t == MB_I16x16);
        if (t == MB_I16x16) {
            s.put(276, 0);
            s.put(c_cl, cl != 0); s.put(c_cc, cc != 0);
            if (cc) s.put(c_cc2, cc == 2);
            s.put(c_m1, m->i16_mode >> 1); s.put(c_m0, m->i16_mode & 1);
        }
    }
    // transform_size_8x8_flag (7.3.5; ctxIdx 399 + the neighbours' flags): 0 for I_NxN (Intra_4x4 only), coded for inter MBs behind the cbp
    const int t8inc = (L && mb_t8(L)) + (T && mb_t8(T));
    if (is_inxn_type(t) && se.t8x8) s.put(399 + t8inc, t == MB_I8x8);
    if (is_inxn_type(t))
        for (int k = 0; k < (t == MB_I8x8 ? 4 : 16); k++) {      // prev_intra{4x4,8x8}_pred_mode_flag / rem_...: same contexts
            const int r = sd->i4_syn[k];
            s.put(68, r == 8);
            if (r != 8) { s.put(69, r & 1); s.put(69, (r >> 1) & 1); s.put(69, (r >> 2) & 1); }
        }
    if (intra) {
        const int inc = (L && is_intra_type(L->mb_type) && L->chroma_mode != 0) + (T && is_intra_type(T->mb_type) && T->chroma_mode != 0), cm = m->chroma_mode;
        s.put(64 + inc, cm != 0);
        if (cm) { s.put(67, cm != 1); if (cm != 1) s.put(67, cm != 2); }
    } else {
        const int np = t == MB_P8x8 ? 4 : 1;
        const bool la = L && !is_intra_type(L->mb_type), ta = T && !is_intra_type(T->mb_type);   // the union holds i4_syn for intra MBs
        for (int q = 0; q < np; q++)
            for (int c = 0; c < 2; c++) {
                const int a = (q & 1) ? abs((int)sd->mvd[q - 1][c]) : la ? abs((int)sL->mvd[q + 1][c]) : 0;
                const int b = (q & 2) ? abs((int)sd->mvd[q - 2][c]) : ta ? abs((int)sT->mvd[q + 2][c]) : 0;
                bin_mvd<W>(s, c ? 47 : 40, a + b, sd->mvd[q][c]);
            }
    }
    if (t != MB_I16x16) {                                                    // coded_block_pattern, 9.3.3.1.1.4
        for (int b8 = 0; b8 < 4; b8++) {
            const int a = (b8 & 1) ? !((cl >> (b8 - 1)) & 1) : L ? !((L->cbp >> (b8 + 1)) & 1) : 0;
            const int b = (b8 & 2) ? !((cl >> (b8 - 2)) & 1) : T ? !((T->cbp >> (b8 + 2)) & 1) : 0;
            s.put(73 + a + 2 * b, (cl >> b8) & 1);
        }
        s.put(77 + (L && (L->cbp >> 4)) + 2 * (T && (T->cbp >> 4)), cc != 0);
        if (cc) s.put(81 + (L && (L->cbp >> 4) == 2) + 2 * (T && (T->cbp >> 4) == 2), cc == 2);
    }
    if (!intra && cl && se.t8x8) s.put(399 + t8inc, mb_t8(m));              // every inter MB here has partitions of 8x8 at least
    if (t == MB_I16x16 || m->cbp) s.put(60, 0);                              // mb_qp_delta = 0
}

struct CabacItem { const int16_t *lv; int n, cat, inc; bool present; };
// the residual block coded by `lane` (1..27) and the ctxIdxInc of its coded_block_flag (9.3.3.1.1.9)
__device__ __forceinline__ CabacItem cabac_item(const Sess &se, const Geom &g, int mx, int my, int lane, const MbInfo *m, const MbCoef *co)
{
    CabacItem it; it.present = false; it.lv = co->luma_dc; it.n = 16; it.cat = 0; it.inc = 0;
    const int t = m->mb_type, cl = m->cbp & 15, cc = m->cbp >> 4;
    if (t == MB_PSKIP) return it;
    const bool i16 = t == MB_I16x16, intra = is_intra_type(t);
    const MbInfo *L = mx > 0 ? m - 1 : nullptr, *T = row_is_slice_top(g, my) ? nullptr : m - g.mbw;
    const MbSide *sd = se.side + (my * g.mbw + mx), *sL = sd - 1, *sT = sd - g.mbw;
    int a = intra, b = intra;                                                // neighbour MB not available: 1 for intra, 0 for inter MBs
    if (lane == 1) {
        if (L) a = L->mb_type == MB_I16x16 && (sL->dc_cbf & 1);
        if (T) b = T->mb_type == MB_I16x16 && (sT->dc_cbf & 1);
        it.present = i16;
    } else if (lane < 18) {
        const int k = lane - 2, bx = blk_x(k), by = blk_y(k);
        if (bx) a = m->nnz[xy2blk(bx - 1, by)] != 0; else if (L) a = L->nnz[xy2blk(3, by)] != 0;
        if (by) b = m->nnz[xy2blk(bx, by - 1)] != 0; else if (T) b = T->nnz[xy2blk(bx, 3)] != 0;
        it.present = (cl >> (k >> 2)) & 1;
        if (i16) { it.lv = co->luma[k] + 1; it.n = 15; it.cat = 1; }
        else if (mb_t8(m)) { it.lv = co->luma[k]; it.n = 64; it.cat = 5; it.present = it.present && !(k & 3); }   // the lane of an 8x8 block's first 4x4 codes all of it
        else { it.lv = co->luma[k]; it.n = 16; it.cat = 2; }
    } else if (lane < 20) {
        const int p = lane - 18;
        if (L) a = (L->cbp >> 4) && ((sL->dc_cbf >> (1 + p)) & 1);
        if (T) b = (T->cbp >> 4) && ((sT->dc_cbf >> (1 + p)) & 1);
        it.present = cc != 0; it.lv = co->chroma_dc[p]; it.n = 4; it.cat = 3;
    } else if (lane < 28) {
        const int p = (lane - 20) >> 2, k = (lane - 20) & 3, bx = k & 1, by = k >> 1, base = 16 + 4 * p;
        if (bx) a = m->nnz[base + k - 1] != 0; else if (L) a = (L->cbp >> 4) == 2 && L->nnz[base + by * 2 + 1] != 0;
        if (by) b = m->nnz[base + k - 2] != 0; else if (T) b = (T->cbp >> 4) == 2 && T->nnz[base + 2 + bx] != 0;
        it.present = cc == 2; it.lv = co->chroma_ac[p][k] + 1; it.n = 15; it.cat = 4;
    }
    it.inc = a + 2 * b;
    return it;
}

#define CABAC_WARPS 8
// Per-MB slot of the bin kernel: every lane (syntax group) writes into its own fixed sub-slot, sized for the worst case of its group --
// at most 8 entries per level (significance 2, magnitude prefix 2, Exp-Golomb suffix + sign 4 for |level| <= 2063) plus the coded_block_flag;
// an 8x8 block (ctxBlockCat 5, 512 entries) is written by the lane of its first 4x4 block over the four sub-slots of its 8x8 quadrant.
//   lane 0 header (<= 98: skip flag, mb_type, 16 mode syntax elements of 4 or 8 mvd components of <= 10, cbp, flags)   112 entries at 0
//   lane 1 Intra16x16DCLevel                                                                                            132 at 112
//   lanes 2-17 luma 4x4 blocks                                                                                     16 x 132 at 244
//   lanes 18-19 chroma DC (4 levels)                                                                                 2 x 40 at 2356
//   lanes 20-27 chroma AC (15 levels)                                                                               8 x 124 at 2436
//   lane 28 end_of_slice_flag                                                                                            1 at 3428
#define CABAC_MB_SLOT 3456
#define CABAC_HDR_SLOT 112
__device__ __forceinline__ int cabac_lane_slot(int lane)
{
    return lane == 0 ? 0 : lane == 1 ? 112 : lane < 18 ? 244 + 132 * (lane - 2) : lane < 20 ? 2356 + 40 * (lane - 18) : lane < 28 ? 2436 + 124 * (lane - 20) : 3428;
}
// grid: (ceil(n_mb / 256), 1, sessions), THREAD per MB (after k_cabac_side: the mvd contexts read the neighbours' side records): the header
// entries of the macroblock (everything of macroblock_layer() before the residual, mb_skip_flag first) into sub-slot 0 and its
// end_of_slice_flag is counted (k_cabac_compact writes it); clears the MB's 32 lane counts and sets those two; mb_bits[mb] = their sum. A header is a short serial
// walk with few entries, which a warp per MB would run on one lane in 32; here the 32 lanes of a warp walk 32 headers.
__global__ void __launch_bounds__(256) k_cabac_hdr(const Sess *ss, Geom g)
{
    const int mb = blockIdx.x * 256 + threadIdx.x;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    int mx, my; mb_xy(g, mb, mx, my);
    // the headers have their own dense array (CABAC_HDR_SLOT entries per MB): neighbouring threads write neighbouring lines, not lines 7 KB apart
    BinSink<1> bs; bs.p = s.bins_hdr + (size_t)mb * CABAC_HDR_SLOT; bs.n = 0;
#ifdef B200_CHECKED
    bs.cap = CABAC_HDR_SLOT;
#endif
    bin_mb_header<1>(bs, s, g, mx, my, s.mbi + mb);
    uint4 *cnt = reinterpret_cast<uint4 *>(s.bin_lane_cnt + (size_t)mb * 32);
    cnt[0] = make_uint4((uint32_t)bs.n, 0u, 0u, 0u); cnt[1] = make_uint4(0u, 0u, 0u, 0u); cnt[2] = make_uint4(0u, 0u, 0u, 0u); cnt[3] = make_uint4(0u, 0u, 1u, 0u);   // lanes 0 and 28
    s.mb_bits[mb] = (uint32_t)bs.n + 1u;
}

// grid: (ceil(n_mb / CABAC_WARPS), 1, sessions), warp per MB: the residual blocks -- lane 1 Intra16x16DCLevel, 2-17 the luma blocks, 18-19 chroma
// DC, 20-27 chroma AC -- ONE evaluation each, into the lane's sub-slot of bins_mb[mb * CABAC_MB_SLOT ..], the count into bin_lane_cnt, the
// total added to mb_bits. Macroblocks without residual (P_Skip, cbp 0 and not Intra_16x16: most of a P picture) leave at once.
// k_cabac_scan turns the totals into offsets and k_cabac_compact moves the sub-slots, in lane order, to bins[slice base + mb_off[mb]] (slice
// base = first_mb * B200_MB_BIN_SLOT), the contiguous list the coder walks.
__global__ void __launch_bounds__(CABAC_WARPS * 32) k_cabac_bins(const Sess *ss, Geom g)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * CABAC_WARPS + warp;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    const MbInfo *mi = s.mbi + mb; const MbCoef *co = s.coef + mb;
    {
        const uint32_t w0 = *reinterpret_cast<const uint32_t *>(mi);
        const int t = w0 & 255, cbp = (int)(w0 >> 24);
        if (t == MB_PSKIP || (cbp == 0 && t != MB_I16x16)) return;
    }
    int mx, my; mb_xy(g, mb, mx, my);
    __align__(16) int16_t lv[16];
    CabacItem it = cabac_item(s, g, mx, my, lane, mi, co);
    const bool mine = lane >= 1 && lane < 28 && it.present;
    if (mine && it.cat != 5)
        for (int i = 0; i < it.n; i++) lv[i] = it.lv[i];
    BinSink<1> bs; bs.p = s.bins_mb + (size_t)mb * CABAC_MB_SLOT + cabac_lane_slot(lane); bs.n = 0;
#ifdef B200_CHECKED
    bs.cap = mine && it.cat == 5 ? 4 * 132 : cabac_lane_slot(lane + 1) - cabac_lane_slot(lane);
#endif
    if (mine) { if (it.cat == 5) bin_residual8<1>(bs, it.lv); else bin_residual<1>(bs, lv, it.n, it.cat, it.inc); }
    if (lane >= 1 && lane < 28) s.bin_lane_cnt[(size_t)mb * 32 + lane] = (uint16_t)bs.n;
    const int total = __reduce_add_sync(0xffffffffu, bs.n);
    if (lane == 0) s.mb_bits[mb] += (uint32_t)total;
}

// grid: (num_slices, 1, sessions), 256 threads: entry offset of every MB inside its slice's bin list, slice total
__global__ void __launch_bounds__(256) k_cabac_scan(const Sess *ss, Geom g)
{
    const Sess &s = ss[blockIdx.z];
    const int sl = blockIdx.x, m0 = g.slice_row0[sl] * g.mbw, m1 = g.slice_row0[sl + 1] * g.mbw;
    __shared__ int wsum[8];
    int carry = 0;
    for (int base = m0; base < m1; base += 256) {
        const int mb = base + threadIdx.x;
        const int len = mb < m1 ? (int)s.mb_bits[mb] : 0;
        int chunk_total; const int off = carry + block_excl_scan(len, &chunk_total, wsum);
        if (mb < m1) s.mb_off[mb] = (uint32_t)off;
        carry += chunk_total;
    }
    if (threadIdx.x == 0) s.slice_nbins[sl] = (uint32_t)carry;
}

// The slices' lists are put together by two kernels after the scan. Most macroblocks of a P picture have no residual (P_Skip: one entry, the
// mb_skip_flag): a warp per MB spent ~290 instructions moving one or two entries (profiles/r02_cabac_ncu.md), so
//   k_cabac_place_hdr : THREAD per MB -- the header entries (dense bins_hdr slot) and the end_of_slice_flag to their place;
//   k_cabac_compact   : warp per MB, only macroblocks WITH residual: lane l (1..27) moves the entries of its sub-slot.
// grid: (ceil(n_mb / 256), 1, sessions), 256 threads
__global__ void __launch_bounds__(256) k_cabac_place_hdr(const Sess *ss, Geom g)
{
    const int mb = blockIdx.x * 256 + threadIdx.x;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    const int my = mb / g.mbw;
    int sl = 0;
    for (int k = 1; k < g.num_slices; k++) sl += (my >= g.slice_row0[k]);
    const int n = (int)s.bin_lane_cnt[(size_t)mb * 32];
    const uint16_t *src = s.bins_hdr + (size_t)mb * CABAC_HDR_SLOT;
    uint16_t *dst = s.bins + (size_t)g.slice_row0[sl] * g.mbw * B200_MB_BIN_SLOT + s.mb_off[mb];
    const int total = (int)s.mb_bits[mb];
    B200_CHECK((size_t)s.mb_off[mb] + (size_t)total <= (size_t)(g.slice_row0[sl + 1] - g.slice_row0[sl]) * g.mbw * B200_MB_BIN_SLOT, 3);
    for (int i = 0; i < n; i++) dst[i] = src[i];
    dst[total - 1] = (uint16_t)(276 | ((mb == g.slice_row0[sl + 1] * g.mbw - 1) << 10));      // end_of_slice_flag
}
// grid: (ceil(n_mb / CABAC_WARPS), 1, sessions)
__global__ void __launch_bounds__(CABAC_WARPS * 32) k_cabac_compact(const Sess *ss, Geom g)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * CABAC_WARPS + warp;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    {
        const uint32_t w0 = *reinterpret_cast<const uint32_t *>(s.mbi + mb);
        const int t = w0 & 255, cbp = (int)(w0 >> 24);
        if (t == MB_PSKIP || (cbp == 0 && t != MB_I16x16)) return;             // the same test as k_cabac_bins: nothing but the header
    }
    const int my = mb / g.mbw;
    int sl = 0;
    for (int k = 1; k < g.num_slices; k++) sl += (my >= g.slice_row0[k]);
    const int n = lane == 28 ? 0 : (int)s.bin_lane_cnt[(size_t)mb * 32 + lane];      // lane 0 = the header (its count is part of the offsets), 28 = the end flag
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 0 || n == 0) return;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(s.bins_mb + (size_t)mb * CABAC_MB_SLOT + cabac_lane_slot(lane));   // sub-slots start at even entries
    uint16_t *dst = s.bins + (size_t)g.slice_row0[sl] * g.mbw * B200_MB_BIN_SLOT + s.mb_off[mb] + (incl - n);
    for (int i = 0; i + 1 < n; i += 2) { const uint32_t w = src[i >> 1]; dst[i] = (uint16_t)w; dst[i + 1] = (uint16_t)(w >> 16); }
    if (n & 1) dst[n - 1] = (uint16_t)src[n >> 1];
}

// ---- the arithmetic coder (9.3.4.2) ----
// `low` keeps the 10 bits of the standard's codILow plus `nb` bits above them that have not been written yet (nb starts at -1:
// the standard drops the first bit). A byte leaves as soon as nb reaches 8; it may carry into the bytes before it, so the last
// byte that is not 0xFF is held back together with the count of 0xFF bytes behind it.
template <bool SWAP> struct CabacOut {
    uint8_t *base; int pos; int hold, n_ff; int cap;       // cap: bytes of room (a slice that does not fit is cut there; k_nal_pack reports the overflow)
    __device__ __forceinline__ void store(int v) { B200_CHECK(pos < cap, 12); if (pos < cap) base[SWAP ? (pos ^ 3) : pos] = (uint8_t)v; pos++; }
    __device__ __forceinline__ void byte(int out)          // 9 bits: a byte and the carry into the earlier ones
    {
        if ((out & 0xff) == 0xff) { n_ff++; return; }
        const int carry = out >> 8;
        if (hold >= 0) store(hold + carry);
        while (n_ff > 0) { store(carry ? 0x00 : 0xff); n_ff--; }
        hold = out & 0xff;
    }
    __device__ __forceinline__ void finish() { if (hold >= 0) store(hold); while (n_ff > 0) { store(0xff); n_ff--; } hold = -1; }
};
struct CabacTables {      // shared-memory copies: (pStateIdx << 1 | valMPS) -> next states, pStateIdx -> the four rangeTabLPS values in one word
    uint32_t range_lps[64];
    uint32_t lps_range[64];       // per pStateIdx: low byte of the RENORMALISED range after an LPS for each of the four quantised ranges (bit 8 is always set)
    uint32_t lps_shift[64];       // per pStateIdx: the four renormalisation shifts after an LPS, one byte each
    uint16_t next[128];           // next state after an MPS (low byte) / after an LPS (high byte)
    uint8_t state[CABAC_NCTX + 4];
};
__device__ __forceinline__ void cabac_tables_init(CabacTables &t, int qp, bool is_p, int lane)
{
    for (int i = lane; i < 64; i += 32) {
        uint32_t w = 0, r = 0, sh = 0;
        for (int q = 0; q < 4; q++) {
            const uint32_t v = c_cabac_range_lps[i * 4 + q]; const int k = __clz(v) - 23;
            w |= v << (8 * q); r |= ((v << k) & 255u) << (8 * q); sh |= (uint32_t)k << (8 * q);
        }
        t.range_lps[i] = w; t.lps_range[i] = r; t.lps_shift[i] = sh;
    }
    for (int i = lane; i < 128; i += 32) {
        const int p = i >> 1, mps = i & 1;
        t.next[i] = (uint16_t)(((c_cabac_next_mps[p] << 1) | mps) | (((c_cabac_next_lps[p] << 1) | (p == 0 ? 1 - mps : mps)) << 8));
    }
    const int q = clip3(0, 51, qp);
    for (int i = lane; i < CABAC_NCTX; i += 32) {                           // 9.3.1.1
        const int m = is_p ? c_cabac_init_p0[2 * i] : c_cabac_init_i[2 * i], n = is_p ? c_cabac_init_p0[2 * i + 1] : c_cabac_init_i[2 * i + 1];
        const int pre = clip3(1, 126, ((m * q) >> 4) + n);
        t.state[i] = (uint8_t)(pre <= 63 ? (63 - pre) << 1 : ((pre - 64) << 1) | 1);
    }
}

// The coder of one slice is a producer / consumer pair of warps around a ring of per-bin records in shared memory:
//  * the probability-state recurrence (9.3.4.2's pStateIdx / valMPS updates) only couples bins of the SAME context, so the producer
//    warp resolves 32 list entries at a time: lanes that share a context (match_any) take turns in list order, all others go at once.
//    Every bin becomes a record {the four rangeTabLPS values of its state, LPS?}; bypass strings and terminate bins pass through;
//  * what is left for the consumer (one lane) is the codIRange / codILow recurrence alone: byte select, subtract, select,
//    count-leading-zeros, shift -- no table walk and no dependent shared-memory load on its critical path.
#define CABAC_RING 2048           /* records; one producer step adds at most 32 entries * 32 repeats = 1024 */
// record (every kind goes through the same branch-free step):
//   regular bin : .x the four rangeTabLPS values of its state, .y bit 31 = the bin is the LPS; if it is: .z the renormalised ranges after
//                 the LPS (low bytes) and .w the shifts, one byte per quantised range; else .z = .w = 0
//   bypass bins : .x = .z = 0 (an "MPS" that leaves the range alone), .y = value, .w = number of bins in every byte
//   terminate 0 : an MPS with rangeTabLPS = 2 (9.3.4.5: codIRange -= 2, RenormE)
// the terminate bin of value 1 that ends the slice is not queued: the consumer flushes when the ring has drained
struct CabacRing { uint4 rec[CABAC_RING]; volatile uint32_t wr, rd; volatile int done, abort; };
// Watchdog of the coder's three spin loops (ring room, ring data / output-queue room, output data): like the wavefront and TMA waits, a wait
// longer than 2 s is an internal error (an inconsistent bin list) that must end the kernel and reach the host as B200ENC_EWAVE, never hang the GPU.
struct CabacSpin {
    unsigned spins = 0; unsigned long long t0 = 0;
    __device__ __forceinline__ bool expired(CabacRing &ring)
    {
        if (ring.abort) return true;
        if ((++spins & 4095u) != 0u) return false;
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (!t0) { t0 = t; return false; }
        if (t - t0 > 2000000000ull) { ring.abort = 1; return true; }
        return false;
    }
};
#ifdef CABAC_TIMING       // phase cycle counts of the coder pair (tools/cabac_coder_bench.py): [0] consumer total [1] consumer waiting [2] records
__device__ long long g_cabac_t[8];   // [3] producer total [4] producer waiting for room [5] producer turn loops [6] steps [7] turns
#define CT_CLK() clock64()
#define CT_ADD(i, v) (g_cabac_t[i] += (v))
#else
#define CT_CLK() 0ll
#define CT_ADD(i, v) ((void)0)
#endif

__device__ __forceinline__ void cabac_produce(CabacRing &ring, CabacTables &t, const uint16_t *bins, int n, int lane)
{
    uint32_t wr = 0;
    [[maybe_unused]] const long long t_begin = CT_CLK(); [[maybe_unused]] long long t_wait = 0, t_turn = 0, n_turn = 0;
    uint32_t vnext = lane < n ? bins[lane] : 0xffffu;
    for (int base = 0; base < n; base += 32) {
        const uint32_t v = vnext;
        vnext = base + 32 + lane < n ? bins[base + 32 + lane] : 0xffffu;    // the next step's entries are in flight during this one
        const bool valid = base + lane < n;
        const int ctx = v & 1023;
        const bool regular = valid && ctx < CABAC_BYPASS0 && ctx != 276, final = valid && ctx == 276 && ((v >> 10) & 1);
        const int rep = regular ? (int)(v >> 11) : 0, nrec = valid && !final ? rep + 1 : 0;
        int incl = nrec;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t at = wr + (uint32_t)(incl - nrec);
        const long long t0 = CT_CLK();
        B200_CHECK(total <= CABAC_RING, 11);
        { CabacSpin sg; while (wr + (uint32_t)total - ring.rd > CABAC_RING) { if (sg.expired(ring)) break; __nanosleep(64); } }       // room in the ring
        if (ring.abort) break;
        const long long t1 = CT_CLK(); t_wait += t1 - t0;
        const uint32_t peers = __match_any_sync(0xffffffffu, regular ? ctx : 0x10000 + lane);
        const uint32_t before = peers & ((1u << lane) - 1u);
        const int rank = __popc(before), pred = before ? 31 - __clz(before) : lane;   // the lane that holds this context just before me
        const int turns = __reduce_max_sync(0xffffffffu, regular ? rank : 0);
        // the lanes of a group take turns in list order; the state travels from lane to lane through a shuffle and only the last
        // lane of a group writes it back
        int st = regular ? t.state[ctx] : 0;
        const int bin = (v >> 10) & 1;
        const bool closes = (peers >> lane) == 1u;                                  // no later lane shares the context: writes the state back
        for (int r = 0; r <= turns; r++) {
            // straight-line body for every lane (the lanes whose turn it is not compute on their own state and store nothing); only an
            // entry that repeats its bin takes the warp through the loop below
            const bool mine = regular && rank == r;
            int lps = bin != (st & 1), p = st >> 1;
            const uint4 rec = make_uint4(t.range_lps[p], (uint32_t)lps << 31, lps ? t.lps_range[p] : 0u, lps ? t.lps_shift[p] : 0u);
            if (mine) ring.rec[at & (CABAC_RING - 1)] = rec;
            int nx = t.next[st], s1 = lps ? nx >> 8 : nx & 255;
            if (__any_sync(0xffffffffu, mine && rep > 0)) {
                if (mine) for (int k = 1; k <= rep; k++) {
                    lps = bin != (s1 & 1); p = s1 >> 1;
                    ring.rec[(at + k) & (CABAC_RING - 1)] = make_uint4(t.range_lps[p], (uint32_t)lps << 31, lps ? t.lps_range[p] : 0u, lps ? t.lps_shift[p] : 0u);
                    nx = t.next[s1]; s1 = lps ? nx >> 8 : nx & 255;
                }
            }
            if (mine) st = s1;
            if (mine && closes) t.state[ctx] = (uint8_t)st;
            const int handed = __shfl_sync(0xffffffffu, st, pred);
            if (regular && rank == r + 1) st = handed;
        }
        t_turn += CT_CLK() - t1; n_turn += turns + 1;
        if (valid && !regular && !final) {
            if (ctx == 276) ring.rec[at & (CABAC_RING - 1)] = make_uint4(0x02020202u, 0u, 0u, 0u);
            else { const int k = ctx - CABAC_BYPASS0; ring.rec[at & (CABAC_RING - 1)] = make_uint4(0u, (v >> 10) & ((1u << k) - 1u), 0u, 0x01010101u * (uint32_t)k); }
        }
        __threadfence_block(); __syncwarp();
        wr += (uint32_t)total;
        if (lane == 0) ring.wr = wr;
    }
    __threadfence_block(); __syncwarp();
    if (lane == 0) { ring.done = 1; CT_ADD(3, CT_CLK() - t_begin); CT_ADD(4, t_wait); CT_ADD(5, t_turn); CT_ADD(6, (n + 31) / 32); CT_ADD(7, n_turn); }
}

// bytes leave the range/low recurrence through a second queue as 16-bit values: the new byte and, above it, the byte before it READ AGAIN
// (low is never masked, so a carry that arrived since shows as a changed upper byte); the writer thread turns that into carries, holds
// back the last byte that is not 0xFF and stores the bytes.
#define CABAC_OUTQ 1024
struct CabacOutQ { uint16_t v[CABAC_OUTQ]; volatile uint32_t wr, rd; volatile int done; };

// one record of the ring through the recurrence, without a branch but the one around the byte store; oq_base = shared-space address of
// the queue, qw = private write index
__device__ __forceinline__ void cabac_step(const uint4 rec, uint32_t &low, uint32_t &range, int &nb, uint32_t oq_base, uint32_t &qw)
{
    const uint32_t sel = range >> 6;                                        // 4..7: byte qCodIRangeIdx of the second PRMT operand
    const uint32_t rmps = range - __byte_perm(0u, rec.x, sel);
    const bool lps = (int)rec.y < 0;
    const uint32_t one = rmps < 256u;                                       // after an MPS at most one shift
    const uint32_t sh = max(__byte_perm(0u, rec.w, sel), one);              // LPS: its shift (>= 1 >= one); bypass: the bin count; MPS: one
    low = ((low + (lps ? rmps : 0u)) << sh) + (rec.y & 63u) * range;        // bypass value * range (0 for regular bins)
    range = lps ? 256u | __byte_perm(0u, rec.z, sel) : rmps << one;
    nb += (int)sh;
    if (nb >= 8) {
        nb -= 8;
        asm volatile("st.shared.u16 [%0], %1;" :: "r"(oq_base + ((qw & (CABAC_OUTQ - 1)) << 1)), "h"((unsigned short)(low >> (nb + 10))) : "memory");
        qw++;
    }
}
// the calling thread drains the ring until the producer is done
__device__ __forceinline__ void cabac_consume(CabacRing &ring, CabacOutQ &oq)
{
    const uint32_t oq_base = (uint32_t)__cvta_generic_to_shared(oq.v);
    uint32_t low = 0, range = 510; int nb = -1;
    uint32_t rd = 0, qw = 0;
    [[maybe_unused]] const long long t_begin = CT_CLK(); [[maybe_unused]] long long t_wait = 0;
    CabacSpin idle;
    for (;;) {
        uint32_t wr = ring.wr;
        if (wr == rd) {
            const long long t0 = CT_CLK();
            if (ring.done) { wr = ring.wr; if (wr == rd) break; } else { if (idle.expired(ring)) break; __nanosleep(32); t_wait += CT_CLK() - t0; continue; }
        }
        __threadfence_block();
        uint32_t avail = min(wr - rd, 256u);
        { CabacSpin sg; while (qw + avail - oq.rd > CABAC_OUTQ) { if (sg.expired(ring)) break; __nanosleep(32); } }            // a record emits at most one byte
        if (ring.abort) break;
        idle.spins = 0; idle.t0 = 0;
        for (; avail >= 8u; avail -= 8u, rd += 8u) {
            uint4 r[8];
#pragma unroll
            for (int i = 0; i < 8; i++) r[i] = ring.rec[(rd + i) & (CABAC_RING - 1)];
#pragma unroll
            for (int i = 0; i < 8; i++) cabac_step(r[i], low, range, nb, oq_base, qw);
        }
        for (; avail; avail--, rd++) cabac_step(ring.rec[rd & (CABAC_RING - 1)], low, range, nb, oq_base, qw);
        __threadfence_block();
        ring.rd = rd; oq.wr = qw;
    }
    {   // the slice's last bin, end_of_slice_flag = 1: EncodeFlush, 9.3.4.5
        range -= 2; low += range; low |= 1u;                                // the last of the ten bits is the rbsp_stop_one_bit
        int width = nb + 10; const int pad = (8 - (width & 7)) & 7;
        low <<= pad; width += pad;                                          // rbsp_alignment_zero_bit
        { CabacSpin sg; while (qw + 8u - oq.rd > CABAC_OUTQ) { if (sg.expired(ring)) break; __nanosleep(32); } }
        while (!ring.abort && width >= 8) { width -= 8; oq.v[qw++ & (CABAC_OUTQ - 1)] = (uint16_t)(low >> width); }
        __threadfence_block();
        oq.wr = qw;
    }
    __threadfence_block();
    oq.done = 1;
    CT_ADD(0, CT_CLK() - t_begin); CT_ADD(1, t_wait); CT_ADD(2, rd);
}
// the calling thread turns the queue into bytes
template <bool SWAP> __device__ __forceinline__ void cabac_write(CabacRing &ring, CabacOutQ &oq, CabacOut<SWAP> &o)
{
    uint32_t rd = 0; int last = -1;                        // the byte taken before this one, as it was then
    CabacSpin idle;
    for (;;) {
        uint32_t wr = oq.wr;
        if (wr == rd) { if (oq.done) { wr = oq.wr; if (wr == rd) break; } else { if (idle.expired(ring)) break; __nanosleep(64); continue; } }
        idle.spins = 0; idle.t0 = 0;
        __threadfence_block();
        for (; rd != wr; rd++) {
            const int v = oq.v[rd & (CABAC_OUTQ - 1)], cur = v & 255;
            const int carry = last >= 0 && (v >> 8) != last;       // the byte before changed (by one): a carry went through it
            o.byte(cur | (carry << 8)); last = cur;
        }
        oq.rd = rd;
    }
    o.finish();
}

// 96 threads: warp 1 produces records, lane 0 of warp 0 runs the recurrence, lane 0 of warp 2 writes the bytes. bins[0..n) is one slice's list.
template <bool SWAP> __device__ __forceinline__ void cabac_code_list(CabacOut<SWAP> &o, CabacTables &t, CabacRing &ring, CabacOutQ &oq, const uint16_t *bins, int n)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1) cabac_produce(ring, t, bins, n, lane);
    else if (warp == 0) { if (lane == 0) cabac_consume(ring, oq); }
    else if (lane == 0) cabac_write<SWAP>(ring, oq, o);
}
#define CABAC_INIT_QUEUES() do { if (threadIdx.x == 0) { ring.wr = 0; ring.rd = 0; ring.done = 0; ring.abort = 0; oq.wr = 0; oq.rd = 0; oq.done = 0; } } while (0)

// grid: (num_slices, 1, sessions), 96 threads
__global__ void __launch_bounds__(96) k_cabac_code(const Sess *ss, Geom g, WaveCtl *ctl)
{
    const Sess &s = ss[blockIdx.z];
    const int sl = blockIdx.x, m0 = g.slice_row0[sl] * g.mbw, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *rb = s.rbsp + (size_t)sl * s.rbsp_words_per_slice;
    __shared__ CabacTables tabs;
    __shared__ CabacRing ring;
    __shared__ CabacOutQ oq;
    __shared__ int hdr_bytes_s;
    CABAC_INIT_QUEUES();
    if (warp == 1) cabac_tables_init(tabs, s.qp, !s.is_idr, lane);
    else if (warp == 0) {
        if (lane < 8) rb[lane] = 0;
        __syncwarp();
        if (lane == 0) {             // slice_header(), 7.3.3, then cabac_alignment_one_bit
            BitSink<1> bs; bs.w = rb; bs.pos = 0; bs.acc = 0ull;
            bs.ue((uint32_t)m0);
            bs.ue(s.is_idr ? 7 : 5);
            bs.ue(0);
            bs.put(8, (uint32_t)(s.frame_num & 255));
            if (s.is_idr) bs.ue((uint32_t)s.idr_pic_id);
            if (!s.is_idr) { bs.put(1, 0); bs.put(1, 0); }
            if (s.is_idr) { bs.put(1, 0); bs.put(1, 0); } else bs.put(1, 0);
            if (!s.is_idr) bs.ue(0);                                        // cabac_init_idc
            bs.se(s.qp - 26);
            bs.ue(0); bs.se(0); bs.se(0);
            const int pad = (8 - (bs.pos & 7)) & 7;
            if (pad) bs.put(pad, (1u << pad) - 1u);
            hdr_bytes_s = bs.pos >> 3;
        }
    }
    __syncthreads();
    CabacOut<true> o; o.base = reinterpret_cast<uint8_t *>(rb); o.pos = hdr_bytes_s; o.hold = -1; o.n_ff = 0; o.cap = (int)s.rbsp_words_per_slice * 4;
    cabac_code_list<true>(o, tabs, ring, oq, s.bins + (size_t)m0 * B200_MB_BIN_SLOT, (int)s.slice_nbins[sl]);
    if (threadIdx.x == 64) { s.slice_bits[sl] = (uint32_t)min(o.pos, o.cap) * 8u; if (ring.abort) atomicExch(&ctl->error, 3); }
}

// test entry: code one bin list into plain bytes (b200k_cabac_code)
__global__ void __launch_bounds__(96) k_cabac_code_test(const uint16_t *bins, int n, int qp, int is_p, uint8_t *out, int *out_len)
{
    __shared__ CabacTables tabs;
    __shared__ CabacRing ring;
    __shared__ CabacOutQ oq;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CABAC_INIT_QUEUES();
    if (warp == 1) cabac_tables_init(tabs, qp, is_p != 0, lane);
    __syncthreads();
    CabacOut<false> o; o.base = out; o.pos = 0; o.hold = -1; o.n_ff = 0; o.cap = 0x7fffffff;
    cabac_code_list<false>(o, tabs, ring, oq, bins, n);
    if (threadIdx.x == 64) *out_len = o.pos;
}

// bench entry: `gridDim.x` independent copies of one list, each coded by its own CTA (b200k_cabac_code_bench)
__global__ void __launch_bounds__(96) k_cabac_code_multi(const uint16_t *bins, int bins_stride, int n, int qp, int is_p, uint8_t *out, int out_stride, int *out_len)
{
    __shared__ CabacTables tabs;
    __shared__ CabacRing ring;
    __shared__ CabacOutQ oq;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CABAC_INIT_QUEUES();
    if (warp == 1) cabac_tables_init(tabs, qp, is_p != 0, lane);
    __syncthreads();
    CabacOut<false> o; o.base = out + (size_t)blockIdx.x * out_stride; o.pos = 0; o.hold = -1; o.n_ff = 0; o.cap = 0x7fffffff;
    cabac_code_list<false>(o, tabs, ring, oq, bins + (size_t)blockIdx.x * bins_stride, n);
    if (threadIdx.x == 64) out_len[blockIdx.x] = o.pos;
}

} // namespace b200
