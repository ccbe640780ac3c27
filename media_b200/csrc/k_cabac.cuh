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
//   k_cabac_code  : slice header + the arithmetic coder over the slice's bin list; the list streams through shared memory by
//                   cp.async double buffering, one lane runs the range/low recurrence                               [warp per slice]
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
    __device__ __forceinline__ void put(int ctx, int bin) { if (WRITE) p[n] = (uint16_t)(ctx | (bin << 10)); n++; }
    __device__ __forceinline__ void run(int ctx, int bin, int rep) { if (WRITE) p[n] = (uint16_t)(ctx | (bin << 10) | ((rep - 1) << 11)); n++; }
    // a string of bypass bins, first bin = most significant of `len` bits: entries of six, the remainder last
    __device__ __forceinline__ void bypass(uint32_t bits, int len)
    {
        while (len > 0) {
            const int k = min(len, 6);
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

__device__ __forceinline__ bool is_intra_type(int t) { return t == MB_I16x16 || t == MB_I4x4; }

// grid: (ceil(n_mb / 256), 1, sessions)
__global__ void __launch_bounds__(256) k_cabac_side(const Sess *ss, Geom g)
{
    const int mb = blockIdx.x * 256 + threadIdx.x;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    const MbInfo *mi = s.mbi + mb; const MbCoef *co = s.coef + mb;
    const int mx = mb % g.mbw, my = mb / g.mbw;
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
        const bool l4 = left && ml->mb_type == MB_I4x4, t4 = top && mt->mb_type == MB_I4x4;
        for (int k = 0; k < 16; k++) {       // predIntra4x4PredMode, 8.3.1.1
            const int bx = blk_x(k), by = blk_y(k);
            const int ma = bx > 0 ? mi->i4_mode[xy2blk(bx - 1, by)] : !left ? -1 : l4 ? ml->i4_mode[xy2blk(3, by)] : 2;
            const int mb_ = by > 0 ? mi->i4_mode[xy2blk(bx, by - 1)] : !top ? -1 : t4 ? mt->i4_mode[xy2blk(bx, 3)] : 2;
            const int pm = (ma < 0 || mb_ < 0) ? 2 : min(ma, mb_), m = mi->i4_mode[k];
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
        else { b0 = 3 + (L && L->mb_type != MB_I4x4) + (T && T->mb_type != MB_I4x4); c_cl = 6; c_cc = 7; c_cc2 = 8; c_m1 = 9; c_m0 = 10; }
        s.put(b0, t == MB_I16x16);
        if (t == MB_I16x16) {
            s.put(276, 0);
            s.put(c_cl, cl != 0); s.put(c_cc, cc != 0);
            if (cc) s.put(c_cc2, cc == 2);
            s.put(c_m1, m->i16_mode >> 1); s.put(c_m0, m->i16_mode & 1);
        }
    }
    if (t == MB_I4x4)
        for (int k = 0; k < 16; k++) {
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
        if (i16) { it.lv = co->luma[k] + 1; it.n = 15; it.cat = 1; } else { it.lv = co->luma[k]; it.n = 16; it.cat = 2; }
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
// grid: (ceil(n_mb / CABAC_WARPS), 1, sessions). WRITE = 0: mb_bits[mb] = number of entries of the MB. WRITE = 1: the entries are
// stored at bins[slice base + mb_off[mb]]; the slice base is first_mb * B200_MB_BIN_SLOT (room for the worst case of every MB).
template <int WRITE> __global__ void __launch_bounds__(CABAC_WARPS * 32) k_cabac_bins(const Sess *ss, Geom g)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * CABAC_WARPS + warp;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    const MbInfo *mi = s.mbi + mb; const MbCoef *co = s.coef + mb;
    const int mx = mb % g.mbw, my = mb / g.mbw;
    __align__(16) int16_t lv[16];
    CabacItem it = cabac_item(s, g, mx, my, lane, mi, co);
    if (lane >= 1 && lane < 28 && it.present)
        for (int i = 0; i < it.n; i++) lv[i] = it.lv[i];
    int sl = 0;
    for (int k = 1; k < g.num_slices; k++) sl += (my >= g.slice_row0[k]);
    const bool last_mb = mb == g.slice_row0[sl + 1] * g.mbw - 1;
    int cnt;
    {
        BinSink<0> bs; bs.p = nullptr; bs.n = 0;
        if (lane == 0) bin_mb_header<0>(bs, s, g, mx, my, mi);
        else if (lane < 28) { if (it.present) bin_residual<0>(bs, lv, it.n, it.cat, it.inc); }
        else if (lane == 28) bs.put(276, last_mb);
        cnt = bs.n;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (!WRITE) { if (lane == 31) s.mb_bits[mb] = (uint32_t)incl; return; }
    BinSink<1> bs; bs.p = s.bins + (size_t)g.slice_row0[sl] * g.mbw * B200_MB_BIN_SLOT + s.mb_off[mb] + (incl - cnt); bs.n = 0;
    if (lane == 0) bin_mb_header<1>(bs, s, g, mx, my, mi);
    else if (lane < 28) { if (it.present) bin_residual<1>(bs, lv, it.n, it.cat, it.inc); }
    else if (lane == 28) bs.put(276, last_mb);
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

// ---- the arithmetic coder (9.3.4.2) ----
// `low` keeps the 10 bits of the standard's codILow plus `nb` bits above them that have not been written yet (nb starts at -1:
// the standard drops the first bit). A byte leaves as soon as nb reaches 8; it may carry into the bytes before it, so the last
// byte that is not 0xFF is held back together with the count of 0xFF bytes behind it.
template <bool SWAP> struct CabacOut {
    uint8_t *base; int pos; int hold, n_ff;
    __device__ __forceinline__ void store(int v) { base[SWAP ? (pos ^ 3) : pos] = (uint8_t)v; pos++; }
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
struct CabacTables {      // shared-memory copies: (pStateIdx << 1 | valMPS) -> next state, pStateIdx -> the four rangeTabLPS values in one word
    uint32_t range_lps[64];
    uint8_t next_mps[128], next_lps[128];
    uint8_t state[CABAC_NCTX + 4];
};
__device__ __forceinline__ void cabac_tables_init(CabacTables &t, int qp, bool is_p, int lane)
{
    for (int i = lane; i < 64; i += 32)
        t.range_lps[i] = c_cabac_range_lps[i * 4] | (c_cabac_range_lps[i * 4 + 1] << 8) | (c_cabac_range_lps[i * 4 + 2] << 16) | (c_cabac_range_lps[i * 4 + 3] << 24);
    for (int i = lane; i < 128; i += 32) {
        const int p = i >> 1, mps = i & 1;
        t.next_mps[i] = (uint8_t)((c_cabac_next_mps[p] << 1) | mps);
        t.next_lps[i] = (uint8_t)((c_cabac_next_lps[p] << 1) | (p == 0 ? 1 - mps : mps));
    }
    const int q = clip3(0, 51, qp);
    for (int i = lane; i < CABAC_NCTX; i += 32) {                           // 9.3.1.1
        const int m = is_p ? c_cabac_init_p0[2 * i] : c_cabac_init_i[2 * i], n = is_p ? c_cabac_init_p0[2 * i + 1] : c_cabac_init_i[2 * i + 1];
        const int pre = clip3(1, 126, ((m * q) >> 4) + n);
        t.state[i] = (uint8_t)(pre <= 63 ? (63 - pre) << 1 : ((pre - 64) << 1) | 1);
    }
}
struct CabacCore { uint32_t low, range; int nb; };
// codes entries e[0..n) with the calling thread
template <bool SWAP> __device__ __forceinline__ void cabac_run(CabacCore &c, CabacOut<SWAP> &o, CabacTables &t, const uint16_t *e, int n)
{
    uint32_t low = c.low, range = c.range; int nb = c.nb;
    for (int i = 0; i < n; i++) {
        const uint32_t v = e[i]; const int ctx = v & 1023;
        if (ctx > CABAC_BYPASS0) {
            const int k = ctx - CABAC_BYPASS0;
            low = (low << k) + ((v >> 10) & ((1u << k) - 1u)) * range; nb += k;
        } else if (ctx == 276) {
            range -= 2;
            if ((v >> 10) & 1) {                                            // end_of_slice_flag = 1: EncodeFlush, 9.3.4.5
                low += range; low |= 1u;                                    // the last of the ten bits is the rbsp_stop_one_bit
                int width = nb + 10; const int pad = (8 - (width & 7)) & 7;
                low <<= pad; width += pad;                                  // rbsp_alignment_zero_bit
                while (width >= 8) { width -= 8; o.byte((int)(low >> width)); low &= (1u << width) - 1u; }
                o.finish(); nb = -1; range = 510; low = 0;
                continue;
            }
            const int sh = range < 256u; range <<= sh; low <<= sh; nb += sh;
        } else {
            int st = t.state[ctx]; const int bin = (v >> 10) & 1;
            for (int rep = (int)(v >> 11); rep >= 0; rep--) {
                const uint32_t rlps = (t.range_lps[st >> 1] >> ((range >> 3) & 24)) & 255u;
                range -= rlps;
                if (bin != (st & 1)) { low += range; range = rlps; st = t.next_lps[st]; } else st = t.next_mps[st];
                const int sh = __clz(range) - 23;                           // range < 512: shift up to bit 8
                range <<= sh; low <<= sh; nb += sh;
                if (nb >= 8) { nb -= 8; o.byte((int)(low >> (nb + 10))); low &= (1u << (nb + 10)) - 1u; }
            }
            t.state[ctx] = (uint8_t)st;
            continue;
        }
        if (nb >= 8) { nb -= 8; o.byte((int)(low >> (nb + 10))); low &= (1u << (nb + 10)) - 1u; }
    }
    c.low = low; c.range = range; c.nb = nb;
}

#define CABAC_CHUNK 1024        /* entries per shared-memory buffer (2 KB); every lane brings in four 16-byte pieces */
__device__ __forceinline__ void cabac_fetch(uint16_t *dst, const uint16_t *src, int lane)
{
#pragma unroll
    for (int k = 0; k < CABAC_CHUNK / 8 / 32; k++) {
        const int i = (k * 32 + lane) * 8;
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + i);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
// one warp codes bins[0..n) (n entries, buffer readable up to the next multiple of CABAC_CHUNK)
template <bool SWAP> __device__ __forceinline__ void cabac_code_list(CabacOut<SWAP> &o, CabacTables &t, uint16_t (*buf)[CABAC_CHUNK], const uint16_t *bins, int n, int lane)
{
    CabacCore c; c.low = 0; c.range = 510; c.nb = -1;
    const int nch = (n + CABAC_CHUNK - 1) / CABAC_CHUNK;
    if (nch > 0) cabac_fetch(buf[0], bins, lane);
    for (int ch = 0; ch < nch; ch++) {
        if (ch + 1 < nch) { cabac_fetch(buf[(ch + 1) & 1], bins + (size_t)(ch + 1) * CABAC_CHUNK, lane); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (lane == 0) cabac_run<SWAP>(c, o, t, buf[ch & 1], min(CABAC_CHUNK, n - ch * CABAC_CHUNK));
        __syncwarp();
    }
}

// grid: (num_slices, 1, sessions), one warp
__global__ void __launch_bounds__(32) k_cabac_code(const Sess *ss, Geom g)
{
    const Sess &s = ss[blockIdx.z];
    const int sl = blockIdx.x, m0 = g.slice_row0[sl] * g.mbw, lane = threadIdx.x;
    uint32_t *rb = s.rbsp + (size_t)sl * s.rbsp_words_per_slice;
    __shared__ CabacTables tabs;
    __shared__ __align__(16) uint16_t buf[2][CABAC_CHUNK];
    __shared__ int hdr_bytes_s;
    cabac_tables_init(tabs, s.qp, !s.is_idr, lane);
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
        if (!s.is_idr) bs.ue(0);                                            // cabac_init_idc
        bs.se(s.qp - 26);
        bs.ue(0); bs.se(0); bs.se(0);
        const int pad = (8 - (bs.pos & 7)) & 7;
        if (pad) bs.put(pad, (1u << pad) - 1u);
        hdr_bytes_s = bs.pos >> 3;
    }
    __syncwarp();
    CabacOut<true> o; o.base = reinterpret_cast<uint8_t *>(rb); o.pos = hdr_bytes_s; o.hold = -1; o.n_ff = 0;
    cabac_code_list<true>(o, tabs, buf, s.bins + (size_t)m0 * B200_MB_BIN_SLOT, (int)s.slice_nbins[sl], lane);
    if (lane == 0) s.slice_bits[sl] = (uint32_t)o.pos * 8u;
}

// test entry: code one bin list into plain bytes (b200k_cabac_code)
__global__ void __launch_bounds__(32) k_cabac_code_test(const uint16_t *bins, int n, int qp, int is_p, uint8_t *out, int *out_len)
{
    __shared__ CabacTables tabs;
    __shared__ __align__(16) uint16_t buf[2][CABAC_CHUNK];
    const int lane = threadIdx.x;
    cabac_tables_init(tabs, qp, is_p != 0, lane);
    __syncwarp();
    CabacOut<false> o; o.base = out; o.pos = 0; o.hold = -1; o.n_ff = 0;
    cabac_code_list<false>(o, tabs, buf, bins, n, lane);
    if (lane == 0) *out_len = o.pos;
}

} // namespace b200
