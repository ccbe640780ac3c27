// media_b200/csrc/k_cavlc.cuh -- entropy coding and NAL packaging (phases D and F of DESIGN.md 3).
//
// Role inside the reference: slice-data writing inside ISVCEncoder::EncodeFrame
// (video_codec/VideoEncoderOpenH264.cpp:344; openh264's WelsSpatialWriteMbSyn, WelsWriteMbResidual,
// WriteBlockResidualCavlc, CavlcParamCal_c in the absent libopenh264) and the layout contract of the output
// buffer (:349-350: all NALs contiguous, 4-byte start codes, SPS+PPS in front of every IDR).
//   k_pskip_scan : MV prediction (8.4.1.3), P_Skip detection (8.4.1.1), mb_skip_run per MB      [CTA per slice]
//   k_cavlc_mb   : macroblock_layer() of every coded MB into its own scratch slot; the 28 syntax groups of an MB
//                  (header, luma DC, 16 luma, 2 chroma DC, 8 chroma AC) are coded by 28 lanes in parallel after a
//                  warp prefix sum of their code lengths                                          [warp per MB]
//   k_slice_scan : slice header + prefix sum of MB lengths -> bit offset of every MB            [CTA per slice]
//   k_slice_copy : bit-exact concatenation: every MB shifts its slot into place              [warp per MB]
//   k_nal_pack   : start codes, NAL headers, emulation prevention (7.4.1), final access unit      [CTA per session]
#pragma once
#include "h264_dev.cuh"

namespace b200 {

__device__ __forceinline__ bool is_inter_type(int t) { return t == MB_P16x16 || t == MB_PSKIP || t == MB_P8x8; }

// Neighbouring partition for MV prediction (6.4.11.7): the 8x8 partition covering luma location (x,y) relative to MB (mx,my).
// Partitions of the current MB with index >= cur_part are not decoded yet. One reference picture: ref is 0 for inter, -1 else.
struct NbMv { int avail, ref, x, y; };
__device__ __forceinline__ NbMv nb_at(const Sess &s, const Geom &g, int mx, int my, int x, int y, int cur_part)
{
    NbMv r; r.avail = 0; r.ref = -1; r.x = 0; r.y = 0;
    const int nx = mx + (x < 0 ? -1 : x > 15 ? 1 : 0), ny = my - (y < 0);
    if (nx < 0 || nx >= g.mbw) return r;
    if (ny != my && row_is_slice_top(g, my)) return r;
    if (ny == my && nx > mx) return r;
    const int part = (((y + 16) & 15) >> 3) * 2 + (((x + 16) & 15) >> 3);
    if (nx == mx && ny == my && part >= cur_part) return r;
    const MbInfo *m = s.mbi + ny * g.mbw + nx;
    r.avail = 1;
    if (is_inter_type(m->mb_type)) {
        const uint32_t v = reinterpret_cast<const uint32_t *>(m)[2 + part];
        r.ref = 0; r.x = (int)(int16_t)(v & 0xffffu); r.y = (int)(int16_t)(v >> 16);
    }
    return r;
}
// 8.4.1.3 for the 16x16 partition (part < 0) or the 8x8 partition `part`; A and B are returned for the P_Skip rule
__device__ __forceinline__ void predict_mv_part(const Sess &s, const Geom &g, int mx, int my, int part, int &pmx, int &pmy, NbMv &A, NbMv &B)
{
    const int px = part < 0 ? 0 : (part & 1) * 8, py = part < 0 ? 0 : (part >> 1) * 8, w = part < 0 ? 16 : 8, cur = part < 0 ? 0 : part;
    A = nb_at(s, g, mx, my, px - 1, py, cur); B = nb_at(s, g, mx, my, px, py - 1, cur);
    NbMv C = nb_at(s, g, mx, my, px + w, py - 1, cur);
    if (!C.avail) C = nb_at(s, g, mx, my, px - 1, py - 1, cur);
    NbMv b2 = B;
    if (!B.avail && !C.avail && A.avail) { b2 = A; C = A; }
    const int n = (A.ref == 0) + (b2.ref == 0) + (C.ref == 0);
    if (n == 1) { const NbMv &o = A.ref == 0 ? A : b2.ref == 0 ? b2 : C; pmx = o.x; pmy = o.y; }
    else { pmx = median3(A.x, b2.x, C.x); pmy = median3(A.y, b2.y, C.y); }
}
// 16x16 prediction and the P_Skip vector of 8.4.1.1
__device__ __forceinline__ void predict_mv(const Sess &s, const Geom &g, int mx, int my, int &pmx, int &pmy, int &skx, int &sky)
{
    NbMv A, B;
    predict_mv_part(s, g, mx, my, -1, pmx, pmy, A, B);
    if (!A.avail || !B.avail || (A.ref == 0 && A.x == 0 && A.y == 0) || (B.ref == 0 && B.x == 0 && B.y == 0)) { skx = 0; sky = 0; }
    else { skx = pmx; sky = pmy; }
}

// grid: (num_slices, 1, sessions), 256 threads. skip_run[mb] = P_Skip MBs immediately before mb inside the slice;
// skip_run[n_mb + slice] = trailing run of the slice.
__global__ void __launch_bounds__(256) k_pskip_scan(const Sess *ss, Geom g)
{
    const Sess &s = ss[blockIdx.z];
    const int sl = blockIdx.x, m0 = g.slice_row0[sl] * g.mbw, m1 = g.slice_row0[sl + 1] * g.mbw, nmb = g.mbw * g.mbh;
    __shared__ int warp_last[8];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = m0 - 1;
    __syncthreads();
    for (int base = m0; base < m1; base += 256) {
        const int mb = base + threadIdx.x;
        int last = -1;                       // index of this MB if it is NOT skipped
        if (mb < m1) {
            MbInfo *mi = s.mbi + mb;
            bool skip = false;
            if (!s.is_idr && mi->mb_type == MB_P16x16 && mi->cbp == 0) {
                int pmx, pmy, skx, sky; predict_mv(s, g, mb % g.mbw, mb / g.mbw, pmx, pmy, skx, sky);
                skip = mi->mv[0] == skx && mi->mv[1] == sky;
                if (skip) mi->mb_type = MB_PSKIP;
            }
            if (!skip) last = mb;
        }
        // inclusive max-scan of `last` over the block
        int v = last;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = max(v, t); }
        if (lane == 31) warp_last[warp] = v;
        __syncthreads();
        int prev = carry_s;
        for (int w = 0; w < warp; w++) prev = max(prev, warp_last[w]);
        int excl = __shfl_up_sync(0xffffffffu, v, 1);
        excl = lane == 0 ? prev : max(prev, excl);
        if (mb < m1) s.skip_run[mb] = mb - 1 - excl;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = max(prev, v);
        __syncthreads();
    }
    if (threadIdx.x == 0) s.skip_run[nmb + sl] = m1 - 1 - carry_s;
}

// ---- bit sink, MSB first. MODE 0: count only; MODE 1: OR into a zeroed word array shared by the lanes of a warp;
// MODE 2: collect up to 64 bits in a register (pos keeps counting past 64, which is how the caller sees an overflow) ----
template <int MODE> struct BitSink {
    uint32_t *w; int pos; unsigned long long acc;
#ifdef B200_CHECKED
    int cap = 0;                                        // bits of room behind w (0: not tracked)
#endif
    __device__ __forceinline__ void put(int n, uint32_t v)
    {
#ifdef B200_CHECKED
        if (MODE == 1) B200_CHECK(cap == 0 || pos + n <= cap, 1);
#endif
        if (MODE == 1 && n > 0) {
            const int o = pos & 31, wi = pos >> 5;
            if (o + n <= 32) atomicOr(w + wi, v << (32 - o - n));
            else { atomicOr(w + wi, v >> (o + n - 32)); atomicOr(w + wi + 1, v << (64 - o - n)); }
        }
        if (MODE == 2 && n > 0 && pos + n <= 64) acc |= (unsigned long long)v << (64 - pos - n);
        pos += n;
    }
    __device__ __forceinline__ void ue(uint32_t v) { const int l = 31 - __clz(v + 1u); put(2 * l + 1, v + 1u); }
    __device__ __forceinline__ void se(int v) { ue(v > 0 ? 2u * v - 1u : (uint32_t)(-2 * v)); }
};

// residual_block_cavlc (7.3.5.3.2, 9.2) of `maxn` levels lv[0..maxn-1] in scan order; nC < 0 selects the chroma DC tables
template <int MODE> __device__ void code_residual(BitSink<MODE> &bs, const int16_t *lv, int maxn, int nC)
{
    int total = 0, last = -1, t1 = 0; bool t1_open = true;
    for (int i = maxn - 1; i >= 0; i--) {
        const int v = lv[i];
        if (!v) continue;
        if (last < 0) last = i;
        total++;
        if (t1_open && t1 < 3 && (v == 1 || v == -1)) t1++; else t1_open = false;
    }
    const int tcls = nC < 2 ? 0 : nC < 4 ? 1 : nC < 8 ? 2 : 3;
    if (nC < 0) bs.put(c_cdc_token_len[4 * total + t1], c_cdc_token_bits[4 * total + t1]);
    else bs.put(c_coeff_token_len[tcls * 68 + 4 * total + t1], c_coeff_token_bits[tcls * 68 + 4 * total + t1]);
    if (!total) return;
    const int zeros = last + 1 - total;
    // levels, highest frequency first
    int k = 0, suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
    for (int i = last; i >= 0; i--) {
        const int v = lv[i];
        if (!v) continue;
        if (k < t1) bs.put(1, v < 0);
        else {
            const int a = abs(v);
            int code = v > 0 ? 2 * a - 2 : 2 * a - 1;
            if (k == t1 && t1 < 3) code -= 2;
            if (suffix_len == 0) {
                if (code < 14) bs.put(code + 1, 1);
                else if (code < 30) { bs.put(15, 1); bs.put(4, (uint32_t)(code - 14)); }
                else { bs.put(16, 1); bs.put(12, (uint32_t)(code - 30)); }
            } else if (code < (15 << suffix_len)) {
                bs.put((code >> suffix_len) + 1, 1);
                bs.put(suffix_len, (uint32_t)(code & ((1 << suffix_len) - 1)));
            } else { bs.put(16, 1); bs.put(12, (uint32_t)(code - (15 << suffix_len))); }
            if (suffix_len == 0) suffix_len = 1;
            if (a > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
        }
        k++;
    }
    if (total < maxn) {
        if (nC < 0) bs.put(c_cdc_total_zeros_len[(total - 1) * 4 + zeros], c_cdc_total_zeros_bits[(total - 1) * 4 + zeros]);
        else bs.put(c_total_zeros_len[(total - 1) * 16 + zeros], c_total_zeros_bits[(total - 1) * 16 + zeros]);
    }
    // run_before for every coefficient but the last (lowest frequency) one
    int zl = zeros, run = 0; k = 0;
    for (int i = last - 1; i >= 0 && zl > 0 && k < total - 1; i--) {
        if (lv[i]) {
            const int zi = min(zl, 7) - 1;
            bs.put(c_run_before_len[zi * 16 + run], c_run_before_bits[zi * 16 + run]);
            zl -= run; run = 0; k++;
        } else run++;
    }
}

struct MbItem { const int16_t *lv; int maxn, nC; bool present; };

// the syntax group coded by `lane` for this MB (lanes 1..27; lane 0 is the MB header)
__device__ __forceinline__ MbItem mb_item(const Sess &s, const Geom &g, int mx, int my, int lane, const MbInfo *mi, const MbCoef *co)
{
    MbItem it; it.present = false; it.lv = co->luma_dc; it.maxn = 16; it.nC = 0;
    const bool i16 = mi->mb_type == MB_I16x16;
    const int cl = mi->cbp & 15, cc = mi->cbp >> 4;
    const bool left = mx > 0, top = !row_is_slice_top(g, my);
    const MbInfo *ml = mi - 1, *mt = mi - g.mbw;
    if (lane == 1 || (lane >= 2 && lane < 18)) {
        const int b = lane == 1 ? 0 : lane - 2, bx = blk_x(b), by = blk_y(b);
        int nA = -1, nB = -1;
        if (bx > 0) nA = mi->nnz[xy2blk(bx - 1, by)]; else if (left) nA = ml->nnz[xy2blk(3, by)];
        if (by > 0) nB = mi->nnz[xy2blk(bx, by - 1)]; else if (top) nB = mt->nnz[xy2blk(bx, 3)];
        it.nC = (nA >= 0 && nB >= 0) ? (nA + nB + 1) >> 1 : (nA >= 0 ? nA : (nB >= 0 ? nB : 0));
        if (lane == 1) { it.present = i16; }
        else if (i16) { it.present = cl != 0; it.lv = co->luma[b] + 1; it.maxn = 15; }
        else { it.present = (cl >> (b >> 2)) & 1; it.lv = co->luma[b]; it.maxn = 16; }
    } else if (lane < 20) {
        it.present = cc != 0; it.lv = co->chroma_dc[lane - 18]; it.maxn = 4; it.nC = -1;
    } else if (lane < 28) {
        const int pl = (lane - 20) >> 2, b = (lane - 20) & 3, bx = b & 1, by = b >> 1, base = 16 + pl * 4;
        int nA = -1, nB = -1;
        if (bx > 0) nA = mi->nnz[base + by * 2]; else if (left) nA = ml->nnz[base + by * 2 + 1];
        if (by > 0) nB = mi->nnz[base + bx]; else if (top) nB = mt->nnz[base + 2 + bx];
        it.nC = (nA >= 0 && nB >= 0) ? (nA + nB + 1) >> 1 : (nA >= 0 ? nA : (nB >= 0 ? nB : 0));
        it.present = cc == 2; it.lv = co->chroma_ac[pl][b] + 1; it.maxn = 15;
    }
    return it;
}

// motion vector differences of an inter MB, predicted once (8.4.1.3) and used by both passes of the header lane:
// mvd[0..1] for P_L0_16x16, mvd[2q..2q+1] for partition q of P_8x8
__device__ __forceinline__ void mb_mvds(const Sess &s, const Geom &g, int mx, int my, const MbInfo *mi, int mvd[8])
{
#pragma unroll
    for (int k = 0; k < 8; k++) mvd[k] = 0;
    if (mi->mb_type == MB_P8x8) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int pmx, pmy; NbMv A, B; predict_mv_part(s, g, mx, my, q, pmx, pmy, A, B);
            mvd[2 * q] = mi->mv8[q][0] - pmx; mvd[2 * q + 1] = mi->mv8[q][1] - pmy;
        }
    } else if (mi->mb_type == MB_P16x16) {
        int pmx, pmy, skx, sky; predict_mv(s, g, mx, my, pmx, pmy, skx, sky);
        mvd[0] = mi->mv[0] - pmx; mvd[1] = mi->mv[1] - pmy;
    }
}

template <int MODE> __device__ void code_mb_header(BitSink<MODE> &bs, const Sess &s, const Geom &g, int mx, int my, const MbInfo *mi, int skip_run, const int mvd[8])
{
    const int cl = mi->cbp & 15, cc = mi->cbp >> 4;
    if (!s.is_idr) bs.ue((uint32_t)skip_run);
    if (mi->mb_type == MB_I16x16) {
        bs.ue((uint32_t)((s.is_idr ? 0 : 5) + 1 + mi->i16_mode + 4 * cc + (cl ? 12 : 0)));
        bs.ue(mi->chroma_mode);
        bs.se(0);
    } else if (mi->mb_type == MB_I4x4) {
        // I_NxN (transform_8x8_mode_flag is 0 in the PPS): prev_intra4x4_pred_mode_flag / rem_intra4x4_pred_mode per block, 7.3.5.1, 8.3.1.1
        bs.ue(s.is_idr ? 0u : 5u);
        const bool left = mx > 0, top = !row_is_slice_top(g, my);
        const MbInfo *ml = mi - 1, *mt = mi - g.mbw;
        const bool l4 = left && ml->mb_type == MB_I4x4, t4 = top && mt->mb_type == MB_I4x4;
        for (int k = 0; k < 16; k++) {
            const int bx = blk_x(k), by = blk_y(k);
            const int ma = bx > 0 ? mi->i4_mode[xy2blk(bx - 1, by)] : !left ? -1 : l4 ? ml->i4_mode[xy2blk(3, by)] : 2;
            const int mb_ = by > 0 ? mi->i4_mode[xy2blk(bx, by - 1)] : !top ? -1 : t4 ? mt->i4_mode[xy2blk(bx, 3)] : 2;
            const int pm = (ma < 0 || mb_ < 0) ? 2 : min(ma, mb_), m = mi->i4_mode[k];
            if (m == pm) bs.put(1, 1); else bs.put(4, (uint32_t)(m < pm ? m : m - 1));
        }
        bs.ue(mi->chroma_mode);
        bs.ue(c_cbp_intra[mi->cbp]);
        if (mi->cbp) bs.se(0);
    } else if (mi->mb_type == MB_P8x8) {
        // P_8x8 with four P_L0_8x8 sub-macroblocks (7.3.5.2); ref_idx_l0 is not coded with one reference picture
        bs.ue(3);
        for (int q = 0; q < 4; q++) bs.ue(0);
#pragma unroll
        for (int q = 0; q < 4; q++) { bs.se(mvd[2 * q]); bs.se(mvd[2 * q + 1]); }
        bs.ue(c_cbp_inter[mi->cbp]);
        if (mi->cbp) bs.se(0);
    } else {
        bs.ue(0);
        bs.se(mvd[0]); bs.se(mvd[1]);
        bs.ue(c_cbp_inter[mi->cbp]);
        if (mi->cbp) bs.se(0);
    }
}

// The header of a macroblock (mb_skip_run, mb_type, prediction modes or mvd, coded_block_pattern, mb_qp_delta) is a serial walk with the MV prediction
// in front of it: as lane 0 of a warp per MB it kept 31 lanes idle for ~200 instructions. Here a THREAD owns a macroblock: the header goes left-aligned
// into words 0-1 of the MB's bit slot and its length into mb_bits. Inter macroblocks without residual (cbp 0) are complete with that -- the warp kernel
// leaves them at once; for the others its lane 0 takes the header from the slot. A header longer than 64 bits (Intra_4x4 with many explicit modes, huge
// skip runs) is marked and coded by the warp kernel's lane 0 as before.
// grid: (ceil(n_mb / 128), 1, sessions), 128 threads
#define CAVLC_HDR_LONG 0xffffffffu
__device__ __forceinline__ bool cavlc_header_only(uint32_t w0) { const uint32_t t = w0 & 255u; return (w0 >> 24) == 0u && (t == MB_P16x16 || t == MB_P8x8); }
__global__ void __launch_bounds__(128) k_cavlc_hdr(const Sess *ss, Geom g)
{
    const int mb = blockIdx.x * 128 + threadIdx.x;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    const MbInfo *mi = s.mbi + mb;
    if (mi->mb_type == MB_PSKIP) { s.mb_bits[mb] = 0; return; }
    int mx, my; mb_xy(g, mb, mx, my);
    int mvd[8];
    mb_mvds(s, g, mx, my, mi, mvd);
    BitSink<2> bs; bs.w = nullptr; bs.pos = 0; bs.acc = 0ull;
    code_mb_header<2>(bs, s, g, mx, my, mi, s.is_idr ? 0 : s.skip_run[mb], mvd);
    if (bs.pos > 64) { s.mb_bits[mb] = CAVLC_HDR_LONG; return; }
    uint32_t *dst = s.mb_slot + (size_t)mb * B200_MB_SLOT_WORDS;
    *reinterpret_cast<uint2 *>(dst) = make_uint2((uint32_t)(bs.acc >> 32), (uint32_t)bs.acc);
    s.mb_bits[mb] = (uint32_t)bs.pos;
}

#define CAVLC_WARPS 8
// grid: (ceil(n_mb / CAVLC_WARPS), 1, sessions); after k_cavlc_hdr
__global__ void __launch_bounds__(CAVLC_WARPS * 32) k_cavlc_mb(const Sess *ss, Geom g)
{
    __shared__ uint32_t slot_all[CAVLC_WARPS][B200_MB_SLOT_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * CAVLC_WARPS + warp;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    const MbInfo *mi = s.mbi + mb; const MbCoef *co = s.coef + mb;
    const uint32_t w0 = *reinterpret_cast<const uint32_t *>(mi);
    if ((w0 & 255u) == MB_PSKIP) return;                          // mb_bits = 0 was written by k_cavlc_hdr
    const uint32_t hdr_len = s.mb_bits[mb];                       // the header's length (its bits are in words 0-1 of the MB's slot), or CAVLC_HDR_LONG
    if (hdr_len != CAVLC_HDR_LONG && cavlc_header_only(w0)) return;     // no residual: the header is the whole macroblock, k_cavlc_hdr wrote it
    int mx, my; mb_xy(g, mb, mx, my);
    uint32_t *slot = slot_all[warp];
    uint32_t *dst = s.mb_slot + (size_t)mb * B200_MB_SLOT_WORDS;
    const int skip_run = s.is_idr ? 0 : s.skip_run[mb];
    __align__(16) int16_t lv[16];
    MbItem it = mb_item(s, g, mx, my, lane, mi, co);
    if (lane >= 1 && lane < 28 && it.present)
        for (int i = 0; i < it.maxn; i++) lv[i] = it.lv[i];
    int mvd[8];
    if (lane == 0 && hdr_len == CAVLC_HDR_LONG) mb_mvds(s, g, mx, my, mi, mvd);
    // pass 1: every lane codes its syntax group into a 64-bit register (and counts its length)
    int len = 0; unsigned long long acc = 0ull;
    {
        BitSink<2> bs; bs.w = nullptr; bs.pos = 0; bs.acc = 0ull;
        if (lane == 0) {
            if (hdr_len == CAVLC_HDR_LONG) code_mb_header<2>(bs, s, g, mx, my, mi, skip_run, mvd);
            else { const uint2 h = *reinterpret_cast<const uint2 *>(dst); bs.pos = (int)hdr_len; bs.acc = ((unsigned long long)h.x << 32) | h.y; }
        } else if (lane < 28 && it.present) code_residual<2>(bs, lv, it.maxn, it.nC);
        len = bs.pos; acc = bs.acc;
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, 31), nwords = (total + 31) >> 5;
    for (int i = lane; i <= nwords && i < B200_MB_SLOT_WORDS; i += 32) slot[i] = 0;
    __syncwarp();
    if (__ballot_sync(0xffffffffu, len > 64) == 0) {
        // the usual case: no group is longer than 64 bits; drop the registers in at the lanes' bit offsets
        if (len > 0) {
            const int off = incl - len, o = off & 31, wi = off >> 5;
            const uint32_t hi = (uint32_t)(acc >> 32), lo = (uint32_t)acc;
            const uint32_t w0 = hi >> o, w1 = o ? (hi << (32 - o)) | (lo >> o) : lo, w2 = o ? lo << (32 - o) : 0u;
            if (w0) atomicOr(slot + wi, w0);
            if (w1) atomicOr(slot + wi + 1, w1);
            if (w2) atomicOr(slot + wi + 2, w2);
        }
    } else {
        // pass 2 (long groups: high-rate intra blocks): code again, writing at the lane's bit offset
        BitSink<1> bs; bs.w = slot; bs.pos = incl - len; bs.acc = 0ull;
#ifdef B200_CHECKED
        bs.cap = B200_MB_SLOT_WORDS * 32;
#endif
        if (lane == 0) {
            if (hdr_len == CAVLC_HDR_LONG) code_mb_header<1>(bs, s, g, mx, my, mi, skip_run, mvd);
            else {      // the precomputed header: len bits, left-aligned in acc
                const int n1 = min(len, 32), n2 = len - n1;
                if (n1) bs.put(n1, (uint32_t)(acc >> (64 - n1)));
                if (n2) bs.put(n2, (uint32_t)(acc >> (32 - n2)) & (n2 == 32 ? 0xffffffffu : (1u << n2) - 1u));
            }
        } else if (lane < 28 && it.present) code_residual<1>(bs, lv, it.maxn, it.nC);
    }
    __syncwarp();
    for (int i = lane; i < nwords; i += 32) dst[i] = slot[i];
    if (lane == 0) s.mb_bits[mb] = (uint32_t)total;
}

// ---- slice assembly ----
__device__ __forceinline__ int block_excl_scan(int v, int *total, int *wsum)   // 256 threads
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { if (w < warp) base += wsum[w]; tot += wsum[w]; }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

// grid: (num_slices, 1, sessions), 256 threads: bit offset of every MB inside its slice's RBSP (prefix sum of the MB
// lengths), slice header, zeroed RBSP words, trailing mb_skip_run and rbsp_trailing_bits.
__global__ void __launch_bounds__(256) k_slice_scan(const Sess *ss, Geom g)
{
    const Sess &s = ss[blockIdx.z];
    const int sl = blockIdx.x, m0 = g.slice_row0[sl] * g.mbw, m1 = g.slice_row0[sl + 1] * g.mbw, nmb = g.mbw * g.mbh;
    uint32_t *rb = s.rbsp + (size_t)sl * s.rbsp_words_per_slice;
    __shared__ int wsum[8];
    __shared__ uint32_t hdr[4];
    __shared__ int hdr_bits_s;
    if (threadIdx.x == 0) {      // slice_header(), 7.3.3
        hdr[0] = hdr[1] = hdr[2] = hdr[3] = 0;
        BitSink<1> bs; bs.w = hdr; bs.pos = 0; bs.acc = 0ull;
        bs.ue((uint32_t)m0);
        bs.ue(s.is_idr ? 7 : 5);
        bs.ue(0);
        bs.put(8, (uint32_t)(s.frame_num & 255));
        if (s.is_idr) bs.ue((uint32_t)s.idr_pic_id);
        if (!s.is_idr) { bs.put(1, 0); bs.put(1, 0); }
        if (s.is_idr) { bs.put(1, 0); bs.put(1, 0); } else bs.put(1, 0);
        bs.se(s.qp - 26);
        bs.ue(0); bs.se(0); bs.se(0);
        hdr_bits_s = bs.pos;
    }
    __syncthreads();
    const int hdr_bits = hdr_bits_s;
    int carry = 0;
    for (int base = m0; base < m1; base += 256) {
        const int mb = base + threadIdx.x;
        const int len = mb < m1 ? (int)s.mb_bits[mb] : 0;
        int chunk_total; const int off = carry + block_excl_scan(len, &chunk_total, wsum);
        if (mb < m1) s.mb_off[mb] = (uint32_t)(hdr_bits + off);
        carry += chunk_total;
    }
    const int total = carry;
    const int trailing_run = s.is_idr ? 0 : s.skip_run[nmb + sl];
    const int tail_bits = trailing_run ? ue_len((uint32_t)trailing_run) : 0;
    const int data_end = hdr_bits + total + tail_bits;               // position of the rbsp_stop_one_bit
    const int all_bits = (data_end + 1 + 7) & ~7;
    const int nwords = (all_bits + 31) >> 5;
    for (int i = threadIdx.x; i <= nwords; i += 256) rb[i] = i < 4 ? hdr[i] : 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        BitSink<1> bs; bs.w = rb; bs.pos = hdr_bits + total; bs.acc = 0ull;
        if (trailing_run) bs.ue((uint32_t)trailing_run);
        bs.put(1, 1);
        s.slice_bits[sl] = (uint32_t)all_bits;
    }
}

// grid: (ceil(n_mb / CAVLC_WARPS), 1, sessions): every coded MB shifts its slot into place inside the slice RBSP
__global__ void __launch_bounds__(CAVLC_WARPS * 32) k_slice_copy(const Sess *ss, Geom g)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * CAVLC_WARPS + warp;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    const int l = (int)s.mb_bits[mb];
    if (!l) return;
    const int my = mb / g.mbw;
    int sl = 0;
    for (int k = 1; k < g.num_slices; k++) sl += (my >= g.slice_row0[k]);
    uint32_t *rb = s.rbsp + (size_t)sl * s.rbsp_words_per_slice;
    const uint32_t *src = s.mb_slot + (size_t)mb * B200_MB_SLOT_WORDS;
    const int D = (int)s.mb_off[mb], nw = (l + 31) >> 5, dw = D >> 5, sh = D & 31;
    B200_CHECK((size_t)dw + (size_t)nw + 1 <= (size_t)s.rbsp_words_per_slice && nw <= B200_MB_SLOT_WORDS, 4);
    for (int i = lane; i <= nw; i += 32) {          // destination word dw + i
        const uint32_t hi = i > 0 ? src[i - 1] : 0u, lo = i < nw ? src[i] : 0u;
        const uint32_t v = sh ? (hi << (32 - sh)) | (lo >> sh) : lo;
        if (v) atomicOr(rb + dw + i, v);
    }
}

// grid: (sessions), 1024 threads. Output: [SPS PPS] then one NAL per slice with emulation prevention.
__global__ void __launch_bounds__(1024) k_nal_pack(const Sess *ss, Geom g)
{
    const Sess &s = ss[blockIdx.x];
    __shared__ int wsum[32];
    __shared__ int out_pos_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) out_pos_s = 0;
    __syncthreads();
    if (s.is_idr) {
        for (int i = threadIdx.x; i < s.hdr_len; i += 1024) s.out[i] = s.hdr[i];
        __syncthreads();
        if (threadIdx.x == 0) out_pos_s = s.hdr_len;
        __syncthreads();
    }
    for (int sl = 0; sl < g.num_slices; sl++) {
        const uint32_t *rb = s.rbsp + (size_t)sl * s.rbsp_words_per_slice;
        const int nbytes = (int)(s.slice_bits[sl] >> 3);
        int o0 = out_pos_s;
        __syncthreads();
        if (threadIdx.x < 5 && o0 + 5 <= (int)s.out_cap) s.out[o0 + threadIdx.x] = threadIdx.x == 3 ? 1 : threadIdx.x == 4 ? (s.is_idr ? 0x65 : 0x61) : 0;
        o0 += 5;
        for (int base = 0; base < nbytes; base += 1024) {
            const int i = base + threadIdx.x;
            int b = 256, flag = 0;
            if (i < nbytes) {
                b = (rb[i >> 2] >> (24 - 8 * (i & 3))) & 255;
                if (b <= 3) {
                    int k = 0;               // zero bytes immediately before i
                    while (k < i && ((rb[(i - 1 - k) >> 2] >> (24 - 8 * ((i - 1 - k) & 3))) & 255) == 0) k++;
                    flag = k >= 2 && !(k & 1);
                }
            }
            int incl = flag;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            int pre = 0, tot = 0;
            for (int w = 0; w < 32; w++) { if (w < warp) pre += wsum[w]; tot += wsum[w]; }
            if (i < nbytes) {
                int o = o0 + i + pre + incl - flag;
                if (o + 2 <= (int)s.out_cap) { if (flag) s.out[o++] = 3; s.out[o] = (uint8_t)b; }
            }
            __syncthreads();
            o0 += tot;
        }
        if (threadIdx.x == 0) out_pos_s = o0 + nbytes;
        __syncthreads();
    }
    if (threadIdx.x == 0) { s.out_size[0] = (uint32_t)min(out_pos_s, (int)s.out_cap); s.out_size[1] = (uint32_t)s.is_idr; }   // size, then the frame kind actually coded
}

} // namespace b200
