// media_b200/csrc/k_intra.cuh -- Intra_16x16 macroblocks on a macroblock wavefront (phase C of DESIGN.md 3).
//
// Role inside the reference: intra prediction, mode decision, transform and reconstruction inside
// ISVCEncoder::EncodeFrame (video_codec/VideoEncoderOpenH264.cpp:344; openh264's WelsMdI16x16, WelsMdI4x4,
// WelsI16x16LumaPred*, WelsI4x4LumaPred*, WelsIChromaPred*, WelsHadamardT4Dc, WelsDequantIHadamard4x4 in the absent libopenh264).
// One warp owns one macroblock row of one session; rows are handed out through an atomic ticket so that a
// warp only ever waits on a row that started earlier. Lanes 0-15 own the 16 luma 4x4 blocks, lanes 16-23 the
// 8 chroma blocks; luma and chroma predictors share one code path parameterised per lane.
#pragma once
#include "h264_dev.cuh"
#include "k_me.cuh"
#include "k_t8.cuh"
#include "i8_tables.cuh"

namespace b200 {

#define WAVE_WARPS 4
#ifdef INTRA_TIMING
__device__ long long g_intra_t[8];
#define INTRA_T(i) do { if (lane == 0) { const long long t_ = clock64(); atomicAdd((unsigned long long *)&g_intra_t[i], (unsigned long long)(t_ - it_last)); it_last = t_; } } while (0)
#else
#define INTRA_T(i) do { } while (0)
#endif
#define WAVE_TIMEOUT_NS 2000000000ull   /* watchdog: a wait longer than 2 s is an internal error, never a hang */


__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// wait until the row above has finished `need` macroblocks; returns false on timeout / global error. EVERY lane performs the acquire load
// (one request per warp: same address), so each lane's later reads of the neighbour row's samples and MbInfo are ordered after the
// producer's release by the memory model itself, not by a shuffle from the one lane that polled.
__device__ __forceinline__ bool wave_wait(const int *prog_above, int need, WaveCtl *ctl, int lane)
{
    (void)lane;
    int cur = ld_acquire(prog_above);
    if (cur >= need) return true;
    const unsigned long long t0 = global_ns();
    int spins = 0;
    do {
        // a row that is d macroblocks short of what we need takes d MB-steps (a few us each): sleep accordingly, so
        // that far-behind rows do not burn the issue slots of the SMs they share with other kernels
        const int d = need - cur;
        __nanosleep(d > 1 ? min(d * 1500, 30000) : 32);
        if ((++spins & 15) == 0 && (ld_acquire(&ctl->error) || global_ns() - t0 > WAVE_TIMEOUT_NS)) { if (lane == 0) atomicExch(&ctl->error, 1); return false; }
    } while ((cur = ld_acquire(prog_above)) < need);
    return true;
}

__device__ __forceinline__ void hadamard16(int v[16])
{
#pragma unroll
    for (int y = 0; y < 4; y++) {
        int a0 = v[y * 4] + v[y * 4 + 3], a1 = v[y * 4 + 1] + v[y * 4 + 2], a2 = v[y * 4 + 1] - v[y * 4 + 2], a3 = v[y * 4] - v[y * 4 + 3];
        v[y * 4] = a0 + a1; v[y * 4 + 1] = a3 + a2; v[y * 4 + 2] = a0 - a1; v[y * 4 + 3] = a3 - a2;
    }
#pragma unroll
    for (int x = 0; x < 4; x++) {
        int a0 = v[x] + v[12 + x], a1 = v[4 + x] + v[8 + x], a2 = v[4 + x] - v[8 + x], a3 = v[x] - v[12 + x];
        v[x] = a0 + a1; v[4 + x] = a3 + a2; v[8 + x] = a0 - a1; v[12 + x] = a3 - a2;
    }
}

// top/left: index 0 = the corner sample p[-1,-1]; luma top also carries the 8 samples of the MB above-right (17..24; Intra_4x4 needs 4 of
// them, Intra_8x8 all). nb: luma reconstruction of the MB being coded as I_NxN with its border, 17 rows x NBP bytes: sample (x,y) at
// (y+1)*NBP + 4 + x, x in [-1,23] on row y = -1, x in [-1,15] below. F: the edge tables of the block in flight (see c_i4_idx / c_i8_idx).
#define NBP 28
struct IntraSmem {
    uint8_t top[3][28], left[3][20];
    uint32_t nb[17 * NBP / 4];
    uint32_t srcw[64];
    uint8_t F[96];
    int8_t mg[28];             // Intra4x4PredMode grid, 5x5: row 0 / column 0 = neighbouring MBs (-1 unavailable, 2 not I_NxN)
    __align__(16) int dq[16];  // dequantised coefficients of the Intra_4x4 block in flight (raster), then its row-transformed values
    int tf[16];
    // Intra_8x8 (High profile): transpose tiles of the 8x8 transform, and the trial's outcome, kept while Intra_4x4 is tried on the same samples
    int t8[4][8][9];
    __align__(16) int16_t keep_lv[4][64];
    uint32_t keep_rec[64];
    uint8_t keep_mode[4], keep_nnz[4];
};
__device__ __forceinline__ bool is_inxn(int t) { return t == MB_I4x4 || t == MB_I8x8; }

// Intra_4x4 predictors (8.3.1.2.1-9) as lookups into the block's filtered edge. Edge E[0..14] = L3 L3 L2 L1 L0 X T0..T7 T7
// (left column bottom-up, corner, top row, end samples doubled); F[i] = E[i], F[16+i] = (E[i]+E[i+1]+1)>>1,
// F[32+i] = (E[i]+2E[i+1]+E[i+2]+2)>>2, F[47] = DC. Entry [mode][y] packs the four F indices of row y, x = 0 in the low byte.
static __device__ __constant__ uint32_t c_i4_idx[9][4] = {
    { 0x09080706u, 0x09080706u, 0x09080706u, 0x09080706u },   // 0 vertical
    { 0x04040404u, 0x03030303u, 0x02020202u, 0x01010101u },   // 1 horizontal
    { 0x2f2f2f2fu, 0x2f2f2f2fu, 0x2f2f2f2fu, 0x2f2f2f2fu },   // 2 DC
    { 0x29282726u, 0x2a292827u, 0x2b2a2928u, 0x2c2b2a29u },   // 3 diagonal down-left
    { 0x27262524u, 0x26252423u, 0x25242322u, 0x24232221u },   // 4 diagonal down-right
    { 0x18171615u, 0x27262524u, 0x17161523u, 0x26252422u },   // 5 vertical-right
    { 0x26252414u, 0x24142313u, 0x23132212u, 0x22122111u },   // 6 horizontal-down
    { 0x19181716u, 0x29282726u, 0x1a191817u, 0x2a292827u },   // 7 vertical-left
    { 0x21122213u, 0x20112112u, 0x01012011u, 0x01010101u },   // 8 horizontal-up
};
#define I4_BIAS_BITS 24       /* fixed cost of choosing Intra_4x4 (16 mode flags), in lambda units */

// the session fields the row loop needs, in registers (every fence / strong access is a compiler memory barrier, and the L1
// invalidation behind the acquire makes re-reading them through `const Sess &` an L2 round trip each)
struct IntraCtx {
    uint8_t *rec0, *rec1, *rec2; const uint8_t *src0, *src1, *src2; MbInfo *mbi; MbCoef *coef; int qp, is_idr, no_i4x4, i8;
    __device__ __forceinline__ uint8_t *rec(int c) const { return c == 0 ? rec0 : c == 1 ? rec1 : rec2; }
    __device__ __forceinline__ const uint8_t *src(int c) const { return c == 0 ? src0 : c == 1 ? src1 : src2; }
};

// Set-up shared by the I_NxN trials of one MB: the reconstruction border, the source MB, and the prediction-mode grid of the neighbours
__device__ __forceinline__ void intra_nxn_setup(const IntraCtx &s, const Geom &g, IntraSmem &sm, int mx, int my, int lane, bool top, bool left)
{
    const int wc = g.wc, mb = my * g.mbw + mx;
    const MbInfo *mi = s.mbi + mb;
    uint8_t *nb = reinterpret_cast<uint8_t *>(sm.nb);
    if (lane < 25) nb[3 + lane] = sm.top[0][lane];
    if (lane < 16) nb[(lane + 1) * NBP + 3] = sm.left[0][lane + 1];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int wi = lane + 32 * i;
        sm.srcw[wi] = *reinterpret_cast<const uint32_t *>(s.src0 + (size_t)(my * 16 + (wi >> 2)) * wc + mx * 16 + (wi & 3) * 4);
    }
    if (lane < 25) {
        const int gy = lane / 5, gx = lane - gy * 5;
        int v = 2;
        if (gy == 0 && gx > 0) { v = -1; if (top) { const MbInfo *mt = mi - g.mbw; v = is_inxn(__ldcg(&mt->mb_type)) ? (int)__ldcg(&mt->i4_mode[xy2blk(gx - 1, 3)]) : 2; } }
        else if (gx == 0 && gy > 0) { v = -1; if (left) { const MbInfo *ml = mi - 1; v = is_inxn(__ldcg(&ml->mb_type)) ? (int)__ldcg(&ml->i4_mode[xy2blk(3, gy - 1)]) : 2; } }
        sm.mg[lane] = (int8_t)v;
    }
    __syncwarp();
}

#define I8_BIAS_BITS 12       /* fixed cost of choosing Intra_8x8 (four mode flags, the transform flag), in lambda units */
// Trial coding of the luma of one MB as Intra_8x8 (High profile, 8.3.2): the four 8x8 blocks in decoding order. Per block the 25 edge
// samples L7..L0 X T0..T15 are filtered (8.3.2.2.1, one three-tap pass with doubled ends) by lanes 0-24, the one- / two- / three-tap tables
// the predictors index (c_i8_idx) go to shared memory, lane = (mode, 4x4 quadrant) evaluates the Hadamard SATD of modes 0-7 and lanes 0-3 that
// of mode 8, and the winner goes through the 8x8 transform chain (t8_code_block: 8 lanes, one row each; the four lane groups run the same
// block on their own tiles). Nothing leaves the warp: levels, modes, nnz and the reconstruction stay in sm.keep_* while Intra_4x4 is tried
// on the same samples. Returns the SATD cost; j8 = 64 SSD + 27 lambda^2 B (B in half bits: 2 + 2 x mode bits + level costs of coded blocks).
__device__ int intra_try_i8x8(const IntraCtx &s, const Geom &g, IntraSmem &sm, const uint32_t *i8idx, int mx, int my, int lane,
                              bool top, bool left, int &cbp_luma, long long &j8)
{
    const int qp = s.qp, lambda = c_lambda[qp];
    const bool topright = top && mx + 1 < g.mbw;
    uint8_t *nb = reinterpret_cast<uint8_t *>(sm.nb);
    int total = lambda * I8_BIAS_BITS, rate = 2, ssd = 0;
    cbp_luma = 0;
    const int qd = lane & 3, qx = qd & 1, qy = qd >> 1, r = lane & 7;
#pragma unroll 1
    for (int b = 0; b < 4; b++) {
        const int bx8 = (b & 1) * 8, by8 = (b >> 1) * 8;
        const bool aT = by8 > 0 || top, aL = bx8 > 0 || left;
        const bool aX = b == 0 ? (top && left) : b == 1 ? top : b == 2 ? left : true;
        const bool aTR = aT && (b == 0 ? top : b == 1 ? topright : b == 2);
        // raw edge sample of lane k: L[7-k] (k < 8), the corner (k = 8), T[k-9] (k > 8; without a top-right neighbour T8..T15 repeat T7)
        int e = 128;
        B200_CHECK((by8 + 8) * NBP + 3 + bx8 < 17 * NBP && by8 * NBP + 4 + bx8 + 15 < 17 * NBP, 9);
        if (lane < 8) { if (aL) e = nb[(by8 + 8 - lane) * NBP + 3 + bx8]; }
        else if (lane == 8) { if (aX) e = nb[by8 * NBP + 3 + bx8]; }
        else if (lane < 25) { const int i = lane - 9; if (aT) e = nb[by8 * NBP + 4 + bx8 + ((i < 8 || aTR) ? i : 7)]; }
        int f;
        {
            int a = __shfl_up_sync(0xffffffffu, e, 1), d = __shfl_down_sync(0xffffffffu, e, 1);
            if (lane == 0 || (lane == 9 && !aX)) a = e;
            if (lane == 24 || (lane == 7 && !aX)) d = e;
            f = (a + 2 * e + d + 2) >> 2;
            if (lane < 8 ? !aL : lane == 8 ? !aX : !aT) f = 128;
        }
        {   // tables of the filtered edge: F[k] = E'[k], F[32 + k] = two-tap, F[64 + k] = three-tap with doubled ends, F[95] = DC
            int up = __shfl_up_sync(0xffffffffu, f, 1), dn = __shfl_down_sync(0xffffffffu, f, 1);
            if (lane == 0) up = f;
            if (lane == 24) dn = f;
            if (lane < 25) { sm.F[lane] = (uint8_t)f; sm.F[64 + lane] = (uint8_t)((up + 2 * f + dn + 2) >> 2); }
            if (lane < 24) sm.F[32 + lane] = (uint8_t)((f + dn + 1) >> 1);
            const int sL = __reduce_add_sync(0xffffffffu, lane < 8 ? f : 0), sT = __reduce_add_sync(0xffffffffu, lane >= 9 && lane < 17 ? f : 0);
            if (lane == 31) sm.F[95] = (uint8_t)(aT && aL ? (sT + sL + 8) >> 4 : aT ? (sT + 4) >> 3 : aL ? (sL + 4) >> 3 : 128);
        }
        // source quadrant of this lane: rows by8 + 4 qy .., word (bx8 >> 2) + qx
        uint32_t Sq[4]; int Ts[16];
        {
#pragma unroll
            for (int y = 0; y < 4; y++) Sq[y] = sm.srcw[(by8 + 4 * qy + y) * 4 + (bx8 >> 2) + qx];
            satd_source_terms(Sq, Ts);
        }
        const int ma = sm.mg[(2 * (b >> 1) + 1) * 5 + 2 * (b & 1)], mb_ = sm.mg[(2 * (b >> 1)) * 5 + 2 * (b & 1) + 1];
        const int pm = (ma < 0 || mb_ < 0) ? 2 : min(ma, mb_);
        __syncwarp();
        uint32_t key = 0xffffffffu;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            const int m = pass ? 8 : lane >> 2;
            uint32_t P[4];
#pragma unroll
            for (int y = 0; y < 4; y++) {
                const uint32_t ix = i8idx[m * 16 + (4 * qy + y) * 2 + qx];
                B200_CHECK((ix & 255) < 96 && ((ix >> 8) & 255) < 96 && ((ix >> 16) & 255) < 96 && (ix >> 24) < 96 && m < 9, 10);
                P[y] = (uint32_t)sm.F[ix & 255] | ((uint32_t)sm.F[(ix >> 8) & 255] << 8) | ((uint32_t)sm.F[(ix >> 16) & 255] << 16) | ((uint32_t)sm.F[ix >> 24] << 24);
            }
            int sat = satd_rows(P, Ts);
            sat += __shfl_xor_sync(0xffffffffu, sat, 1); sat += __shfl_xor_sync(0xffffffffu, sat, 2);
            const bool ok = (pass == 0 || lane < 4) && (m == 2 || ((m == 0 || m == 3 || m == 7) ? aT : (m == 1 || m == 8) ? aL : (aT && aL && aX)));
            if (ok) key = min(key, ((uint32_t)(sat + lambda * (m == pm ? 1 : 4)) << 4) | (uint32_t)m);
        }
        key = warp_min(key);
        const int wm = key & 15;
        total += (int)(key >> 4);
        rate += 2 * (wm == pm ? 1 : 4);
        // the winner through the 8x8 transform chain: lane group g8 = lane >> 3 works row r of the block on its own transpose tile
        int pr[8], sr[8], res[8], lv[8], rr[8];
        {
            const uint32_t i0 = i8idx[wm * 16 + r * 2], i1 = i8idx[wm * 16 + r * 2 + 1];
            const uint32_t s0 = sm.srcw[(by8 + r) * 4 + (bx8 >> 2)], s1 = sm.srcw[(by8 + r) * 4 + (bx8 >> 2) + 1];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                pr[k] = sm.F[((k < 4 ? i0 : i1) >> (8 * (k & 3))) & 255];
                sr[k] = (int)(((k < 4 ? s0 : s1) >> (8 * (k & 3))) & 255);
                res[k] = sr[k] - pr[k];
            }
        }
        t8_code_block(sm.t8[lane >> 3], r, qp, 3, res, lv, rr);
        uint32_t w0 = 0, w1 = 0; int sd = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int pix = clip255(pr[k] + rr[k]), d = sr[k] - pix;
            sd += d * d;
            if (k < 4) w0 |= (uint32_t)pix << (8 * k); else w1 |= (uint32_t)pix << (8 * (k - 4));
        }
        T8Cost a = { 0, 0, 0 };
#pragma unroll
        for (int k = 0; k < 8; k++) t8_cost_add(a, lv[k], c_izigzag8[k * 8 + r]);
        int n = a.nz;
        n += __shfl_xor_sync(0xffffffffu, n, 1); n += __shfl_xor_sync(0xffffffffu, n, 2); n += __shfl_xor_sync(0xffffffffu, n, 4);
        const int lcost = t8_cost_fold<8>(a);
        if (n) { rate += lcost; cbp_luma |= 1 << b; }
        if (lane < 8) {
            ssd += sd;
            sm.nb[((by8 + r + 1) * NBP + 4 + bx8) >> 2] = w0; sm.nb[(((by8 + r + 1) * NBP + 4 + bx8) >> 2) + 1] = w1;
#pragma unroll
            for (int k = 0; k < 8; k++) sm.keep_lv[b][c_izigzag8[k * 8 + r]] = (int16_t)lv[k];
        } else if (lane == 8) { sm.keep_mode[b] = (uint8_t)wm; sm.keep_nnz[b] = (uint8_t)n; }
        else if (lane < 13) {     // the four grid entries of this block
            const int k = lane - 9;
            sm.mg[(2 * (b >> 1) + 1 + (k >> 1)) * 5 + 2 * (b & 1) + 1 + (k & 1)] = (int8_t)wm;
        }
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int wi = lane + 32 * i, rw = wi >> 2, cw4 = wi & 3;
        sm.keep_rec[wi] = sm.nb[((rw + 1) * NBP + 4 + cw4 * 4) >> 2];
    }
    j8 = 64ll * __reduce_add_sync(0xffffffffu, ssd) + 27ll * lambda * lambda * rate;
    __syncwarp();
    return total;
}

// Trial coding of the luma of one MB as Intra_4x4: the 16 blocks in decoding order; lanes 0-8 evaluate the nine predictors of a
// block (SATD + lambda * mode bits, key = cost << 4 | mode), the winner is transformed, quantised and reconstructed at once
// (the next block predicts from it). Gives up as soon as the running cost reaches `limit` (the Intra_16x16 SATD).
// On success the levels, nnz and reconstruction are in place; returns true, the luma cbp and the 16 modes (4 bits each).
// full_trial (High profile, where Intra_8x8 competes): never gives up, and also returns the SATD cost and J = 64 SSD + 27 lambda^2 B of the
// finished coding (B in half bits: 2 x mode bits, and 1 + the level cost of every block of a coded 8x8 quadrant).
__device__ bool intra_try_i4x4(const IntraCtx &s, const Geom &g, IntraSmem &sm, int mx, int my, int lane, int limit,
                               bool top, bool left, int &cbp_luma, unsigned long long &modes, bool full_trial, int &cost4, long long &j4)
{
    const int wc = g.wc, mb = my * g.mbw + mx, qp = s.qp, lambda = c_lambda[qp];
    const bool topright = top && mx + 1 < g.mbw;
    MbInfo *mi = s.mbi + mb; MbCoef *co = s.coef + mb;
    uint8_t *nb = reinterpret_cast<uint8_t *>(sm.nb);
    int mode_rate = 0; unsigned long long lcq = 0ull;      // full trial: 2 x mode bits; per 8x8 quadrant the sum of (1 + level cost) in 16-bit fields
    const QParam q = make_qparam(qp);
    uint32_t ix[4];
#pragma unroll
    for (int y = 0; y < 4; y++) ix[y] = c_i4_idx[min(lane, 8)][y];
    // per-lane transform constants: the zig-zag position this lane quantises, its horizontal basis as signed bytes for IDP.4A
    // (rows of Cf = {1,1,1,1},{2,1,-1,-2},{1,-1,-1,1},{1,-2,2,-1}), its vertical basis, and its quantiser entries
    const int zpos = c_zigzag[lane & 15], fi = zpos >> 2, fj = zpos & 3;
    const uint32_t WX = fj == 0 ? 0x01010101u : fj == 1 ? 0xFEFF0102u : fj == 2 ? 0x01FFFF01u : 0xFF02FE01u;
    const int wy0 = fi == 1 ? 2 : 1, wy1 = fi == 0 ? 1 : fi == 1 ? 1 : fi == 2 ? -1 : -2, wy2 = fi == 0 ? 1 : fi == 1 ? -1 : fi == 2 ? -1 : 2, wy3 = fi == 0 ? 1 : fi == 1 ? -2 : fi == 2 ? 1 : -1;
    const int qcl = pos_class(zpos), qmf = qcl == 0 ? q.mf[0] : qcl == 1 ? q.mf[1] : q.mf[2], qv = qcl == 0 ? q.v[0] : qcl == 1 ? q.v[1] : q.v[2];
    // edge gather offsets of this lane relative to the block origin (lanes 0-4: L3 L3 L2 L1 L0, 5: corner, 6-14: T0..T7 T7;
    // without a top-right neighbour T4..T7 repeat T3)
    const int rel_tr = lane <= 4 ? (lane == 0 ? 4 : 5 - lane) * NBP + 3 : lane == 5 ? 3 : 4 + min(lane - 6, 7);
    const int rel_notr = lane <= 5 ? rel_tr : 4 + min(lane - 6, 3);
    int total = lambda * I4_BIAS_BITS;
    cbp_luma = 0; modes = 0ull;
#ifdef INTRA_TIMING
    long long it_last = clock64();
#endif
#pragma unroll 1
    for (int b = 0; b < 16; b++) {
        const int bxb = blk_x(b), byb = blk_y(b), bx = bxb * 4, by = byb * 4;
        const bool aT = byb > 0 || top, aL = bxb > 0 || left, aX = aT && aL;
        const bool aTR = byb == 0 ? (bxb < 3 ? top : topright) : !((0xA888u >> b) & 1u);
        // work that does not depend on the previous block's reconstruction first: source rows, their horizontal Hadamard
        // transforms, the predicted mode and this lane's cost terms
        uint32_t S[4]; int Ts[16];
        {
#pragma unroll
            for (int y = 0; y < 4; y++) S[y] = sm.srcw[(by + y) * 4 + bxb];
            satd_source_terms(S, Ts);
        }
        const int ma = sm.mg[(byb + 1) * 5 + bxb], mb_ = sm.mg[byb * 5 + bxb + 1];
        const int pm = (ma < 0 || mb_ < 0) ? 2 : min(ma, mb_);
        // (block 5 never predicts from the upper-right MACROBLOCK outside High-profile sessions: modes 3 and 7 are left out there, which is what lets the
        // wavefront run at a lag of one macroblock -- see k_intra_wave and oracle/orc_encoder.c: code_intra4x4_luma)
        const bool ok = lane < 9 && (lane == 2 || ((lane == 0 || lane == 3 || lane == 7) ? aT : (lane == 1 || lane == 8) ? aL : aX)) &&
                        !(b == 5 && topright && !s.i8 && (lane == 3 || lane == 7));
        const int mode_cost = lambda * (lane == pm ? 1 : 4);
        // filtered edge of this block
        int e = 128;
        if (lane < 15) e = nb[by * NBP + bx + (aTR ? rel_tr : rel_notr)];
        const int e1 = __shfl_down_sync(0xffffffffu, e, 1), e2 = __shfl_down_sync(0xffffffffu, e, 2);
        if (lane < 15) sm.F[lane] = (uint8_t)e;
        if (lane < 14) sm.F[16 + lane] = (uint8_t)((e + e1 + 1) >> 1);
        if (lane < 13) sm.F[32 + lane] = (uint8_t)((e + 2 * e1 + e2 + 2) >> 2);
        {
            const int sT = dp4a_us(sm.nb[(by * NBP + 4 + bx) >> 2], 0x01010101u, 0);
            const int sL = nb[(by + 1) * NBP + 3 + bx] + nb[(by + 2) * NBP + 3 + bx] + nb[(by + 3) * NBP + 3 + bx] + nb[(by + 4) * NBP + 3 + bx];
            const int dc = aT && aL ? (sT + sL + 4) >> 3 : aT ? (sT + 2) >> 2 : aL ? (sL + 2) >> 2 : 128;
            if (lane == 15) sm.F[47] = (uint8_t)dc;
        }
        __syncwarp();
        INTRA_T(4);
        uint32_t P[4];
#pragma unroll
        for (int y = 0; y < 4; y++)
            P[y] = (uint32_t)sm.F[ix[y] & 255] | ((uint32_t)sm.F[(ix[y] >> 8) & 255] << 8) | ((uint32_t)sm.F[(ix[y] >> 16) & 255] << 16) | ((uint32_t)sm.F[ix[y] >> 24] << 24);
        uint32_t key = 0xffffffffu;
        {
            const int sat = satd_rows(P, Ts);
            if (ok) key = ((uint32_t)(sat + mode_cost) << 4) | (uint32_t)lane;
        }
        key = warp_min(key);
        const int wm = key & 15;
        total += (int)(key >> 4);
        if (!full_trial && total >= limit) return false;
#pragma unroll
        for (int y = 0; y < 4; y++) P[y] = __shfl_sync(0xffffffffu, P[y], wm);
        INTRA_T(5);
        // transform, quantise, reconstruct, spread over 16 lanes (lanes 16-31 mirror them): lane l owns the coefficient at
        // zig-zag index l for the forward transform + quantiser (its value is one IDP.4A per row against the lane's horizontal
        // basis, then four multiply-adds with its vertical basis), and the sample (y,x) = (l>>2, l&3) for the inverse transform
        // (8.5.12.2, rows then columns, exchanged through shared memory) and the reconstruction.
        int level, dq;
        {
            const int t0 = dp4a_us(S[0], WX, 0) - dp4a_us(P[0], WX, 0), t1 = dp4a_us(S[1], WX, 0) - dp4a_us(P[1], WX, 0);
            const int t2 = dp4a_us(S[2], WX, 0) - dp4a_us(P[2], WX, 0), t3 = dp4a_us(S[3], WX, 0) - dp4a_us(P[3], WX, 0);
            const int cf = wy0 * t0 + wy1 * t1 + wy2 * t2 + wy3 * t3;
            const int l = min((int)(((unsigned)abs(cf) * (unsigned)qmf + (unsigned)q.f_intra) >> q.qbits), B200_MAX_LEVEL);
            level = cf < 0 ? -l : l;
            dq = (level * qv) << q.sh;
        }
        const uint32_t nzm = __ballot_sync(0xffffffffu, level != 0) & 0xffffu;
        if (full_trial) {      // rate terms of the J comparison with Intra_8x8 (oracle level_cost2)
            const int al = abs(level);
            const int lc = __reduce_add_sync(0xffffffffu, lane < 16 && al ? 2 * (3 + min(al, 16)) : 0) + (nzm ? 32 - __clz(nzm) : 0) - __popc(nzm);
            lcq += (unsigned long long)(1 + lc) << (16 * (b >> 2));
            mode_rate += 2 * (wm == pm ? 1 : 4);
        }
        if (lane < 16) { sm.dq[zpos] = dq; co->luma[b][lane] = (int16_t)level; }
        __syncwarp();
        const int py = (lane >> 2) & 3, pxl = lane & 3;
        {
            const int4 d = *reinterpret_cast<const int4 *>(&sm.dq[py * 4]);
            const int e0 = d.x + d.z, e1 = d.x - d.z, e2 = (d.y >> 1) - d.w, e3 = d.y + (d.w >> 1);
            const int f = pxl == 0 ? e0 + e3 : pxl == 1 ? e1 + e2 : pxl == 2 ? e1 - e2 : e0 - e3;
            if (lane < 16) sm.tf[lane] = f;
        }
        __syncwarp();
        {
            const int f0 = sm.tf[pxl], f1 = sm.tf[4 + pxl], f2 = sm.tf[8 + pxl], f3 = sm.tf[12 + pxl];
            const int g0 = f0 + f2, g1 = f0 - f2, g2 = (f1 >> 1) - f3, g3 = f1 + (f3 >> 1);
            const int r = ((py == 0 ? g0 + g3 : py == 1 ? g1 + g2 : py == 2 ? g1 - g2 : g0 - g3) + 32) >> 6;
            const uint32_t Pr = py == 0 ? P[0] : py == 1 ? P[1] : py == 2 ? P[2] : P[3];
            const int pix = clip255((int)((Pr >> (8 * pxl)) & 255u) + r);
            if (lane < 16) nb[(by + py + 1) * NBP + 4 + bx + pxl] = (uint8_t)pix;
            else if (lane == 16) { mi->nnz[b] = (uint8_t)__popc(nzm); sm.mg[(byb + 1) * 5 + bxb + 1] = (int8_t)wm; }
        }
        if (nzm) cbp_luma |= 1 << (b >> 2);
        modes |= (unsigned long long)wm << (4 * b);
        __syncwarp();
        INTRA_T(6);
    }
    // the reconstruction of the whole MB goes out at once: 16 rows x 4 words
    int ssd = 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int wi = lane + 32 * i, r = wi >> 2, cw4 = wi & 3;
        const uint32_t rw = sm.nb[((r + 1) * NBP + 4 + cw4 * 4) >> 2];
        *reinterpret_cast<uint32_t *>(s.rec0 + (size_t)(my * 16 + r) * wc + mx * 16 + cw4 * 4) = rw;
        const uint32_t d = __vabsdiffu4(rw, sm.srcw[wi]);
        ssd = (int)__dp4a(d, d, (unsigned)ssd);
    }
    cost4 = total;
    if (full_trial) {
        int rate = mode_rate;
#pragma unroll
        for (int q = 0; q < 4; q++) if (cbp_luma & (1 << q)) rate += (int)((lcq >> (16 * q)) & 0xffffu);
        j4 = 64ll * __reduce_add_sync(0xffffffffu, ssd) + 27ll * lambda * lambda * rate;
    }
    return true;
}

// predictor kinds shared by luma and chroma: 0 vertical, 1 horizontal, 2 DC, 3 plane (8.3.3 / 8.3.4)
__device__ __forceinline__ void intra_pred_block(int kind, const uint8_t *T, const uint8_t *L, int bx, int by, int dcv,
                                                 int pa, int pb, int pc, int off, int p[16])
{
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
        for (int x = 0; x < 4; x++) {
            int v;
            if (kind == 0) v = T[bx + x];
            else if (kind == 1) v = L[by + y];
            else if (kind == 2) v = dcv;
            else v = clip255((pa + pb * (bx + x - off) + pc * (by + y - off) + 16) >> 5);
            p[y * 4 + x] = v;
        }
}

__device__ void intra_code_mb(const IntraCtx &s, const Geom &g, IntraSmem &sm, const uint32_t *i8idx, int mx, int my, int lane)
{
    const int wc = g.wc, cw = wc / 2, mb = my * g.mbw + mx, qp = s.qp;
    const bool top = !row_is_slice_top(g, my), left = mx > 0;
#ifdef INTRA_TIMING
    long long it_last = clock64();
#endif
    // source block of this lane and the neighbours from the (pre-deblock) reconstruction (written by other warps / kernels ->
    // L2 loads): all global loads of the MB are issued back to back, then consumed
    const bool is_luma = lane < 16, active = lane < 24;
    const int comp = is_luma ? 0 : (lane < 20 ? 1 : 2);
    const int cb = lane & 3, b = lane & 15;
    const int bx = is_luma ? blk_x(b) * 4 : (cb & 1) * 4, by = is_luma ? blk_y(b) * 4 : (cb >> 1) * 4;
    const int n = is_luma ? 16 : 8, half = n / 2, st = is_luma ? wc : cw;
    uint32_t spw[4];
    {
        const uint8_t *spx = s.src(comp) + (size_t)(my * n + by) * st + mx * n + bx;
#pragma unroll
        for (int y = 0; y < 4; y++) spw[y] = *reinterpret_cast<const uint32_t *>(spx + (size_t)y * st);
    }
    __syncwarp();
    int nbv[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int i = lane + 32 * k;
        int ncomp, idx, is_top;
        if (i < 41) { ncomp = 0; is_top = i < 25; idx = is_top ? i : i - 25; }
        else { int j = i - 41; ncomp = 1 + j / 17; j %= 17; is_top = j < 9; idx = is_top ? j : j - 9; }
        const int nn = ncomp ? 8 : 16, nst = ncomp ? cw : wc, px0 = mx * nn, py0 = my * nn;
        const uint8_t *r = s.rec(ncomp);
        int v = 0;
        if (i < 25 + 16 + 2 * (9 + 8)) {
            if (is_top) { if (top && (idx > 0 || left) && (idx <= nn || mx + 1 < g.mbw)) v = __ldcg(r + (size_t)(py0 - 1) * nst + px0 + idx - 1); }
            else if (left) v = __ldcg(r + (size_t)(py0 + idx) * nst + px0 - 1);
        }
        nbv[k] = v;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int i = lane + 32 * k;
        int ncomp, idx, is_top;
        if (i < 41) { ncomp = 0; is_top = i < 25; idx = is_top ? i : i - 25; }
        else { int j = i - 41; ncomp = 1 + j / 17; j %= 17; is_top = j < 9; idx = is_top ? j : j - 9; }
        if (i < 25 + 16 + 2 * (9 + 8)) {
            if (is_top) { sm.top[ncomp][idx] = (uint8_t)nbv[k]; if (idx == 0) sm.left[ncomp][0] = (uint8_t)nbv[k]; }
            else sm.left[ncomp][idx + 1] = (uint8_t)nbv[k];
        }
    }
    __syncwarp();
    const uint8_t *T = sm.top[comp] + 1, *L = sm.left[comp] + 1;
    int sp[16];
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
        for (int x = 0; x < 4; x++) sp[y * 4 + x] = (spw[y] >> (8 * x)) & 255;
    // DC value and plane parameters of this lane's component
    int dcv, pa, pb, pc;
    {
        int sT = 0, sL = 0;
        const int t0 = is_luma ? 0 : bx, l0 = is_luma ? 0 : by, cnt = is_luma ? 16 : 4;
        for (int k = 0; k < cnt; k++) { sT += T[t0 + k]; sL += L[l0 + k]; }
        if (is_luma) dcv = top && left ? (sT + sL + 16) >> 5 : top ? (sT + 8) >> 4 : left ? (sL + 8) >> 4 : 128;
        else if (cb == 0 || cb == 3) dcv = top && left ? (sT + sL + 4) >> 3 : top ? (sT + 2) >> 2 : left ? (sL + 2) >> 2 : 128;
        else if (cb == 1) dcv = top ? (sT + 2) >> 2 : left ? (sL + 2) >> 2 : 128;
        else dcv = left ? (sL + 2) >> 2 : top ? (sT + 2) >> 2 : 128;
        int H = 0, V = 0;
        for (int k = 0; k < half; k++) { H += (k + 1) * (T[half + k] - T[half - 2 - k]); V += (k + 1) * (L[half + k] - L[half - 2 - k]); }
        const int coef = is_luma ? 5 : 34;
        pa = 16 * (L[n - 1] + T[n - 1]); pb = (coef * H + 32) >> 6; pc = (coef * V + 32) >> 6;
    }
    const int off = half - 1;
    INTRA_T(0);
    // mode decision: key = (SATD << 2) | mode id; luma ids V0 H1 DC2 P3, chroma ids DC0 H1 V2 P3
    uint32_t best = 0xffffffffu;
#pragma unroll 1
    for (int kind = 0; kind < 4; kind++) {
        const bool allowed = kind == 0 ? top : kind == 1 ? left : kind == 2 ? true : (top && left);
        int p[16]; intra_pred_block(kind, T, L, bx, by, dcv, pa, pb, pc, off, p);
#pragma unroll
        for (int k = 0; k < 16; k++) p[k] = sp[k] - p[k];
        int cost = active ? satd4x4(p) : 0;
#pragma unroll
        for (int o = 8; o; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
        const int id = is_luma ? kind : (kind == 0 ? 2 : kind == 2 ? 0 : kind);
        if (allowed) best = min(best, ((uint32_t)cost << 2) | (uint32_t)id);
    }
    const int mode_id = best & 3, kind = is_luma ? mode_id : (mode_id == 0 ? 2 : mode_id == 2 ? 0 : mode_id);
    const int chroma_mode = __shfl_sync(0xffffffffu, mode_id, 16);
    // I_NxN on trial against the Intra_16x16 SATD. High profile: Intra_8x8 is tried first and kept aside, Intra_4x4 is then coded
    // to the end on the same samples, the two meet by J = 64 SSD + 27 lambda^2 B, and the winner's SATD cost meets Intra_16x16's.
    int cbp_luma4 = 0; unsigned long long modes4 = 0ull;
    INTRA_T(1);
    const int cost16 = (int)(__shfl_sync(0xffffffffu, best, 0) >> 2);
    bool use_i4 = false, use_i8 = false; int cbp_luma8 = 0;
    if (!s.no_i4x4 || s.i8) intra_nxn_setup(s, g, sm, mx, my, lane, top, left);
    if (s.i8) {
        long long j8 = 0, j4 = 0; int cost4 = 1 << 30;
        const int cost8 = intra_try_i8x8(s, g, sm, i8idx, mx, my, lane, top, left, cbp_luma8, j8);
        if (!s.no_i4x4) intra_try_i4x4(s, g, sm, mx, my, lane, 0, top, left, cbp_luma4, modes4, true, cost4, j4);
        const bool pick8 = s.no_i4x4 || j8 < j4;
        const bool nxn = (pick8 ? cost8 : cost4) < cost16;
        use_i8 = nxn && pick8; use_i4 = nxn && !pick8;
        if (use_i8) {      // put the kept outcome in place: levels (512 bytes, same linear layout as MbCoef::luma), reconstruction, nnz
            MbCoef *co8 = s.coef + mb; MbInfo *mi8 = s.mbi + mb;
            reinterpret_cast<uint4 *>(co8->luma)[lane] = reinterpret_cast<const uint4 *>(sm.keep_lv)[lane];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int wi = lane + 32 * i, r = wi >> 2, cw4 = wi & 3;
                *reinterpret_cast<uint32_t *>(s.rec0 + (size_t)(my * 16 + r) * wc + mx * 16 + cw4 * 4) = sm.keep_rec[wi];
            }
            if (lane < 16) mi8->nnz[lane] = sm.keep_nnz[lane >> 2];
        }
    } else if (!s.no_i4x4) {
        long long j4 = 0; int cost4 = 0;
        use_i4 = intra_try_i4x4(s, g, sm, mx, my, lane, cost16, top, left, cbp_luma4, modes4, false, cost4, j4);
    }
    const bool use_nxn = use_i4 || use_i8;

    int p[16], c[16]; intra_pred_block(kind, T, L, bx, by, dcv, pa, pb, pc, off, p);
#pragma unroll
    for (int k = 0; k < 16; k++) c[k] = sp[k] - p[k];
    INTRA_T(2);
    fdct4x4(c);
    MbCoef *co = s.coef + mb; MbInfo *mi = s.mbi + mb;
    int nnz = 0; bool dc_nz = false;
    __align__(16) int16_t lz[16];
    if (is_luma) {
        const QParam q = make_qparam(qp);
        int m[16];
#pragma unroll
        for (int k = 0; k < 16; k++) m[k] = __shfl_sync(0x0000ffffu, c[0], xy2blk(k & 3, k >> 2));
        hadamard16(m);
        const int ps = blk_y(b) * 4 + blk_x(b);
        int hv = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) if (k == ps) hv = m[k];
        const int lev = quant_dc((hv + 1) >> 1, q, q.f_intra);
#pragma unroll
        for (int k = 0; k < 16; k++) m[k] = __shfl_sync(0x0000ffffu, lev, xy2blk(k & 3, k >> 2));
        hadamard16(m);
        int f = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) if (k == ps) f = m[k];
        const int LS = 16 * q.v[0];
        const int dcY = qp >= 36 ? (f * LS) << (q.sh - 6) : (f * LS + (1 << (5 - q.sh))) >> (6 - q.sh);
        nnz = quant_dequant4x4(c, lz, q, q.f_intra, true);
        c[0] = dcY;
        const uint8_t inv_zz[16] = { 0, 1, 5, 6, 2, 4, 7, 12, 3, 8, 11, 13, 9, 10, 14, 15 };
        co->luma_dc[inv_zz[ps]] = use_nxn ? (int16_t)0 : (int16_t)lev;
    } else if (active) {
        const QParam qc = make_qparam(c_chroma_qp[qp]);
        const int pl = comp - 1;
        int dcs[4], lv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) dcs[k] = __shfl_sync(0x00ff0000u, c[0], 16 + pl * 4 + k);
        const int hd[4] = { dcs[0] + dcs[1] + dcs[2] + dcs[3], dcs[0] - dcs[1] + dcs[2] - dcs[3], dcs[0] + dcs[1] - dcs[2] - dcs[3], dcs[0] - dcs[1] - dcs[2] + dcs[3] };
#pragma unroll
        for (int k = 0; k < 4; k++) { lv[k] = quant_dc(hd[k], qc, qc.f_intra); dc_nz |= lv[k] != 0; }
        const int fi[4] = { lv[0] + lv[1] + lv[2] + lv[3], lv[0] - lv[1] + lv[2] - lv[3], lv[0] + lv[1] - lv[2] - lv[3], lv[0] - lv[1] - lv[2] + lv[3] };
        nnz = quant_dequant4x4(c, lz, qc, qc.f_intra, true);
        c[0] = ((fi[cb] * 16 * qc.v[0]) << qc.sh) >> 5;
        if (cb == 0) *reinterpret_cast<uint2 *>(co->chroma_dc[pl]) = make_uint2((uint32_t)(uint16_t)lv[0] | ((uint32_t)(uint16_t)lv[1] << 16),
                                                                              (uint32_t)(uint16_t)lv[2] | ((uint32_t)(uint16_t)lv[3] << 16));
    }
    if (active && !(is_luma && use_nxn)) {
        idct4x4(c);
        uint4 *dst = reinterpret_cast<uint4 *>(is_luma ? co->luma[b] : co->chroma_ac[comp - 1][cb]);
        dst[0] = reinterpret_cast<uint4 *>(lz)[0]; dst[1] = reinterpret_cast<uint4 *>(lz)[1];
        uint8_t *rp = s.rec(comp) + (size_t)(my * n + by) * st + mx * n + bx;
#pragma unroll
        for (int y = 0; y < 4; y++) {
            uint32_t w = 0;
#pragma unroll
            for (int x = 0; x < 4; x++) w |= (uint32_t)clip255(p[y * 4 + x] + c[y * 4 + x]) << (8 * x);
            *reinterpret_cast<uint32_t *>(rp + (size_t)y * st) = w;
        }
        mi->nnz[lane] = (uint8_t)nnz;
    }
    INTRA_T(3);
#ifdef INTRA_TIMING
    if (lane == 0) atomicAdd((unsigned long long *)&g_intra_t[7], 1ull);
#endif
    const uint32_t nzmask = __ballot_sync(0xffffffffu, nnz != 0), dcmask = __ballot_sync(0xffffffffu, dc_nz);
    if (lane == 0) {
        int cbp = use_i8 ? cbp_luma8 : use_i4 ? cbp_luma4 : ((nzmask & 0xffff) ? 15 : 0);
        cbp |= ((nzmask >> 16) & 255) ? 32 : (dcmask ? 16 : 0);
        // Intra_8x8: I_NxN with transform_size_8x8_flag (bit 2 of i16_mode)
        reinterpret_cast<uint32_t *>(mi)[0] = (use_i8 ? (uint32_t)MB_I8x8 | (4u << 8) : use_i4 ? (uint32_t)MB_I4x4 : (uint32_t)MB_I16x16 | ((uint32_t)mode_id << 8)) |
                                              ((uint32_t)chroma_mode << 16) | ((uint32_t)cbp << 24);
    } else if (lane == 1) reinterpret_cast<uint32_t *>(mi)[1] = 0;
    else if (lane < 6) {     // i4_mode[16]: one byte per 4x4 block (an Intra_8x8 block's mode in all four of its entries)
        const uint32_t nib = use_i4 ? (uint32_t)(modes4 >> (16 * (lane - 2))) & 0xffffu : 0u;
        reinterpret_cast<uint32_t *>(mi)[lane] = use_i8 ? 0x01010101u * sm.keep_mode[lane - 2]
                                                        : (nib & 15u) | ((nib & 0xf0u) << 4) | ((nib & 0xf00u) << 8) | ((nib & 0xf000u) << 12);
    }
}

// grid: ceil(sessions * mbh / WAVE_WARPS) CTAs of WAVE_WARPS warps
__global__ void __launch_bounds__(WAVE_WARPS * 32) k_intra_wave(const Sess *ss, Geom g, int nsess, WaveCtl *ctl)
{
    __shared__ IntraSmem sm_all[WAVE_WARPS];
    __shared__ uint32_t i8idx[9 * 16];          // the Intra_8x8 predictor index table: lanes read different entries at once (constant memory would serialise)
    for (int i = threadIdx.x; i < 9 * 16; i += blockDim.x) i8idx[i] = c_i8_idx[i / 16][i % 16];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = 0;
    if (lane == 0) t = atomicAdd(&ctl->ticket_intra, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= nsess * g.mbh) return;
    const int my = t / nsess;
    const Sess &sg = ss[t % nsess];
    IntraSmem &sm = sm_all[warp];
    int *prog = sg.row_prog_intra;
    IntraCtx s; s.rec0 = sg.rec[0]; s.rec1 = sg.rec[1]; s.rec2 = sg.rec[2]; s.src0 = sg.src[0]; s.src1 = sg.src[1]; s.src2 = sg.src[2];
    s.mbi = sg.mbi; s.coef = sg.coef; s.qp = sg.qp; s.is_idr = sg.is_idr; s.no_i4x4 = sg.no_i4x4; s.i8 = sg.t8x8;
    const bool slice_top = row_is_slice_top(g, my);
    int mx = 0;
    while (mx < g.mbw) {
        int nx = mx;
        if (!s.is_idr) {       // P picture: only the MBs phase A marked intra are coded here
            nx = g.mbw;
            for (int base = mx; base < g.mbw; base += 32) {
                const int x = base + lane;
                const int ty = x < g.mbw ? s.mbi[my * g.mbw + x].mb_type : 0;
                const uint32_t m = __ballot_sync(0xffffffffu, ty == MB_I16x16);
                if (m) { nx = base + __ffs(m) - 1; break; }
            }
            if (nx >= g.mbw) break;
            if (nx > mx && lane == 0) { __threadfence(); st_release(prog + my, nx); }
        }
        // the row above must be past the upper neighbour -- and past the upper-right one only where a predictor reads it (Intra_8x8: High profile)
        if (!slice_top && !wave_wait(prog + my - 1, min(nx + (s.i8 ? 2 : 1), g.mbw), ctl, lane)) return;
        intra_code_mb(s, g, sm, i8idx, nx, my, lane);
        fence_acq_rel_gpu();
        __syncwarp();
        if (lane == 0) st_relaxed(prog + my, nx + 1);
        mx = nx + 1;
    }
    __syncwarp();
    if (lane == 0) { __threadfence(); st_release(prog + my, g.mbw); }
}

} // namespace b200
