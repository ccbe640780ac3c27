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

// wait until the row above has finished `need` macroblocks; returns false on timeout / global error
__device__ __forceinline__ bool wave_wait(const int *prog_above, int need, WaveCtl *ctl, int lane)
{
    int ok = 1;
    if (lane == 0) {
        int cur = ld_acquire(prog_above);
        if (cur < need) {
            const unsigned long long t0 = global_ns();
            int spins = 0;
            do {
                // a row that is d macroblocks short of what we need takes d MB-steps (a few us each): sleep accordingly, so
                // that far-behind rows do not burn the issue slots of the SMs they share with other kernels
                const int d = need - cur;
                __nanosleep(d > 1 ? min(d * 1500, 30000) : 32);
                if ((++spins & 15) == 0 && (ld_acquire(&ctl->error) || global_ns() - t0 > WAVE_TIMEOUT_NS)) { atomicExch(&ctl->error, 1); ok = 0; break; }
            } while ((cur = ld_acquire(prog_above)) < need);
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    return ok != 0;
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

// top/left: index 0 = the corner sample p[-1,-1]; luma top also carries the 4 samples of the MB above-right (17..20).
// nb: luma reconstruction of the MB being coded as Intra_4x4 with its border, 17 rows x 24 bytes: sample (x,y) at
// (y+1)*24 + 4 + x, x in [-1,19] on row y = -1, x in [-1,15] below. F: the filtered edge of one 4x4 block (see c_i4_idx).
struct IntraSmem {
    uint8_t top[3][24], left[3][20];
    uint32_t nb[17 * 6];
    uint32_t srcw[64];
    uint8_t F[48];
    int8_t mg[28];
    __align__(16) int dq[16];  // dequantised coefficients of the Intra_4x4 block in flight (raster), then its row-transformed values
    int tf[16];            // Intra4x4PredMode grid, 5x5: row 0 / column 0 = neighbouring MBs (-1 unavailable, 2 not Intra_4x4)
};

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
    uint8_t *rec0, *rec1, *rec2; const uint8_t *src0, *src1, *src2; MbInfo *mbi; MbCoef *coef; int qp, is_idr, no_i4x4;
    __device__ __forceinline__ uint8_t *rec(int c) const { return c == 0 ? rec0 : c == 1 ? rec1 : rec2; }
    __device__ __forceinline__ const uint8_t *src(int c) const { return c == 0 ? src0 : c == 1 ? src1 : src2; }
};

// Trial coding of the luma of one MB as Intra_4x4: the 16 blocks in decoding order; lanes 0-8 evaluate the nine predictors of a
// block (SATD + lambda * mode bits, key = cost << 4 | mode), the winner is transformed, quantised and reconstructed at once
// (the next block predicts from it). Gives up as soon as the running cost reaches `limit` (the Intra_16x16 SATD).
// On success the levels, nnz and reconstruction are in place; returns true, the luma cbp and the 16 modes (4 bits each).
__device__ bool intra_try_i4x4(const IntraCtx &s, const Geom &g, IntraSmem &sm, int mx, int my, int lane, int limit,
                               bool top, bool left, int &cbp_luma, unsigned long long &modes)
{
    const int wc = g.wc, mb = my * g.mbw + mx, qp = s.qp, lambda = c_lambda[qp];
    const bool topright = top && mx + 1 < g.mbw;
    MbInfo *mi = s.mbi + mb; MbCoef *co = s.coef + mb;
    uint8_t *nb = reinterpret_cast<uint8_t *>(sm.nb);
    if (lane < 21) nb[3 + lane] = sm.top[0][lane];
    if (lane < 16) nb[(lane + 1) * 24 + 3] = sm.left[0][lane + 1];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int wi = lane + 32 * i;
        sm.srcw[wi] = *reinterpret_cast<const uint32_t *>(s.src0 + (size_t)(my * 16 + (wi >> 2)) * wc + mx * 16 + (wi & 3) * 4);
    }
    if (lane < 25) {
        const int gy = lane / 5, gx = lane - gy * 5;
        int v = 2;
        if (gy == 0 && gx > 0) { v = -1; if (top) { const MbInfo *mt = mi - g.mbw; v = __ldcg(&mt->mb_type) == MB_I4x4 ? (int)__ldcg(&mt->i4_mode[xy2blk(gx - 1, 3)]) : 2; } }
        else if (gx == 0 && gy > 0) { v = -1; if (left) { const MbInfo *ml = mi - 1; v = __ldcg(&ml->mb_type) == MB_I4x4 ? (int)__ldcg(&ml->i4_mode[xy2blk(3, gy - 1)]) : 2; } }
        sm.mg[lane] = (int8_t)v;
    }
    __syncwarp();
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
    const int rel_tr = lane <= 4 ? (lane == 0 ? 4 : 5 - lane) * 24 + 3 : lane == 5 ? 3 : 4 + min(lane - 6, 7);
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
            const uint32_t H[4] = { 0x01010101u, 0xffff0101u, 0x01ffff01u, 0xff01ff01u };
#pragma unroll
            for (int y = 0; y < 4; y++) {
                S[y] = sm.srcw[(by + y) * 4 + bxb];
#pragma unroll
                for (int k = 0; k < 4; k++) Ts[y * 4 + k] = dp4a_us(S[y], H[k], 0);
            }
        }
        const int ma = sm.mg[(byb + 1) * 5 + bxb], mb_ = sm.mg[byb * 5 + bxb + 1];
        const int pm = (ma < 0 || mb_ < 0) ? 2 : min(ma, mb_);
        const bool ok = lane < 9 && (lane == 2 || ((lane == 0 || lane == 3 || lane == 7) ? aT : (lane == 1 || lane == 8) ? aL : aX));
        const int mode_cost = lambda * (lane == pm ? 1 : 4);
        // filtered edge of this block
        int e = 128;
        if (lane < 15) e = nb[by * 24 + bx + (aTR ? rel_tr : rel_notr)];
        const int e1 = __shfl_down_sync(0xffffffffu, e, 1), e2 = __shfl_down_sync(0xffffffffu, e, 2);
        if (lane < 15) sm.F[lane] = (uint8_t)e;
        if (lane < 14) sm.F[16 + lane] = (uint8_t)((e + e1 + 1) >> 1);
        if (lane < 13) sm.F[32 + lane] = (uint8_t)((e + 2 * e1 + e2 + 2) >> 2);
        {
            const int sT = dp4a_us(sm.nb[(by * 24 + 4 + bx) >> 2], 0x01010101u, 0);
            const int sL = nb[(by + 1) * 24 + 3 + bx] + nb[(by + 2) * 24 + 3 + bx] + nb[(by + 3) * 24 + 3 + bx] + nb[(by + 4) * 24 + 3 + bx];
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
        if (total >= limit) return false;
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
            if (lane < 16) nb[(by + py + 1) * 24 + 4 + bx + pxl] = (uint8_t)pix;
            else if (lane == 16) { mi->nnz[b] = (uint8_t)__popc(nzm); sm.mg[(byb + 1) * 5 + bxb + 1] = (int8_t)wm; }
        }
        if (nzm) cbp_luma |= 1 << (b >> 2);
        modes |= (unsigned long long)wm << (4 * b);
        __syncwarp();
        INTRA_T(6);
    }
    // the reconstruction of the whole MB goes out at once: 16 rows x 4 words
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int wi = lane + 32 * i, r = wi >> 2, cw4 = wi & 3;
        *reinterpret_cast<uint32_t *>(s.rec0 + (size_t)(my * 16 + r) * wc + mx * 16 + cw4 * 4) = sm.nb[((r + 1) * 24 + 4 + cw4 * 4) >> 2];
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

__device__ void intra_code_mb(const IntraCtx &s, const Geom &g, IntraSmem &sm, int mx, int my, int lane)
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
        if (i < 37) { ncomp = 0; is_top = i < 21; idx = is_top ? i : i - 21; }
        else { int j = i - 37; ncomp = 1 + j / 17; j %= 17; is_top = j < 9; idx = is_top ? j : j - 9; }
        const int nn = ncomp ? 8 : 16, nst = ncomp ? cw : wc, px0 = mx * nn, py0 = my * nn;
        const uint8_t *r = s.rec(ncomp);
        int v = 0;
        if (i < 21 + 16 + 2 * (9 + 8)) {
            if (is_top) { if (top && (idx > 0 || left) && (idx <= nn || mx + 1 < g.mbw)) v = __ldcg(r + (size_t)(py0 - 1) * nst + px0 + idx - 1); }
            else if (left) v = __ldcg(r + (size_t)(py0 + idx) * nst + px0 - 1);
        }
        nbv[k] = v;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int i = lane + 32 * k;
        int ncomp, idx, is_top;
        if (i < 37) { ncomp = 0; is_top = i < 21; idx = is_top ? i : i - 21; }
        else { int j = i - 37; ncomp = 1 + j / 17; j %= 17; is_top = j < 9; idx = is_top ? j : j - 9; }
        if (i < 21 + 16 + 2 * (9 + 8)) {
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
    // Intra_4x4 on trial against the Intra_16x16 SATD
    int cbp_luma4 = 0; unsigned long long modes4 = 0ull;
    INTRA_T(1);
    const bool use_i4 = !s.no_i4x4 && intra_try_i4x4(s, g, sm, mx, my, lane, (int)(__shfl_sync(0xffffffffu, best, 0) >> 2), top, left, cbp_luma4, modes4);

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
        co->luma_dc[inv_zz[ps]] = use_i4 ? (int16_t)0 : (int16_t)lev;
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
    if (active && !(is_luma && use_i4)) {
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
        int cbp = use_i4 ? cbp_luma4 : ((nzmask & 0xffff) ? 15 : 0);
        cbp |= ((nzmask >> 16) & 255) ? 32 : (dcmask ? 16 : 0);
        reinterpret_cast<uint32_t *>(mi)[0] = (use_i4 ? (uint32_t)MB_I4x4 : (uint32_t)MB_I16x16 | ((uint32_t)mode_id << 8)) | ((uint32_t)chroma_mode << 16) | ((uint32_t)cbp << 24);
    } else if (lane == 1) reinterpret_cast<uint32_t *>(mi)[1] = 0;
    else if (lane < 6) {     // i4_mode[16]: one byte per block
        const uint32_t nib = use_i4 ? (uint32_t)(modes4 >> (16 * (lane - 2))) & 0xffffu : 0u;
        reinterpret_cast<uint32_t *>(mi)[lane] = (nib & 15u) | ((nib & 0xf0u) << 4) | ((nib & 0xf00u) << 8) | ((nib & 0xf000u) << 12);
    }
}

// grid: ceil(sessions * mbh / WAVE_WARPS) CTAs of WAVE_WARPS warps
__global__ void __launch_bounds__(WAVE_WARPS * 32) k_intra_wave(const Sess *ss, Geom g, int nsess, WaveCtl *ctl)
{
    __shared__ IntraSmem sm_all[WAVE_WARPS];
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
    s.mbi = sg.mbi; s.coef = sg.coef; s.qp = sg.qp; s.is_idr = sg.is_idr; s.no_i4x4 = sg.no_i4x4;
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
        if (!slice_top && !wave_wait(prog + my - 1, min(nx + 2, g.mbw), ctl, lane)) return;
        intra_code_mb(s, g, sm, nx, my, lane);
        fence_acq_rel_gpu();
        __syncwarp();
        if (lane == 0) st_relaxed(prog + my, nx + 1);
        mx = nx + 1;
    }
    __syncwarp();
    if (lane == 0) { __threadfence(); st_release(prog + my, g.mbw); }
}

} // namespace b200
