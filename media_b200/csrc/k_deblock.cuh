// media_b200/csrc/k_deblock.cuh -- in-loop deblocking filter (H.264 clause 8.7), phase E of DESIGN.md 3.
//
// Role inside the reference: the loop filter inside ISVCEncoder::EncodeFrame with iLoopFilterDisableIdc = 0
// (video_codec/VideoEncoderOpenH264.cpp:295,344; openh264's DeblockingBSCalcEnc_c, DeblockLumaLt4/Eq4,
// DeblockChromaLt4/Eq4 in the absent libopenh264). The standard filters macroblocks in raster order and each
// MB reads samples its left, upper and upper-right neighbours have already filtered, so the kernel runs the
// same 2-MB-lag wavefront as the intra kernel: one warp per MB row, tile staged in shared memory, vertical
// edges by 16 row-lanes (+16 chroma row-lanes), then horizontal edges by column-lanes.
#pragma once
#include "h264_dev.cuh"
#include "k_intra.cuh"
#include <cstdio>

namespace b200 {

#ifndef DBK_MIN_CTAS
#define DBK_MIN_CTAS 4      /* resident CTAs per SM the register allocation must allow (128 registers: the column of 20 samples of the horizontal pass lives in registers) */
#endif
#ifdef DBK_TIMING
#define DBK_T(i) do { const long long t_ = clock64(); dbk_t[i] += t_ - dbk_last; dbk_last = t_; } while (0)
#else
#define DBK_T(i) do { } while (0)
#endif

// the few session fields the row loop needs, held in registers: every fence / strong access in the loop is a compiler memory
// barrier, so reading them through `const Sess &` re-fetched them from L2 several times per macroblock (1 800 cycles measured)
struct DbkCtx { uint8_t *rec[3]; const uint4 *bs; int qp; uint32_t seq; };
// per-lane filter constants of a row (one QP per picture): lanes 0-15 the luma values, lanes 16-31 the chroma ones; tc0 for bS 1, 2, 3
struct DbkK { int alpha, beta, tc0[3]; };

// Row-to-row hand-over ("flag in the data", the scheme of NCCL's LL protocol). The only samples a macroblock row takes from the row above are
// its 4 bottom luma rows and 2 bottom chroma rows (the p side of the horizontal MB edge). The upper row publishes them per macroblock as 24
// 8-byte messages {4 samples, sequence number of this launch}: an aligned 64-bit store is single-copy atomic, so a reader that sees the
// sequence number has the samples -- no fence on the writer's side, no acquire + second load on the reader's, and the messages can be
// fetched one macroblock ahead like the rest of the prefetch. Message chunk k of a row = columns 16k-4 .. 16k+11 (chroma 8k-4 .. 8k+3):
// exactly what is final once macroblock k has been filtered (its left edge changed up to 3 columns of k-1, its right 3 are still open).
// chunk mbw carries the last 4 columns of the row. Layout: [row][chunk 0..mbw][24].
#define DBK_LL_PER_MB 24
__device__ __forceinline__ uint2 ld_ll(const uint2 *p)
{
    unsigned long long v; asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return make_uint2((uint32_t)v, (uint32_t)(v >> 32));
}
__device__ __forceinline__ void st_ll(uint2 *p, uint32_t data, uint32_t seq)
{
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" :: "l"(p), "l"((unsigned long long)data | ((unsigned long long)seq << 32)) : "memory");
}

// write-back slots of a lane: word lane + 32k of the luma tile (20 rows x 5 words) / of the two chroma tiles (2 x 12 rows x 3 words);
// flags: 1 slot exists and is ever stored, 2 it lies in the rows above the MB, 4 it is not in the 4 left columns, 8 (chroma) plane,
// 16 it lies in the bottom rows the MB row below may still filter (luma 13-15, chroma 6-7).
// Every sample has ONE writer: the rows above an MB are stored by that MB iff its upper edge is filtered at all (any bS != 0 there) -- and
// then the MB above leaves its bottom rows alone; otherwise the MB above stores them itself. No two warps ever store the same word, so the
// global stores need no ordering between rows.
struct DbkWb { int oy[4], oc[3]; int fy[4], fc[3]; };
__device__ __forceinline__ void dbk_wb_init(DbkWb &wb, int wc, int lane)
{
    const int cw = wc / 2;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = lane + 32 * k, r = i / 5 - 4, c4 = (i % 5) * 4 - 4;
        wb.oy[k] = r * wc + c4;
        wb.fy[k] = ((i < 100 && r >= -3) ? 1 : 0) | (r < 0 ? 2 : 0) | (c4 >= 0 ? 4 : 0) | (r >= 13 ? 16 : 0);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int i = lane + 32 * k, pl = i / 36, j = i - pl * 36, r = j / 3 - 4, c4 = (j % 3) * 4 - 4;
        wb.oc[k] = r * cw + c4;
        wb.fc[k] = ((i < 72 && r >= -2) ? 1 : 0) | (r < 0 ? 2 : 0) | (c4 >= 0 ? 4 : 0) | (pl ? 8 : 0) | (r >= 6 ? 16 : 0);
    }
}

struct DbkSmem {
    uint32_t y[20 * 5];        // rows -4..15, cols -4..15 (stride 20 bytes)
    uint32_t c[2][12 * 3];     // rows -4..7, cols -4..7 (stride 12 bytes)
};

// One line of samples across one edge, in registers and branch-free (8.7.2.3 / 8.7.2.4): p3 p2 p1 p0 | q0 q1 q2 q3. The 16 luma lines and the
// 16 chroma lines of a macroblock edge are filtered by the 32 lanes in one instruction stream. Chroma uses only p1..q1, tc = tc0 + 1 and the
// weak bS = 4 filter; the selects fold that in. ANY_STRONG: some lane of the warp has bS = 4 (warp-uniform, so the strong filter's
// arithmetic is only issued on macroblock edges of intra macroblocks).
template <bool ANY_STRONG>
__device__ __forceinline__ void filter_line(int p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int q3, int bs, int alpha, int beta, int tc0, bool chroma)
{
    const int d = abs(p0 - q0);
    const bool on = bs != 0 && d < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta;
    const bool ap = !chroma && abs(p2 - p0) < beta, aq = !chroma && abs(q2 - q0) < beta;
    const int tc = chroma ? tc0 + 1 : tc0 + (int)ap + (int)aq;
    const int delta = clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
    const int avg = (p0 + q0 + 1) >> 1;
    int n0 = clip255(p0 + delta), m0 = clip255(q0 - delta);
    int n1 = ap ? p1 + clip3(-tc0, tc0, (p2 + avg - (p1 << 1)) >> 1) : p1;
    int m1 = aq ? q1 + clip3(-tc0, tc0, (q2 + avg - (q1 << 1)) >> 1) : q1;
    int n2 = p2, m2 = q2;
    if (ANY_STRONG) {
        const bool strong = bs == 4, small = d < ((alpha >> 2) + 2), sp = ap && small, sq = aq && small;
        const int s0 = sp ? (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3 : (2 * p1 + p0 + q1 + 2) >> 2;
        const int s1 = sp ? (p2 + p1 + p0 + q0 + 2) >> 2 : p1;
        const int s2 = sp ? (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3 : p2;
        const int t0 = sq ? (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3 : (2 * q1 + q0 + p1 + 2) >> 2;
        const int t1 = sq ? (p0 + q0 + q1 + q2 + 2) >> 2 : q1;
        const int t2 = sq ? (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3 : q2;
        if (strong) { n0 = s0; n1 = s1; n2 = s2; m0 = t0; m1 = t1; m2 = t2; }
    }
    if (on) { p0 = n0; p1 = n1; p2 = n2; q0 = m0; q1 = m1; q2 = m2; }
}
// the filter across the boundary between two words of a row: wl = .. p3 p2 p1 p0 (low to high byte), wr = q0 q1 q2 q3
template <bool ANY_STRONG>
__device__ __forceinline__ void filter_words(uint32_t &wl, uint32_t &wr, int bs, int alpha, int beta, int tc0, bool chroma)
{
    int p3 = wl & 255, p2 = (wl >> 8) & 255, p1 = (wl >> 16) & 255, p0 = wl >> 24;
    int q0 = wr & 255, q1 = (wr >> 8) & 255, q2 = (wr >> 16) & 255, q3 = wr >> 24;
    filter_line<ANY_STRONG>(p3, p2, p1, p0, q0, q1, q2, q3, bs, alpha, beta, tc0, chroma);
    wl = (uint32_t)p3 | ((uint32_t)p2 << 8) | ((uint32_t)p1 << 16) | ((uint32_t)p0 << 24);
    wr = (uint32_t)q0 | ((uint32_t)q1 << 8) | ((uint32_t)q2 << 16) | ((uint32_t)q3 << 24);
}

// boundary strength between 4x4 block (bxp,byp) of MB p and block (bxq,byq) of MB q (8.7.2.1, frame pictures, one reference)
__device__ __forceinline__ int bs_of(const MbInfo *p, int bxp, int byp, const MbInfo *q, int bxq, int byq, bool mb_edge)
{
    const bool ip = p->mb_type == MB_I16x16 || p->mb_type == MB_I4x4 || p->mb_type == MB_I8x8, iq = q->mb_type == MB_I16x16 || q->mb_type == MB_I4x4 || q->mb_type == MB_I8x8;
    if (ip || iq) return mb_edge ? 4 : 3;
    if (p->nnz[xy2blk(bxp, byp)] || q->nnz[xy2blk(bxq, byq)]) return 2;
    const int16_t *vp = p->mv8[(byp >> 1) * 2 + (bxp >> 1)], *vq = q->mv8[(byq >> 1) * 2 + (bxq >> 1)];
    if (abs(vp[0] - vq[0]) >= 4 || abs(vp[1] - vq[1]) >= 4) return 1;
    return 0;
}

// Boundary strengths of every macroblock, ahead of the wavefront: they only depend on the MbInfo records (types, nnz, vectors), not on
// samples, so they are computed fully in parallel and stored as three bit planes of the 32 values (index 4e + k = vertical edge e, segment k;
// 16 + 4e + k horizontal; x, y, z = bits 0, 1, 2 of bS; w = their OR). In the row loop of k_deblock_wave this was 48 % of the executed
// instructions (ncu, profiles/r01_ncu_summary.md); as a warp per MB with a lane per value it still was 331 warp-instructions per MB
// (profiles/r02a_ncu_summary.md). Here a THREAD owns a macroblock and works on whole bit masks: the nonzero flags of the 16 luma blocks as a
// raster 4x4 mask (an OR with its shifts gives every internal edge at once), vector differences only where partitions meet (edge 2 and the
// MB edges), the intra / 8x8-transform rules as mask selects.
// grid: (ceil(n_mb / 128), 1, sessions), 128 threads
__device__ __forceinline__ uint32_t nz_bits4(uint32_t w)     // bit i = byte i of w is nonzero
{
    const uint32_t m = ((w & 0x7f7f7f7fu) + 0x7f7f7f7fu | w) & 0x80808080u;
    return ((m >> 7) * 0x00204081u >> 21) & 15u;
}
__device__ __forceinline__ uint32_t blk_to_raster(uint32_t m)   // 16 flags in blkIdx order (Figure 6-10) -> raster order y * 4 + x: index bits 1 and 2 swap
{
    return (m & 0xC3C3u) | ((m & 0x0C0Cu) << 2) | ((m & 0x3030u) >> 2);
}
__device__ __forceinline__ bool mv_far(uint32_t a, uint32_t b)  // packed (x, y) int16 vectors at least 4 quarter-samples apart in a component
{
    const int dx = (int)(short)(a & 0xffffu) - (int)(short)(b & 0xffffu), dy = ((int)a >> 16) - ((int)b >> 16);
    return abs(dx) >= 4 || abs(dy) >= 4;
}
__device__ __forceinline__ bool type_is_intra(uint32_t w0) { const uint32_t t = w0 & 255u; return t == MB_I16x16 || t == MB_I4x4 || t == MB_I8x8; }
__global__ void __launch_bounds__(128) k_deblock_bs(const Sess *ss, Geom g)
{
    const int mb = blockIdx.x * 128 + threadIdx.x;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    int mx, my; mb_xy(g, mb, mx, my);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(s.mbi + mb);
    const uint4 a = *reinterpret_cast<const uint4 *>(q), b = *reinterpret_cast<const uint4 *>(q + 4), c = *reinterpret_cast<const uint4 *>(q + 8);
    // a.x type word, a.z a.w b.x b.y the four partition vectors, b.z b.w c.x c.y the 16 luma nnz
    const bool intra = type_is_intra(a.x), t8 = (a.x >> 10) & 1u;
    const uint32_t nzr = blk_to_raster(nz_bits4(b.z) | (nz_bits4(b.w) << 4) | (nz_bits4(c.x) << 8) | (nz_bits4(c.y) << 12));
    uint32_t p1 = 0, p2 = 0, p4 = 0;       // planes: value 1 (vector), value 2 (coefficients) or 3 (intra, both), value 4
    if (intra) { p1 = p2 = 0xfff0fff0u; }
    else {
        // internal edges: a block or its neighbour across the edge has coefficients
        const uint32_t V = (nzr | (nzr << 1)) & 0xeeeeu;             // bit 4y + e: vertical edge e >= 1, row y
        // transpose the 4x4 bit matrix (index 4y + e -> 4e + y)
        uint32_t t = V;
        t = (t & 0xA5A5u) | ((t & 0x0A0Au) << 3) | ((t & 0x5050u) >> 3);
        t = (t & 0xCC33u) | ((t & 0x00CCu) << 6) | ((t & 0x3300u) >> 6);
        const uint32_t H = (nzr | (nzr << 4)) & 0xfff0u;             // bit 4e + x: horizontal edge e >= 1, column x
        p2 = t | (H << 16);
        // vectors only differ where 8x8 partitions meet: edge 2
        const uint32_t v2 = (mv_far(a.z, a.w) ? 0x300u : 0u) | (mv_far(b.x, b.y) ? 0xC00u : 0u);
        const uint32_t h2 = (mv_far(a.z, b.x) ? 0x300u : 0u) | (mv_far(a.w, b.y) ? 0xC00u : 0u);
        p1 = (v2 | (h2 << 16)) & ~p2;
    }
    if (mx > 0) {
        const uint32_t *l = q - 12;
        if (intra || type_is_intra(l[0])) p4 |= 0xfu;
        else {
            // left MB: blocks (3, y) = blkIdx 5, 7, 13, 15; this MB: blocks (0, y) = raster bits 0, 4, 8, 12
            const uint32_t n1 = l[7], n3 = l[9];
            const uint32_t ln = ((n1 >> 8) & 255u ? 1u : 0u) | ((n1 >> 24) ? 2u : 0u) | ((n3 >> 8) & 255u ? 4u : 0u) | ((n3 >> 24) ? 8u : 0u);
            const uint32_t cn = (nzr & 1u) | ((nzr >> 3) & 2u) | ((nzr >> 6) & 4u) | ((nzr >> 9) & 8u);
            const uint32_t e2 = ln | cn;
            const uint32_t e1 = (mv_far(l[3], a.z) ? 0x3u : 0u) | (mv_far(l[5], b.x) ? 0xCu : 0u);
            p2 |= e2; p1 |= e1 & ~e2;
        }
    }
    if (my > 0) {
        const uint32_t *u = q - 12 * g.mbw;
        if (intra || type_is_intra(u[0])) p4 |= 0xf0000u;
        else {
            // upper MB: blocks (x, 3) = blkIdx 10, 11, 14, 15; this MB: blocks (x, 0) = raster bits 0..3
            const uint32_t n2 = u[8], n3 = u[9];
            const uint32_t un = ((n2 >> 16) & 255u ? 1u : 0u) | ((n2 >> 24) ? 2u : 0u) | ((n3 >> 16) & 255u ? 4u : 0u) | ((n3 >> 24) ? 8u : 0u);
            const uint32_t e2 = un | (nzr & 15u);
            const uint32_t e1 = (mv_far(u[4], a.z) ? 0x3u : 0u) | (mv_far(u[5], a.w) ? 0xCu : 0u);
            p2 |= e2 << 16; p1 |= (e1 & ~e2) << 16;
        }
    }
    if (t8) { p1 &= 0x0f0f0f0fu; p2 &= 0x0f0f0f0fu; }          // transform_size_8x8_flag: only the 8x8 transform edges are filtered
    s.dbk_bs[mb] = make_uint4(p1, p2, p4, p1 | p2 | p4);
}

// One MB of the row. Software pipeline of the row loop: the MB's own samples and its MbInfo were prefetched into
// registers one iteration earlier (nobody else touches them before this MB runs), the 4 left columns are carried over
// from the previous tile in shared memory, and only the rows above are loaded after the wavefront wait.
struct DbkPrefetch { uint32_t y0, y1, c, below; uint4 bs; uint2 ll; };

// ll_above: the messages of the row above (null on row 0); ll_off: this lane's message inside the two chunks macroblock mx reads
__device__ __forceinline__ void dbk_prefetch(const DbkCtx &s, const Geom &g, int mx, int my, int lane, DbkPrefetch &pf, const uint2 *ll_above, int ll_off)
{
    const int wc = g.wc, cw = wc / 2, mb = my * g.mbw + mx;
    const uint8_t *Y = s.rec[0] + (size_t)my * 16 * wc + mx * 16;
    // luma: 64 words, lane l takes words l and l + 32 (row = w / 4, col word = w % 4)
    pf.y0 = __ldcg(reinterpret_cast<const uint32_t *>(Y + (size_t)(lane >> 2) * wc + (lane & 3) * 4));
    pf.y1 = __ldcg(reinterpret_cast<const uint32_t *>(Y + (size_t)(8 + (lane >> 2)) * wc + (lane & 3) * 4));
    // chroma: 2 planes x 8 rows x 2 words
    const uint8_t *C = ((lane >> 4) ? s.rec[2] : s.rec[1]) + (size_t)(my * 8 + ((lane >> 1) & 7)) * cw + mx * 8 + (lane & 1) * 4;
    pf.c = __ldcg(reinterpret_cast<const uint32_t *>(C));
    pf.bs = __ldg(s.bs + mb);        // the MB's 32 boundary strengths (k_deblock_bs), one broadcast load
    // does the MB below filter its upper edge (bits 16-19 = horizontal edge 0)? Then the bottom rows of this MB are its to store.
    pf.below = my + 1 < g.mbh ? __ldg(&s.bs[mb + g.mbw].w) : 0u;      // (used as (below >> 16) & 15 -- not here: that would wait for the load)
    pf.ll = make_uint2(0u, 0u);
    if (ll_above && lane < DBK_LL_PER_MB) pf.ll = ld_ll(ll_above + mx * DBK_LL_PER_MB + ll_off);     // maybe not published yet: checked (and repeated) at use
}

// returns true when the MB wrote samples (a fence is needed before publishing)
__device__ bool deblock_mb(const DbkCtx &s, const DbkK &k, const Geom &g, DbkSmem &sm, const DbkWb &wb, int mx, int my, int lane, const DbkPrefetch &pf,
                           const uint2 *ll_above, int ll_off, const int *prog_above, uint32_t below_prev, WaveCtl *ctl, bool &ok
#ifdef DBK_TIMING
                           , long long *dbk_t, long long &dbk_last
#endif
                           )
{
    const int wc = g.wc, cw = wc / 2;
    ok = true;
    __syncwarp();
    // carry the previous tile's right columns / MbInfo over as this MB's left neighbour, then drop in the prefetched data
    if (mx > 0) {
        if (lane < 16) sm.y[(lane + 4) * 5] = sm.y[(lane + 4) * 5 + 4];
        else sm.c[(lane >> 3) & 1][((lane & 7) + 4) * 3] = sm.c[(lane >> 3) & 1][((lane & 7) + 4) * 3 + 2];
    }
    __syncwarp();
    sm.y[((lane >> 2) + 4) * 5 + 1 + (lane & 3)] = pf.y0;
    sm.y[((lane >> 2) + 12) * 5 + 1 + (lane & 3)] = pf.y1;
    sm.c[lane >> 4][(((lane >> 1) & 7) + 4) * 3 + 1 + (lane & 1)] = pf.c;
    __syncwarp();
    DBK_T(1);
    if (pf.bs.w == 0u) return false;

    uint8_t *Y = s.rec[0] + (size_t)my * 16 * wc + mx * 16;
    uint8_t *C[2] = { s.rec[1] + (size_t)my * 8 * cw + mx * 8, s.rec[2] + (size_t)my * 8 * cw + mx * 8 };
    const bool topf = ((pf.bs.w >> 16) & 15u) != 0u;      // the upper MB edge is filtered: the only case that needs (and then stores) the rows above
    if (topf) {
        // the rows above are final once the upper-right neighbour is done: their messages carry this launch's sequence number then
        uint2 v = pf.ll;
        const bool mine = lane < DBK_LL_PER_MB;
        if (__ballot_sync(0xffffffffu, mine && v.y != s.seq)) {
            const uint2 *p = ll_above + mx * DBK_LL_PER_MB + ll_off;
            unsigned long long t0 = 0; int spins = 0;
            for (;;) {
                if (mine && v.y != s.seq) v = ld_ll(p);
                if (!__ballot_sync(0xffffffffu, mine && v.y != s.seq)) break;
                // a row that is d macroblocks short of what we need takes d MB-steps: sleep accordingly, so that far-behind rows do not burn
                // the issue slots of the SMs they share with other kernels (the counter is only this hint, it orders nothing)
                const int d = min(mx + 2, g.mbw) - ld_relaxed(prog_above);
                __nanosleep(d > 1 ? min(d * 700, 20000) : 20);
                if ((++spins & 15) == 0) {
                    const unsigned long long t = global_ns();
                    if (!t0) t0 = t;
                    if (ld_relaxed(&ctl->error) || t - t0 > WAVE_TIMEOUT_NS) { if (lane == 0) atomicExch(&ctl->error, 1); ok = false; return false; }
                }
            }
        }
        if (lane < 16) sm.y[(lane >> 2) * 5 + 1 + (lane & 3)] = v.x;
        else if (lane < 24) sm.c[(lane >> 2) & 1][(2 + ((lane >> 1) & 1)) * 3 + 1 + (lane & 1)] = v.x;     // chroma rows -2, -1
    }
    __syncwarp();
    DBK_T(2);
    // lane's boundary strength at edge e of the given direction (bit planes of k_deblock_bs: lane 4e + seg vertical, 16 + 4e + seg horizontal)
    const bool isc = lane >= 16;
    const int seg = isc ? (lane & 7) >> 1 : lane >> 2;
    auto bs_at = [&](int i) { return (int)(((pf.bs.x >> i) & 1u) | (((pf.bs.y >> i) & 1u) << 1) | (((pf.bs.z >> i) & 1u) << 2)); };
    auto tc0_of = [&](int b) { return b == 1 ? k.tc0[0] : b == 2 ? k.tc0[1] : k.tc0[2]; };
    // vertical edges (filtering across columns), left to right: lane = row (lanes 0-15 luma rows, 16-31 the 2 x 8 chroma rows), the row in
    // registers across all four edges. Chroma rows take part in edges 0 and 2 (their columns 0 and 4): word pairs (W0,W1) and (W2,W3), where
    // W2 is the chroma row's middle word again (refreshed after edge 0) and W3 its last.
    if (pf.bs.w & 0xffffu) {
        uint32_t *row = isc ? sm.c[(lane >> 3) & 1] + ((lane & 7) + 4) * 3 : sm.y + (lane + 4) * 5;
        uint32_t W0 = row[0], W1 = row[1], W2 = row[2], W3 = 0, W4 = 0;
        if (!isc) { W3 = row[3]; W4 = row[4]; } else { W3 = W2; W2 = W1; }
        {
            const int b = bs_at(seg);
            if ((pf.bs.w & 0xfu) != 0u) {
                if (__any_sync(0xffffffffu, b == 4)) filter_words<true>(W0, W1, b, k.alpha, k.beta, tc0_of(b), isc);
                else filter_words<false>(W0, W1, b, k.alpha, k.beta, tc0_of(b), isc);
            }
            if (isc) W2 = W1;
        }
        if ((pf.bs.w & 0xf0u) != 0u) { const int b = isc ? 0 : bs_at(4 + seg); filter_words<false>(W1, W2, b, k.alpha, k.beta, tc0_of(b), false); }
        if ((pf.bs.w & 0xf00u) != 0u) { const int b = bs_at(8 + seg); filter_words<false>(W2, W3, b, k.alpha, k.beta, tc0_of(b), isc); }
        if ((pf.bs.w & 0xf000u) != 0u) { const int b = isc ? 0 : bs_at(12 + seg); filter_words<false>(W3, W4, b, k.alpha, k.beta, tc0_of(b), false); }
        row[0] = W0;
        if (!isc) { row[1] = W1; row[2] = W2; row[3] = W3; row[4] = W4; } else { row[1] = W2; row[2] = W3; }
    }
    __syncwarp();
    DBK_T(3);
    // horizontal edges (filtering across rows), top to bottom: lane = column (lanes 0-15 luma, 16-31 the 2 x 8 chroma columns), the column in
    // registers across all four edges. v[i] = luma row i - 4; a chroma column keeps rows -2 .. 1 in v[2..5] (edge 0) and rows 2 .. 5 in
    // v[10..13] (edge 2), exactly where the luma edges' p1 p0 q0 q1 sit, so both share one instruction stream.
    if (pf.bs.w & 0xffff0000u) {
        uint8_t *colA = isc ? reinterpret_cast<uint8_t *>(sm.c[(lane >> 3) & 1]) + 4 + (lane & 7) : reinterpret_cast<uint8_t *>(sm.y) + 4 + lane;
        uint8_t *colB = isc ? colA + 4 * 12 : colA + 8 * 20;          // v[8 + j] = row j of colB
        const int st = isc ? 12 : 20;
        int v[20];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (!isc || (i >= 2 && i < 6)) ? colA[i * st] : 0;
#pragma unroll
        for (int i = 8; i < 16; i++) v[i] = (!isc || (i >= 10 && i < 14)) ? colB[(i - 8) * st] : 0;
#pragma unroll
        for (int i = 16; i < 20; i++) v[i] = !isc ? colB[(i - 8) * st] : 0;
        {
            const int b = bs_at(16 + seg);
            if ((pf.bs.w & 0xf0000u) != 0u) {
                if (__any_sync(0xffffffffu, b == 4)) filter_line<true>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], b, k.alpha, k.beta, tc0_of(b), isc);
                else filter_line<false>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], b, k.alpha, k.beta, tc0_of(b), isc);
            }
        }
        if ((pf.bs.w & 0xf00000u) != 0u) { const int b = isc ? 0 : bs_at(20 + seg); filter_line<false>(v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], b, k.alpha, k.beta, tc0_of(b), false); }
        if ((pf.bs.w & 0xf000000u) != 0u) { const int b = bs_at(24 + seg); filter_line<false>(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15], b, k.alpha, k.beta, tc0_of(b), isc); }
        if ((pf.bs.w & 0xf0000000u) != 0u) { const int b = isc ? 0 : bs_at(28 + seg); filter_line<false>(v[12], v[13], v[14], v[15], v[16], v[17], v[18], v[19], b, k.alpha, k.beta, tc0_of(b), false); }
        // a luma column stores rows -3 .. 14 (bS < 4 inside the MB: row 15 is nobody's p or q side here); a chroma column rows -1, 0, 3, 4
        if (!isc) {
#pragma unroll
            for (int i = 1; i < 8; i++) colA[i * 20] = (uint8_t)v[i];
#pragma unroll
            for (int i = 8; i < 19; i++) colB[(i - 8) * 20] = (uint8_t)v[i];
        } else {
            colA[3 * 12] = (uint8_t)v[3]; colA[4 * 12] = (uint8_t)v[4]; colB[3 * 12] = (uint8_t)v[11]; colB[4 * 12] = (uint8_t)v[12];
        }
    }
    __syncwarp();
    DBK_T(4);
    // write back: the MB with its 4 left columns, and the 3 rows above it (per-lane store slots precomputed once per row)
    // (one writer per sample, see DbkWb: rows above only when this MB filters its upper edge; bottom rows only when the MB below them does not)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int f = wb.fy[k];
        bool st = (f & 1) && ((f & 2) ? ((f & 4) && topf) : ((f & 4) || mx > 0));
        if ((f & 16) && (((f & 4) ? pf.below : below_prev) & 0xf0000u)) st = false;
        if (st) *reinterpret_cast<uint32_t *>(Y + wb.oy[k]) = sm.y[lane + 32 * k];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int f = wb.fc[k];
        bool st = (f & 1) && ((f & 2) ? ((f & 4) && topf) : ((f & 4) || mx > 0));
        if ((f & 16) && (((f & 4) ? pf.below : below_prev) & 0xf0000u)) st = false;
        if (st) *reinterpret_cast<uint32_t *>(C[(f >> 3) & 1] + wb.oc[k]) = sm.c[0][lane + 32 * k];
    }
    DBK_T(5);
    return true;
}

// grid: up to ceil(sessions * mbh / WAVE_WARPS) CTAs of WAVE_WARPS persistent warps. Slices do not break the wavefront:
// disable_deblocking_filter_idc = 0 filters across slice boundaries.
__global__ void __launch_bounds__(WAVE_WARPS * 32, DBK_MIN_CTAS) k_deblock_wave(const Sess *ss, Geom g, int nsess, WaveCtl *ctl)
{
    __shared__ DbkSmem sm_all[WAVE_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DbkWb wb; dbk_wb_init(wb, g.wc, lane);
    // this lane's message among the two chunks a macroblock reads: luma lane = (row r, word w) takes columns 4w .. 4w+3 = word w + 1 of the MB's
    // own chunk, or word 0 of the next chunk for w = 3; chroma lane = (plane, row, word) alike with two words per row
    const int ll_off = lane < 16 ? ((lane & 3) < 3 ? (lane >> 2) * 4 + (lane & 3) + 1 : DBK_LL_PER_MB + (lane >> 2) * 4)
                                 : ((lane & 1) == 0 ? 16 + ((lane >> 2) & 1) * 4 + ((lane >> 1) & 1) * 2 + 1 : DBK_LL_PER_MB + 16 + ((lane >> 2) & 1) * 4 + ((lane >> 1) & 1) * 2);
    // persistent warps: a warp takes the next (row, session) ticket until none is left. Tickets are handed out in wavefront order and a
    // row only ever waits on a row with an earlier ticket, so any number of resident warps makes progress; the grid is sized to the rows a
    // wavefront keeps busy at once (engine.cu) instead of one warp per row holding its registers while it waits for its turn.
    for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&ctl->ticket_dbk, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= nsess * g.mbh) return;
    const int my = t / nsess;
    const Sess &sg = ss[t % nsess];
    int *prog = sg.row_prog_dbk;
    DbkCtx s; s.rec[0] = sg.rec[0]; s.rec[1] = sg.rec[1]; s.rec[2] = sg.rec[2]; s.bs = sg.dbk_bs; s.qp = sg.qp; s.seq = sg.dbk_seq;
    uint2 *ll_row = sg.dbk_ll + (size_t)my * (g.mbw + 1) * DBK_LL_PER_MB;
    const uint2 *ll_above = my > 0 ? ll_row - (size_t)(g.mbw + 1) * DBK_LL_PER_MB : nullptr;
    const bool ll_out = my + 1 < g.mbh;
    DbkSmem &sm = sm_all[warp];
    DbkK kc;
    { const int q_ = lane < 16 ? s.qp : (int)c_chroma_qp[s.qp]; kc.alpha = c_alpha[q_]; kc.beta = c_beta[q_]; kc.tc0[0] = c_tc0[q_][0]; kc.tc0[1] = c_tc0[q_][1]; kc.tc0[2] = c_tc0[q_][2]; }
    DbkPrefetch cur, nxt;
    dbk_prefetch(s, g, 0, my, lane, cur, ll_above, ll_off);
    uint32_t below_prev = 0;
#ifdef DBK_TIMING
    long long dbk_t[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, dbk_last = clock64(); int dbk_n = 0;
#endif
    for (int mx = 0; mx < g.mbw; mx++) {
        if (mx + 1 < g.mbw) dbk_prefetch(s, g, mx + 1, my, lane, nxt, ll_above, ll_off);
        DBK_T(0);
        bool ok;
#ifdef DBK_TIMING
        const bool wrote = deblock_mb(s, kc, g, sm, wb, mx, my, lane, cur, ll_above, ll_off, prog + my - 1, below_prev, ctl, ok, dbk_t, dbk_last); dbk_n += wrote;
#else
        deblock_mb(s, kc, g, sm, wb, mx, my, lane, cur, ll_above, ll_off, prog + my - 1, below_prev, ctl, ok);
#endif
        if (!ok) return;
        // publish what became final with this MB for the row below (the tile is complete in shared memory, filtered or not)
        if (ll_out) {
            uint2 *o = ll_row + mx * DBK_LL_PER_MB;
            if (lane < DBK_LL_PER_MB) st_ll(o + lane, lane < 16 ? sm.y[(16 + (lane >> 2)) * 5 + (lane & 3)] : sm.c[(lane >> 2) & 1][(10 + ((lane >> 1) & 1)) * 3 + (lane & 1)], s.seq);
            if (mx + 1 == g.mbw) {      // the last four columns of the row: word 0 of the extra chunk
                if (lane < 16 && (lane & 3) == 0) st_ll(o + DBK_LL_PER_MB + lane, sm.y[(16 + (lane >> 2)) * 5 + 4], s.seq);
                else if (lane >= 16 && lane < DBK_LL_PER_MB && (lane & 1) == 0) st_ll(o + DBK_LL_PER_MB + lane, sm.c[(lane >> 2) & 1][(10 + ((lane >> 1) & 1)) * 3 + 2], s.seq);
            }
            if (lane == 0) st_relaxed(prog + my, mx + 1);       // distance hint for the sleeping readers
        }
        DBK_T(6);
        below_prev = cur.below;
        cur = nxt;
    }
#ifdef DBK_TIMING
    if (lane == 0 && my == 0 && t % nsess == 0)
        printf("dbk row0: %d MBs (%d filtered) cycles/MB: prefetch %lld stage+bs %lld wait+above %lld vert %lld horz %lld writeback %lld fence+publish %lld\n", g.mbw, dbk_n,
               dbk_t[0] / g.mbw, dbk_t[1] / g.mbw, dbk_t[2] / g.mbw, dbk_t[3] / g.mbw, dbk_t[4] / g.mbw, dbk_t[5] / g.mbw, dbk_t[6] / g.mbw);
#endif
    __syncwarp();
    }
}

} // namespace b200
