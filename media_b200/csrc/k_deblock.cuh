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
#define DBK_MIN_CTAS 6      /* resident CTAs per SM the register allocation must allow: 76 registers, no spills (8 -> 64 registers spills and is slower) */
#endif
#ifdef DBK_TIMING
#define DBK_T(i) do { const long long t_ = clock64(); dbk_t[i] += t_ - dbk_last; dbk_last = t_; } while (0)
#else
#define DBK_T(i) do { } while (0)
#endif

// the few session fields the row loop needs, held in registers: every fence / strong access in the loop is a compiler memory
// barrier, so reading them through `const Sess &` re-fetched them from L2 several times per macroblock (1 800 cycles measured)
struct DbkCtx { uint8_t *rec[3]; const uint4 *bs; int qp; uint32_t seq; };

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

// One edge position of one line of samples, luma or chroma in the same instruction stream (8.7.2.3 / 8.7.2.4): the 16 luma
// lines and the 16 chroma lines of a macroblock edge are filtered by the 32 lanes at once instead of one after the other.
// Chroma uses only p1..q1, tc = tc0 + 1 and the weak bS = 4 filter; the selects below fold that in.
__device__ __forceinline__ void filter_edge(uint8_t *p, int step, int bs, int alpha, int beta, int tc0, bool chroma)
{
    const int p0 = p[-step], p1 = p[-2 * step], q0 = p[0], q1 = p[step];
    if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
    int p2 = 0, q2 = 0;
    if (!chroma) { p2 = p[-3 * step]; q2 = p[2 * step]; }
    const bool ap = !chroma && abs(p2 - p0) < beta, aq = !chroma && abs(q2 - q0) < beta;
    if (bs < 4) {
        const int tc = chroma ? tc0 + 1 : tc0 + (int)ap + (int)aq;
        const int delta = clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        p[-step] = (uint8_t)clip255(p0 + delta);
        p[0] = (uint8_t)clip255(q0 - delta);
        if (ap) p[-2 * step] = (uint8_t)(p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1));
        if (aq) p[step] = (uint8_t)(q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1));
    } else {
        const bool small = abs(p0 - q0) < ((alpha >> 2) + 2);
        if (ap && small) {
            const int p3 = p[-4 * step];
            p[-step] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            p[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            p[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else p[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        if (aq && small) {
            const int q3 = p[3 * step];
            p[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            p[step] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
            p[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else p[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
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
// samples, so they are computed fully in parallel -- warp per MB, lanes 0-15 = vertical edge e, segment k (lane = 4e + k), lanes 16-31 the
// horizontal ones -- and stored as three bit planes of the 32 values (x, y, z = bits 0, 1, 2 of bS by lane; w = their OR). In the row loop
// of k_deblock_wave this was 48 % of the executed instructions and a third of the stall samples (ncu, profiles/r01_ncu_summary.md).
// grid: (ceil(n_mb / 8), 1, sessions), 256 threads
__global__ void __launch_bounds__(256) k_deblock_bs(const Sess *ss, Geom g)
{
    __shared__ uint32_t info_all[8][3][12];                           // per warp: MbInfo of the current, left and upper MB
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, mb = blockIdx.x * 8 + warp;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    const int mx = mb % g.mbw, my = mb / g.mbw;
    uint32_t (*info)[12] = info_all[warp];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(s.mbi + mb);
        if (lane < 12) info[0][lane] = src[lane];
        else if (lane < 24) { if (mx > 0) info[1][lane - 12] = src[lane - 24]; }                       // (mb - 1) * 12 + lane - 12
        if (lane < 12 && my > 0) info[2][lane] = src[lane - 12 * g.mbw];
    }
    __syncwarp();
    const MbInfo *q = reinterpret_cast<const MbInfo *>(info[0]), *ql = reinterpret_cast<const MbInfo *>(info[1]), *qt = reinterpret_cast<const MbInfo *>(info[2]);
    const int e = (lane >> 2) & 3, k = lane & 3; const bool vert = lane < 16;
    int bs;
    if (e == 0) {
        if (vert) bs = mx > 0 ? bs_of(ql, 3, k, q, 0, k, true) : 0;
        else bs = my > 0 ? bs_of(qt, k, 3, q, k, 0, true) : 0;
    } else if ((e & 1) && mb_t8(q)) bs = 0;                           // transform_size_8x8_flag: only the 8x8 transform edges are filtered
    else bs = vert ? bs_of(q, e - 1, k, q, e, k, false) : bs_of(q, k, e - 1, q, k, e, false);
    const uint32_t b0 = __ballot_sync(0xffffffffu, bs & 1), b1 = __ballot_sync(0xffffffffu, bs & 2), b2 = __ballot_sync(0xffffffffu, bs & 4);
    if (lane == 0) s.dbk_bs[mb] = make_uint4(b0, b1, b2, b0 | b1 | b2);
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
    pf.below = my + 1 < g.mbh ? (__ldg(&s.bs[mb + g.mbw].w) >> 16) & 15u : 0u;
    pf.ll = make_uint2(0u, 0u);
    if (ll_above && lane < DBK_LL_PER_MB) pf.ll = ld_ll(ll_above + mx * DBK_LL_PER_MB + ll_off);     // maybe not published yet: checked (and repeated) at use
}

// returns true when the MB wrote samples (a fence is needed before publishing)
__device__ bool deblock_mb(const DbkCtx &s, const Geom &g, DbkSmem &sm, const DbkWb &wb, int mx, int my, int lane, const DbkPrefetch &pf,
                           const uint2 *ll_above, int ll_off, const int *prog_above, uint32_t below_prev, WaveCtl *ctl, bool &ok
#ifdef DBK_TIMING
                           , long long *dbk_t, long long &dbk_last
#endif
                           )
{
    const int wc = g.wc, cw = wc / 2, qp = s.qp, qpc = c_chroma_qp[qp];
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
    // lanes 0-15: bS of vertical edge e, segment k; lanes 16-31: horizontal edge e, segment k (precomputed by k_deblock_bs)
    const int bs = (int)(((pf.bs.x >> lane) & 1u) | (((pf.bs.y >> lane) & 1u) << 1) | (((pf.bs.z >> lane) & 1u) << 2));
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
    const int alphaY = c_alpha[qp], betaY = c_beta[qp], alphaC = c_alpha[qpc], betaC = c_beta[qpc];
    uint8_t *ty = reinterpret_cast<uint8_t *>(sm.y) + 4 * 20 + 4;            // sample (0,0) of the MB
    uint8_t *tc = reinterpret_cast<uint8_t *>(sm.c[(lane >> 3) & 1]) + 4 * 12 + 4;
    const bool isc = lane >= 16;
    const int qpl = isc ? qpc : qp, alphaL = isc ? alphaC : alphaY, betaL = isc ? betaC : betaY;
    const int seg = isc ? (lane & 7) >> 1 : lane >> 2;
    // vertical edges (filtering across columns), left to right; chroma lines take part in the even ones
#pragma unroll 1
    for (int e = 0; e < 4; e++) {
        const int b = __shfl_sync(0xffffffffu, bs, e * 4 + seg);
        uint8_t *p = isc ? tc + (lane & 7) * 12 + 2 * e : ty + lane * 20 + 4 * e;
        if (b && !(isc && (e & 1))) filter_edge(p, 1, b, alphaL, betaL, b < 4 ? c_tc0[qpl][b - 1] : 0, isc);
    }
    __syncwarp();
    DBK_T(3);
    // horizontal edges (filtering across rows), top to bottom
#pragma unroll 1
    for (int e = 0; e < 4; e++) {
        const int b = __shfl_sync(0xffffffffu, bs, 16 + e * 4 + seg);
        uint8_t *p = isc ? tc + 2 * e * 12 + (lane & 7) : ty + 4 * e * 20 + lane;
        if (b && !(isc && (e & 1))) filter_edge(p, isc ? 12 : 20, b, alphaL, betaL, b < 4 ? c_tc0[qpl][b - 1] : 0, isc);
    }
    __syncwarp();
    DBK_T(4);
    // write back: the MB with its 4 left columns, and the 3 rows above it (per-lane store slots precomputed once per row)
    // (one writer per sample, see DbkWb: rows above only when this MB filters its upper edge; bottom rows only when the MB below them does not)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int f = wb.fy[k];
        bool st = (f & 1) && ((f & 2) ? ((f & 4) && topf) : ((f & 4) || mx > 0));
        if ((f & 16) && ((f & 4) ? pf.below : below_prev)) st = false;
        if (st) *reinterpret_cast<uint32_t *>(Y + wb.oy[k]) = sm.y[lane + 32 * k];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int f = wb.fc[k];
        bool st = (f & 1) && ((f & 2) ? ((f & 4) && topf) : ((f & 4) || mx > 0));
        if ((f & 16) && ((f & 4) ? pf.below : below_prev)) st = false;
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
        const bool wrote = deblock_mb(s, g, sm, wb, mx, my, lane, cur, ll_above, ll_off, prog + my - 1, below_prev, ctl, ok, dbk_t, dbk_last); dbk_n += wrote;
#else
        deblock_mb(s, g, sm, wb, mx, my, lane, cur, ll_above, ll_off, prog + my - 1, below_prev, ctl, ok);
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
