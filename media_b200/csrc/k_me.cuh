// media_b200/csrc/k_me.cuh -- motion estimation and inter macroblock coding (phases A and B of DESIGN.md 3).
//
// Role inside the reference: the ME / MC / transform part of ISVCEncoder::EncodeFrame
// (video_codec/VideoEncoderOpenH264.cpp:344; openh264's WelsMotionEstimateSearch, MeRefineFracPixel, McHorVer*,
// WelsDctT4, WelsQuant4x4, WelsIDctT4Rec live in the absent libopenh264). Everything here is independent per
// macroblock, so the grid is one warp per MB across all sessions of the batch.
#pragma once
#include "h264_dev.cuh"
#include <cstddef>

namespace b200 {

#ifndef ME_WARPS
#define ME_WARPS 8
#endif
#ifndef ME_FINE_MIN_CTAS
#define ME_FINE_MIN_CTAS 4     /* 64 registers: 4 CTAs of 8 warps per SM (measured 8 % faster than 75-94 registers) */
#endif

__device__ __forceinline__ uint32_t warp_min(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }   // one REDUX.MIN
__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

// Window staging from a PADDED plane (`org` addresses sample (0,0); the border makes every search window addressable, so
// there is no clamping): aligned 32-bit loads + funnel shift. ww must be a multiple of 4.
template <int WW, int WH>
__device__ __forceinline__ void stage_window(uint32_t *dstw, const uint8_t *__restrict__ org, int stride, int x0, int y0, int lane)
{
    constexpr int WPR = WW / 4;
    const int sh = (x0 & 3) * 8, sw = stride >> 2;
    const uint32_t *base = reinterpret_cast<const uint32_t *>(org + (ptrdiff_t)y0 * stride + (x0 & ~3));
    for (int i = lane; i < WPR * WH; i += 32) {
        const int r = i / WPR, j = i - r * WPR;
        const uint32_t *p = base + r * sw + j;
        const uint32_t lo = __ldg(p), hi = sh ? __ldg(p + 1) : 0u;
        dstw[i] = __funnelshift_r(lo, hi, sh);
    }
}
__device__ __forceinline__ void stage_window_rt(uint32_t *dstw, int W, const uint8_t *__restrict__ org, int stride, int x0, int y0, int lane)
{
    const int wpr = W >> 2, sh = (x0 & 3) * 8, sw = stride >> 2;
    const uint32_t *base = reinterpret_cast<const uint32_t *>(org + (ptrdiff_t)y0 * stride + (x0 & ~3));
    for (int r = 0; r < W; r++)
        for (int j = lane; j < wpr; j += 32) {
            const uint32_t *p = base + r * sw + j;
            const uint32_t lo = __ldg(p), hi = sh ? __ldg(p + 1) : 0u;
            dstw[r * wpr + j] = __funnelshift_r(lo, hi, sh);
        }
}

// ---- coarse levels: 1/4 resolution exhaustive search, 1/2 resolution refinement ----
// The search windows are kept as four byte-shifted copies (copy k, word j of a row = bytes 4j+k .. 4j+k+3), so every
// candidate row is read with aligned 32-bit loads straight into VABSDIFF4.
template <int MAXR4> struct CoarseSmem {
    static constexpr int MAXW = 8 + 2 * MAXR4, CW = MAXW * MAXW / 4 + 8;   // +8: fewest bank conflicts of the candidate reads at R = 16 (modelled)
    uint32_t win[4][CW]; uint32_t src[16]; uint32_t win1[4][12 * 3 + 2];
};
// build copies 1..3 from copy 0 (n words each, rows are contiguous so word j+1 is the right neighbour)
__device__ __forceinline__ void make_shifted_copies(uint32_t *c0, int copy_stride, int n, int lane)
{
    for (int i = lane; i < n; i += 32) {
        const uint32_t a = c0[i], b = c0[i + 1];
        c0[copy_stride + i] = __funnelshift_r(a, b, 8);
        c0[2 * copy_stride + i] = __funnelshift_r(a, b, 16);
        c0[3 * copy_stride + i] = __funnelshift_r(a, b, 24);
    }
}

// EXACT: the search range equals 4 * MAXR4 (the usual 16 / 32 / 64), so every window dimension and divisor is a compile-time constant
template <int MAXR4, int WARPS, bool EXACT>
__global__ void __launch_bounds__(WARPS * 32) k_me_coarse(const Sess *ss, Geom g)
{
    typedef CoarseSmem<MAXR4> Smem;
    __shared__ Smem sm_all[WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * WARPS + warp;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    if (s.is_idr) return;
    Smem &sm = sm_all[warp];
    int mx, my; mb_xy(g, mb, mx, my);
    const int R4 = EXACT ? MAXR4 : g.search_range / 4, span = 2 * R4 + 1, W = 8 + 2 * R4, wpr = W >> 2;

    // level 2: 8x8 block centred on the MB (origin 4mx-2, 4my-2), all (2R4+1)^2 displacements
    stage_window<8, 8>(sm.src, s.srcL2, g.s2, 4 * mx - 2, 4 * my - 2, lane);
    if (R4 == MAXR4) stage_window<8 + 2 * MAXR4, 8 + 2 * MAXR4>(sm.win[0], s.refL2, g.s2, 4 * mx - 2 - R4, 4 * my - 2 - R4, lane);
    else stage_window_rt(sm.win[0], W, s.refL2, g.s2, 4 * mx - 2 - R4, 4 * my - 2 - R4, lane);
    __syncwarp();
    make_shifted_copies(sm.win[0], Smem::CW, W * wpr, lane);
    uint32_t sw[16];
#pragma unroll
    for (int i = 0; i < 16; i++) sw[i] = sm.src[i];
    __syncwarp();
    uint32_t best = 0xffffffffu;
    {
        int dy = lane / span, dx = lane - dy * span;              // candidate = lane + 32k, walked without divisions
        const int sdy = 32 / span, sdx = 32 - sdy * span;
        for (int cand = lane; cand < span * span; cand += 32) {
            const uint32_t *p = sm.win[dx & 3] + dy * wpr + (dx >> 2);
            uint32_t sad = 0;
#pragma unroll
            for (int r = 0; r < 8; r++) { sad = sad4(sw[2 * r], p[r * wpr], sad); sad = sad4(sw[2 * r + 1], p[r * wpr + 1], sad); }
            best = min(best, ((sad + abs(dx - R4) + abs(dy - R4)) << 11) | (uint32_t)cand);
            dy += sdy; dx += sdx;
            if (dx >= span) { dx -= span; dy++; }
        }
    }
    best = warp_min(best);
    const int c2 = best & 2047, v2y = c2 / span - R4, v2x = c2 - (v2y + R4) * span - R4;

    // level 1: 8x8 block at (8mx, 8my), +-2 around 2*mv2
    const int cx = 2 * v2x, cy = 2 * v2y;
    __syncwarp();
    stage_window<8, 8>(sm.src, s.srcL1, g.s1, 8 * mx, 8 * my, lane);
    stage_window<12, 12>(sm.win1[0], s.refL1, g.s1, 8 * mx + cx - 2, 8 * my + cy - 2, lane);
    __syncwarp();
    make_shifted_copies(sm.win1[0], 12 * 3 + 2, 36, lane);
    __syncwarp();
    best = 0xffffffffu;
    if (lane < 25) {
        const int dy = lane / 5, dx = lane - dy * 5;
        const uint32_t *p = sm.win1[dx & 3] + dy * 3 + (dx >> 2);
        uint32_t sad = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) { sad = sad4(sm.src[2 * r], p[r * 3], sad); sad = sad4(sm.src[2 * r + 1], p[r * 3 + 1], sad); }
        best = ((sad + abs(cx + dx - 2) + abs(cy + dy - 2)) << 5) | (uint32_t)lane;
    }
    best = warp_min(best);
    if (lane == 0) {
        const int c1 = best & 31;
        s.me2[mb * 2] = (int16_t)v2x; s.me2[mb * 2 + 1] = (int16_t)v2y;
        s.me1[mb * 2] = (int16_t)(cx + c1 % 5 - 2); s.me1[mb * 2 + 1] = (int16_t)(cy + c1 / 5 - 2);
    }
}

// ---- fine level: full-pel refinement, half/quarter-pel SATD refinement, intra estimate, inter coding ----
// TMA (cp.async.bulk.tensor) moves the three tiles of a macroblock into shared memory: the source MB, the 20x20 full-pel
// window, and the [-1,16]^2 windows of the four reference planes G, b, h, j in one 3-D box. The innermost TMA coordinate
// must be 16-byte aligned (measured on B200: any other value raises an illegal-instruction fault), so boxes are 48 bytes
// wide, start at the aligned column below the window and the kernel carries the 0..15 byte offset.
#define PL_STRIDE 48                     /* bytes per row of the TMA boxes */
#define PL_ROWS 18
struct __align__(128) FineSmem {
    uint32_t plane[4][PL_ROWS * PL_STRIDE / 4];   // G, b, h, j: sample (x,y) of the best full-pel block at (y+1)*48 + o1 + 4 + x (3456 B)
    uint32_t src[64];                             // source MB, 16x16 (256 B)
    uint32_t win[20 * PL_STRIDE / 4];             // 20 rows of plane G around the level-0 centre; window column c at byte o0 + c (960 B)
    uint32_t nb_top[4], nb_left[4];               // SOURCE neighbours for the intra estimate
    unsigned long long bar[2];                    // mbarriers: [0] source + window, [1] planes
    uint32_t pad_[20];
};
static_assert(sizeof(FineSmem) % 128 == 0 && offsetof(FineSmem, src) % 128 == 0 && offsetof(FineSmem, win) % 128 == 0, "TMA destinations must be 128-byte aligned");
// the five byte-shifted copies of the 20-row window live in the plane area until the planes are fetched
#define WIN_COPY 336            /* bytes between the byte-shifted copies of the full-pel window: 20 rows x 16 bytes + 16 */
#define P8X8_BIAS_BITS 8        /* extra header bits of P_8x8 over P_L0_16x16: mb_type ue(3) vs ue(0), four sub_mb_type ue(0) */

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait (phase 0): a transaction that never completes must not hang the GPU; returns false after 2 s
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar)
{
    uint32_t ok = 0; unsigned long long t0 = 0;
    for (unsigned spin = 0; !ok; spin++) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
        if (!ok && (spin & 1023u) == 1023u) {
            unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (!t0) t0 = t; else if (t - t0 > 2000000000ull) return false;
        }
    }
    return true;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int x, int y, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const void *tmap, int x, int y, int z, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// The 16 quarter-pel positions as the average of two samples out of {G,b,h,j} (8.4.2.2.1, Table 8-12):
// entry = {planeA, dxA, dyA, planeB, dxB, dyB}
static __device__ __constant__ uint8_t c_qpel_tab[16][6] = {
    { 0, 0, 0, 0, 0, 0 }, { 0, 0, 0, 1, 0, 0 }, { 1, 0, 0, 1, 0, 0 }, { 0, 1, 0, 1, 0, 0 },
    { 0, 0, 0, 2, 0, 0 }, { 1, 0, 0, 2, 0, 0 }, { 1, 0, 0, 3, 0, 0 }, { 1, 0, 0, 2, 1, 0 },
    { 2, 0, 0, 2, 0, 0 }, { 2, 0, 0, 3, 0, 0 }, { 3, 0, 0, 3, 0, 0 }, { 3, 0, 0, 2, 1, 0 },
    { 0, 0, 1, 2, 0, 0 }, { 2, 0, 0, 1, 0, 1 }, { 3, 0, 0, 1, 0, 1 }, { 2, 1, 0, 1, 0, 1 },
};

// the same table as byte offsets into FineSmem::plane (plane * 18 * 48 + dy * 48 + dx): sample A in the low, sample B in the high half
static __device__ __constant__ uint32_t c_qpel_off[16] = { 0x00000000u, 0x03600000u, 0x03600360u, 0x03600001u, 0x06c00000u, 0x06c00360u, 0x0a200360u, 0x06c10360u, 0x06c006c0u, 0x0a2006c0u, 0x0a200a20u, 0x06c10a20u, 0x06c00030u, 0x039006c0u, 0x03900a20u, 0x039006c1u };
static_assert(PL_ROWS * PL_STRIDE == 864, "c_qpel_off was generated for 18 rows of 48 bytes per plane");

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)   // sum of u8(a_i) * s8(b_i) + c: one IDP.4A
{
    int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t avg4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) & 0xfefefefeu) >> 1); }  // per-byte (a+b+1)>>1
__device__ __forceinline__ uint32_t plane_row(const uint32_t *pl, int o)   // 4 samples starting at byte offset o of a plane
{
    const uint32_t *w = pl + (o >> 2);
    return __funnelshift_r(w[0], w[1], (o & 3) * 8);
}
// four rows (packed bytes) of the prediction of one 4x4 block at quarter-pel offset (ox,oy) in [-3,3] from the best full-pel
// block; o1 = byte offset of the window inside the TMA box; P[y] receives row (y ^ rx): the callers permute the rows of the
// lower two block rows (rx = 1) so that the 16 block lanes of a half-warp hit 16 different banks at the 48-byte row pitch,
// and the Hadamard SATD is invariant under that (dyadic) permutation as long as the source rows are permuted alike.
__device__ __forceinline__ void pred_rows_qpel(const FineSmem &sm, int o1, int bx, int by, int ox, int oy, int rx, uint32_t P[4])
{
    const uint32_t t = c_qpel_off[(oy & 3) * 4 + (ox & 3)];
    const int common = (by + (oy >> 2) + 1) * PL_STRIDE + o1 + bx + (ox >> 2) + 4;
    const uint32_t *base = sm.plane[0];
    const int oa = common + (int)(t & 0xffffu);
    // The row pitch (48 bytes) is a multiple of 4, so the byte shift is the same for all four rows, and rows y ^ rx are words
    // {rx*12, 12 - rx*12} and the same + 24: two base pointers per plane, the rest are immediate offsets.
    const int rxw = rx * (PL_STRIDE / 4);
    const uint32_t *a0 = base + (oa >> 2) + rxw, *a1 = base + (oa >> 2) + (PL_STRIDE / 4) - rxw;
    const int sa = (oa & 3) * 8;
    const uint32_t A0 = __funnelshift_r(a0[0], a0[1], sa), A1 = __funnelshift_r(a1[0], a1[1], sa);
    const uint32_t A2 = __funnelshift_r(a0[PL_STRIDE / 2], a0[PL_STRIDE / 2 + 1], sa), A3 = __funnelshift_r(a1[PL_STRIDE / 2], a1[PL_STRIDE / 2 + 1], sa);
    if (((ox | oy) & 1) == 0) {          // full/half-pel positions are a single plane (both table entries coincide)
        P[0] = A0; P[1] = A1; P[2] = A2; P[3] = A3;
    } else {
        const int ob = common + (int)(t >> 16);
        const uint32_t *b0 = base + (ob >> 2) + rxw, *b1 = base + (ob >> 2) + (PL_STRIDE / 4) - rxw;
        const int sb = (ob & 3) * 8;
        P[0] = avg4(A0, __funnelshift_r(b0[0], b0[1], sb)); P[1] = avg4(A1, __funnelshift_r(b1[0], b1[1], sb));
        P[2] = avg4(A2, __funnelshift_r(b0[PL_STRIDE / 2], b0[PL_STRIDE / 2 + 1], sb)); P[3] = avg4(A3, __funnelshift_r(b1[PL_STRIDE / 2], b1[PL_STRIDE / 2 + 1], sb));
    }
}
// 4x4 Hadamard SATD of (source block - P). Ts holds the source's part of the first vertical butterfly stage: for basis k,
// Ts[k] / Ts[4+k] = horizontal transform of source row 0 +/- row 1, Ts[8+k] / Ts[12+k] = row 2 +/- row 3. The prediction rows are folded in
// with chained dp4a against the negated / plain +-1 basis, so the butterfly's additions run as IDP.4A on the FMA pipe (the kernel is bound
// by the ALU pipe: math-pipe throttle was its second stall reason, profiles/r02a_ncu_summary.md); |x+y|+|x-y| = 2 max(|x|,|y|) finishes the columns.
__device__ __forceinline__ void satd_source_terms(const uint32_t S[4], int Ts[16])    // S: the four source rows (packed samples)
{
    const uint32_t H[4] = { 0x01010101u, 0xffff0101u, 0x01ffff01u, 0xff01ff01u }, NH[4] = { 0xffffffffu, 0x0101ffffu, 0xff0101ffu, 0x01ff01ffu };
#pragma unroll
    for (int y = 0; y < 4; y += 2)
#pragma unroll
        for (int k = 0; k < 4; k++) { const int h0 = dp4a_us(S[y], H[k], 0); Ts[y * 4 + k] = dp4a_us(S[y + 1], H[k], h0); Ts[y * 4 + 4 + k] = dp4a_us(S[y + 1], NH[k], h0); }
}
__device__ __forceinline__ int satd_rows(const uint32_t P[4], const int Ts[16])
{
    const uint32_t NH[4] = { 0xffffffffu, 0x0101ffffu, 0xff0101ffu, 0x01ff01ffu }, PH[4] = { 0x01010101u, 0xffff0101u, 0x01ffff01u, 0xff01ff01u };
    int s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int a0 = dp4a_us(P[0], NH[k], dp4a_us(P[1], NH[k], Ts[k])), a1 = dp4a_us(P[0], NH[k], dp4a_us(P[1], PH[k], Ts[4 + k]));
        const int a2 = dp4a_us(P[2], NH[k], dp4a_us(P[3], NH[k], Ts[8 + k])), a3 = dp4a_us(P[2], NH[k], dp4a_us(P[3], PH[k], Ts[12 + k]));
        s += max(abs(a0), abs(a2)) + max(abs(a1), abs(a3));
    }
    return s;
}
__device__ __forceinline__ int half_reduce16(int v)   // sum over the 16 lanes of a half-warp
{
#pragma unroll
    for (int o = 8; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(ME_WARPS * 32, ME_FINE_MIN_CTAS) k_me_fine(const Sess *ss, Geom g, WaveCtl *err)
{
    __shared__ FineSmem sm_all[ME_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * ME_WARPS + warp;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    if (s.is_idr) return;
    FineSmem &sm = sm_all[warp];
    int mx, my; mb_xy(g, mb, mx, my);
    const int x0 = mx * 16, y0 = my * 16, wc = g.wc;
    const int qp = s.qp, lambda = c_lambda[qp];
    const char *tmaps = static_cast<const char *>(s.tmaps);

    // source MB and the +-2 window around 2*mv1, fetched by TMA
    const int cx = 2 * s.me1[mb * 2], cy = 2 * s.me1[mb * 2 + 1];
    const int X0 = g.lp + x0 + cx - 2, o0 = X0 & 15;
    const bool top = !row_is_slice_top(g, my), left = mx > 0;
    // estimate of the MV predictor for the rate terms below: 8.4.1.3 median over the LEVEL-1 vectors of the neighbours (x8)
    int ppx, ppy;
    {
        int ax = 0, ay = 0, tx = 0, ty = 0, rx = 0, ry = 0;
        const int16_t *m1 = s.me1;
        if (left) { ax = m1[(mb - 1) * 2]; ay = m1[(mb - 1) * 2 + 1]; }
        if (top) { tx = m1[(mb - g.mbw) * 2]; ty = m1[(mb - g.mbw) * 2 + 1]; }
        const int nc = top && mx + 1 < g.mbw ? mb - g.mbw + 1 : (top && left ? mb - g.mbw - 1 : -1);
        if (nc >= 0) { rx = m1[nc * 2]; ry = m1[nc * 2 + 1]; }
        ppx = 8 * median3(ax, tx, rx); ppy = 8 * median3(ay, ty, ry);
    }
    if (lane == 0) {
        mbar_init(&sm.bar[0]); mbar_init(&sm.bar[1]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sm.bar[0], 256 + 20 * PL_STRIDE);
        tma_load_2d(sm.src, tmaps + s.src_tmap, x0, y0, &sm.bar[0]);
        tma_load_2d(sm.win, tmaps + 128, X0 & ~15, g.lp + y0 + cy - 2, &sm.bar[0]);
    }
    // zero-vector candidate, computed cooperatively straight from HBM while the tiles are in flight
    uint32_t zsad;
    uint2 zref, zsrc;
    {
        const uint8_t *rp = s.ref[0] + (size_t)(y0 + (lane >> 1)) * wc + x0 + (lane & 1) * 8;
        const uint8_t *sp = s.src[0] + (size_t)(y0 + (lane >> 1)) * wc + x0 + (lane & 1) * 8;
        uint2 a = *reinterpret_cast<const uint2 *>(rp), b = *reinterpret_cast<const uint2 *>(sp);
        zsad = sad4(a.x, b.x, sad4(a.y, b.y, 0)); zref = a; zsrc = b;
        zsad = __reduce_add_sync(0xffffffffu, zsad);         // one REDUX.SUM instead of five shuffle + add steps
    }
    __syncwarp();
    bool tma_ok = mbar_wait(&sm.bar[0]);
    // EARLY SKIP: with a zero predictor estimate, a macroblock whose zero-vector residual quantises to nothing in
    // luma and chroma is final -- P_L0_16x16, vector (0,0), cbp 0, reconstruction = reference. Static screen content (the bulk of
    // a cloud-phone framebuffer) never enters the search. Lanes 0-15 test the luma blocks, 16-23 the chroma blocks.
    if (ppx == 0 && ppy == 0) {
        const bool isl = lane < 16, act = lane < 24;
        const int cw = wc / 2, pl = (lane >> 2) & 1, cb = lane & 3, b = lane & 15;
        const int st = isl ? wc : cw;
        const size_t off = isl ? (size_t)(y0 + blk_y(b) * 4) * wc + x0 + blk_x(b) * 4 : (size_t)(my * 8 + (cb >> 1) * 4) * cw + mx * 8 + (cb & 1) * 4;
        const uint8_t *sp = s.src[isl ? 0 : 1 + pl] + off, *rp = s.ref[isl ? 0 : 1 + pl] + off;
        int c[16];
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const uint32_t ws = *reinterpret_cast<const uint32_t *>(sp + (size_t)y * st), wr = *reinterpret_cast<const uint32_t *>(rp + (size_t)y * st);
#pragma unroll
            for (int x = 0; x < 4; x++) c[y * 4 + x] = (int)((ws >> (8 * x)) & 255) - (int)((wr >> (8 * x)) & 255);
        }
        fdct4x4(c);
        const QParam q = make_qparam(isl ? qp : (int)c_chroma_qp[qp]);
        bool nz = quant_any_nonzero(c, q, q.f_inter, !isl);
        {   // chroma DC: 2x2 Hadamard of the four DC terms of the lane's plane (lanes 16+4pl .. 19+4pl)
            int dcs[4];
#pragma unroll
            for (int k = 0; k < 4; k++) dcs[k] = __shfl_sync(0xffffffffu, c[0], 16 + pl * 4 + k);
            const int hd[4] = { dcs[0] + dcs[1] + dcs[2] + dcs[3], dcs[0] - dcs[1] + dcs[2] - dcs[3], dcs[0] + dcs[1] - dcs[2] - dcs[3], dcs[0] - dcs[1] - dcs[2] + dcs[3] };
            if (!isl) for (int k = 0; k < 4; k++) nz |= quant_dc(hd[k], q, q.f_inter) != 0;
        }
        // BACKGROUND DETECTION (the wrapper's bEnableBackgroundDetection, VideoEncoderOpenH264.cpp:282): static against the PREVIOUS
        // SOURCE picture -- every 8x8 unit of luma and both chroma blocks with SAD <= 128 and no sample off by more than 12 -- and close enough to
        // the reference (zero-vector luma SAD <= 64 (8 + lambda)): skipped like a macroblock whose residual vanishes
        bool bg = false;
        if (s.bgd && s.src_prev[0]) {
            const uint2 pv = *reinterpret_cast<const uint2 *>(s.src_prev[0] + (size_t)(y0 + (lane >> 1)) * wc + x0 + (lane & 1) * 8);
            uint32_t ou = sad4(zsrc.x, pv.x, sad4(zsrc.y, pv.y, 0));       // lane = (row, half): its unit's lanes differ in bits 1-3
            ou += __shfl_xor_sync(0xffffffffu, ou, 2); ou += __shfl_xor_sync(0xffffffffu, ou, 4); ou += __shfl_xor_sync(0xffffffffu, ou, 8);
            const uint32_t d0 = __vabsdiffu4(zsrc.x, pv.x), d1 = __vabsdiffu4(zsrc.y, pv.y);
            uint32_t big = ((((d0 & 0x7f7f7f7fu) + 0x73737373u) | d0) | (((d1 & 0x7f7f7f7fu) + 0x73737373u) | d1)) & 0x80808080u;     // a byte above 12
            const int cp_ = (lane >> 3) & 1, cr_ = lane & 7;
            const size_t coff = (size_t)(my * 8 + cr_) * cw + mx * 8;
            const uint2 cc = *reinterpret_cast<const uint2 *>(s.src[1 + cp_] + coff), cq = *reinterpret_cast<const uint2 *>(s.src_prev[1 + cp_] + coff);
            uint32_t cs_ = sad4(cc.x, cq.x, sad4(cc.y, cq.y, 0));
            cs_ += __shfl_xor_sync(0xffffffffu, cs_, 1); cs_ += __shfl_xor_sync(0xffffffffu, cs_, 2); cs_ += __shfl_xor_sync(0xffffffffu, cs_, 4);
            const uint32_t e0 = __vabsdiffu4(cc.x, cq.x), e1 = __vabsdiffu4(cc.y, cq.y);
            big |= ((((e0 & 0x7f7f7f7fu) + 0x73737373u) | e0) | (((e1 & 0x7f7f7f7fu) + 0x73737373u) | e1)) & 0x80808080u;
            bg = __ballot_sync(0xffffffffu, ou > 128u || cs_ > 128u || big != 0u) == 0u && zsad <= 64u * (8u + (uint32_t)lambda);
        }
        if (bg || __ballot_sync(0xffffffffu, act && nz) == 0) {
            const int cpl = lane >> 4, crow = (lane >> 1) & 7, chalf = lane & 1;
            const size_t co_ = (size_t)(my * 8 + crow) * cw + mx * 8 + chalf * 4;
            *reinterpret_cast<uint2 *>(s.rec[0] + (size_t)(y0 + (lane >> 1)) * wc + x0 + (lane & 1) * 8) = zref;
            *reinterpret_cast<uint32_t *>(s.rec[1 + cpl] + co_) = *reinterpret_cast<const uint32_t *>(s.ref[1 + cpl] + co_);
            if (s.dump) {      // nobody reads the levels of a cbp-0 macroblock except the stage dumps of the parity tests
                uint4 *cz = reinterpret_cast<uint4 *>(s.coef + mb);
                cz[lane] = make_uint4(0, 0, 0, 0);
                if (lane < 51 - 32) cz[32 + lane] = make_uint4(0, 0, 0, 0);
            }
            if (lane < 12) reinterpret_cast<uint32_t *>(s.mbi + mb)[lane] = 0u;      // P_L0_16x16, cbp 0, zero vectors, nnz 0
            if (lane == 0) { s.me0[mb * 2] = 0; s.me0[mb * 2 + 1] = 0; s.inter_cost[mb] = 0; }
            return;
        }
    }
    // Five byte-shifted copies of the 20-row window, one per candidate column (copy k = the window from column k: 16 bytes per row), so a
    // candidate row is ONE aligned 128-bit load against one 128-bit load of the source row. Copies are WIN_COPY bytes apart: the quarter-warps
    // of both the 128-bit stores here and the 25 candidate loads below then touch distinct 16-byte bank groups (k * 336 mod 128 = 0, 80, 32, 112, 64).
    uint8_t *cpb = reinterpret_cast<uint8_t *>(sm.plane[0]);
    // One lane per window row: the row's 48 bytes as three conflict-free 128-bit loads (a quarter-warp's rows start at 16-byte groups 3r mod 8),
    // the five words from byte o0 by ONE data-dependent funnel shift each (u_j = bytes o0 + 4j ..), copies 1-3 from those by constant shifts,
    // copy 4 = the same words one further: 5 + 12 shifts and 5 stores per row instead of 4 shifts, 5 loads and a division per (row, copy).
    if (lane < 20) {
        const uint4 *rowv = reinterpret_cast<const uint4 *>(sm.win + lane * (PL_STRIDE / 4));
        const uint4 q0 = rowv[0], q1 = rowv[1], q2 = rowv[2];
        const uint32_t w[12] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w };
        const int sh = (o0 & 3) * 8;
        uint32_t u0, u1, u2, u3, u4;
#define ME_WSEL(A) { u0 = __funnelshift_r(w[A], w[A + 1], sh); u1 = __funnelshift_r(w[A + 1], w[A + 2], sh); u2 = __funnelshift_r(w[A + 2], w[A + 3], sh); \
                     u3 = __funnelshift_r(w[A + 3], w[A + 4], sh); u4 = __funnelshift_r(w[A + 4], w[A + 5], sh); }
        switch (o0 >> 2) { case 0: ME_WSEL(0) break; case 1: ME_WSEL(1) break; case 2: ME_WSEL(2) break; default: ME_WSEL(3) break; }
#undef ME_WSEL
        uint8_t *d = cpb + lane * 16;
        B200_CHECK((o0 >> 2) + 5 < PL_STRIDE / 4 && lane * 16 + 4 * WIN_COPY + 16 <= (int)sizeof(sm.plane), 8);
        *reinterpret_cast<uint4 *>(d) = make_uint4(u0, u1, u2, u3);
        *reinterpret_cast<uint4 *>(d + WIN_COPY) = make_uint4(__funnelshift_r(u0, u1, 8), __funnelshift_r(u1, u2, 8), __funnelshift_r(u2, u3, 8), __funnelshift_r(u3, u4, 8));
        *reinterpret_cast<uint4 *>(d + 2 * WIN_COPY) = make_uint4(__funnelshift_r(u0, u1, 16), __funnelshift_r(u1, u2, 16), __funnelshift_r(u2, u3, 16), __funnelshift_r(u3, u4, 16));
        *reinterpret_cast<uint4 *>(d + 3 * WIN_COPY) = make_uint4(__funnelshift_r(u0, u1, 24), __funnelshift_r(u1, u2, 24), __funnelshift_r(u2, u3, 24), __funnelshift_r(u3, u4, 24));
        *reinterpret_cast<uint4 *>(d + 4 * WIN_COPY) = make_uint4(u1, u2, u3, u4);
    }
    __syncwarp();
    uint32_t best = 0xffffffffu;
    if (lane < 25) {
        const int dy = lane / 5, dx = lane - dy * 5;
        const uint4 *p = reinterpret_cast<const uint4 *>(cpb + dx * WIN_COPY + dy * 16), *sp = reinterpret_cast<const uint4 *>(sm.src);
        uint32_t sad = 0;
#pragma unroll 4
        for (int r = 0; r < 16; r++) {
            const uint4 a = sp[r], w = p[r];
            sad = sad4(a.x, w.x, sad); sad = sad4(a.y, w.y, sad); sad = sad4(a.z, w.z, sad); sad = sad4(a.w, w.w, sad);
        }
        best = ((sad + lambda * (se_len(4 * (cx + dx - 2) - ppx) + se_len(4 * (cy + dy - 2) - ppy))) << 5) | (uint32_t)lane;
    } else if (lane == 25) best = ((zsad + lambda * (se_len(-ppx) + se_len(-ppy))) << 5) | 25u;
    best = warp_min(best);
    const int c0 = best & 31;
    const int fx = c0 < 25 ? cx + c0 % 5 - 2 : 0, fy = c0 < 25 ? cy + c0 / 5 - 2 : 0;   // best full-pel vector

    // the four reference planes around the winner: one 3-D TMA box (the shifted copies above are dead by now)
    const int X1 = g.lp + x0 + fx - 4, o1 = X1 & 15;
    __syncwarp();
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&sm.bar[1], 4 * PL_ROWS * PL_STRIDE);
        tma_load_3d(sm.plane, tmaps + 256, X1 & ~15, g.lp + y0 + fy - 1, 0, &sm.bar[1]);
    }
    // sub-pel refinement by SATD: lane = (half-warp hw, 4x4 block b). A half-warp evaluates one candidate at a time over its 16 blocks, the
    // two half-warps evaluate different candidates in the same instruction stream.
    const int hw = lane >> 4, b = lane & 15, bx = blk_x(b) * 4, by = blk_y(b) * 4;
    int Ts[16];                                         // horizontal Hadamard of the source rows of this lane's block, rows 0 +/- 1 and 2 +/- 3 (see satd_rows)
    {
        uint32_t S[4];
#pragma unroll
        for (int y = 0; y < 4; y++) S[y] = sm.src[(by + y) * 4 + (bx >> 2)];
        satd_source_terms(S, Ts);
    }
    {
        uint8_t *nt = reinterpret_cast<uint8_t *>(sm.nb_top), *nl = reinterpret_cast<uint8_t *>(sm.nb_left);
        if (lane < 16) nt[lane] = top ? s.src[0][(size_t)(y0 - 1) * wc + x0 + lane] : 0;
        else nl[lane - 16] = left ? s.src[0][(size_t)(y0 + lane - 16) * wc + x0 - 1] : 0;
    }
    tma_ok &= mbar_wait(&sm.bar[1]);
    if (!tma_ok && lane == 0) atomicExch(&err->error, 2);
    __syncwarp();
    // se(v) lengths of the 7 possible vector components per axis: lane l holds x offset l-3 (l < 8) or y offset l-11 (l >= 8)
    const int mvb = se_len(((lane & 8) ? 4 * fy - ppy : 4 * fx - ppx) + (lane & 7) - 3);
    // DC predictor of the intra estimate (source neighbours): evaluated by the second half-warp in the slot the first one spends on the centre
    uint32_t dcw;
    {
        int sum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) sum = dp4a_us(sm.nb_top[k], 0x01010101u, dp4a_us(sm.nb_left[k], 0x01010101u, sum));
        const int dc = top && left ? (sum + 16) >> 5 : (top || left) ? (sum + 8) >> 4 : 128;
        dcw = 0x01010101u * (uint32_t)dc;
    }
    const uint32_t *pw = sm.plane[0];                   // word view of the planes G, b, h, j (PLW words each)
    constexpr int PLW = PL_ROWS * PL_STRIDE / 4, RW = PL_STRIDE / 4;
    const int ob = by * PL_STRIDE + o1 + bx + 3;         // byte offset inside a plane of sample (bx - 1, by - 1) of the best full-pel block
    B200_CHECK(((ob + 1 + PL_STRIDE) >> 2) + 3 * (PL_STRIDE / 4) + 1 < PL_ROWS * PL_STRIDE / 4 && (ob >> 2) + 4 * (PL_STRIDE / 4) + 1 < PL_ROWS * PL_STRIDE / 4, 6);
    // P_8x8: every candidate's SATD is also summed per 8x8 quadrant (the first two steps of the 16-lane reduction) and each quadrant keeps its
    // own best candidate: key = (SATD8x8 + lambda * bits) << 5 | sequence number (0..8 half-pel ring, 9..16 quarter-pel ring).
    uint32_t bk = 0xffffffffu, bq = 0xffffffffu;
    // one candidate per half-warp: SATD of the lane's block, quadrant and macroblock sums, rate term, the two running minima
    auto slot = [&](const uint32_t P[4], int idx, int ox, int oy, int seq, bool live) -> int {
        int sq = satd_rows(P, Ts);
        sq += __shfl_xor_sync(0xffffffffu, sq, 1); sq += __shfl_xor_sync(0xffffffffu, sq, 2);       // this lane's 8x8 quadrant
        int sat = sq + __shfl_xor_sync(0xffffffffu, sq, 4); sat += __shfl_xor_sync(0xffffffffu, sat, 8);
        const int bits = __shfl_sync(0xffffffffu, mvb, ox + 3) + __shfl_sync(0xffffffffu, mvb, oy + 11);
        if (live) {
            bk = min(bk, ((uint32_t)(sat + lambda * bits) << 4) | (uint32_t)idx);
            bq = min(bq, ((uint32_t)(sq + lambda * bits) << 5) | (uint32_t)seq);
        }
        return sat;
    };
    // ---- half-pel ring + centre: the nine candidates are single planes at integer offsets (-1 | 0): i = 1 j(-1,-1), 2 h(0,-1), 3 j(0,-1),
    // 4 b(-1,0), 5 b(0,0), 6 j(-1,0), 7 h(0,0), 8 j(0,0), 0 G(0,0). A lane fetches the 5 x 5 samples around its block once per plane
    // (rows by-1 .. by+3; A = the four bytes from x-1, B = from x) and takes every candidate of that plane from those registers.
    int ie_dc;
    {
        uint32_t A[5], B[5];
        {   // region 1: b for the first half-warp, j for the second
            const uint32_t *w = pw + (hw ? 3 : 1) * PLW + (ob >> 2); const int sa = (ob & 3) * 8;
#pragma unroll
            for (int r = 0; r < 5; r++) { const uint32_t lo = w[r * RW], hi = w[r * RW + 1]; A[r] = __funnelshift_r(lo, hi, sa); B[r] = __funnelshift_rc(lo, hi, sa + 8); }
        }
        {   // b(-1,0) [4] | j(-1,-1) [1]
            const uint32_t P[4] = { hw ? A[0] : A[1], hw ? A[1] : A[2], hw ? A[2] : A[3], hw ? A[3] : A[4] };
            slot(P, hw ? 1 : 4, -2, hw ? -2 : 0, hw ? 1 : 4, true);
        }
        {   // b(0,0) [5] | j(-1,0) [6]
            const uint32_t P[4] = { hw ? A[1] : B[1], hw ? A[2] : B[2], hw ? A[3] : B[3], hw ? A[4] : B[4] };
            slot(P, hw ? 6 : 5, hw ? -2 : 2, hw ? 2 : 0, hw ? 6 : 5, true);
        }
        {   // region 2: h for the first half-warp, j again for the second; only the bytes from x
            const uint32_t *w = pw + (hw ? 3 : 2) * PLW + ((ob + 1) >> 2); const int sb = ((ob + 1) & 3) * 8;
#pragma unroll
            for (int r = 0; r < 5; r++) B[r] = __funnelshift_r(w[r * RW], w[r * RW + 1], sb);
        }
        {   // h(0,-1) [2] | j(0,-1) [3]
            const uint32_t P[4] = { B[0], B[1], B[2], B[3] };
            slot(P, hw ? 3 : 2, hw ? 2 : 0, -2, hw ? 3 : 2, true);
        }
        {   // h(0,0) [7] | j(0,0) [8]
            const uint32_t P[4] = { B[1], B[2], B[3], B[4] };
            slot(P, hw ? 8 : 7, hw ? 2 : 0, 2, hw ? 8 : 7, true);
        }
        {   // G(0,0) [0] | the DC predictor of the intra estimate
            const uint32_t *w = pw + ((ob + 1 + PL_STRIDE) >> 2); const int sb = ((ob + 1) & 3) * 8;
            uint32_t P[4];
#pragma unroll
            for (int r = 0; r < 4; r++) P[r] = hw ? dcw : __funnelshift_r(w[r * RW], w[r * RW + 1], sb);
            const int sat = slot(P, 0, 0, 0, 0, hw == 0);
            ie_dc = __shfl_sync(0xffffffffu, sat, 16);
        }
    }
    bk = min(bk, __shfl_xor_sync(0xffffffffu, bk, 16)); bq = min(bq, __shfl_xor_sync(0xffffffffu, bq, 16));
    // candidate i: 0 centre, then (-1,-1),(0,-1),(1,-1),(-1,0),(1,0),(-1,1),(0,1),(1,1); packed 2-bit (offset + 1) tables
    const uint32_t OXP = 0x24891u, OYP = 0x2A501u;
    const int hx = 2 * ((int)((OXP >> (2 * (bk & 15))) & 3) - 1), hy = 2 * ((int)((OYP >> (2 * (bk & 15))) & 3) - 1);   // half-pel winner = centre of the quarter-pel ring
    // ---- quarter-pel ring: every candidate is the rounded average of two of the planes' samples (8.4.2.2.1, c_qpel_off)
    bk &= ~15u;                                         // the centre competes as candidate 0 of this ring
#pragma unroll 1
    for (int pass = 0; pass < 4; pass++) {
        const int i = 2 * pass + hw + 1;
        const int ox = hx + (int)((OXP >> (2 * i)) & 3) - 1, oy = hy + (int)((OYP >> (2 * i)) & 3) - 1;
        const uint32_t t = c_qpel_off[(oy & 3) * 4 + (ox & 3)];
        const int common = (by + (oy >> 2) + 1) * PL_STRIDE + o1 + bx + (ox >> 2) + 4;
        const int oa = common + (int)(t & 0xffffu), obb = common + (int)(t >> 16);
        const uint32_t *wa = pw + (oa >> 2), *wb = pw + (obb >> 2); const int sa = (oa & 3) * 8, sb = (obb & 3) * 8;
        B200_CHECK(oa >= 0 && obb >= 0 && (oa >> 2) + 3 * RW + 1 < 4 * PLW && (obb >> 2) + 3 * RW + 1 < 4 * PLW, 7);
        uint32_t P[4];
#pragma unroll
        for (int r = 0; r < 4; r++) P[r] = avg4(__funnelshift_r(wa[r * RW], wa[r * RW + 1], sa), __funnelshift_r(wb[r * RW], wb[r * RW + 1], sb));
        slot(P, i, ox, oy, 8 + i, true);
    }
    bk = min(bk, __shfl_xor_sync(0xffffffffu, bk, 16)); bq = min(bq, __shfl_xor_sync(0xffffffffu, bq, 16));
    const int qx = hx + (int)((OXP >> (2 * (bk & 15))) & 3) - 1, qy = hy + (int)((OYP >> (2 * (bk & 15))) & 3) - 1;     // offset from 4*(fx,fy), quarter-pel units
    const int cost16 = (int)(bk >> 4);
    // this lane's quadrant vector (offset from the full-pel winner) and the P_8x8 cost
    int lx, ly;
    {
        const int sq_ = bq & 31, ci = sq_ < 9 ? sq_ : sq_ - 8, st = sq_ < 9 ? 2 : 1;
        lx = (sq_ < 9 ? 0 : hx) + st * ((int)((OXP >> (2 * ci)) & 3) - 1); ly = (sq_ < 9 ? 0 : hy) + st * ((int)((OYP >> (2 * ci)) & 3) - 1);
    }
    const int c8q = (int)(bq >> 5);
    const int cost8 = __shfl_sync(0xffffffffu, c8q, 0) + __shfl_sync(0xffffffffu, c8q, 4) + __shfl_sync(0xffffffffu, c8q, 8) +
                      __shfl_sync(0xffffffffu, c8q, 12) + lambda * P8X8_BIAS_BITS;
    const bool use8 = !s.no_p8x8 && cost8 < cost16;
    if (!use8) { lx = qx; ly = qy; }
    const int inter_cost = use8 ? cost8 : cost16;

    // intra estimate from source neighbours (staged above): V, H, DC 16x16 by SATD (the DC predictor was evaluated beside the centre candidate)
    int ie = 1 << 30;
    {
        const uint8_t *nl = reinterpret_cast<const uint8_t *>(sm.nb_left);
        uint32_t P[4];
#pragma unroll
        for (int y = 0; y < 4; y++) P[y] = hw == 0 ? sm.nb_top[bx >> 2] : 0x01010101u * nl[by + y];
        const int sat = half_reduce16(satd_rows(P, Ts));
        const int other = __shfl_xor_sync(0xffffffffu, sat, 16);
        const int sv = hw == 0 ? sat : other, sh = hw == 0 ? other : sat;
        if (top) ie = min(ie, sv);
        if (left) ie = min(ie, sh);
        ie = min(ie, ie_dc);
    }
    const bool intra = ie + lambda * 16 < inter_cost;

    MbInfo *mi = s.mbi + mb;
    if (lane == 0) {
        s.me0[mb * 2] = (int16_t)fx; s.me0[mb * 2 + 1] = (int16_t)fy; s.inter_cost[mb] = inter_cost;
    }
    if (intra) {
        if (lane < 12) reinterpret_cast<uint32_t *>(mi)[lane] = lane == 0 ? (uint32_t)MB_I16x16 : 0u;
        return;
    }

    // ---- phase B: code the inter macroblock ----
    // lanes 0-15 own the luma 4x4 blocks, lanes 16-23 the chroma blocks. Prediction and source stay packed (four samples per word) through the
    // forward transform (IDP.4A against the transform's rows) and the reconstruction (16x2 adds and clamps); only the coefficients are scalars.
    MbCoef *co = s.coef + mb;
    const bool is_luma = lane < 16, active = lane < 24;
    const int pl = (lane >> 2) & 1, cb = lane & 3;                 // chroma plane / block of lanes 16-23
    const int cw = wc / 2;
    const int cx0 = mx * 8 + (cb & 1) * 4, cy0 = my * 8 + (cb >> 1) * 4;
    // the vector of this lane's 8x8 partition (lanes 16-31 mirror 0-15)
    const int plx = __shfl_sync(0xffffffffu, lx, lane & 15), ply = __shfl_sync(0xffffffffu, ly, lane & 15);
    const int mvx = 4 * fx + plx, mvy = 4 * fy + ply;
    uint32_t P[4], S[4];
    pred_rows_qpel(sm, o1, bx, by, plx, ply, 0, P);
    {
        // chroma motion compensation, 1/8-pel bilinear (8.4.2.2.2), from the edge-extended reference chroma planes, by all 32 lanes: lane =
        // (plane, row 0..7, half): four samples = two IDP.4A each against the packed weights; chroma block cb lies under luma partition cb
        const int cpl = lane >> 4, crow = (lane >> 1) & 7, chalf = lane & 1, cbk = (crow >> 2) * 2 + chalf;
        const int cvx = 4 * fx + __shfl_sync(0xffffffffu, lx, 4 * cbk), cvy = 4 * fy + __shfl_sync(0xffffffffu, ly, 4 * cbk);
        const int fxc = cvx & 7, fyc = cvy & 7;
        const uint8_t *rp = s.rpc[cpl] + (ptrdiff_t)(my * 8 + crow + (cvy >> 3)) * g.cs + mx * 8 + chalf * 4 + (cvx >> 3);
        const uint32_t *q0 = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(rp) & ~(uintptr_t)3), *q1 = q0 + (g.cs >> 2);
        const int sh = (int)(reinterpret_cast<uintptr_t>(rp) & 3) * 8;
        const uint32_t a_lo = __ldg(q0), a_hi = __ldg(q0 + 1), b_lo = __ldg(q1), b_hi = __ldg(q1 + 1);
        const uint32_t R0 = __funnelshift_r(a_lo, a_hi, sh), S0 = __funnelshift_rc(a_lo, a_hi, sh + 8);     // samples x .. x+3 and x+1 .. x+4 of the upper row
        const uint32_t R1 = __funnelshift_r(b_lo, b_hi, sh), S1 = __funnelshift_rc(b_lo, b_hi, sh + 8);     // ... of the lower row
        const uint32_t WA = (uint32_t)((8 - fxc) * (8 - fyc)) | ((uint32_t)(fxc * (8 - fyc)) << 8), WB = (uint32_t)((8 - fxc) * fyc) | ((uint32_t)(fxc * fyc) << 8);
        const uint32_t TA0 = __byte_perm(R0, S0, 0x5140), TA1 = __byte_perm(R0, S0, 0x7362), TB0 = __byte_perm(R1, S1, 0x5140), TB1 = __byte_perm(R1, S1, 0x7362);
        const int p0 = dp4a_us(TB0, WB, dp4a_us(TA0, WA, 32)) >> 6, p1 = dp4a_us(TB0, WB << 16, dp4a_us(TA0, WA << 16, 32)) >> 6;
        const int p2 = dp4a_us(TB1, WB, dp4a_us(TA1, WA, 32)) >> 6, p3 = dp4a_us(TB1, WB << 16, dp4a_us(TA1, WA << 16, 32)) >> 6;
        const uint32_t cword = (uint32_t)p0 | ((uint32_t)p1 << 8) | ((uint32_t)p2 << 16) | ((uint32_t)p3 << 24);
        // the chroma block lanes collect their four rows; source rows: luma from the staged macroblock, chroma from the source planes
        const int srcl = pl * 16 + (cb >> 1) * 8 + (cb & 1);
        const uint8_t *sp_c = s.src[1 + pl] + (size_t)cy0 * cw + cx0;
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const uint32_t gw = __shfl_sync(0xffffffffu, cword, srcl + 2 * y);
            if (!is_luma) { P[y] = gw; S[y] = *reinterpret_cast<const uint32_t *>(sp_c + (size_t)y * cw); }
            else S[y] = sm.src[(by + y) * 4 + (bx >> 2)];
        }
    }
    // forward core transform of (S - P): rows as IDP.4A of the packed samples against the transform rows {1,1,1,1}, {2,1,-1,-2}, {1,-1,-1,1}, {1,-2,2,-1}
    // (the prediction against their negation), then the column butterflies
    int c[16];
    {
        const uint32_t BK[4] = { 0x01010101u, 0xFEFF0102u, 0x01FFFF01u, 0xFF02FE01u }, NK[4] = { 0xFFFFFFFFu, 0x0201FFFEu, 0xFF0101FFu, 0x01FE02FFu };
#pragma unroll
        for (int y = 0; y < 4; y++)
#pragma unroll
            for (int k = 0; k < 4; k++) c[y * 4 + k] = dp4a_us(S[y], BK[k], dp4a_us(P[y], NK[k], 0));
#pragma unroll
        for (int x = 0; x < 4; x++) {
            const int a0 = c[x] + c[12 + x], a1 = c[4 + x] + c[8 + x], a2 = c[4 + x] - c[8 + x], a3 = c[x] - c[12 + x];
            c[x] = a0 + a1; c[4 + x] = 2 * a3 + a2; c[8 + x] = a0 - a1; c[12 + x] = a3 - 2 * a2;
        }
    }
    const QParam q = make_qparam(is_luma ? qp : (int)c_chroma_qp[qp]);
    // 2x2 Hadamard of the four DC terms of this lane's chroma plane (lanes 16+4pl .. 19+4pl)
    int hd[4];
    {
        int dcs[4];
#pragma unroll
        for (int k = 0; k < 4; k++) dcs[k] = __shfl_sync(0xffffffffu, c[0], 16 + pl * 4 + k);
        hd[0] = dcs[0] + dcs[1] + dcs[2] + dcs[3]; hd[1] = dcs[0] - dcs[1] + dcs[2] - dcs[3]; hd[2] = dcs[0] + dcs[1] - dcs[2] - dcs[3]; hd[3] = dcs[0] - dcs[1] - dcs[2] + dcs[3];
    }
    // Does anything of this macroblock quantise to a nonzero level? The quantiser is monotone in |coefficient|, so the largest magnitude per
    // multiplier class decides (position 0 of a chroma block is coded through the DC Hadamard). Most macroblocks of a P picture at the
    // session bitrates answer no: their reconstruction is the prediction and no level leaves the warp.
    bool nzl;
    {
        const int m0 = max(max(is_luma ? abs(c[0]) : 0, abs(c[2])), max(abs(c[8]), abs(c[10])));
        const int m1 = max(max(abs(c[5]), abs(c[7])), max(abs(c[13]), abs(c[15])));
        const int m2 = max(max(max(abs(c[1]), abs(c[3])), max(abs(c[4]), abs(c[6]))), max(max(abs(c[9]), abs(c[11])), max(abs(c[12]), abs(c[14]))));
        unsigned t = (((unsigned)m0 * (unsigned)q.mf[0] + (unsigned)q.f_inter) | ((unsigned)m1 * (unsigned)q.mf[1] + (unsigned)q.f_inter) |
                      ((unsigned)m2 * (unsigned)q.mf[2] + (unsigned)q.f_inter)) >> q.qbits;
        const int md = max(max(abs(hd[0]), abs(hd[1])), max(abs(hd[2]), abs(hd[3])));
        if (!is_luma) t |= ((unsigned)md * (unsigned)q.mf[0] + 2u * (unsigned)q.f_inter) >> (q.qbits + 1);
        nzl = t != 0u;
    }
    uint8_t *recp = is_luma ? s.rec[0] + (size_t)(y0 + by) * wc + x0 + bx : s.rec[1 + pl] + (size_t)cy0 * cw + cx0;
    const int rst = is_luma ? wc : cw;
    const uint32_t mvw = (uint32_t)(uint16_t)mvx | ((uint32_t)(uint16_t)mvy << 16);
    const uint32_t mvq = __shfl_sync(0xffffffffu, mvw, lane == 0 ? 0 : 4 * ((lane - 1) & 3));
    if (__ballot_sync(0xffffffffu, active && nzl) == 0u) {
        if (active) {
#pragma unroll
            for (int y = 0; y < 4; y++) *reinterpret_cast<uint32_t *>(recp + (size_t)y * rst) = P[y];
        }
        // nobody reads the levels of a macroblock whose cbp is 0 (CAVLC / CABAC / the 8x8 pass all go by cbp); the stage dumps of the parity tests do
        if (s.dump) { uint4 *cz = reinterpret_cast<uint4 *>(co); cz[lane] = make_uint4(0, 0, 0, 0); if (lane < 51 - 32) cz[32 + lane] = make_uint4(0, 0, 0, 0); }
        // word 0: type, cbp 0; word 1: vector of partition 0; words 2-5: the four partition vectors; words 6-11: nnz 0
        uint32_t *miw = reinterpret_cast<uint32_t *>(mi);
        if (lane == 0) { miw[0] = use8 ? (uint32_t)MB_P8x8 : (uint32_t)MB_P16x16; miw[1] = mvq; }
        else if (lane < 5) miw[1 + lane] = mvq;
        else if (lane < 11) miw[1 + lane] = 0u;
        return;
    }
    int nnz = 0; bool dc_nz = false;
    int dcC = 0;
    if (!is_luma && active) {
        // chroma DC: quantise, and the normative inverse (8.5.11)
        int lv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { lv[k] = quant_dc(hd[k], q, q.f_inter); dc_nz |= lv[k] != 0; }
        const int fi[4] = { lv[0] + lv[1] + lv[2] + lv[3], lv[0] - lv[1] + lv[2] - lv[3], lv[0] + lv[1] - lv[2] - lv[3], lv[0] - lv[1] - lv[2] + lv[3] };
        const int f = cb == 0 ? fi[0] : cb == 1 ? fi[1] : cb == 2 ? fi[2] : fi[3];
        dcC = ((f * 16 * q.v[0]) << q.sh) >> 5;
        if (cb == 0) *reinterpret_cast<uint2 *>(co->chroma_dc[pl]) = make_uint2((uint32_t)(uint16_t)lv[0] | ((uint32_t)(uint16_t)lv[1] << 16),
                                                                              (uint32_t)(uint16_t)lv[2] | ((uint32_t)(uint16_t)lv[3] << 16));
    }
    {
        __align__(16) int16_t lz[16];
        nnz = quant_dequant4x4(c, lz, q, q.f_inter, !is_luma);
        if (!is_luma) c[0] = dcC;
        idct4x4(c);
        if (active) {
            uint4 *dst = reinterpret_cast<uint4 *>(is_luma ? co->luma[b] : co->chroma_ac[pl][cb]);
            dst[0] = reinterpret_cast<uint4 *>(lz)[0]; dst[1] = reinterpret_cast<uint4 *>(lz)[1];
            // reconstruction on 16x2 lanes: prediction bytes widened, residual pairs packed, one add and one clamp to [0, 255] per pair
#pragma unroll
            for (int y = 0; y < 4; y++) {
                const uint32_t lo = __vimin_s16x2_relu(__vadd2(__byte_perm(P[y], 0u, 0x4140), __byte_perm((uint32_t)c[y * 4], (uint32_t)c[y * 4 + 1], 0x5410)), 0x00ff00ffu);
                const uint32_t hi = __vimin_s16x2_relu(__vadd2(__byte_perm(P[y], 0u, 0x4342), __byte_perm((uint32_t)c[y * 4 + 2], (uint32_t)c[y * 4 + 3], 0x5410)), 0x00ff00ffu);
                *reinterpret_cast<uint32_t *>(recp + (size_t)y * rst) = __byte_perm(lo, hi, 0x6420);
            }
        } else {
            nnz = 0;
            if (lane < 26) reinterpret_cast<uint4 *>(co->luma_dc)[lane - 24] = make_uint4(0, 0, 0, 0);
        }
    }
    const uint32_t nzmask = __ballot_sync(0xffffffffu, nnz != 0), dcmask = __ballot_sync(0xffffffffu, dc_nz);
    int cbp = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) if ((nzmask >> (4 * k)) & 15) cbp |= 1 << k;
    cbp |= ((nzmask >> 16) & 255) ? 32 : ((dcmask ? 16 : 0));
    if (lane < 24) mi->nnz[lane] = (uint8_t)nnz;
    // word 0: type, cbp; word 1: vector of partition 0 (= the 16x16 vector); words 2-5: the four partition vectors
    if (lane == 0) {
        reinterpret_cast<uint32_t *>(mi)[0] = (use8 ? (uint32_t)MB_P8x8 : (uint32_t)MB_P16x16) | ((uint32_t)cbp << 24);
        reinterpret_cast<uint32_t *>(mi)[1] = mvq;
    } else if (lane < 5) reinterpret_cast<uint32_t *>(mi)[1 + lane] = mvq;
}

// Scene change (phase A'): the wrapper asks openh264 for bEnableSceneChangeDetect (video_codec/VideoEncoderOpenH264.cpp:283). Here
// a P picture whose macroblocks came out >= 2/5 intra from the motion search (normal P pictures: a few percent) is coded as an IDR instead: the session's descriptor
// is rewritten on the device (is_idr, frame_num), every MB is marked intra, and the host reads the kind back with the size.
// grid: (sessions), 256 threads
__global__ void __launch_bounds__(256) k_scene_change(Sess *ss, Geom g)
{
    Sess &s = ss[blockIdx.x];
    if (s.is_idr || !s.scene_change) return;
    __shared__ int cnt_s;
    if (threadIdx.x == 0) cnt_s = 0;
    __syncthreads();
    const int nmb = g.mbw * g.mbh;
    int c = 0;
    for (int mb = threadIdx.x; mb < nmb; mb += 256) c += s.mbi[mb].mb_type == MB_I16x16;
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&cnt_s, c);
    __syncthreads();
    if (5 * cnt_s < 2 * nmb) return;
    for (int i = threadIdx.x; i < nmb * 12; i += 256) reinterpret_cast<uint32_t *>(s.mbi)[i] = (i % 12) == 0 ? (uint32_t)MB_I16x16 : 0u;
    if (threadIdx.x == 0) { s.is_idr = 1; s.frame_num = 0; }
}

} // namespace b200
