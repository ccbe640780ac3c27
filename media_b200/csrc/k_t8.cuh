// media_b200/csrc/k_t8.cuh -- the 8x8 transform of the High profile (transform_size_8x8_flag = 1) for inter macroblocks.
//
// Role inside the reference: the wrapper's profile property accepts "high" (video_codec/VideoEncoderOpenH264.cpp:186-188,248-253);
// openh264's 8x8 residual path lives in the absent libopenh264. Specification: DESIGN.md 3.3 / oracle/orc_encoder.c
// code_inter_mb(): after k_me_fine has coded an inter macroblock with the 4x4 transform, its luma residual is coded again with
// the 8x8 transform (normative inverse: 8.5.13) and the macroblock keeps the cheaper of the two by J = 64 SSD + 27 lambda^2 B.
// One warp per macroblock; 8 lanes own one 8x8 block, one ROW of it each; the separable transforms alternate between row and
// column ownership through a conflict-free shared-memory transpose.
#pragma once
#include "h264_dev.cuh"
#include "k_me.cuh"

namespace b200 {

// 8x8 zig-zag scan (frame, Figure 8-9): scan index -> raster position, and its inverse
static __device__ __constant__ uint8_t c_zigzag8[64] = {
     0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
static __device__ __constant__ uint8_t c_izigzag8[64] = {
     0,  1,  5,  6, 14, 15, 27, 28,  2,  4,  7, 13, 16, 26, 29, 42,  3,  8, 12, 17, 25, 30, 41, 43,  9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63 };
// normAdjust8x8 v(m, class) of 8.5.9 and the matching forward multipliers (JM quant_coef8, qbits = 16 + qp / 6)
static __device__ __constant__ uint8_t c_dequant8_v[6][6] = {
    { 20, 18, 32, 19, 25, 24 }, { 22, 19, 35, 21, 28, 26 }, { 26, 23, 42, 24, 33, 31 },
    { 28, 25, 45, 26, 35, 33 }, { 32, 28, 51, 30, 40, 38 }, { 36, 32, 58, 34, 46, 43 } };
static __device__ __constant__ uint16_t c_quant8_mf[6][6] = {
    { 13107, 11428, 20972, 12222, 16777, 15481 }, { 11916, 10826, 19174, 11058, 14980, 14290 }, { 10082, 8943, 15978, 9675, 12710, 11985 },
    {  9362,  8228, 14913,  8931, 11984, 11259 }, {  8192,  7346, 13159,  7740, 10486,  9777 }, {  7282, 6428, 11570, 6830,  9118,  8640 } };
__device__ __forceinline__ int pos_class8(int i, int j)
{
    if (!(i & 3) && !(j & 3)) return 0;
    if ((i & 1) && (j & 1)) return 1;
    if ((i & 3) == 2 && (j & 3) == 2) return 2;
    if ((!(i & 3) && (j & 1)) || ((i & 1) && !(j & 3))) return 3;
    if ((!(i & 3) && (j & 3) == 2) || ((i & 3) == 2 && !(j & 3))) return 4;
    return 5;
}

// forward 8-point transform (encoder side; rows first, then columns) and the inverse of 8.5.13, in place
__device__ __forceinline__ void fdct8_1d(int x[8])
{
    const int a0 = x[0] + x[7], a1 = x[1] + x[6], a2 = x[2] + x[5], a3 = x[3] + x[4];
    const int a4 = x[0] - x[7], a5 = x[1] - x[6], a6 = x[2] - x[5], a7 = x[3] - x[4];
    const int b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2;
    const int b4 = a5 + a6 + ((a4 >> 1) + a4), b5 = a4 - a7 - ((a6 >> 1) + a6), b6 = a4 + a7 - ((a5 >> 1) + a5), b7 = a5 - a6 + ((a7 >> 1) + a7);
    x[0] = b0 + b1; x[1] = b4 + (b7 >> 2); x[2] = b2 + (b3 >> 1); x[3] = b5 + (b6 >> 2);
    x[4] = b0 - b1; x[5] = b6 - (b5 >> 2); x[6] = (b2 >> 1) - b3; x[7] = (b4 >> 2) - b7;
}
__device__ __forceinline__ void idct8_1d(int d[8])
{
    const int a0 = d[0] + d[4], a1 = -d[3] + d[5] - d[7] - (d[7] >> 1), a2 = d[0] - d[4], a3 = d[1] + d[7] - d[3] - (d[3] >> 1);
    const int a4 = (d[2] >> 1) - d[6], a5 = -d[1] + d[7] + d[5] + (d[5] >> 1), a6 = d[2] + (d[6] >> 1), a7 = d[3] + d[5] + d[1] + (d[1] >> 1);
    const int b0 = a0 + a6, b1 = a1 + (a7 >> 2), b2 = a2 + a4, b3 = a3 + (a5 >> 2), b4 = a2 - a4, b5 = (a3 >> 2) - a5, b6 = a0 - a6, b7 = a7 - (a1 >> 2);
    d[0] = b0 + b7; d[1] = b2 + b5; d[2] = b4 + b3; d[3] = b6 + b1; d[4] = b6 - b1; d[5] = b4 - b3; d[6] = b2 - b5; d[7] = b0 - b7;
}

// per warp: four 8x8 tiles with a row pitch of 9 words (the 32 lanes of a transpose step hit 32 different banks)
struct T8Smem { int t[4][8][9]; };
// lane r of a block's 8 lanes holds v[k] = X(r, k); afterwards it holds X(k, r). All 32 lanes of the warp must call it together.
__device__ __forceinline__ void t8_xpose(int (*t)[9], int r, int v[8])
{
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) t[r][k] = v[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = t[k][r];
}
// One 8x8 block by its 8 lanes. in: res[] = row r of the residual. out: lv[k] = level at (row k, column r), rr[] = row r of the
// reconstructed residual ((x + 32) >> 6 applied). dz = 3 (intra) or 6 (inter): dead zone f = 2^qbits / dz.
__device__ __forceinline__ void t8_code_block(int (*t)[9], int r, int qp, int dz, const int res[8], int lv[8], int rr[8])
{
    int v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = res[k];
    fdct8_1d(v); t8_xpose(t, r, v); fdct8_1d(v);                  // now v[k] = coefficient (k, r)
    const int m = qp % 6, sh = qp / 6, qbits = 16 + sh; const unsigned f = (1u << qbits) / (unsigned)dz;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int cls = pos_class8(k, r);
        int l = min((int)(((unsigned)abs(v[k]) * (unsigned)c_quant8_mf[m][cls] + f) >> qbits), B200_MAX_LEVEL);
        l = v[k] < 0 ? -l : l;
        lv[k] = l;
        const int ls = 16 * c_dequant8_v[m][cls];                 // 8.5.13, flat scaling list
        v[k] = qp >= 36 ? (l * ls) << (sh - 6) : (l * ls + (1 << (5 - sh))) >> (6 - sh);
    }
    t8_xpose(t, r, v); idct8_1d(v);                               // rows first (8.5.13), then columns
    t8_xpose(t, r, v); idct8_1d(v);
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (v[k] + 32) >> 6;
    t8_xpose(t, r, v);
#pragma unroll
    for (int k = 0; k < 8; k++) rr[k] = v[k];
}
// rate estimate in half bits (oracle level_cost2): 2 * (3 + min(|level|, 16)) per nonzero level, 1 per zero before the last
// nonzero one. The caller passes per-lane partial sums; this folds them over the `width` lanes that share a block.
struct T8Cost { int c, nz, last1; };
__device__ __forceinline__ void t8_cost_add(T8Cost &a, int level, int scan_idx)
{
    const int v = abs(level);
    if (v) { a.c += 2 * (3 + min(v, 16)); a.nz++; a.last1 = max(a.last1, scan_idx + 1); }
}
template <int WIDTH> __device__ __forceinline__ int t8_cost_fold(T8Cost a)
{
#pragma unroll
    for (int o = 1; o < WIDTH; o <<= 1) {
        a.c += __shfl_xor_sync(0xffffffffu, a.c, o); a.nz += __shfl_xor_sync(0xffffffffu, a.nz, o);
        a.last1 = max(a.last1, __shfl_xor_sync(0xffffffffu, a.last1, o));
    }
    return a.c + a.last1 - a.nz;
}

#define T8_WARPS 8
// grid: (ceil(n_mb / T8_WARPS), 1, sessions); runs after k_me_fine / k_scene_change, before the intra wavefront
__global__ void __launch_bounds__(T8_WARPS * 32) k_inter_t8(const Sess *ss, Geom g)
{
    __shared__ T8Smem sm_all[T8_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * T8_WARPS + warp;
    if (mb >= g.mbw * g.mbh) return;
    const Sess &s = ss[blockIdx.z];
    if (s.is_idr || !s.t8x8) return;
    MbInfo *mi = s.mbi + mb;
    const uint32_t w0 = reinterpret_cast<const uint32_t *>(mi)[0];
    const int type = w0 & 255, cbp = (int)(w0 >> 24);
    if ((type != MB_P16x16 && type != MB_P8x8) || !(cbp & 15)) return;        // early-skip and all-zero MBs keep the flag 0 (not coded)
    int mx, my; mb_xy(g, mb, mx, my);
    const int x0 = mx * 16, y0 = my * 16, wc = g.wc, qp = s.qp;
    const int b8 = lane >> 3, r = lane & 7, ox = (b8 & 1) * 8, oy = (b8 >> 1) * 8;
    int (*t)[9] = sm_all[warp].t[b8];
    MbCoef *co = s.coef + mb;

    // the 4x4 coding k_me_fine left behind: rate of its levels (lane = half a 4x4 block) and its distortion (lane = 8 samples)
    int rate4;
    {
        const uint4 q = reinterpret_cast<const uint4 *>(co->luma)[lane];
        const uint32_t w[4] = { q.x, q.y, q.z, q.w };
        T8Cost a = { 0, 0, 0 };
#pragma unroll
        for (int i = 0; i < 8; i++) t8_cost_add(a, (int)(int16_t)(w[i >> 1] >> (16 * (i & 1))), (lane & 1) * 8 + i);
        int cb = t8_cost_fold<2>(a) + 1;                                       // + the block's coded_block_flag
        if ((lane & 1) || !((cbp >> (lane >> 3)) & 1)) cb = 0;                 // blocks of an 8x8 group without its cbp bit are not coded
#pragma unroll
        for (int o = 16; o; o >>= 1) cb += __shfl_xor_sync(0xffffffffu, cb, o);
        rate4 = cb;
    }
    const size_t row_off = (size_t)(y0 + oy + r) * wc + x0 + ox;
    const uint2 sw = *reinterpret_cast<const uint2 *>(s.src[0] + row_off), r4 = *reinterpret_cast<const uint2 *>(s.rec[0] + row_off);
    int src[8], pred[8], res[8];
#pragma unroll
    for (int k = 0; k < 8; k++) src[k] = (int)(((k < 4 ? sw.x : sw.y) >> (8 * (k & 3))) & 255);
    int d4 = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { const int d = src[k] - (int)(((k < 4 ? r4.x : r4.y) >> (8 * (k & 3))) & 255); d4 += d * d; }

    // prediction of this lane's row from the padded reference planes G, b, h, j (Table 8-12: the average of two plane samples)
    {
        const int mvx = mi->mv8[b8][0], mvy = mi->mv8[b8][1];
        const uint8_t *e = c_qpel_tab[(mvy & 3) * 4 + (mvx & 3)];
        const int xi = x0 + ox + (mvx >> 2), yi = y0 + oy + r + (mvy >> 2);
        const uint8_t *pa = s.rpl[e[0]] + (ptrdiff_t)(yi + e[2]) * g.ls + xi + e[1], *pb = s.rpl[e[3]] + (ptrdiff_t)(yi + e[5]) * g.ls + xi + e[4];
#pragma unroll
        for (int k = 0; k < 8; k++) { pred[k] = ((int)__ldg(pa + k) + (int)__ldg(pb + k) + 1) >> 1; res[k] = src[k] - pred[k]; }
    }
    int lv[8], rr[8];
    t8_code_block(t, r, qp, 6, res, lv, rr);
    int rec[8], d8 = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { rec[k] = clip255(pred[k] + rr[k]); const int d = src[k] - rec[k]; d8 += d * d; }
    T8Cost a = { 0, 0, 0 };
#pragma unroll
    for (int k = 0; k < 8; k++) t8_cost_add(a, lv[k], c_izigzag8[k * 8 + r]);
    int n8 = a.nz;                                                             // nonzero levels of this lane's 8x8 block
    n8 += __shfl_xor_sync(0xffffffffu, n8, 1); n8 += __shfl_xor_sync(0xffffffffu, n8, 2); n8 += __shfl_xor_sync(0xffffffffu, n8, 4);
    int rate8 = t8_cost_fold<8>(a);
    if (r) rate8 = 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) { rate8 += __shfl_xor_sync(0xffffffffu, rate8, o); d4 += __shfl_xor_sync(0xffffffffu, d4, o); d8 += __shfl_xor_sync(0xffffffffu, d8, o); }
    const uint32_t nzmask = __ballot_sync(0xffffffffu, n8 != 0);
    const long long l2 = 27ll * c_lambda[qp] * c_lambda[qp];
    if (!nzmask || !(64ll * d8 + l2 * (rate8 + 4) < 64ll * d4 + l2 * rate4)) return;      // keep the 4x4 coding

    int16_t *dst = &co->luma[4 * b8][0];
#pragma unroll
    for (int k = 0; k < 8; k++) dst[c_izigzag8[k * 8 + r]] = (int16_t)lv[k];
    uint2 ow;
    ow.x = (uint32_t)rec[0] | ((uint32_t)rec[1] << 8) | ((uint32_t)rec[2] << 16) | ((uint32_t)rec[3] << 24);
    ow.y = (uint32_t)rec[4] | ((uint32_t)rec[5] << 8) | ((uint32_t)rec[6] << 16) | ((uint32_t)rec[7] << 24);
    *reinterpret_cast<uint2 *>(s.rec[0] + row_off) = ow;
    if (r < 4) mi->nnz[4 * b8 + r] = (uint8_t)n8;
    if (lane == 0) {
        const int cl = ((nzmask & 0xffu) ? 1 : 0) | ((nzmask & 0xff00u) ? 2 : 0) | ((nzmask & 0xff0000u) ? 4 : 0) | ((nzmask & 0xff000000u) ? 8 : 0);
        reinterpret_cast<uint32_t *>(mi)[0] = (w0 & 0x00ff00ffu) | (4u << 8) | ((uint32_t)((cbp & 0x30) | cl) << 24);   // bit 2 of i16_mode = transform_size_8x8_flag
    }
}

// test kernel: n 8x8 residual blocks (raster) -> levels in 8x8 zig-zag order + reconstructed residual; four blocks per warp
__global__ void __launch_bounds__(T8_WARPS * 32) k_test_transform8(const int16_t *res_in, int n, int qp, int intra, int16_t *levels, int *recon)
{
    __shared__ T8Smem sm_all[T8_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b8 = lane >> 3, r = lane & 7;
    const int blk = (blockIdx.x * T8_WARPS + warp) * 4 + b8;
    const bool act = blk < n;                                      // whole warps stay together for the transposes
    int res[8], lv[8], rr[8];
#pragma unroll
    for (int k = 0; k < 8; k++) res[k] = act ? res_in[(size_t)blk * 64 + r * 8 + k] : 0;
    t8_code_block(sm_all[warp].t[b8], r, qp, intra ? 3 : 6, res, lv, rr);
    if (!act) return;
#pragma unroll
    for (int k = 0; k < 8; k++) { levels[(size_t)blk * 64 + c_izigzag8[k * 8 + r]] = (int16_t)lv[k]; recon[(size_t)blk * 64 + r * 8 + k] = rr[k]; }
}

} // namespace b200
