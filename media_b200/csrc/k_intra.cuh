// media_b200/csrc/k_intra.cuh -- Intra_16x16 macroblocks on a macroblock wavefront (phase C of DESIGN.md 3).
//
// Role inside the reference: intra prediction, mode decision, transform and reconstruction inside
// ISVCEncoder::EncodeFrame (video_codec/VideoEncoderOpenH264.cpp:344; openh264's WelsMdI16x16,
// WelsI16x16LumaPred*, WelsIChromaPred*, WelsHadamardT4Dc, WelsDequantIHadamard4x4 in the absent libopenh264).
// One warp owns one macroblock row of one session; rows are handed out through an atomic ticket so that a
// warp only ever waits on a row that started earlier. Lanes 0-15 own the 16 luma 4x4 blocks, lanes 16-23 the
// 8 chroma blocks; luma and chroma predictors share one code path parameterised per lane.
#pragma once
#include "h264_dev.cuh"

namespace b200 {

#define WAVE_WARPS 4
#define WAVE_TIMEOUT_NS 2000000000ull   /* watchdog: a wait longer than 2 s is an internal error, never a hang */

struct WaveCtl { int ticket_intra, ticket_dbk, error, pad; };

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

struct IntraSmem { uint8_t top[3][20], left[3][20]; };   // index 0 = the corner sample p[-1,-1]; [comp]

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

__device__ void intra_code_mb(const Sess &s, const Geom &g, IntraSmem &sm, int mx, int my, int lane)
{
    const int wc = g.wc, cw = wc / 2, mb = my * g.mbw + mx, qp = s.qp;
    const bool top = !row_is_slice_top(g, my), left = mx > 0;
    // neighbours from the (pre-deblock) reconstruction; written by other warps / kernels -> L2 loads
    __syncwarp();
    for (int i = lane; i < 17 + 16 + 2 * (9 + 8); i += 32) {
        int comp, idx, is_top;
        if (i < 33) { comp = 0; is_top = i < 17; idx = is_top ? i : i - 17; }
        else { int j = i - 33; comp = 1 + j / 17; j %= 17; is_top = j < 9; idx = is_top ? j : j - 9; }
        const int n = comp ? 8 : 16, st = comp ? cw : wc, px0 = mx * n, py0 = my * n;
        const uint8_t *r = s.rec[comp];
        if (is_top) {     // idx 0 = corner, 1..n = row above
            int v = 0;
            if (top && (idx > 0 || left)) v = __ldcg(r + (size_t)(py0 - 1) * st + px0 + idx - 1);
            sm.top[comp][idx] = (uint8_t)v;
            if (idx == 0) sm.left[comp][0] = (uint8_t)v;
        } else {
            sm.left[comp][idx + 1] = left ? __ldcg(r + (size_t)(py0 + idx) * st + px0 - 1) : 0;
        }
    }
    __syncwarp();
    const bool is_luma = lane < 16, active = lane < 24;
    const int comp = is_luma ? 0 : (lane < 20 ? 1 : 2);
    const int cb = lane & 3, b = lane & 15;
    const int bx = is_luma ? blk_x(b) * 4 : (cb & 1) * 4, by = is_luma ? blk_y(b) * 4 : (cb >> 1) * 4;
    const int n = is_luma ? 16 : 8, half = n / 2, st = is_luma ? wc : cw;
    const uint8_t *T = sm.top[comp] + 1, *L = sm.left[comp] + 1;
    // source block
    int sp[16];
    {
        const uint8_t *spx = s.src[comp] + (size_t)(my * n + by) * st + mx * n + bx;
#pragma unroll
        for (int y = 0; y < 4; y++) {
            uint32_t w = *reinterpret_cast<const uint32_t *>(spx + (size_t)y * st);
#pragma unroll
            for (int x = 0; x < 4; x++) sp[y * 4 + x] = (w >> (8 * x)) & 255;
        }
    }
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

    int p[16], c[16]; intra_pred_block(kind, T, L, bx, by, dcv, pa, pb, pc, off, p);
#pragma unroll
    for (int k = 0; k < 16; k++) c[k] = sp[k] - p[k];
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
        co->luma_dc[inv_zz[ps]] = (int16_t)lev;
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
    if (active) {
        idct4x4(c);
        uint4 *dst = reinterpret_cast<uint4 *>(is_luma ? co->luma[b] : co->chroma_ac[comp - 1][cb]);
        dst[0] = reinterpret_cast<uint4 *>(lz)[0]; dst[1] = reinterpret_cast<uint4 *>(lz)[1];
        uint8_t *rp = s.rec[comp] + (size_t)(my * n + by) * st + mx * n + bx;
#pragma unroll
        for (int y = 0; y < 4; y++) {
            uint32_t w = 0;
#pragma unroll
            for (int x = 0; x < 4; x++) w |= (uint32_t)clip255(p[y * 4 + x] + c[y * 4 + x]) << (8 * x);
            *reinterpret_cast<uint32_t *>(rp + (size_t)y * st) = w;
        }
        mi->nnz[lane] = (uint8_t)nnz;
    }
    const uint32_t nzmask = __ballot_sync(0xffffffffu, nnz != 0), dcmask = __ballot_sync(0xffffffffu, dc_nz);
    if (lane == 0) {
        int cbp = (nzmask & 0xffff) ? 15 : 0;
        cbp |= ((nzmask >> 16) & 255) ? 32 : (dcmask ? 16 : 0);
        reinterpret_cast<uint32_t *>(mi)[0] = (uint32_t)MB_I16x16 | ((uint32_t)mode_id << 8) | ((uint32_t)chroma_mode << 16) | ((uint32_t)cbp << 24);
    } else if (lane < 6) reinterpret_cast<uint32_t *>(mi)[lane] = 0;
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
    const Sess &s = ss[t % nsess];
    IntraSmem &sm = sm_all[warp];
    int *prog = s.row_prog_intra;
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
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release(prog + my, nx + 1);
        mx = nx + 1;
    }
    __syncwarp();
    if (lane == 0) { __threadfence(); st_release(prog + my, g.mbw); }
}

} // namespace b200
