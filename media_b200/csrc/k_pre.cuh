// media_b200/csrc/k_pre.cuh -- input stage: I420 / NV12 / RGBA -> coded-size I420 planes, and the 2x2 pyramid.
//
// The reference hands tightly packed I420 with stride = width (video_codec/VideoEncoderOpenH264.cpp:354-365);
// RGBA and NV12 inputs are an extension of this sibling (BASELINE.json config 3). All three
// kernels are pure HBM streams: one pass over the input, one over the output (algorithmic bytes per luma pixel:
// I420 3.0, NV12 3.0, RGBA 5.5; pyramid 1.3125).
#pragma once
#include "h264_dev.cuh"

namespace b200 {

enum { FMT_I420 = 0, FMT_NV12 = 1, FMT_RGBA = 2 };

// grid: (ceil(units/256), 1, sessions); a unit is 16 output bytes of one plane row (8 at the end of a chroma row whose coded width is 8 mod 16):
// one 128-bit load and one 128-bit store per thread wherever the input row and the address allow it.
#define INGEST_UNITS(wc, hc) (((wc) / 16) * (hc) + 2 * (((wc) / 2 + 15) / 16) * ((hc) / 2))
__global__ void __launch_bounds__(256) k_ingest_planar(const Sess *ss, Geom g)
{
    const Sess &s = ss[blockIdx.z];
    const int w = g.width, h = g.height, wc = g.wc, hc = g.hc;
    const int ly = (wc / 16) * hc, upc = (wc / 2 + 15) / 16, lc = upc * (hc / 2);
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= ly + 2 * lc) return;
    const int comp = u < ly ? 0 : (u < ly + lc ? 1 : 2);
    if (comp) u -= ly + (comp - 1) * lc;
    const int cw = comp ? wc / 2 : wc, pw = comp ? w / 2 : w, ph = comp ? h / 2 : h;
    const int upr = comp ? upc : wc / 16, y = u / upr, x = (u - y * upr) * 16, sy = min(y, ph - 1);
    uint32_t v[4];
    if (s.input_format == FMT_I420 || comp == 0) {
        const uint8_t *in = s.input + (comp == 0 ? 0 : (size_t)w * h + (comp == 2 ? (size_t)pw * ph : 0)) + (size_t)sy * pw;
        if (x + 16 <= pw && (reinterpret_cast<uintptr_t>(in + x) & 15) == 0) {
            const uint4 q = *reinterpret_cast<const uint4 *>(in + x);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t a = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) a |= (uint32_t)in[min(x + 4 * k + i, pw - 1)] << (8 * i);
                v[k] = a;
            }
        }
    } else {   // NV12 chroma: de-interleave 32 bytes of UV pairs
        const uint8_t *in = s.input + (size_t)w * h + (size_t)sy * w + (comp - 1);
        if (x + 16 <= pw && (reinterpret_cast<uintptr_t>(in - (comp - 1) + 2 * x) & 15) == 0) {
            const uint4 q0 = *reinterpret_cast<const uint4 *>(in - (comp - 1) + 2 * x), q1 = *reinterpret_cast<const uint4 *>(in - (comp - 1) + 2 * x + 16);
            const uint32_t sel = comp == 1 ? 0x6420 : 0x7531;
            v[0] = __byte_perm(q0.x, q0.y, sel); v[1] = __byte_perm(q0.z, q0.w, sel); v[2] = __byte_perm(q1.x, q1.y, sel); v[3] = __byte_perm(q1.z, q1.w, sel);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t a = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) a |= (uint32_t)in[2 * min(x + 4 * k + i, pw - 1)] << (8 * i);
                v[k] = a;
            }
        }
    }
    uint8_t *out = s.src[comp] + (size_t)y * cw + x;
    if (x + 16 <= cw && (reinterpret_cast<uintptr_t>(out) & 15) == 0) *reinterpret_cast<uint4 *>(out) = make_uint4(v[0], v[1], v[2], v[3]);
    else {
        *reinterpret_cast<uint2 *>(out) = make_uint2(v[0], v[1]);
        if (x + 16 <= cw) *reinterpret_cast<uint2 *>(out + 8) = make_uint2(v[2], v[3]);
    }
}

// BT.601 limited range, 8-bit fixed point; chroma from the rounded 2x2 mean RGB (no reference
// function computes this). Each thread converts an 8x2 pixel tile: four 128-bit loads, two 64-bit luma stores,
// one 32-bit store per chroma plane. grid: (ceil((wc/8)*(hc/2)/256), 1, sessions).
__global__ void __launch_bounds__(256) k_ingest_rgba(const Sess *ss, Geom g)
{
    const Sess &s = ss[blockIdx.z];
    const int w = g.width, h = g.height, wc = g.wc, hc = g.hc;
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (wc / 8) * (hc / 2)) return;
    const int x = (u % (wc / 8)) * 8, y = (u / (wc / 8)) * 2;
    uint32_t px[2][8], cpx[2][8];      // luma taps and chroma taps (they differ only inside the padding)
    const bool fast = x + 8 <= w && y + 2 <= h && (w & 3) == 0;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        if (fast) {
            const uint8_t *row = s.input + (size_t)(y + r) * w * 4;
            uint4 a = *reinterpret_cast<const uint4 *>(row + 4 * x), b = *reinterpret_cast<const uint4 *>(row + 4 * x + 16);
            px[r][0] = a.x; px[r][1] = a.y; px[r][2] = a.z; px[r][3] = a.w; px[r][4] = b.x; px[r][5] = b.y; px[r][6] = b.z; px[r][7] = b.w;
#pragma unroll
            for (int i = 0; i < 8; i++) cpx[r][i] = px[r][i];
        } else {
            // padding replicates the last luma sample and the last CHROMA sample (not the chroma of a replicated pixel)
            const uint8_t *row = s.input + (size_t)min(y + r, h - 1) * w * 4;
            const uint8_t *crow = s.input + (size_t)(2 * min(y / 2, h / 2 - 1) + r) * w * 4;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                px[r][i] = *reinterpret_cast<const uint32_t *>(row + 4 * min(x + i, w - 1));
                cpx[r][i] = *reinterpret_cast<const uint32_t *>(crow + 4 * (2 * min(x / 2 + i / 2, w / 2 - 1) + (i & 1)));
            }
        }
    }
    uint32_t yw[2][2] = { { 0, 0 }, { 0, 0 } }, uw = 0, vw = 0;
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            int R = px[r][i] & 255, G = (px[r][i] >> 8) & 255, B = (px[r][i] >> 16) & 255;
            yw[r][i >> 2] |= (uint32_t)(((66 * R + 129 * G + 25 * B + 128) >> 8) + 16) << (8 * (i & 3));
        }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int R = 0, G = 0, B = 0;
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int k = 0; k < 2; k++) { uint32_t p = cpx[r][2 * i + k]; R += p & 255; G += (p >> 8) & 255; B += (p >> 16) & 255; }
        R = (R + 2) >> 2; G = (G + 2) >> 2; B = (B + 2) >> 2;
        uw |= (uint32_t)(((-38 * R - 74 * G + 112 * B + 128) >> 8) + 128) << (8 * i);
        vw |= (uint32_t)(((112 * R - 94 * G - 18 * B + 128) >> 8) + 128) << (8 * i);
    }
    *reinterpret_cast<uint2 *>(s.src[0] + (size_t)y * wc + x) = make_uint2(yw[0][0], yw[0][1]);
    *reinterpret_cast<uint2 *>(s.src[0] + (size_t)(y + 1) * wc + x) = make_uint2(yw[1][0], yw[1][1]);
    *reinterpret_cast<uint32_t *>(s.src[1] + (size_t)(y / 2) * (wc / 2) + x / 2) = uw;
    *reinterpret_cast<uint32_t *>(s.src[2] + (size_t)(y / 2) * (wc / 2) + x / 2) = vw;
}

// 2x2 box filter with rounding (role of DyadicBilinearDownsampler_c). level 0: full -> 1/2, level 1: 1/2 -> 1/4.
// The output planes carry an edge-replicated border of g.p1 / g.p2 samples (the coarse search windows reach outside the
// picture and the oracle clamps coordinates per level), so the kernel covers the padded area and clamps into the interior.
// grid: (ceil(units/256), 2 {src, ref}, sessions); a unit is 8 output pixels (two 64-bit loads per input row, one 64-bit store).
#define DOWNSAMPLE_UNITS(ow, oh, po) ((((ow) + 2 * (po) + 7) / 8) * ((oh) + 2 * (po)))
__device__ __forceinline__ uint32_t box4(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1)     // four outputs from 8 + 8 input samples
{
    // horizontal pair sums on 16x2 lanes: (even bytes) + (odd bytes), both rows, then + 2 and >> 2 per lane
    const uint32_t s0 = (a0 & 0x00ff00ffu) + ((a0 >> 8) & 0x00ff00ffu) + (b0 & 0x00ff00ffu) + ((b0 >> 8) & 0x00ff00ffu) + 0x00020002u;
    const uint32_t s1 = (a1 & 0x00ff00ffu) + ((a1 >> 8) & 0x00ff00ffu) + (b1 & 0x00ff00ffu) + ((b1 >> 8) & 0x00ff00ffu) + 0x00020002u;
    return __byte_perm((s0 >> 2) & 0x00ff00ffu, (s1 >> 2) & 0x00ff00ffu, 0x6420);
}
__global__ void __launch_bounds__(256) k_downsample(const Sess *ss, Geom g, int level)
{
    const Sess &s = ss[blockIdx.z];
    if (s.is_idr) return;
    const int iw = g.wc >> level, ih = g.hc >> level, ow = iw / 2, oh = ih / 2;
    const int is = level == 0 ? g.wc : g.s1, os = level == 0 ? g.s1 : g.s2, po = level == 0 ? g.p1 : g.p2;
    const uint8_t *in = blockIdx.y == 0 ? (level == 0 ? s.src[0] : s.srcL1) : (level == 0 ? s.ref[0] : s.refL1);
    uint8_t *out = blockIdx.y == 0 ? (level == 0 ? s.srcL1 : s.srcL2) : (level == 0 ? s.refL1 : s.refL2);
    const int upr = (ow + 2 * po + 7) / 8;
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= upr * (oh + 2 * po)) return;
    const int y = u / upr - po, x = (u - (y + po) * upr) * 8 - po, cy = min(max(y, 0), oh - 1);
    const int n = min(8, ow + po - x);                 // outputs of this unit that exist (the padded row need not be a multiple of 8)
    uint32_t o[2];
    const uint8_t *r0 = in + (size_t)(2 * cy) * is, *r1 = r0 + is;
    if (x >= 0 && x + 8 <= ow && ((reinterpret_cast<uintptr_t>(r0 + 2 * x) | (uintptr_t)is) & 7) == 0) {
        const uint2 a0 = *reinterpret_cast<const uint2 *>(r0 + 2 * x), a1 = *reinterpret_cast<const uint2 *>(r0 + 2 * x + 8);
        const uint2 b0 = *reinterpret_cast<const uint2 *>(r1 + 2 * x), b1 = *reinterpret_cast<const uint2 *>(r1 + 2 * x + 8);
        o[0] = box4(a0.x, a0.y, b0.x, b0.y); o[1] = box4(a1.x, a1.y, b1.x, b1.y);
    } else {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int cx = min(max(x + 4 * k + i, 0), ow - 1);
                w |= (uint32_t)((r0[2 * cx] + r0[2 * cx + 1] + r1[2 * cx] + r1[2 * cx + 1] + 2) >> 2) << (8 * i);
            }
            o[k] = w;
        }
    }
    uint8_t *dst = out + (ptrdiff_t)y * os + x;
    if (n == 8 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) *reinterpret_cast<uint2 *>(dst) = make_uint2(o[0], o[1]);
    else { *reinterpret_cast<uint32_t *>(dst) = o[0]; if (n > 4) *reinterpret_cast<uint32_t *>(dst + 4) = o[1]; }
}

// Reference planes of a P picture, built once per frame from the previous deblocked reconstruction: plane G is the
// edge-extended picture itself, b / h / j are the half-sample planes of 8.4.2.2.1 (6-tap horizontally, vertically, and both
// from the unrounded intermediate). All carry a border of g.lp samples, so sub-pel search and motion compensation read them
// without clamping and the search windows can be fetched by TMA. HBM stream: 1 B read, 4 B written per padded sample.
#define RP_TW 64
#define RP_TH 32
__device__ __forceinline__ int rp_dp4a(uint32_t a, uint32_t b, int c) { int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int rp_dp2a(uint32_t a, uint32_t b, int c) { int d; asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
// grid: (ceil(ls / RP_TW), ceil((hc + 2 lp) / RP_TH), sessions), 256 threads; a CTA produces a 64x32 tile of all four planes.
//   stage 1  the (32+5) x 72-byte source tile as words (plain 32-bit loads when the tile lies inside the picture)
//   stage 2  unrounded horizontal 6-tap sums by two IDP.4A per sample ({1,-5,20,20} and {-5,1,0,0} against funnel-shifted
//            words), stored as vertical pairs (row r in the low, row r+1 in the high 16 bits)
//   stage 3  j = three IDP.2A over the pairs; h = the vertical 6-tap on bytes unpacked to 16x2 lanes (two samples per
//            integer op, clamped with VIMNMX.S16x2); b = rounded stage-2 sums; four 32-bit stores per thread
__global__ void __launch_bounds__(256) k_refplanes(const Sess *ss, Geom g)
{
    const Sess &s = ss[blockIdx.z];
    if (s.is_idr) return;
    __shared__ uint32_t tile[RP_TH + 5][18];                  // row r: y = Y0 - 2 + r; word j: x = X0 - 4 + 4j ..
    __shared__ __align__(16) uint32_t pair[RP_TH + 4][RP_TW];
    const int X0 = blockIdx.x * RP_TW - g.lp, Y0 = blockIdx.y * RP_TH - g.lp, wc = g.wc, hc = g.hc;
    const uint8_t *ref = s.ref[0];
    const bool inside = X0 - 4 >= 0 && X0 + 68 <= wc && Y0 - 2 >= 0 && Y0 + RP_TH + 3 <= hc;
    for (int i = threadIdx.x; i < (RP_TH + 5) * 18; i += 256) {
        const int r = i / 18, j = i - r * 18, x = X0 - 4 + 4 * j;
        uint32_t w;
        if (inside) w = *reinterpret_cast<const uint32_t *>(ref + (size_t)(Y0 - 2 + r) * wc + x);
        else {
            const uint8_t *row = ref + (size_t)min(max(Y0 - 2 + r, 0), hc - 1) * wc;
            w = (uint32_t)row[min(max(x, 0), wc - 1)] | ((uint32_t)row[min(max(x + 1, 0), wc - 1)] << 8) |
                ((uint32_t)row[min(max(x + 2, 0), wc - 1)] << 16) | ((uint32_t)row[min(max(x + 3, 0), wc - 1)] << 24);
        }
        tile[r][j] = w;
    }
    __syncthreads();
    const uint32_t WLO = 0x1414FB01u, WHI = 0x000001FBu;     // {1,-5,20,20}, {-5,1,0,0} as signed bytes
    for (int i = threadIdx.x; i < (RP_TH + 4) * 16; i += 256) {
        const int r = i >> 4, j = i & 15;
        int v[2][4];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const uint32_t A = tile[r + q][j], B = tile[r + q][j + 1], C = tile[r + q][j + 2];
            v[q][0] = rp_dp4a(__funnelshift_r(A, B, 16), WLO, rp_dp4a(__funnelshift_r(B, C, 16), WHI, 0));
            v[q][1] = rp_dp4a(__funnelshift_r(A, B, 24), WLO, rp_dp4a(__funnelshift_r(B, C, 24), WHI, 0));
            v[q][2] = rp_dp4a(B, WLO, rp_dp4a(C, WHI, 0));
            v[q][3] = rp_dp4a(__funnelshift_r(B, C, 8), WLO, rp_dp4a(C >> 8, WHI, 0));
        }
        uint4 o;
        o.x = __byte_perm((uint32_t)v[0][0], (uint32_t)v[1][0], 0x5410); o.y = __byte_perm((uint32_t)v[0][1], (uint32_t)v[1][1], 0x5410);
        o.z = __byte_perm((uint32_t)v[0][2], (uint32_t)v[1][2], 0x5410); o.w = __byte_perm((uint32_t)v[0][3], (uint32_t)v[1][3], 0x5410);
        *reinterpret_cast<uint4 *>(&pair[r][4 * j]) = o;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RP_TH * 16; i += 256) {
        const int y = i >> 4, c = i & 15;
        if (X0 + 4 * c >= wc + g.lp || Y0 + y >= hc + g.lp) continue;
        const uint4 Q0 = *reinterpret_cast<const uint4 *>(&pair[y][4 * c]), Q1 = *reinterpret_cast<const uint4 *>(&pair[y + 2][4 * c]);
        const uint4 Q2 = *reinterpret_cast<const uint4 *>(&pair[y + 4][4 * c]);
        const uint32_t q0[4] = { Q0.x, Q0.y, Q0.z, Q0.w }, q1[4] = { Q1.x, Q1.y, Q1.z, Q1.w }, q2[4] = { Q2.x, Q2.y, Q2.z, Q2.w };
        uint32_t wb = 0, wj = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int j = rp_dp2a(q0[k], 0xFB01u, rp_dp2a(q1[k], 0x1414u, rp_dp2a(q2[k], 0x01FBu, 512))) >> 10;
            const int b = ((int)(short)(q1[k] & 0xffffu) + 16) >> 5;
            wj |= (uint32_t)__vimin_s32_relu(j, 255) << (8 * k);
            wb |= (uint32_t)__vimin_s32_relu(b, 255) << (8 * k);
        }
        uint32_t e[6], o[6];
#pragma unroll
        for (int r = 0; r < 6; r++) { const uint32_t w = tile[y + r][c + 1]; e[r] = w & 0x00ff00ffu; o[r] = (w >> 8) & 0x00ff00ffu; }
        // per 16-bit lane: tap + 16 + 2560 stays in [26, 13286] (no borrow between lanes); 2560 = 80 << 5 is removed after the shift
        const uint32_t K = (16u + 2560u) * 0x00010001u;
        uint32_t he = (e[0] + e[5]) + 20u * (e[2] + e[3]) + K - 5u * (e[1] + e[4]);
        uint32_t ho = (o[0] + o[5]) + 20u * (o[2] + o[3]) + K - 5u * (o[1] + o[4]);
        he = __vmins2(__vmaxs2((he >> 5) & 0x07ff07ffu, 0x00500050u), 0x014f014fu) - 0x00500050u;
        ho = __vmins2(__vmaxs2((ho >> 5) & 0x07ff07ffu, 0x00500050u), 0x014f014fu) - 0x00500050u;
        const ptrdiff_t off = (ptrdiff_t)(Y0 + y) * g.ls + X0 + 4 * c;
        *reinterpret_cast<uint32_t *>(s.rpl[0] + off) = tile[y + 2][c + 1];
        *reinterpret_cast<uint32_t *>(s.rpl[1] + off) = wb;
        *reinterpret_cast<uint32_t *>(s.rpl[2] + off) = he | (ho << 8);
        *reinterpret_cast<uint32_t *>(s.rpl[3] + off) = wj;
    }
}

// edge-extended copies of the reference chroma planes (border g.cp). grid: (ceil(units/256), 2 {Cb, Cr}, sessions); unit = 16 samples
// (the padded row g.cs is a multiple of 16 and so is every unit's destination address)
__global__ void __launch_bounds__(256) k_refchroma(const Sess *ss, Geom g)
{
    const Sess &s = ss[blockIdx.z];
    if (s.is_idr) return;
    const int cw = g.wc / 2, ch = g.hc / 2, upr = g.cs / 16;
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= upr * (ch + 2 * g.cp)) return;
    const int y = u / upr - g.cp, x = (u - (y + g.cp) * upr) * 16 - g.cp;
    const uint8_t *row = s.ref[1 + blockIdx.y] + (size_t)min(max(y, 0), ch - 1) * cw;
    uint32_t o[4];
    if (x >= 0 && x + 16 <= cw && (reinterpret_cast<uintptr_t>(row + x) & 7) == 0) {
        const uint2 a = *reinterpret_cast<const uint2 *>(row + x), b = *reinterpret_cast<const uint2 *>(row + x + 8);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) w |= (uint32_t)row[min(max(x + 4 * k + i, 0), cw - 1)] << (8 * i);
            o[k] = w;
        }
    }
    *reinterpret_cast<uint4 *>(s.rpc[blockIdx.y] + (ptrdiff_t)y * g.cs + x) = make_uint4(o[0], o[1], o[2], o[3]);
}

} // namespace b200
