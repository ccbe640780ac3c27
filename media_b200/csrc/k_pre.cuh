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

// grid: (ceil(units/256), 1, sessions); a unit is 8 output bytes of one plane row.
__global__ void __launch_bounds__(256) k_ingest_planar(const Sess *ss, Geom g)
{
    const Sess &s = ss[blockIdx.z];
    const int w = g.width, h = g.height, wc = g.wc, hc = g.hc;
    const int ly = (wc / 8) * hc, lc = (wc / 16) * (hc / 2);
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= ly + 2 * lc) return;
    int comp = u < ly ? 0 : (u < ly + lc ? 1 : 2);
    if (comp) u -= ly + (comp - 1) * lc;
    const int cw = comp ? wc / 2 : wc, pw = comp ? w / 2 : w, ph = comp ? h / 2 : h;
    const int upr = cw / 8, y = u / upr, x = (u % upr) * 8, sy = min(y, ph - 1);
    uint2 v;
    if (s.input_format == FMT_I420 || comp == 0) {
        const uint8_t *in = s.input + (comp == 0 ? 0 : (size_t)w * h + (comp == 2 ? (size_t)pw * ph : 0)) + (size_t)sy * pw;
        if (x + 8 <= pw && (pw & 7) == 0) v = *reinterpret_cast<const uint2 *>(in + x);
        else {
            uint32_t a = 0, b = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) { a |= (uint32_t)in[min(x + i, pw - 1)] << (8 * i); b |= (uint32_t)in[min(x + 4 + i, pw - 1)] << (8 * i); }
            v = make_uint2(a, b);
        }
    } else {   // NV12 chroma: de-interleave 16 bytes of UV pairs
        const uint8_t *in = s.input + (size_t)w * h + (size_t)sy * w + (comp - 1);
        uint32_t a = 0, b = 0;
        if (x + 8 <= pw && (w & 15) == 0) {
            uint4 q = *reinterpret_cast<const uint4 *>(in - (comp - 1) + 2 * x);
            uint32_t sel = comp == 1 ? 0x6420 : 0x7531;
            a = __byte_perm(q.x, q.y, sel); b = __byte_perm(q.z, q.w, sel);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) { a |= (uint32_t)in[2 * min(x + i, pw - 1)] << (8 * i); b |= (uint32_t)in[2 * min(x + 4 + i, pw - 1)] << (8 * i); }
        }
        v = make_uint2(a, b);
    }
    *reinterpret_cast<uint2 *>(s.src[comp] + (size_t)y * cw + x) = v;
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
// grid: (ceil(units/256), 2 {src, ref}, sessions); a unit is 4 output pixels.
__global__ void __launch_bounds__(256) k_downsample(const Sess *ss, Geom g, int level)
{
    const Sess &s = ss[blockIdx.z];
    if (s.is_idr) return;
    const int iw = g.wc >> level, ih = g.hc >> level, ow = iw / 2, oh = ih / 2;
    const uint8_t *in = blockIdx.y == 0 ? (level == 0 ? s.src[0] : s.srcL1) : (level == 0 ? s.ref[0] : s.refL1);
    uint8_t *out = blockIdx.y == 0 ? (level == 0 ? s.srcL1 : s.srcL2) : (level == 0 ? s.refL1 : s.refL2);
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (ow / 4) * oh) return;
    const int x = (u % (ow / 4)) * 4, y = u / (ow / 4);
    uint2 a = *reinterpret_cast<const uint2 *>(in + (size_t)(2 * y) * iw + 2 * x);
    uint2 b = *reinterpret_cast<const uint2 *>(in + (size_t)(2 * y + 1) * iw + 2 * x);
    uint32_t aw[2] = { a.x, a.y }, bw[2] = { b.x, b.y }, o = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t p = aw[i >> 1] >> (16 * (i & 1)), q = bw[i >> 1] >> (16 * (i & 1));
        o |= (((p & 255) + ((p >> 8) & 255) + (q & 255) + ((q >> 8) & 255) + 2) >> 2) << (8 * i);
    }
    *reinterpret_cast<uint32_t *>(out + (size_t)y * ow + x) = o;
}

} // namespace b200
