// media_b200/csrc/k_test.cuh -- thin kernels that expose the production device functions (SAD, SATD, transform chain)
// one block at a time, for the per-kernel parity tests against oracle/ (SURVEY.md section 2b lists the openh264
// C functions in the same roles), plus the VABSDIFF4 issue-rate microbenchmark that defines the ME roofline.
#pragma once
#include "h264_dev.cuh"
#include "k_me.cuh"

namespace b200 {

// one thread per 16x16 block pair; xy = {cx, cy, rx, ry}; coordinates must keep both blocks inside the planes
__global__ void k_test_sad16(const uint8_t *cur, const uint8_t *ref, int stride, int n, const int *xy, int *out, int satd)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *a = cur + (size_t)xy[4 * i + 1] * stride + xy[4 * i], *b = ref + (size_t)xy[4 * i + 3] * stride + xy[4 * i + 2];
    int acc = 0;
    if (!satd) {
        for (int y = 0; y < 16; y++)
            for (int x = 0; x < 16; x += 4) {
                uint32_t wa = a[y * stride + x] | (a[y * stride + x + 1] << 8) | (a[y * stride + x + 2] << 16) | ((uint32_t)a[y * stride + x + 3] << 24);
                uint32_t wb = b[y * stride + x] | (b[y * stride + x + 1] << 8) | (b[y * stride + x + 2] << 16) | ((uint32_t)b[y * stride + x + 3] << 24);
                acc = sad4(wa, wb, acc);
            }
    } else {
        for (int by = 0; by < 16; by += 4)
            for (int bx = 0; bx < 16; bx += 4) {
                int d[16];
                for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) d[y * 4 + x] = (int)a[(by + y) * stride + bx + x] - (int)b[(by + y) * stride + bx + x];
                acc += satd4x4(d);
            }
    }
    out[i] = acc;
}

// residual -> fdct -> quant (zig-zag levels) -> dequant -> idct, one thread per 4x4 block
__global__ void k_test_transform(const int16_t *res, int n, int qp, int intra, int16_t *levels, int *recon)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[16]; int16_t lz[16];
    for (int k = 0; k < 16; k++) c[k] = res[i * 16 + k];
    fdct4x4(c);
    const QParam q = make_qparam(qp);
    quant_dequant4x4(c, lz, q, intra ? q.f_intra : q.f_inter, false);
    idct4x4(c);
    for (int k = 0; k < 16; k++) { levels[i * 16 + k] = lz[k]; recon[i * 16 + k] = c[k]; }
}

// register-resident VABSDIFF4.U8.ACC chain: 8 independent accumulators per thread, `iters` x 8 instructions
__global__ void __launch_bounds__(256) k_vabsdiff4_peak(uint32_t seed, int iters, uint32_t *sink, long long *clocks)
{
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    uint32_t s0 = 0, s1 = 1, s2 = 2, s3 = 3, s4 = 4, s5 = 5, s6 = 6, s7 = 7;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            s0 = sad4(a0, s1, s0); s1 = sad4(a1, s2, s1); s2 = sad4(a2, s3, s2); s3 = sad4(a3, s4, s3);
            s4 = sad4(a4, s5, s4); s5 = sad4(a5, s6, s5); s6 = sad4(a6, s7, s6); s7 = sad4(a7, s0, s7);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s0 ^ s1 ^ s2 ^ s3 ^ s4 ^ s5 ^ s6 ^ s7;
}

// Issue-rate microbenchmarks of the integer instructions the motion search and the transform chain are made of (the denominators of the
// INT roofline): register-resident, 8 independent dependency chains per thread, 2 048 resident threads per SM.
// KIND 0 VABSDIFF4.U8.ACC, 1 IADD3, 2 LOP3, 3 IDP.4A, 4 IMAD, 5 VIMNMX (max), 6 IABS + IADD (abs as the kernels use it), 7 SHF (funnel shift),
// 8 = the kernel's own 4x4 Hadamard SATD (satd_rows of k_me.cuh: 16 IDP.4A + butterflies + abs / max) counted as 64 lane-operations.
// Every block records its own cycle count (clock64) and its own duration in ns (globaltimer): the SM clock of the run is their ratio,
// measured inside the kernel and not read from the driver afterwards.
template <int KIND>
__device__ __forceinline__ uint32_t int_peak_op(uint32_t a, uint32_t b, uint32_t c)
{
    if (KIND == 0) return sad4(a, b, c);
    if (KIND == 1) return a + b + c;
    if (KIND == 2) return (a & b) ^ c;
    if (KIND == 3) return (uint32_t)dp4a_us(a, b, (int)c);
    if (KIND == 4) return a * b + c;
    if (KIND == 5) return (uint32_t)max((int)a ^ (int)c, (int)b);
    if (KIND == 6) return (uint32_t)abs((int)(c - a)) + b;
    return __funnelshift_r(a, c, b);
}
template <int KIND>
__global__ void __launch_bounds__(256) k_int_peak(uint32_t seed, int iters, uint32_t *sink, long long *clocks, unsigned long long *ns)
{
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    uint32_t s0 = 0, s1 = 1, s2 = 2, s3 = 3, s4 = 4, s5 = 5, s6 = 6, s7 = 7;
    unsigned long long n0, n1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n0));
    const long long t0 = clock64();
    if (KIND == 8) {
        // 8 independent SATD evaluations per iteration on rotating prediction words
        int Ts[16];
#pragma unroll
        for (int k = 0; k < 16; k++) Ts[k] = (int)(a0 >> k) & 1023;
        uint32_t P[4] = { a0, a1, a2, a3 };
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int v = satd_rows(P, Ts);
                s0 += (uint32_t)v;
                // every prediction word depends on the result, so no part of an evaluation can be reused by the next (+9 instructions, not counted)
                P[0] = P[0] * 5u + s0; P[1] = P[1] * 3u + s0; P[2] += s0 + (uint32_t)u; P[3] ^= s0;
            }
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                s0 = int_peak_op<KIND>(a0, s1, s0); s1 = int_peak_op<KIND>(a1, s2, s1); s2 = int_peak_op<KIND>(a2, s3, s2); s3 = int_peak_op<KIND>(a3, s4, s3);
                s4 = int_peak_op<KIND>(a4, s5, s4); s5 = int_peak_op<KIND>(a5, s6, s5); s6 = int_peak_op<KIND>(a6, s7, s6); s7 = int_peak_op<KIND>(a7, s0, s7);
            }
        }
    }
    const long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
    if (threadIdx.x == 0) { clocks[blockIdx.x] = t1 - t0; ns[blockIdx.x] = n1 - n0; }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s0 ^ s1 ^ s2 ^ s3 ^ s4 ^ s5 ^ s6 ^ s7;
}

} // namespace b200
