// tools/probes/tma_probe.cu -- checks the TMA usage pattern the ME kernels rely on: tensor maps kept in GLOBAL memory,
// 2D and 3D u8 tile loads at byte-granular (unaligned, partly out-of-bounds) coordinates, one mbarrier per warp.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cstdlib>
typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const CUtensorMap *maps, int x, int y, uint8_t *out2d, uint8_t *out3d)
{
    __shared__ __align__(128) uint8_t buf2[32 * 20];
    __shared__ __align__(128) uint8_t buf3[4 * 18 * 32];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(32 * 20 + 4 * 18 * 32) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(smem_u32(buf2)), "l"(maps), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(smem_u32(buf3)), "l"(maps + 1), "r"(x), "r"(y), "r"(0), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 32 * 20; i += blockDim.x) out2d[i] = buf2[i];
    for (int i = threadIdx.x; i < 4 * 18 * 32; i += blockDim.x) out3d[i] = buf3[i];
}
int main(int argc, char **argv)
{
    const int W = 96, H = 64, P = 4;
    std::vector<uint8_t> h((size_t)P * W * H);
    for (int p = 0; p < P; p++) for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) h[((size_t)p * H + y) * W + x] = (uint8_t)(p * 64 + ((x * 7 + y * 13) & 63));
    uint8_t *d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    EncodeTiled enc = nullptr; cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qr) != cudaSuccess || !enc) { printf("no entry point\n"); return 1; }
    CUtensorMap hm[2];
    { cuuint64_t dims[2] = { W, H }, strides[1] = { W }; cuuint32_t box[2] = { 32, 20 }, es[2] = { 1, 1 };
      CUresult r = enc(&hm[0], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode2d %d\n", (int)r); return 1; } }
    { cuuint64_t dims[3] = { W, H, P }, strides[2] = { W, (cuuint64_t)W * H }; cuuint32_t box[3] = { 32, 18, 4 }, es[3] = { 1, 1, 1 };
      CUresult r = enc(&hm[1], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode3d %d\n", (int)r); return 1; } }
    CUtensorMap *dm; cudaMalloc(&dm, sizeof hm); cudaMemcpy(dm, hm, sizeof hm, cudaMemcpyHostToDevice);
    uint8_t *o2, *o3; cudaMalloc(&o2, 32 * 20); cudaMalloc(&o3, 4 * 18 * 32);
    int bad = 0;
    std::vector<int> xs, ys;
    if (argc >= 3) { xs.push_back(atoi(argv[1])); ys.push_back(atoi(argv[2])); } else { xs = { 0, 16, -16, 80 }; ys = { 0, 7, -2, 50, 60 }; }
    for (int x : xs) for (int y : ys) {
        probe<<<1, 64>>>(dm, x, y, o2, o3);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed at (%d,%d): %s\n", x, y, cudaGetErrorString(e)); return 2; }
        std::vector<uint8_t> r2(32 * 20), r3(4 * 18 * 32);
        cudaMemcpy(r2.data(), o2, r2.size(), cudaMemcpyDeviceToHost); cudaMemcpy(r3.data(), o3, r3.size(), cudaMemcpyDeviceToHost);
        for (int r = 0; r < 20; r++) for (int c = 0; c < 32; c++) {
            int gx = x + c, gy = y + r; uint8_t want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0 : h[(size_t)gy * W + gx];
            if (r2[r * 32 + c] != want) bad++;
        }
        for (int p = 0; p < P; p++) for (int r = 0; r < 18; r++) for (int c = 0; c < 32; c++) {
            int gx = x + c, gy = y + r; uint8_t want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0 : h[((size_t)p * H + gy) * W + gx];
            if (r3[(p * 18 + r) * 32 + c] != want) bad++;
        }
    }
    printf("tma probe x=%d: %d mismatching bytes\n", xs[0], bad);
    return bad ? 3 : 0;
}
