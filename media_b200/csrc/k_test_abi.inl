// media_b200/csrc/k_test_abi.inl -- b200k_* C entry points (include/b200enc.h): host buffers in, production kernels, host buffers out.
namespace {
struct DevBuf {
    void *p = nullptr;
    explicit DevBuf(size_t n) { if (cudaMalloc(&p, n ? n : 16) != cudaSuccess) p = nullptr; }
    ~DevBuf() { if (p) cudaFree(p); }
    template <class T> T *as() { return static_cast<T *>(p); }
};
void fill_geom(Geom &g, int w, int h, int slices, int range)
{
    memset(&g, 0, sizeof g);
    g.width = w; g.height = h; g.mbw = (w + 15) / 16; g.mbh = (h + 15) / 16; g.wc = g.mbw * 16; g.hc = g.mbh * 16;
    geom_set_magic(g);
    g.num_slices = slices; g.search_range = range;
    for (int i = 1; i <= B200_MAX_SLICES; i++) g.slice_row0[i] = g.mbh;
    g.slice_top[0] = 1u;
}
#define K_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { g_last_cuda_error = (int)e_; return B200ENC_ECUDA; } } while (0)
}

extern "C" {

int b200k_convert_to_i420(int device, int fmt, const uint8_t *in, int w, int h, uint8_t *out, int *coded_w, int *coded_h)
{
    if (!in || !out || w < 2 || h < 2 || (w & 1) || (h & 1) || fmt < 0 || fmt > 2) return B200ENC_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return B200ENC_ENODEV;
    Geom g; fill_geom(g, w, h, 1, 16);
    const size_t in_bytes = fmt == B200ENC_FMT_RGBA ? (size_t)w * h * 4 : (size_t)w * h * 3 / 2, ny = (size_t)g.wc * g.hc;
    DevBuf din(in_bytes), dout(ny * 3 / 2), dsess(sizeof(Sess));
    if (!din.p || !dout.p || !dsess.p) return B200ENC_ENOMEM;
    Sess s; memset(&s, 0, sizeof s);
    s.input = din.as<uint8_t>(); s.src[0] = dout.as<uint8_t>(); s.src[1] = s.src[0] + ny; s.src[2] = s.src[1] + ny / 4; s.input_format = fmt;
    K_TRY(cudaMemcpy(din.p, in, in_bytes, cudaMemcpyHostToDevice));
    K_TRY(cudaMemcpy(dsess.p, &s, sizeof s, cudaMemcpyHostToDevice));
    if (fmt == B200ENC_FMT_RGBA) k_ingest_rgba<<<dim3(((g.wc / 8) * (g.hc / 2) + 255) / 256, 1, 1), 256>>>(dsess.as<Sess>(), g);
    else k_ingest_planar<<<dim3((INGEST_UNITS(g.wc, g.hc) + 255) / 256, 1, 1), 256>>>(dsess.as<Sess>(), g);
    K_TRY(cudaGetLastError());
    K_TRY(cudaMemcpy(out, dout.p, ny * 3 / 2, cudaMemcpyDeviceToHost));
    if (coded_w) *coded_w = g.wc;
    if (coded_h) *coded_h = g.hc;
    return B200ENC_OK;
}

int b200k_downsample2(int device, const uint8_t *in, int w, int h, uint8_t *out)
{
    if (!in || !out || w < 8 || h < 2 || (w & 7) || (h & 1)) return B200ENC_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return B200ENC_ENODEV;
    Geom g; memset(&g, 0, sizeof g); g.wc = w; g.hc = h; g.p1 = 0; g.s1 = w / 2;
    DevBuf din((size_t)w * h), dout((size_t)w * h / 4), dsess(sizeof(Sess));
    if (!din.p || !dout.p || !dsess.p) return B200ENC_ENOMEM;
    Sess s; memset(&s, 0, sizeof s);
    s.src[0] = din.as<uint8_t>(); s.srcL1 = dout.as<uint8_t>();
    K_TRY(cudaMemcpy(din.p, in, (size_t)w * h, cudaMemcpyHostToDevice));
    K_TRY(cudaMemcpy(dsess.p, &s, sizeof s, cudaMemcpyHostToDevice));
    k_downsample<<<dim3((DOWNSAMPLE_UNITS(w / 2, h / 2, 0) + 255) / 256, 1, 1), 256>>>(dsess.as<Sess>(), g, 0);
    K_TRY(cudaGetLastError());
    K_TRY(cudaMemcpy(out, dout.p, (size_t)w * h / 4, cudaMemcpyDeviceToHost));
    return B200ENC_OK;
}

static int sad_common(int device, const uint8_t *cur, const uint8_t *ref, int stride, int n, const int32_t *xy, int32_t *out, int satd)
{
    if (!cur || !ref || !xy || !out || n <= 0 || stride < 16) return B200ENC_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return B200ENC_ENODEV;
    int maxy = 0;
    for (int i = 0; i < n; i++) { maxy = std::max(maxy, std::max(xy[4 * i + 1], xy[4 * i + 3])); }
    const size_t bytes = (size_t)stride * (maxy + 16);
    DevBuf dc(bytes), dr(bytes), dxy((size_t)n * 16), dout((size_t)n * 4);
    if (!dc.p || !dr.p || !dxy.p || !dout.p) return B200ENC_ENOMEM;
    K_TRY(cudaMemcpy(dc.p, cur, bytes, cudaMemcpyHostToDevice));
    K_TRY(cudaMemcpy(dr.p, ref, bytes, cudaMemcpyHostToDevice));
    K_TRY(cudaMemcpy(dxy.p, xy, (size_t)n * 16, cudaMemcpyHostToDevice));
    k_test_sad16<<<(n + 127) / 128, 128>>>(dc.as<uint8_t>(), dr.as<uint8_t>(), stride, n, dxy.as<int>(), dout.as<int>(), satd);
    K_TRY(cudaGetLastError());
    K_TRY(cudaMemcpy(out, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return B200ENC_OK;
}
int b200k_sad16x16(int device, const uint8_t *cur, const uint8_t *ref, int stride, int n, const int32_t *xy, int32_t *sad) { return sad_common(device, cur, ref, stride, n, xy, sad, 0); }
int b200k_satd16x16(int device, const uint8_t *cur, const uint8_t *ref, int stride, int n, const int32_t *xy, int32_t *satd) { return sad_common(device, cur, ref, stride, n, xy, satd, 1); }

int b200k_transform_block(int device, const int16_t *res, int n, int qp, int intra, int16_t *levels, int32_t *recon)
{
    if (!res || !levels || !recon || n <= 0 || qp < 0 || qp > 51) return B200ENC_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return B200ENC_ENODEV;
    DevBuf dres((size_t)n * 32), dlev((size_t)n * 32), drec((size_t)n * 64);
    if (!dres.p || !dlev.p || !drec.p) return B200ENC_ENOMEM;
    K_TRY(cudaMemcpy(dres.p, res, (size_t)n * 32, cudaMemcpyHostToDevice));
    k_test_transform<<<(n + 127) / 128, 128>>>(dres.as<int16_t>(), n, qp, intra, dlev.as<int16_t>(), drec.as<int>());
    K_TRY(cudaGetLastError());
    K_TRY(cudaMemcpy(levels, dlev.p, (size_t)n * 32, cudaMemcpyDeviceToHost));
    K_TRY(cudaMemcpy(recon, drec.p, (size_t)n * 64, cudaMemcpyDeviceToHost));
    return B200ENC_OK;
}

int b200k_transform_block8(int device, const int16_t *res, int n, int qp, int intra, int16_t *levels, int32_t *recon)
{
    if (!res || !levels || !recon || n <= 0 || qp < 0 || qp > 51) return B200ENC_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return B200ENC_ENODEV;
    DevBuf dres((size_t)n * 128), dlev((size_t)n * 128), drec((size_t)n * 256);
    if (!dres.p || !dlev.p || !drec.p) return B200ENC_ENOMEM;
    K_TRY(cudaMemcpy(dres.p, res, (size_t)n * 128, cudaMemcpyHostToDevice));
    k_test_transform8<<<(n + 4 * T8_WARPS - 1) / (4 * T8_WARPS), T8_WARPS * 32>>>(dres.as<int16_t>(), n, qp, intra, dlev.as<int16_t>(), drec.as<int>());
    K_TRY(cudaGetLastError());
    K_TRY(cudaMemcpy(levels, dlev.p, (size_t)n * 128, cudaMemcpyDeviceToHost));
    K_TRY(cudaMemcpy(recon, drec.p, (size_t)n * 256, cudaMemcpyDeviceToHost));
    return B200ENC_OK;
}

int b200k_deblock(int device, uint8_t *i420, int mbw, int mbh, const void *mbinfo, int qp)
{
    if (!i420 || !mbinfo || mbw <= 0 || mbh <= 0 || qp < 0 || qp > 51) return B200ENC_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return B200ENC_ENODEV;
    Geom g; fill_geom(g, mbw * 16, mbh * 16, 1, 16);
    const size_t ny = (size_t)g.wc * g.hc, nmb = (size_t)mbw * mbh;
    DevBuf dpix(ny * 3 / 2), dmbi(nmb * sizeof(MbInfo)), dbs(nmb * sizeof(uint4)), dprog((size_t)mbh * 8), dsess(sizeof(Sess)), dctl(sizeof(WaveCtl));
    DevBuf dll((size_t)mbh * (mbw + 1) * DBK_LL_PER_MB * sizeof(uint2));
    if (!dpix.p || !dmbi.p || !dbs.p || !dprog.p || !dsess.p || !dctl.p || !dll.p) return B200ENC_ENOMEM;
    K_TRY(cudaMemset(dll.p, 0, (size_t)mbh * (mbw + 1) * DBK_LL_PER_MB * sizeof(uint2)));
    Sess s; memset(&s, 0, sizeof s);
    s.rec[0] = dpix.as<uint8_t>(); s.rec[1] = s.rec[0] + ny; s.rec[2] = s.rec[1] + ny / 4; s.mbi = dmbi.as<MbInfo>(); s.dbk_bs = dbs.as<uint4>();
    s.row_prog_intra = dprog.as<int>(); s.row_prog_dbk = dprog.as<int>() + mbh; s.qp = qp;
    s.dbk_ll = dll.as<uint2>(); s.dbk_seq = 1;
    K_TRY(cudaMemcpy(dpix.p, i420, ny * 3 / 2, cudaMemcpyHostToDevice));
    K_TRY(cudaMemcpy(dmbi.p, mbinfo, nmb * sizeof(MbInfo), cudaMemcpyHostToDevice));
    K_TRY(cudaMemcpy(dsess.p, &s, sizeof s, cudaMemcpyHostToDevice));
    k_reset<<<(mbh + 255) / 256, 256>>>(dsess.as<Sess>(), g, 1, dctl.as<WaveCtl>());
    k_deblock_bs<<<dim3((unsigned)((nmb + 127) / 128), 1, 1), 128>>>(dsess.as<Sess>(), g);
    k_deblock_wave<<<(mbh + WAVE_WARPS - 1) / WAVE_WARPS, WAVE_WARPS * 32>>>(dsess.as<Sess>(), g, 1, dctl.as<WaveCtl>());
    K_TRY(cudaGetLastError());
    K_TRY(cudaMemcpy(i420, dpix.p, ny * 3 / 2, cudaMemcpyDeviceToHost));
    WaveCtl ctl; K_TRY(cudaMemcpy(&ctl, dctl.p, sizeof ctl, cudaMemcpyDeviceToHost));
    return ctl.error ? B200ENC_EWAVE : B200ENC_OK;
}

int b200k_vabsdiff4_peak(int device, double *ginstr_per_s, int *sm_clock_mhz)
{
    if (!ginstr_per_s) return B200ENC_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return B200ENC_ENODEV;
    cudaDeviceProp prop; K_TRY(cudaGetDeviceProperties(&prop, device));
    const int ctas = prop.multiProcessorCount * 8, iters = 1 << 16;
    DevBuf dsink((size_t)ctas * 256 * 4), dclk((size_t)ctas * 8);
    if (!dsink.p || !dclk.p) return B200ENC_ENOMEM;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_vabsdiff4_peak<<<ctas, 256>>>(1u, 1 << 10, dsink.as<uint32_t>(), dclk.as<long long>());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_vabsdiff4_peak<<<ctas, 256>>>(rep + 2u, iters, dsink.as<uint32_t>(), dclk.as<long long>());
        cudaEventRecord(e1);
        K_TRY(cudaEventSynchronize(e1));
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1); best_ms = std::min(best_ms, ms);
    }
    long long clk = 0; K_TRY(cudaMemcpy(&clk, dclk.p, 8, cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double instr = (double)ctas * 256 * iters * 32.0;
    *ginstr_per_s = instr / (best_ms * 1e-3) / 1e9;
    if (sm_clock_mhz) *sm_clock_mhz = (int)((double)clk / (best_ms * 1e-3) / 1e6);
    return B200ENC_OK;
}


} // extern "C"
// out[kind * 4 + {0,1,2,3}] = G warp-instructions/s over the whole GPU (CUDA events), warp-instructions per clock per SM, SM clock in MHz of the
// run (clock64 / globaltimer inside the kernel, median block), lane-operations per counted unit (32, or 64 x 32 for the SATD kind)
template <int KIND> static int run_int_peak(int sms, double *out)
{
    const int ctas = sms * 8, iters = KIND == 8 ? 1 << 11 : 1 << 15;
    DevBuf dsink((size_t)ctas * 256 * 4), dclk((size_t)ctas * 8), dns((size_t)ctas * 8);
    if (!dsink.p || !dclk.p || !dns.p) return B200ENC_ENOMEM;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_int_peak<KIND><<<ctas, 256>>>(1u, 64, dsink.as<uint32_t>(), dclk.as<long long>(), dns.as<unsigned long long>());
    float best_ms = 1e30f;
    std::vector<long long> clk(ctas); std::vector<unsigned long long> ns(ctas);
    double mhz = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_int_peak<KIND><<<ctas, 256>>>(rep + 2u, iters, dsink.as<uint32_t>(), dclk.as<long long>(), dns.as<unsigned long long>());
        cudaEventRecord(e1);
        K_TRY(cudaEventSynchronize(e1));
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best_ms) {
            best_ms = ms;
            K_TRY(cudaMemcpy(clk.data(), dclk.p, (size_t)ctas * 8, cudaMemcpyDeviceToHost));
            K_TRY(cudaMemcpy(ns.data(), dns.p, (size_t)ctas * 8, cudaMemcpyDeviceToHost));
            std::vector<double> r;
            for (int i = 0; i < ctas; i++) if (ns[i] > 1000) r.push_back((double)clk[i] / (double)ns[i] * 1e3);
            std::sort(r.begin(), r.end());
            mhz = r.empty() ? 0 : r[r.size() / 2];
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double units_per_thread = (double)iters * (KIND == 8 ? 8.0 * 64.0 : 32.0);      // SATD: 8 evaluations of 64 lane-operations per iteration
    const double warp_instr = (double)ctas * 8 * units_per_thread;
    out[0] = warp_instr / (best_ms * 1e-3) / 1e9;
    out[2] = mhz;
    out[1] = mhz > 0 ? out[0] * 1e9 / sms / (mhz * 1e6) : 0;
    out[3] = KIND == 8 ? 64.0 : 1.0;
    return B200ENC_OK;
}
extern "C" {
// checked build: number of device-side bound-check failures since the library was loaded and the id of the first failing site; -1 = not a checked build
// host-only: the macroblock-index arithmetic of every warp-per-MB kernel (mb_xy) against / and % for mb in [0, n); returns the number of mismatches
int b200k_mb_xy_mismatches(int mbw, int n)
{
    if (mbw < 1 || n < 0) return -1;
    Geom g; memset(&g, 0, sizeof g); g.mbw = mbw; geom_set_magic(g);
    int bad = 0;
    for (int mb = 0; mb < n; mb++) { int mx, my; mb_xy_core(g.mbw_magic, g.mbw, mb, mx, my); bad += mx != mb % mbw || my != mb / mbw; }
    return bad;
}
int b200k_check_failures(int device, int *first_site)
{
#ifdef B200_CHECKED
    int v[2] = { 0, 0 };
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(v, g_check_fail, sizeof v) != cudaSuccess) return -2;
    if (first_site) *first_site = v[1];
    return v[0];
#else
    (void)device; if (first_site) *first_site = 0;
    return -1;
#endif
}
int b200k_int_peaks(int device, double *out, int kinds)
{
    if (!out || kinds < 1) return B200ENC_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return B200ENC_ENODEV;
    cudaDeviceProp prop; K_TRY(cudaGetDeviceProperties(&prop, device));
    const int sms = prop.multiProcessorCount;
    typedef int (*Fn)(int, double *);
    const Fn fns[9] = { run_int_peak<0>, run_int_peak<1>, run_int_peak<2>, run_int_peak<3>, run_int_peak<4>, run_int_peak<5>, run_int_peak<6>, run_int_peak<7>, run_int_peak<8> };
    for (int k = 0; k < kinds && k < 9; k++) { const int rc = fns[k](sms, out + 4 * k); if (rc != B200ENC_OK) return rc; }
    return B200ENC_OK;
}

int b200k_cabac_code(int device, const uint16_t *bins, int n, int qp, int is_p, uint8_t *out, int cap, int *out_len)
{
    if (!bins || n <= 0 || !out || !out_len || cap <= 0) return B200ENC_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return B200ENC_ENODEV;
    const size_t padded = (size_t)n + 64;
    DevBuf dbins(padded * 2), dout((size_t)cap + 16), dlen(sizeof(int));
    if (!dbins.p || !dout.p || !dlen.p) return B200ENC_ENOMEM;
    K_TRY(cudaMemset(dbins.p, 0, padded * 2));
    K_TRY(cudaMemcpy(dbins.p, bins, (size_t)n * 2, cudaMemcpyHostToDevice));
    k_cabac_code_test<<<1, 96>>>(dbins.as<uint16_t>(), n, qp, is_p, dout.as<uint8_t>(), dlen.as<int>());
    K_TRY(cudaGetLastError());
    K_TRY(cudaMemcpy(out_len, dlen.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (*out_len > cap) return B200ENC_EOVERFLOW;
    K_TRY(cudaMemcpy(out, dout.p, (size_t)*out_len, cudaMemcpyDeviceToHost));
    return B200ENC_OK;
}

// the coder kernel `reps` times on one bin list: average device time per run (CUDA events) and, in a -DCABAC_TIMING build, the phase
// cycle counts of the producer / consumer pair summed over the runs (tools/cabac_coder_bench.py)
int b200k_cabac_code_bench(int device, const uint16_t *bins, int n, int qp, int is_p, int reps, int copies, float *ms, long long *stats)
{
    if (!bins || n <= 0 || reps <= 0 || copies <= 0 || !ms) return B200ENC_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return B200ENC_ENODEV;
    const size_t padded = (size_t)n + 64;
    DevBuf dbins(padded * 2 * copies), dout(((size_t)n * 8 + 64) * copies), dlen(sizeof(int) * copies);
    if (!dbins.p || !dout.p || !dlen.p) return B200ENC_ENOMEM;
    for (int c = 0; c < copies; c++) K_TRY(cudaMemcpy(dbins.as<uint16_t>() + c * padded, bins, (size_t)n * 2, cudaMemcpyHostToDevice));
#ifdef CABAC_TIMING
    { long long z[8] = { 0 }; K_TRY(cudaMemcpyToSymbol(g_cabac_t, z, sizeof z)); }
#endif
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++)
        k_cabac_code_multi<<<copies, 96>>>(dbins.as<uint16_t>(), (int)padded, n, qp, is_p, dout.as<uint8_t>(), n * 8 + 64, dlen.as<int>());
    cudaEventRecord(e1);
    K_TRY(cudaEventSynchronize(e1));
    K_TRY(cudaGetLastError());
    cudaEventElapsedTime(ms, e0, e1); *ms /= reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (stats) {
        memset(stats, 0, 8 * sizeof(long long));
#ifdef CABAC_TIMING
        K_TRY(cudaMemcpyFromSymbol(stats, g_cabac_t, 8 * sizeof(long long)));
#endif
    }
    return B200ENC_OK;
}
} // extern "C"

#ifdef INTRA_TIMING
extern "C" int b200k_intra_timing(long long *out8, int reset)
{
    if (cudaMemcpyFromSymbol(out8, b200::g_intra_t, sizeof(long long) * 8) != cudaSuccess) return -1;
    if (reset) { long long z[8] = { 0 }; cudaMemcpyToSymbol(b200::g_intra_t, z, sizeof z); }
    return 0;
}
#endif
