// flo_selftest_umma: on-device checks of the tcgen05/TMEM/TMA building blocks.
//   1. descriptor micro-GEMMs: canonical no-swizzle K-major operands, shifted start addresses and
//      non-128-byte group strides (the forms the implicit-GEMM convolution relies on) vs a host GEMM;
//   2. the full tcgen05 convolution kernel vs the CUDA-core convolution on identical bf16 inputs,
//      for every layer shape class of the U-Net (strip / flattened tiles, concat, n-tiling, batch tail).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "flo_internal.h"

namespace flo {

struct Report {
    std::string text;
    int fails = 0;
    void line(bool ok, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        text += ok ? "PASS " : "FAIL ";
        text += buf;
        text += "\n";
        if (!ok) ++fails;
    }
};

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }
static float rnd(uint32_t& s) { return ((lcg(s) >> 8) & 0xFFFF) / 32768.0f - 1.0f; }   // [-1, 1)

#define ST_CUDA(expr)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            set_error("selftest: %s failed: %s", #expr, cudaGetErrorString(_e));           \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

static int micro_case(Report& rep, cudaStream_t st, int N, int K, int a_lbo, int a_sbo, int a_shift) {
    uint32_t seed = 1234u + N * 7 + K;
    std::vector<__nv_bfloat16> a((size_t)128 * K), b((size_t)N * K);
    for (auto& v : a) v = __float2bfloat16(rnd(seed));
    for (auto& v : b) v = __float2bfloat16(rnd(seed));
    __nv_bfloat16 *da, *db;
    float* dd;
    ST_CUDA(cudaMalloc(&da, a.size() * 2));
    ST_CUDA(cudaMalloc(&db, b.size() * 2));
    ST_CUDA(cudaMalloc(&dd, (size_t)128 * N * 4));
    ST_CUDA(cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
    ST_CUDA(cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    ST_CUDA(cudaMemset(dd, 0, (size_t)128 * N * 4));
    ST_CUDA(launch_umma_micro(da, db, dd, N, K, a_lbo, a_sbo, a_shift, N * 16, 128, st));
    ST_CUDA(cudaStreamSynchronize(st));
    std::vector<float> d((size_t)128 * N);
    ST_CUDA(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)__bfloat162float(a[(size_t)m * K + k]) * __bfloat162float(b[(size_t)n * K + k]);
            maxerr = fmax(maxerr, fabs(r - d[(size_t)m * N + n]));
            maxref = fmax(maxref, fabs(r));
        }
    rep.line(maxerr <= 1e-3 * fmax(1.0, maxref), "umma_micro N=%d K=%d a_lbo=%d a_sbo=%d a_shift=%d  max|err|=%.3g max|ref|=%.3g", N, K,
             a_lbo, a_sbo, a_shift, maxerr, maxref);
    cudaFree(da); cudaFree(db); cudaFree(dd);
    return 0;
}


static int micro2_case(Report& rep, cudaStream_t st, const char* what, int N, int K, int layout, int row_bytes, int a_sbo, int a_shift,
                       int a_lbo, int use_bo, int reps) {
    uint32_t seed = 777u + N * 3 + K + layout;
    std::vector<__nv_bfloat16> a((size_t)128 * K), b((size_t)N * K);
    for (auto& v : a) v = __float2bfloat16(rnd(seed));
    for (auto& v : b) v = __float2bfloat16(rnd(seed));
    __nv_bfloat16 *da, *db; float* dd; long long* dc;
    ST_CUDA(cudaMalloc(&da, a.size() * 2)); ST_CUDA(cudaMalloc(&db, b.size() * 2));
    ST_CUDA(cudaMalloc(&dd, (size_t)128 * N * 4)); ST_CUDA(cudaMalloc(&dc, 8));
    ST_CUDA(cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
    ST_CUDA(cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    ST_CUDA(cudaMemset(dd, 0, (size_t)128 * N * 4));
    ST_CUDA(launch_umma_micro2(da, db, dd, dc, N, K, layout, row_bytes, a_sbo, a_shift, a_lbo, use_bo, reps, st));
    ST_CUDA(cudaStreamSynchronize(st));
    std::vector<float> d((size_t)128 * N);
    long long cyc = 0;
    ST_CUDA(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
    ST_CUDA(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)__bfloat162float(a[(size_t)m * K + k]) * __bfloat162float(b[(size_t)n * K + k]);
            r *= reps;
            maxerr = fmax(maxerr, fabs(r - d[(size_t)m * N + n]));
            maxref = fmax(maxref, fabs(r));
        }
    const bool ok = maxerr <= 2e-3 * fmax(1.0, maxref);
    rep.text += ok ? "INFO-PASS " : "INFO-FAIL ";
    char buf[512];
    snprintf(buf, sizeof(buf), "umma_swizzle %-28s N=%d K=%d layout=%d row=%d sbo=%d shift=%d base_off=%d reps=%d  max|err|=%.3g max|ref|=%.3g  cycles=%lld (%.1f per MMA)\n",
             what, N, K, layout, row_bytes, a_sbo, a_shift, use_bo, reps, maxerr, maxref, cyc, (double)cyc / (reps * (K / 16)));
    rep.text += buf;
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc);
    return 0;
}

static int conv_case(Report& rep, cudaStream_t st, const char* name, int B, int H, int W, int C0, int C1, int cout,
                     int ksize, int n_tile, bool with_res) {
    const int cin = C0 + C1, taps = ksize * ksize, HW = H * W;
    uint32_t seed = 99u + H * 131 + cin * 7 + cout;
    std::vector<float> w((size_t)cout * cin * taps), bias(cout);
    const float wscale = 1.0f / sqrtf((float)(cin * taps));
    for (auto& v : w) v = __bfloat162float(__float2bfloat16(rnd(seed) * wscale));
    for (auto& v : bias) v = rnd(seed);
    auto make_in = [&](int C, std::vector<__nv_bfloat16>& v) {
        v.resize((size_t)C * B * HW);
        for (auto& x : v) x = __float2bfloat16(rnd(seed));
    };
    std::vector<__nv_bfloat16> in0, in1;
    make_in(C0, in0);
    if (C1) make_in(C1, in1);
    std::vector<float> res((size_t)cout * B * HW);
    for (auto& v : res) v = rnd(seed);
    // weights: SIMT [tap][cin][cout] fp32, UMMA packed stream
    std::vector<float> wsimt((size_t)taps * cin * cout);
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
            for (int t = 0; t < taps; ++t) wsimt[((size_t)t * cin + ci) * cout + co] = w[((size_t)co * cin + ci) * taps + t];
    std::vector<__nv_bfloat16> wumma;
    pack_umma_weights(w.data(), cout, cin, ksize, nullptr, n_tile, false, wumma);

    __nv_bfloat16 *d_in0 = nullptr, *d_in1 = nullptr, *d_wu = nullptr, *d_oo = nullptr;
    float *d_ws = nullptr, *d_bias = nullptr, *d_res = nullptr, *d_ref = nullptr, *d_om = nullptr;
    const size_t out_n = (size_t)cout * B * HW;
    ST_CUDA(cudaMalloc(&d_in0, in0.size() * 2));
    ST_CUDA(cudaMemcpy(d_in0, in0.data(), in0.size() * 2, cudaMemcpyHostToDevice));
    if (C1) {
        ST_CUDA(cudaMalloc(&d_in1, in1.size() * 2));
        ST_CUDA(cudaMemcpy(d_in1, in1.data(), in1.size() * 2, cudaMemcpyHostToDevice));
    }
    ST_CUDA(cudaMalloc(&d_wu, wumma.size() * 2));
    ST_CUDA(cudaMemcpy(d_wu, wumma.data(), wumma.size() * 2, cudaMemcpyHostToDevice));
    ST_CUDA(cudaMalloc(&d_ws, wsimt.size() * 4));
    ST_CUDA(cudaMemcpy(d_ws, wsimt.data(), wsimt.size() * 4, cudaMemcpyHostToDevice));
    ST_CUDA(cudaMalloc(&d_bias, bias.size() * 4));
    ST_CUDA(cudaMemcpy(d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    ST_CUDA(cudaMalloc(&d_res, res.size() * 4));
    ST_CUDA(cudaMemcpy(d_res, res.data(), res.size() * 4, cudaMemcpyHostToDevice));
    ST_CUDA(cudaMalloc(&d_ref, out_n * 4));
    ST_CUDA(cudaMalloc(&d_om, out_n * 4));
    ST_CUDA(cudaMalloc(&d_oo, out_n * 2));
    ST_CUDA(cudaMemset(d_om, 0xFF, out_n * 4));   // NaN pattern: unwritten outputs are caught
    ST_CUDA(cudaMemset(d_oo, 0xFF, out_n * 2));

    ConvSimtParams sp{};
    sp.in0 = d_in0; sp.in1 = d_in1; sp.ncb0 = C0 / 8; sp.ncb1 = C1 / 8; sp.in_is_bf16 = 1; sp.w = d_ws; sp.bias = d_bias;
    sp.res = with_res ? d_res : nullptr; sp.out_m = d_ref; sp.out_o = nullptr; sp.o_is_bf16 = 0;
    sp.B = B; sp.H = H; sp.W = W; sp.cout = cout; sp.ksize = ksize;
    ST_CUDA(launch_conv_simt(sp, st));

    ConvUmmaParams up{};
    ConvShape cs{name, H, W, ksize, C0 / 8, C1 / 8, cout, n_tile};
    if (plan_umma(cs, B, up)) return -1;
    CUtensorMap t0, t1;
    if (make_tmap(&t0, d_in0, C0 / 8, B, H, W, ksize / 2, up.nb)) return -1;
    t1 = t0;
    if (C1 && make_tmap(&t1, d_in1, C1 / 8, B, H, W, ksize / 2, up.nb)) return -1;
    up.w = d_wu; up.bias = d_bias; up.res = with_res ? d_res : nullptr; up.out_m = d_om; up.out_o = d_oo;
    ST_CUDA(conv_umma_configure());
    ST_CUDA(launch_conv_umma(up, t0, t1, st));
    ST_CUDA(cudaStreamSynchronize(st));

    std::vector<float> ref(out_n), om(out_n);
    std::vector<__nv_bfloat16> oo(out_n);
    ST_CUDA(cudaMemcpy(ref.data(), d_ref, out_n * 4, cudaMemcpyDeviceToHost));
    ST_CUDA(cudaMemcpy(om.data(), d_om, out_n * 4, cudaMemcpyDeviceToHost));
    ST_CUDA(cudaMemcpy(oo.data(), d_oo, out_n * 2, cudaMemcpyDeviceToHost));
    double e_m = 0, e_o = 0, mref = 0;
    size_t bad = 0;
    for (size_t i = 0; i < out_n; ++i) {
        const double r = ref[i];
        mref = fmax(mref, fabs(r));
        const double dm = fabs(om[i] - r), dob = fabs(__bfloat162float(oo[i]) - r);
        if (!(dm == dm) || !(dob == dob)) { ++bad; continue; }
        e_m = fmax(e_m, dm);
        e_o = fmax(e_o, dob);
    }
    const bool ok = bad == 0 && e_m <= 2e-4 * fmax(1.0, mref) && e_o <= 1e-2 * fmax(1.0, mref);
    rep.line(ok, "conv %-22s B=%d %dx%d cin=%d+%d cout=%d k=%d | nb=%d mt=%d n_tile=%d S=%d stages=%d smem=%d tmem=%d | "
                 "max|f32 err|=%.3g max|bf16 err|=%.3g max|ref|=%.3g nan=%zu",
             name, B, H, W, C0, C1, cout, ksize, up.nb, up.n_mtiles, up.n_tile, up.slices_per_stage, up.n_wstages, up.smem_bytes,
             up.tmem_cols, e_m, e_o, mref, bad);
    cudaFree(d_in0); if (d_in1) cudaFree(d_in1);
    cudaFree(d_wu); cudaFree(d_ws); cudaFree(d_bias); cudaFree(d_res); cudaFree(d_ref); cudaFree(d_om); cudaFree(d_oo);
    return 0;
}

}  // namespace flo

using namespace flo;

extern "C" int flo_selftest_umma(char* report, int report_cap, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    Report rep;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device"); return -1; }
    int rc = 0;
    // ---- descriptor micro-GEMMs
    rc |= micro_case(rep, st, 16, 16, 2048, 128, 0);
    rc |= micro_case(rep, st, 64, 64, 2048, 128, 0);
    rc |= micro_case(rep, st, 128, 144, 2048, 128, 0);
    rc |= micro_case(rep, st, 256, 32, 2048, 128, 0);
    rc |= micro_case(rep, st, 32, 32, 2560, 128, 16 * 19);        // odd-pixel start shift
    rc |= micro_case(rep, st, 16, 32, 5184, 288, 304);            // strip mode: group stride = padded row pitch
    rc |= micro_case(rep, st, 48, 16, 2048, 128, 0);              // N multiple of 16 but not a power of two
    // ---- informational probes (not counted as failures): swizzled operands, shifted starts, issue rate
    micro2_case(rep, st, "none aligned rate", 128, 64, 0, 16, 128, 0, 2048, 0, 16);
    micro2_case(rep, st, "sw128 aligned", 128, 64, 2, 128, 1024, 0, 0, 0, 1);
    micro2_case(rep, st, "sw128 aligned rate", 128, 64, 2, 128, 1024, 0, 0, 0, 16);
    micro2_case(rep, st, "sw128 shift3 bo=0", 128, 64, 2, 128, 1024, 3 * 128, 0, 0, 1);
    micro2_case(rep, st, "sw128 strip sbo=2304 bo=0", 32, 64, 2, 128, 2304, 19 * 128, 0, 0, 1);
    micro2_case(rep, st, "sw32 aligned", 16, 16, 6, 32, 256, 0, 0, 0, 1);
    micro2_case(rep, st, "sw32 aligned rate", 16, 16, 6, 32, 256, 0, 0, 0, 64);
    micro2_case(rep, st, "none N=16 rate", 16, 16, 0, 16, 128, 0, 2048, 0, 64);
    micro2_case(rep, st, "sw32 shift19 bo=0", 16, 16, 6, 32, 256, 19 * 32, 0, 0, 1);
    micro2_case(rep, st, "sw32 strip sbo=576 bo=0", 16, 16, 6, 32, 18 * 32, 19 * 32, 0, 0, 1);
    micro2_case(rep, st, "sw64 shift5 bo=0", 32, 32, 4, 64, 512, 5 * 64, 0, 0, 1);
    {   // back-to-back MMA throughput (informational)
        long long* dc; long long hc[2];
        ST_CUDA(cudaMalloc(&dc, 16));
        const int Ns[4] = {16, 32, 64, 128}, counts[4] = {1, 8, 32, 128};
        // acc -1: two issuing warps, one accumulator each; 1: one issuer.  Then, one issuer: A start one pixel (16 bytes)
        // and one padded row + pixel (16*19 bytes) off the line, as the 3x3 tap descriptors are; and A read from TMEM.
        // {accumulators, a_mode, a_shift}: a_mode 0 = A in shared memory (contiguous row groups), 1 = A in tensor memory,
        // >= 2 = A in shared memory with row groups a_mode*16 bytes apart (18: padded 16-pixel rows; 24: rows padded to 384 B; 10: 8-pixel rows)
        const int modes[12][3] = {{-1, 0, 0}, {1, 0, 0}, {1, 0, 16}, {1, 0, 16 * 19}, {1, 1, 0}, {1, 18, 0}, {1, 18, 16}, {1, 18, 32},
                                  {1, 24, 0}, {1, 24, 16}, {1, 10, 0}, {1, 10, 16}};
        for (int m = 0; m < 12; ++m)
            for (int ni = 0; ni < 4; ++ni) {
                char buf[320]; int o = snprintf(buf, sizeof(buf), "INFO umma_rate N=%-3d acc=%d a=%s shift=%d :", Ns[ni], modes[m][0],
                                                modes[m][1] == 1 ? "tmem" : (modes[m][1] == 0 ? "smem" : (modes[m][1] == 18 ? "smem/sbo288" : (modes[m][1] == 24 ? "smem/sbo384" : "smem/sbo160"))), modes[m][2]);
                for (int ci = 0; ci < 4; ++ci) {
                    ST_CUDA(launch_umma_rate(dc, Ns[ni], counts[ci], modes[m][0], st, modes[m][1], modes[m][2]));
                    ST_CUDA(cudaStreamSynchronize(st));
                    ST_CUDA(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
                    o += snprintf(buf + o, sizeof(buf) - o, "  n=%d issue %lld done %lld", counts[ci], hc[0], hc[1]);
                }
                rep.text += buf; rep.text += "\n";
            }
        // weight streaming from L2 into one CTA's shared-memory ring (1 MiB per CTA, warm in L2): bytes per SM clock
        {
            const int total = 1 << 20;
            uint8_t* dsrc = nullptr;
            ST_CUDA(cudaMalloc(&dsrc, (size_t)148 * total));
            ST_CUDA(cudaMemset(dsrc, 1, (size_t)148 * total));
            // {CTAs, chunk bytes, ring slots, copies per chunk (< 0: that many issuing lanes, one copy per chunk), same source}
            const int cfgs[14][5] = {{1, 16384, 4, 4, 1}, {1, 16384, 8, 1, 1}, {1, 16384, 8, 16, 1}, {1, 4096, 32, 1, 1}, {1, 32768, 4, 1, 1},
                                     {1, 65536, 2, 1, 1}, {1, 16384, 8, -2, 1}, {1, 16384, 8, -4, 1}, {1, 16384, 8, -8, 1}, {1, 4096, 32, -8, 1},
                                     {1, 4096, 32, -32, 1}, {128, 16384, 8, 1, 1}, {128, 16384, 8, -4, 1}, {128, 16384, 8, -4, 0}};
            for (int k = 0; k < 14; ++k) {
                long long best = 1LL << 60;
                for (int rep2 = 0; rep2 < 3; ++rep2) {
                    ST_CUDA(launch_stream_rate(dsrc, dc, cfgs[k][0], total, cfgs[k][1], cfgs[k][2], cfgs[k][3], cfgs[k][4], st));
                    ST_CUDA(cudaStreamSynchronize(st));
                    ST_CUDA(cudaMemcpy(hc, dc, 8, cudaMemcpyDeviceToHost));
                    if (hc[0] < best) best = hc[0];
                }
                char buf[256];
                snprintf(buf, sizeof(buf), "INFO stream_rate ctas=%-3d chunk=%-5d ring=%-2d pieces=%-2d same_src=%d : %lld cycles per MiB -> %.1f B/clk",
                         cfgs[k][0], cfgs[k][1], cfgs[k][2], cfgs[k][3], cfgs[k][4], best, (double)total / (double)best);
                rep.text += buf; rep.text += "\n";
            }
            cudaFree(dsrc);
        }
        cudaFree(dc);
    }
    auto flush = [&]() { if (report && report_cap > 0) snprintf(report, report_cap, "%s", rep.text.c_str()); };
    if (rc) { flush(); return -1; }
    // ---- tcgen05 convolution vs CUDA-core convolution
    rc |= conv_case(rep, st, "3x3 16->16 @16 strip", 5, 16, 16, 16, 0, 16, 3, 16, false);
    rc |= conv_case(rep, st, "3x3 16+16->16 @16", 5, 16, 16, 16, 16, 16, 3, 16, true);
    rc |= conv_case(rep, st, "1x1 16->384 @16 qkv", 3, 16, 16, 16, 0, 384, 1, 128, false);
    rc |= conv_case(rep, st, "1x1 128->16 @16", 3, 16, 16, 128, 0, 16, 1, 16, true);
    rc |= conv_case(rep, st, "3x3 16->16 @8 flat", 7, 8, 8, 16, 0, 16, 3, 16, false);
    rc |= conv_case(rep, st, "1x1 64->16 @8 down", 7, 8, 8, 64, 0, 16, 1, 16, false);
    rc |= conv_case(rep, st, "3x3 32->32 @4", 37, 4, 4, 32, 0, 32, 3, 32, true);
    rc |= conv_case(rep, st, "3x3 64+32->64 @4", 37, 4, 4, 64, 32, 64, 3, 64, false);
    rc |= conv_case(rep, st, "3x3 64->128 @2", 70, 2, 2, 64, 0, 128, 3, 128, false);
    rc |= conv_case(rep, st, "3x3 128+64->128 @2", 70, 2, 2, 128, 64, 128, 3, 128, true);
    rc |= conv_case(rep, st, "1x1 128->384 @2 qkv", 70, 2, 2, 128, 0, 384, 1, 128, false);
    rc |= conv_case(rep, st, "3x3 32->16 @16 up", 300, 16, 16, 32, 0, 16, 3, 16, false);
    flush();
    if (rc) return -1;
    return rep.fails;
}
