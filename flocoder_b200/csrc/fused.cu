// Fused stage kernels (sm_100a): whole ResnetBlock chains and the attention blocks of one resolution
// level in ONE kernel each.  A CTA owns `nb` complete samples; activations never leave the SM inside a
// stage -- 16-bit operands live in shared memory in the blocked/padded layout that doubles as the tcgen05
// no-swizzle K-major operand (3x3 taps = shifted descriptors, torch.cat = two operand slots), fp32
// accumulators live in TMEM, weights are streamed through a shared-memory ring by bulk TMA copies.
//
// k_chain   [conv -> GroupNorm+FiLM+SiLU(+residual)]* with the ResnetBlock 1x1 res_conv accumulated into a
//           second TMEM region, the PreNorm of the following attention block fused into the last epilogue,
//           init_conv (unet.py:295) as the first epilogue of the first stage and final_conv + the RK4/Euler
//           /CFG stage update (unet.py:372, sampling.py:43-48,69-74) as the last epilogue of the last stage.
// k_attn    Residual(PreNorm(LinearAttention)) (unet.py:125-161): q/k/v 1x1 convs, both softmaxes,
//           context and output contractions and the to_out conv on tcgen05, GroupNorm + residual epilogue;
//           and the mid-block softmax attention (unet.py:99-122).
//
// Three decoupled loops per CTA (192 threads): warp 4 lane 0 = TMA producer (input tiles, weight ring),
// warp 5 lane 0 = tcgen05.mma issuer, warps 0-3 = epilogue (TMEM lane quadrants).  Steps alternate
// MMA(i) -> EPI(i) -> MMA(i+1) ... through two mbarriers (bar_mma: tcgen05.commit, bar_epi: 128 arrivals).
#include <cuda_fp16.h>

#include "flo_internal.h"
#include "umma_common.cuh"
#include "fused_common.cuh"

namespace flo {

// one thread's row of an M tile: padded pixel -> (sample, h, w)
struct RowInfo {
    int pp;        // flattened padded pixel index inside the CTA's planes
    int s, px;     // sample within the CTA, unpadded pixel index h*W+w
    int h, w;
    bool valid;
};
__device__ __forceinline__ int chain_tile_row0(const ChainParams& p, int t) {
    const int Wp = p.W + 2;
    if (p.strips) {
        const int tx_n = p.W >> 3;
        return ((t / tx_n) * 16 + 1) * Wp + 1 + 8 * (t % tx_n);
    }
    return Wp + 1 + t * 128;
}
__device__ __forceinline__ RowInfo chain_row(const ChainParams& p, int t, int r, int b0) {
    const int Wp = p.W + 2, PP = Wp * (p.H + 2);
    RowInfo ri;
    ri.pp = chain_tile_row0(p, t) + (r >> 3) * (p.strips ? Wp : 8) + (r & 7);
    ri.s = ri.pp / PP;
    const int rem = ri.pp - ri.s * PP;
    const int hh = rem / Wp, ww = rem - hh * Wp;
    ri.h = hh - 1; ri.w = ww - 1;
    ri.px = ri.h * p.W + ri.w;
    ri.valid = (ri.s < p.nb) && (b0 + ri.s < p.B) && hh >= 1 && hh <= p.H && ww >= 1 && ww <= p.W;
    return ri;
}

// issue one convolution (all taps, all K slices, all M tiles) whose weights arrive through the ring
struct RingState { int cc; };   // global chunk counter (producer and issuer advance in lock step)

__device__ __forceinline__ void issue_conv(const ChainParams& p, uint32_t smem_base, uint32_t tmem_base, uint32_t bar_full,
                                           uint32_t bar_empty, RingState& rs, int a0_off, int a0_ncb, int a1_off, int a1_ncb,
                                           int ksize, int n, int col, int n_chunks, int S) {
    const int Wp = p.W + 2;
    const uint32_t plane_bytes = (uint32_t)p.plane_px * 16u;
    const uint32_t a_sbo = (uint32_t)(p.strips ? Wp : 8) * 16u;
    const uint32_t idesc = make_idesc16(128, n, p.fmt, 0, 0);
    const int cpT = (a0_ncb + a1_ncb) >> 1;
    for (int ci = 0; ci < n_chunks; ++ci) {
        const int slot = rs.cc % p.n_ring;
        mbar_wait(bar_full + 8 * slot, (rs.cc / p.n_ring) & 1);
        tc_fence_after();
        const uint32_t bstage = smem_base + p.ring_off + slot * p.ring_slot_bytes;
        for (int s = 0; s < S; ++s) {
            const int ks = ci * S + s;
            const int tap = ks / cpT, cp = ks - tap * cpT;
            const int shift = (ksize == 3) ? ((tap / 3 - 1) * Wp + (tap % 3 - 1)) : 0;
            const uint32_t a_plane = (2 * cp < a0_ncb) ? (uint32_t)a0_off + (uint32_t)(2 * cp) * plane_bytes
                                                       : (uint32_t)a1_off + (uint32_t)(2 * cp - a0_ncb) * plane_bytes;
            const uint64_t bdesc = make_smem_desc(bstage + (uint32_t)s * (uint32_t)n * 32u, (uint32_t)n * 16u, 128u);
            for (int t = 0; t < p.n_mtiles; ++t) {
                const uint32_t a_addr = smem_base + a_plane + (uint32_t)(chain_tile_row0(p, t) + shift) * 16u;
                umma_bf16(tmem_base + (uint32_t)(col + t * n), make_smem_desc(a_addr, plane_bytes, a_sbo), bdesc, idesc,
                          ks > 0 ? 1u : 0u);
            }
        }
        umma_commit(bar_empty + 8 * slot);
        ++rs.cc;
    }
}
__device__ __forceinline__ void stream_weights(const ChainParams& p, uint32_t smem_base, uint32_t bar_full, uint32_t bar_empty,
                                               RingState& rs, const uint16_t* w, int n_chunks, int chunk_bytes) {
    for (int ci = 0; ci < n_chunks; ++ci) {
        const int slot = rs.cc % p.n_ring;
        if (rs.cc >= p.n_ring) mbar_wait(bar_empty + 8 * slot, ((rs.cc / p.n_ring) - 1) & 1);
        mbar_expect_tx(bar_full + 8 * slot, (uint32_t)chunk_bytes);
        bulk_load_1d(smem_base + p.ring_off + slot * p.ring_slot_bytes,
                     reinterpret_cast<const uint8_t*>(w) + (size_t)ci * chunk_bytes, (uint32_t)chunk_bytes, bar_full + 8 * slot);
        ++rs.cc;
    }
}

// write one 16-channel chunk of an output row to every requested destination
__device__ __forceinline__ void write_outputs(const ChainParams& p, const ChainStep& st, uint8_t* smem, const RowInfo& ri, int b,
                                              int c16, const float* v) {
    const int HW = p.H * p.W;
    const uint32_t plane_bytes = (uint32_t)p.plane_px * 16u;
    const int ncb = st.C >> 3;
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
        const int cb = (c16 >> 3) + hb;
        const uint4 u = pack8(v + hb * 8, p.fmt);
        if (st.out_slot_off >= 0)
            *reinterpret_cast<uint4*>(smem + st.out_slot_off + (uint32_t)cb * plane_bytes + (uint32_t)ri.pp * 16u) = u;
        if (st.out_g >= 0)
            reinterpret_cast<uint4*>(p.gt[st.out_g])[(size_t)(cb * p.B + b) * HW + ri.px] = u;
        if (st.out_un_g >= 0) {     // 'b c (h p1) (w p2) -> b (c p1 p2) h w' with our channel order (p1 p2 c)  (unet.py:52)
            const int plane = ((ri.h & 1) * 2 + (ri.w & 1)) * ncb + cb;
            const int q = (ri.h >> 1) * (p.W >> 1) + (ri.w >> 1);
            reinterpret_cast<uint4*>(p.gt[st.out_un_g])[(size_t)(plane * p.B + b) * (HW >> 2) + q] = u;
        }
        if (st.out_up_g >= 0) {     // nearest x2 (unet.py:44)
            const int W2 = p.W * 2;
            uint4* dst = reinterpret_cast<uint4*>(p.gt[st.out_up_g]) + (size_t)(cb * p.B + b) * (HW * 4);
#pragma unroll
            for (int d = 0; d < 4; ++d) dst[(2 * ri.h + (d >> 1)) * W2 + 2 * ri.w + (d & 1)] = u;
        }
    }
}

// deterministic per-(sample, group) reduction of per-row (sum, sumsq) pairs held in shared memory
//   rowstat[row * G + g], rows = n_mtiles*128;  result stat[s*G+g] = (mean, rstd)
__device__ void reduce_stats(const ChainParams& p, float2* rowstat, float2* partial, float2* stat, int G, float count,
                             int tid) {
    const int Wp = p.W + 2, PP = Wp * (p.H + 2);
    const int R = p.n_mtiles * 128;
    const int combos = p.nb * G;
    int parts = 1;
    while (parts * 2 * combos <= EPI_THREADS && parts < 16) parts *= 2;
    epi_sync();
    if (tid < combos * parts) {
        const int part = tid % parts, sg = tid / parts, g = sg % G, s = sg / G;
        int lo = 0, hi = R;
        if (!p.strips) {
            lo = min(max(s * PP - (Wp + 1), 0), R);
            hi = min(max((s + 1) * PP - (Wp + 1), 0), R);
        }
        const int per = (hi - lo + parts - 1) / parts;
        const int a = lo + part * per, b = min(hi, a + per);
        float sx = 0.f, sq = 0.f;
        for (int r = a; r < b; ++r) { const float2 v = rowstat[r * G + g]; sx += v.x; sq += v.y; }
        partial[tid] = make_float2(sx, sq);
    }
    epi_sync();
    if (tid < combos) {
        float sx = 0.f, sq = 0.f;
        for (int k = 0; k < parts; ++k) { const float2 v = partial[tid * parts + k]; sx += v.x; sq += v.y; }
        const float mean = sx / count;
        const float var = fmaxf(sq / count - mean * mean, 0.f);
        stat[tid] = make_float2(mean, 1.0f / sqrtf(var + 1e-5f));
    }
    epi_sync();
}

// ------------------------------------------------------------------------------------------------
// k_chain
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FUSED_THREADS) k_chain(const __grid_constant__ CUtensorMap tm0,
                                                         const __grid_constant__ CUtensorMap tm1,
                                                         const __grid_constant__ CUtensorMap tm2,
                                                         const __grid_constant__ CUtensorMap tm3,
                                                         const __grid_constant__ ChainParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_full = smem_base + p.bar_off;
    const uint32_t bar_empty = bar_full + 8 * MAX_WSTAGES;
    const uint32_t bar_load = bar_empty + 8 * MAX_WSTAGES;
    const uint32_t bar_mma = bar_load + 8;
    const uint32_t bar_epi = bar_mma + 8;
    const uint32_t tmem_slot = bar_epi + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + p.bar_off + 16 * MAX_WSTAGES + 24);
    const int b0 = blockIdx.x * p.nb;
    const uint32_t plane_bytes = (uint32_t)p.plane_px * 16u;

    if (warp == 4 && lane == 0) {
        for (int i = 0; i < p.n_ring; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        mbar_init(bar_load, 1);
        mbar_init(bar_mma, 1);
        mbar_init(bar_epi, EPI_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    // clear the slots the epilogues write (their halo must read as zero)
    for (int i = tid * 16; i < p.zero_bytes; i += FUSED_THREADS * 16)
        *reinterpret_cast<uint4*>(smem + p.zero_off + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 4) {
        // ============================ producer ============================
        if (lane == 0) {
            if (p.n_loads > 0) {
                uint32_t bytes = 0;
                for (int i = 0; i < p.n_loads; ++i) bytes += (uint32_t)p.load_ncb[i] * plane_bytes;
                mbar_expect_tx(bar_load, bytes);
                const CUtensorMap* maps[4] = {&tm0, &tm1, &tm2, &tm3};
                for (int i = 0; i < p.n_loads; ++i) tma_load_5d(smem_base + p.load_off[i], maps[i], bar_load, 0, -1, -1, b0, 0);
            }
            RingState rs{0};
            for (int i = 0; i < p.n_steps; ++i) {
                const ChainStep& st = p.st[i];
                if (!st.has_conv) continue;
                stream_weights(p, smem_base, bar_full, bar_empty, rs, p.wblob + st.w_off, st.n_chunks,
                               st.slices_per_chunk * st.n * 32);
                if (st.has_res)
                    stream_weights(p, smem_base, bar_full, bar_empty, rs, p.wblob + st.wres_off, st.res_chunks,
                                   st.res_slices_per_chunk * st.n * 32);
            }
        }
    } else if (warp == 5) {
        // ============================ MMA issuer ============================
        if (lane == 0) {
            RingState rs{0};
            if (p.n_loads > 0) mbar_wait(bar_load, 0);
            for (int i = 0; i < p.n_steps; ++i) {
                const ChainStep& st = p.st[i];
                if (i > 0) mbar_wait(bar_epi, (i - 1) & 1);
                tc_fence_after();
                if (st.has_conv) {
                    issue_conv(p, smem_base, tmem_base, bar_full, bar_empty, rs, st.a0_off, st.a0_ncb, st.a1_off, st.a1_ncb,
                               st.ksize, st.n, st.acc_col, st.n_chunks, st.slices_per_chunk);
                    if (st.has_res)
                        issue_conv(p, smem_base, tmem_base, bar_full, bar_empty, rs, st.a0_off, st.a0_ncb, st.a1_off, st.a1_ncb,
                                   1, st.n, st.res_col, st.res_chunks, st.res_slices_per_chunk);
                    umma_commit(bar_mma);
                } else {
                    mbar_arrive(bar_mma);
                }
            }
        }
    } else {
        // ============================ epilogue warps ============================
        const int r = warp * 32 + lane;                       // row inside every M tile == TMEM lane
        const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
        float2* rowstat = reinterpret_cast<float2*>(smem + p.stats_off);
        float2* partial = rowstat + p.n_mtiles * 128 * 8;
        float2* stat = partial + EPI_THREADS;
        const Ctrl* ctrl = p.ctrl;
        const int HW = p.H * p.W;
        if (p.n_loads > 0) mbar_wait(bar_load, 0);
        for (int i = 0; i < p.n_steps; ++i) {
            const ChainStep& st = p.st[i];
            mbar_wait(bar_mma, i & 1);
            tc_fence_after();
            const int C = st.C;
            if (st.epi == CE_INIT) {
                // ---- init_conv 1x1 from the NCHW fp32 integrator state (unet.py:295)
                const float* xs = ctrl->xs;
                const float* w = p.fblob + p.init_w_off;
                const float* bias = p.fblob + p.init_b_off;
                for (int t = 0; t < p.n_mtiles; ++t) {
                    const RowInfo ri = chain_row(p, t, r, b0);
                    if (!ri.valid) continue;
                    const int b = b0 + ri.s;
                    float xin[16];
                    for (int ci = 0; ci < p.cin0; ++ci) xin[ci] = xs[((size_t)b * p.cin0 + ci) * HW + ri.px];
                    for (int c16 = 0; c16 < C; c16 += 16) {
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float a = 0.f;
                            for (int ci = 0; ci < p.cin0; ++ci) a = fmaf(xin[ci], w[(c16 + j) * p.cin0 + ci], a);
                            v[j] = a + bias[c16 + j];
                        }
                        write_outputs(p, st, smem, ri, b, c16, v);
                    }
                }
            } else if (st.epi == CE_BIAS) {
                // ---- conv + bias (+ residual from a shared-memory slot)
                const float* bias = p.fblob + st.bias_off;
                for (int t = 0; t < p.n_mtiles; ++t) {
                    const RowInfo ri = chain_row(p, t, r, b0);
                    const int b = b0 + ri.s;
                    for (int c16 = 0; c16 < C; c16 += 16) {
                        float v[16];
                        tmem_ld16(tlane + (uint32_t)(st.acc_col + t * C + c16), v);
                        if (!ri.valid) continue;
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += bias[c16 + j];
                        if (st.res_mode == 2) {
                            float rr[16];
                            const uint8_t* src = smem + st.res_slot_off + (uint32_t)(c16 >> 3) * plane_bytes + (uint32_t)ri.pp * 16u;
                            unpack8(*reinterpret_cast<const uint4*>(src), rr, p.fmt);
                            unpack8(*reinterpret_cast<const uint4*>(src + plane_bytes), rr + 8, p.fmt);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] += rr[j];
                        }
                        write_outputs(p, st, smem, ri, b, c16, v);
                    }
                }
            } else {
                // ---- conv + bias -> GroupNorm -> FiLM -> SiLU -> + residual     (unet.py:64-73,96)
                const int G = st.groups, cpg = C / G;
                const float* bias = p.fblob + st.bias_off;
                const float* gamma = p.fblob + st.gamma_off;
                const float* beta = p.fblob + st.beta_off;
                // pass 1: per-row (sum, sumsq) per group
                for (int t = 0; t < p.n_mtiles; ++t) {
                    const RowInfo ri = chain_row(p, t, r, b0);
                    float2* rs_row = rowstat + (size_t)(t * 128 + r) * G;
                    if (cpg >= 16) {
                        for (int g = 0; g < G; ++g) {
                            float sx = 0.f, sq = 0.f;
                            for (int c16 = g * cpg; c16 < (g + 1) * cpg; c16 += 16) {
                                float v[16];
                                tmem_ld16(tlane + (uint32_t)(st.acc_col + t * C + c16), v);
#pragma unroll
                                for (int j = 0; j < 16; ++j) { const float x = v[j] + bias[c16 + j]; sx += x; sq += x * x; }
                            }
                            rs_row[g] = ri.valid ? make_float2(sx, sq) : make_float2(0.f, 0.f);
                        }
                    } else {
                        for (int c16 = 0; c16 < C; c16 += 16) {
                            float v[16];
                            tmem_ld16(tlane + (uint32_t)(st.acc_col + t * C + c16), v);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] += bias[c16 + j];
                            if (cpg == 4) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    float sx = 0.f, sq = 0.f;
#pragma unroll
                                    for (int j = 0; j < 4; ++j) { const float x = v[q * 4 + j]; sx += x; sq += x * x; }
                                    rs_row[(c16 >> 2) + q] = ri.valid ? make_float2(sx, sq) : make_float2(0.f, 0.f);
                                }
                            } else {   // cpg == 8
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    float sx = 0.f, sq = 0.f;
#pragma unroll
                                    for (int j = 0; j < 8; ++j) { const float x = v[q * 8 + j]; sx += x; sq += x * x; }
                                    rs_row[(c16 >> 3) + q] = ri.valid ? make_float2(sx, sq) : make_float2(0.f, 0.f);
                                }
                            }
                        }
                    }
                }
                reduce_stats(p, rowstat, partial, stat, G, (float)(cpg * HW), r);
                // pass 2: normalise, modulate, activate, add residual, write
                const Stage sg = ctrl->stages[ctrl->step];
                for (int t = 0; t < p.n_mtiles; ++t) {
                    float kacc[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) kacc[j] = 0.f;
                    const RowInfo ri = chain_row(p, t, r, b0);
                    const int b = b0 + ri.s;
                    const float* film = nullptr;
                    if (st.film_off >= 0 && ri.valid) {
                        const int row = ctrl->film_per_sample ? b : sg.film_row;
                        film = ctrl->film + (size_t)row * p.film_dim + st.film_off;
                    }
                    float psx = 0.f, psq = 0.f;
                    for (int c16 = 0; c16 < C; c16 += 16) {
                        float v[16], rr[16];
                        tmem_ld16(tlane + (uint32_t)(st.acc_col + t * C + c16), v);
                        if (st.res_mode == 1) tmem_ld16(tlane + (uint32_t)(st.res_col + t * C + c16), rr);
                        if (!ri.valid) continue;
                        if (st.res_mode == 1) {
                            const float* rb = p.fblob + st.res_bias_off;
#pragma unroll
                            for (int j = 0; j < 16; ++j) rr[j] += rb[c16 + j];
                        } else if (st.res_mode == 2) {
                            const uint8_t* src = smem + st.res_slot_off + (uint32_t)(c16 >> 3) * plane_bytes + (uint32_t)ri.pp * 16u;
                            unpack8(*reinterpret_cast<const uint4*>(src), rr, p.fmt);
                            unpack8(*reinterpret_cast<const uint4*>(src + plane_bytes), rr + 8, p.fmt);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int c = c16 + j;
                            const float2 ms = stat[ri.s * G + c / cpg];
                            float y = (v[j] + bias[c] - ms.x) * ms.y * gamma[c] + beta[c];
                            if (film) y = y * (film[c] + 1.0f) + film[C + c];
                            if (st.silu) y = y / (1.0f + __expf(-y));
                            if (st.res_mode) y += rr[j];
                            v[j] = y;
                            psx += y; psq += y * y;
                        }
                        if (st.final) {
                            const float* fw = p.fblob + p.final_w_off;
#pragma unroll
                            for (int co = 0; co < 16; ++co) {
                                if (co < p.channels) {
                                    float a = kacc[co];
#pragma unroll
                                    for (int j = 0; j < 16; ++j) a = fmaf(v[j], fw[co * p.dim + c16 + j], a);
                                    kacc[co] = a;
                                }
                            }
                        } else {
                            write_outputs(p, st, smem, ri, b, c16, v);
                        }
                    }
                    if (st.pn_g >= 0) rowstat[(size_t)(t * 128 + r)] = ri.valid ? make_float2(psx, psq) : make_float2(0.f, 0.f);
                    if (st.final && ri.valid) {
                        // ---- final_conv bias + RK4 / Euler / CFG stage update (sampling.py:43-48,69-74)
                        Ctrl* c = p.ctrl;
                        const float* fb = p.fblob + p.final_b_off;
                        const size_t plane = (size_t)p.B * p.channels * HW;
#pragma unroll
                        for (int co = 0; co < 16; ++co) {
                            if (co >= p.channels) continue;
                            float k = kacc[co] + fb[co];
                            const size_t o = ((size_t)b * p.channels + co) * HW + ri.px;
                            if (sg.flags & SF_CFG_COMBINE) k = __fadd_rn(k, __fmul_rn(c->cfg, __fsub_rn(c->vcond[o], k)));
                            if (c->vtrace && sg.eval_idx >= 0) c->vtrace[(size_t)sg.eval_idx * plane + o] = k;
                            switch (sg.kind) {
                                case ST_PLAIN: c->vout[o] = k; break;
                                case ST_CFG_COND: c->vcond[o] = k; break;
                                case ST_RK1:
                                    c->acc[o] = k;
                                    c->xs[o] = __fadd_rn(c->y[o], __fmul_rn(__fmul_rn(sg.dt, k), 0.5f));
                                    break;
                                case ST_RK2:
                                    c->acc[o] = __fadd_rn(c->acc[o], __fmul_rn(2.0f, k));
                                    c->xs[o] = __fadd_rn(c->y[o], __fmul_rn(__fmul_rn(sg.dt, k), 0.5f));
                                    break;
                                case ST_RK3:
                                    c->acc[o] = __fadd_rn(c->acc[o], __fmul_rn(2.0f, k));
                                    c->xs[o] = __fadd_rn(c->y[o], __fmul_rn(sg.dt, k));
                                    break;
                                case ST_RK4: {
                                    const float yn = __fadd_rn(c->y[o], __fmul_rn(sg.dt6, __fadd_rn(c->acc[o], k)));
                                    c->y[o] = yn; c->xs[o] = yn;
                                } break;
                                case ST_EULER: {
                                    const float yn = __fadd_rn(c->y[o], __fmul_rn(k, sg.dt));
                                    c->y[o] = yn; c->xs[o] = yn;
                                } break;
                                default: break;
                            }
                        }
                    }
                }
                if (st.pn_g >= 0) {
                    // ---- fused PreNorm of the following attention block: GroupNorm(1, C) of the 16-bit result
                    reduce_stats(p, rowstat, partial, stat, 1, (float)(C * HW), r);
                    const float* g2 = p.fblob + st.pn_gamma_off;
                    const float* b2 = p.fblob + st.pn_beta_off;
                    for (int t = 0; t < p.n_mtiles; ++t) {
                        const RowInfo ri = chain_row(p, t, r, b0);
                        if (!ri.valid) continue;
                        const int b = b0 + ri.s;
                        const float2 ms = stat[ri.s];
                        for (int cb = 0; cb < (C >> 3); ++cb) {
                            float x[8];
                            unpack8(*reinterpret_cast<const uint4*>(smem + st.out_slot_off + (uint32_t)cb * plane_bytes +
                                                                    (uint32_t)ri.pp * 16u), x, p.fmt);
#pragma unroll
                            for (int j = 0; j < 8; ++j) x[j] = (x[j] - ms.x) * ms.y * g2[cb * 8 + j] + b2[cb * 8 + j];
                            reinterpret_cast<uint4*>(p.gt[st.pn_g])[(size_t)(cb * p.B + b) * HW + ri.px] = pack8(x, p.fmt);
                        }
                    }
                }
                if (st.final) {
                    // the last CTA to finish advances the stage counter (every CTA has read ctrl->step by now)
                    epi_sync();
                    if (r == 0) {
                        Ctrl* c = p.ctrl;
                        __threadfence();
                        const int done = atomicAdd(&c->done_ctr, 1);
                        if (done == (int)gridDim.x - 1) {
                            c->done_ctr = 0;
                            c->step = c->step + 1;
                            __threadfence();
                        }
                    }
                }
            }
            fence_proxy_async();      // shared-memory results -> visible to the next step's tcgen05.mma
            tc_fence_before();
            mbar_arrive(bar_epi);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

cudaError_t attn_configure();
cudaError_t fused_configure() {
    cudaError_t e = cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    return attn_configure();
}

cudaError_t launch_chain(const ChainParams& p, const CUtensorMap* maps, int grid, cudaStream_t s) {
    k_chain<<<grid, FUSED_THREADS, p.smem_bytes, s>>>(maps[0], maps[1], maps[2], maps[3], p);
    return cudaGetLastError();
}

}  // namespace flo
