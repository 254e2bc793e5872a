// Fused backward of a pointwise (1x1) ConvBlock, bf16 NHWC: the BatchNorm-backward elementwise pass, conv
// backward-data (+ the residual skip gradient), conv backward-weight and the BatchNorm-backward REDUCTIONS of the block
// that produced this block's input, in ONE pass over
//     G (gradient w.r.t. this block's ReLU output, M x Cout), Z (this block's raw conv output, M x Cout),
//     X (input, M x Cin: raw output of the producing block + its BN scale / shift, or a materialised activation)
//  -> dX (M x Cin), dW (Cout x Cin), sum(dX'), sum(dX' x).
// Replaces, for the MBConv expand / project blocks of the 112x112 and 56x56 stages (src/models/mnasnet.py:116-129 under
// autograd): threshold_backward + native_batch_norm_backward + convolution_backward, 2*Cout + 2*Cin channel passes
// instead of 5*Cout + 2*Cin (+ 2*Cin for the producer's reduction).  Math contract: SURVEY.md appendix F.
//
// A 1x1 conv has no spatial structure: the tensors are [M][C] matrices.  A CTA walks row tiles of R = 16 * WARPS rows;
// per tile one thread issues the TMA boxes (double-buffered: the next tile's boxes are in flight during this tile's
// math), every thread turns its share of (G, Z) vectors into dZ = a*G*[scale*Z+shift>0] + b*Z + c in place, then each
// warp owns 16 rows:
//   backward-data  dX[16 x Cin] = dZ[16 x Cout] W          A = dZ via ldmatrix, B = weight fragments in registers
//   backward-weight dW[Cout x Cin] += dZ^T[Cout x 16] A_x   both operands via ldmatrix.trans (the reduction index,
//                  the pixel row, is the slow index of both tiles); A_x = relu(s_in*x+t_in) applied to the fragments;
//                  accumulators stay in registers for the CTA's life
// and the finished dX rows are staged for the TMA store and reduced against the raw X for the producing block.
#include <algorithm>

#include "dw_mma.cuh"

namespace mnb {

typedef unsigned long long pf2_t;
__device__ __forceinline__ pf2_t pf2_pack(float lo, float hi) {
    pf2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void pf2_unpack(pf2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ pf2_t pf2_fma(pf2_t a, pf2_t b, pf2_t c) {
    pf2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ pf2_t pf2_from_bf16x2(uint32_t u) { return pf2_pack(bf_lo(u), bf_hi(u)); }
__device__ __forceinline__ uint32_t pf2_relu_bf16x2(pf2_t v) {
    float lo, hi;
    pf2_unpack(v, lo, hi);
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

struct PwbP {
    const float* scale;         // this block's BN scale / shift (ReLU mask)
    const float* shift;
    const double* sums;         // [2][Cout] sum(G'), sum(G' z) of this block
    const float* mean;
    const float* invstd;
    double m;
    float* dgamma;              // += (NULL = frozen)
    float* dbeta;
    float* dbias;
    const float* in_scale;      // producing block's BN scale / shift (NULL = X is the activation itself)
    const float* in_shift;
    const float* w;             // [Cout][Cin] fp32
    float* dw;                  // [Cout][Cin] += (NULL = frozen)
    double* nsums;              // [2][Cin] reductions for the producing block (NULL = none)
    int has_add;                // residual skip gradient added into dX
    int cin_total;              // Cin of the layer; a CTA column (blockIdx.y) owns the CI-channel slice blockIdx.y * CI
    long long M;
};

template <int CO, int CI>
struct PwbCfg {
    static constexpr int NCO = CO / 8, NCI = CI / 8;
    // a thread keeps one 8-channel chunk of the (G, Z) tile for life: THREADS is a multiple of 32 and of NCO
    static constexpr int THREADS = (NCO == 9) ? 288 : ((NCO == 5) ? 160 : 96);
    static constexpr int WARPS = THREADS / 32;
    static constexpr int R = 16 * WARPS;                    // rows per tile
    static constexpr int KS16 = CO / 16, KS8 = (CO % 16) / 8;     // backward-data k-steps (k = Cout)
    static constexpr int MT = (CO + 15) / 16;               // backward-weight m-tiles (m = Cout)
    static constexpr int GB = R * CO * 2, XBB = R * CI * 2;
    static constexpr int STAGE = 2 * GB + 2 * XBB;          // G, Z, X, skip
    static constexpr int CO_PAD = MT * 16;
    static constexpr int SMEM = 2 * STAGE + XBB /*dX staging*/ + CO_PAD * CI * 4 /*dW*/ + 7 * CO * 4 + 2 * CI * 4 +
                                2 * CI * 4 + 64;
};

template <int CO, int CI>
__global__ void __launch_bounds__(PwbCfg<CO, CI>::THREADS)
    pw_bwd_fused_k(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_z,
                   const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_add,
                   const __grid_constant__ CUtensorMap tm_dx, const PwbP p) {
    using Cfg = PwbCfg<CO, CI>;
    constexpr int NCO = Cfg::NCO, NCI = Cfg::NCI, THREADS = Cfg::THREADS, R = Cfg::R, MT = Cfg::MT;
    constexpr int KS16 = Cfg::KS16, KS8 = Cfg::KS8, GB = Cfg::GB, XBB = Cfg::XBB, STAGE = Cfg::STAGE;
    constexpr int GP = CO * 2, XP = CI * 2;                 // row pitches in bytes
    extern __shared__ __align__(128) unsigned char dsm[];
    const uint32_t S0 = smem_u32(dsm);
    const uint32_t OUT = S0 + 2 * STAGE;
    float* dwacc = reinterpret_cast<float*>(dsm + 2 * STAGE + XBB);          // [CO_PAD][CI]
    float* coef = dwacc + Cfg::CO_PAD * CI;                                  // a, b, c, scale, shift [5][CO] (+2 spare)
    float* icoef = coef + 7 * CO;                                            // in_scale, in_shift [2][CI]
    float* red = icoef + 2 * CI;                                             // [2][CI]
    const uint32_t bar0 = smem_u32(red + 2 * CI);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const bool act_in = p.in_scale != nullptr;
    const bool do_red = p.nsums != nullptr && act_in, do_wgrad = p.dw != nullptr;
    // Input-channel slicing for wide inputs (240 -> 40: five 48-channel slices): backward-data, backward-weight, the
    // skip add and the producer reductions are all independent per input channel; only the (narrow) G / Z tiles are
    // re-read by every slice, from L2.
    const int ci0 = blockIdx.y * CI, CIT = p.cin_total;

    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    for (int i = tid; i < Cfg::CO_PAD * CI; i += THREADS) dwacc[i] = 0.f;
    for (int c = tid; c < CO; c += THREADS) {
        // the per-channel work of BatchNorm backward, redundantly per CTA (as bn_bwd_apply_fused_k)
        const double sg = p.sums[c], sgz = p.sums[CO + c];
        const double mean = p.mean[c], invstd = p.invstd[c], s = p.scale[c];
        const double dga = invstd * (sgz - mean * sg);
        const double bb = -s * invstd * dga / p.m;
        const double cc = -s * sg / p.m - bb * mean;
        coef[c] = (float)s; coef[CO + c] = (float)bb; coef[2 * CO + c] = (float)cc;
        coef[3 * CO + c] = p.scale[c]; coef[4 * CO + c] = p.shift[c];
        if (blockIdx.x == 0 && blockIdx.y == 0) {
            if (p.dgamma) p.dgamma[c] += (float)dga;
            if (p.dbeta) p.dbeta[c] += (float)sg;
            if (p.dbias) p.dbias[c] += (float)(s * sg + bb * mean * p.m + cc * p.m);   // analytically 0
        }
    }
    for (int c = tid; c < CI; c += THREADS) {
        icoef[c] = act_in ? p.in_scale[ci0 + c] : 1.f;
        icoef[CI + c] = act_in ? p.in_shift[ci0 + c] : 0.f;
        red[c] = 0.f; red[CI + c] = 0.f;
    }
    // backward-data B fragments: B[k = co][n = ci] = w[co][ci]; register (ks, nt, h) = rows k = 16ks + 8h + 2t, +1 of
    // column n = 8nt + g
    uint32_t wb[KS16 + KS8][NCI][2];
#pragma unroll
    for (int ks = 0; ks < KS16 + KS8; ++ks)
#pragma unroll
        for (int nt = 0; nt < NCI; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int co = 16 * ks + 8 * h + 2 * t, ci = 8 * nt + g;
                float w0 = 0.f, w1 = 0.f;
                if (co + 1 < CO) { w0 = p.w[co * CIT + ci0 + ci]; w1 = p.w[(co + 1) * CIT + ci0 + ci]; }
                wb[ks][nt][h] = pack_bf16x2(w0, w1);
            }
    // transform mapping: one 8-channel chunk of (G, Z) per thread for life
    const int tchunk = tid % NCO, trow0 = tid / NCO;
    constexpr int TRS = THREADS / NCO;
    // ldmatrix lane addresses inside a 16-row group (relative to the group's first row)
    const int mi = lane >> 3, r8 = lane & 7;
    // A of backward-data (x4: (rows 0-7, ch 0-7), (rows 8-15, ch 0-7), (rows 0-7, ch 8-15), (rows 8-15, ch 8-15))
    const uint32_t a_dg = (uint32_t)(((mi & 1) * 8 + r8) * GP + (mi >> 1) * 16);
    // A of backward-weight, transposed (x4: (rows 0-7, co 0-7), (rows 0-7, co 8-15), (rows 8-15, co 0-7), (rows 8-15, co 8-15))
    const uint32_t a_wg = (uint32_t)(((mi >> 1) * 8 + r8) * GP + (mi & 1) * 16);
    // B of backward-weight, transposed (x4: n-tiles (nt, nt+1) x row halves; x2 for a last odd n-tile)
    const uint32_t b_wg = (uint32_t)(((mi & 1) * 8 + r8) * XP + (mi >> 1) * 16);
    float wacc[MT][NCI][4];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < NCI; ++b) wacc[a][b][0] = wacc[a][b][1] = wacc[a][b][2] = wacc[a][b][3] = 0.f;
    pf2_t rs[NCI][2];           // per n-tile: (sum dX', sum dX' x) of the lane's channel pair
#pragma unroll
    for (int b = 0; b < NCI; ++b) rs[b][0] = rs[b][1] = pf2_pack(0.f, 0.f);
    uint32_t ph = 0;
    __syncthreads();

    const long long ntiles = (p.M + R - 1) / R;
    auto issue = [&](long long tile, int b) {
        const uint32_t sb = S0 + b * STAGE;
        const int row0 = (int)(tile * R);           // M < 2^31 rows (checked by the launcher)
        mbar_expect_tx(bar0 + 8 * b, (uint32_t)(2 * GB + XBB + (p.has_add ? XBB : 0)));
        tma_load4(sb, &tm_g, 0, row0, 0, 0, bar0 + 8 * b);
        tma_load4(sb + GB, &tm_z, 0, row0, 0, 0, bar0 + 8 * b);
        tma_load4(sb + 2 * GB, &tm_x, ci0, row0, 0, 0, bar0 + 8 * b);
        if (p.has_add) tma_load4(sb + 2 * GB + XBB, &tm_add, ci0, row0, 0, 0, bar0 + 8 * b);
    };
    long long tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) issue(tile, 0);
    int b = 0;
    for (; tile < ntiles; tile += gridDim.x) {
        const long long nxt = tile + gridDim.x;
        if (tid == 0) {
            tma_store_wait_read();                  // the previous tile's dX store has finished reading OUT
            if (nxt < ntiles) issue(nxt, b ^ 1);    // stage b^1 was released by the previous tile's last barrier
        }
        const uint32_t sb = S0 + b * STAGE;
        const uint32_t GBs = sb, ZBs = sb + GB, XBs = sb + 2 * GB, ABs = sb + 2 * GB + XBB;
        const long long row0 = tile * R;
        mbar_wait(bar0 + 8 * b, (ph >> b) & 1);
        ph ^= 1u << b;
        {   // dZ = a * G * [scale*Z + shift > 0] + b * Z + c in place over G; rows beyond M stay zero
            pf2_t ca[4], cb[4], cc[4], cs[4], ct[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = tchunk * 8 + 2 * i;
                ca[i] = pf2_pack(coef[c], coef[c + 1]);
                cb[i] = pf2_pack(coef[CO + c], coef[CO + c + 1]);
                cc[i] = pf2_pack(coef[2 * CO + c], coef[2 * CO + c + 1]);
                cs[i] = pf2_pack(coef[3 * CO + c], coef[3 * CO + c + 1]);
                ct[i] = pf2_pack(coef[4 * CO + c], coef[4 * CO + c + 1]);
            }
#pragma unroll 2
            for (int r = trow0; r < R; r += TRS) {
                if (row0 + r < p.M) {
                    const uint32_t o = (uint32_t)(r * GP + tchunk * 16);
                    const uint4 ug = lds128(GBs + o), uz = lds128(ZBs + o);
                    const uint32_t gg[4] = {ug.x, ug.y, ug.z, ug.w}, zz[4] = {uz.x, uz.y, uz.z, uz.w};
                    uint32_t q[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const pf2_t z2 = pf2_from_bf16x2(zz[i]);
                        float y0, y1, g0, g1, d0, d1;
                        pf2_unpack(pf2_fma(cs[i], z2, ct[i]), y0, y1);
                        pf2_unpack(pf2_from_bf16x2(gg[i]), g0, g1);
                        const pf2_t gm = pf2_pack(y0 > 0.f ? g0 : 0.f, y1 > 0.f ? g1 : 0.f);
                        pf2_unpack(pf2_fma(ca[i], gm, pf2_fma(cb[i], z2, cc[i])), d0, d1);
                        q[i] = pack_bf16x2(d0, d1);
                    }
                    sts128(GBs + o, make_uint4(q[0], q[1], q[2], q[3]));
                }
            }
        }
        __syncthreads();
        if (row0 + 16 * warp < p.M) {
            const uint32_t grow = GBs + (uint32_t)(16 * warp * GP);
            const uint32_t xrow = XBs + (uint32_t)(16 * warp * XP);
            // ---- backward-data: dX[16 x CI] = dZ[16 x CO] W ----
            {
                float acc[NCI][4];
#pragma unroll
                for (int nt = 0; nt < NCI; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
                for (int ks = 0; ks < KS16; ++ks) {
                    uint32_t a0, a1, a2, a3;
                    ldsm4(grow + a_dg + ks * 32, a0, a1, a2, a3);
#pragma unroll
                    for (int nt = 0; nt < NCI; ++nt) mma16816(acc[nt], a0, a1, a2, a3, wb[ks][nt][0], wb[ks][nt][1]);
                }
                if constexpr (KS8 == 1) {
                    uint32_t a0, a1;
                    ldsm2(grow + a_dg + KS16 * 32, a0, a1);         // lanes 0-15: (rows 0-7 | 8-15, ch 16*KS16 ..)
#pragma unroll
                    for (int nt = 0; nt < NCI; ++nt) mma1688(acc[nt], a0, a1, wb[KS16][nt][0]);
                }
                // epilogue: + skip gradient, stage for the TMA store, reduce against the raw input
                const bool v0 = row0 + 16 * warp + g < p.M, v1 = row0 + 16 * warp + g + 8 < p.M;
                const pf2_t one = pf2_pack(1.f, 1.f);
#pragma unroll
                for (int nt = 0; nt < NCI; ++nt) {
                    const uint32_t o0 = (uint32_t)((16 * warp + g) * XP + nt * 16 + t * 4), o1 = o0 + 8 * XP;
                    float d0 = acc[nt][0], d1 = acc[nt][1], d2 = acc[nt][2], d3 = acc[nt][3];
                    if (p.has_add) {
                        const uint32_t s0 = lds32(ABs + o0), s1 = lds32(ABs + o1);
                        d0 += bf_lo(s0); d1 += bf_hi(s0); d2 += bf_lo(s1); d3 += bf_hi(s1);
                    }
                    const uint32_t u0 = pack_bf16x2(d0, d1), u1 = pack_bf16x2(d2, d3);
                    sts32(OUT + o0, u0);
                    sts32(OUT + o1, u1);
                    if (do_red) {
                        const int c = 8 * nt + 2 * t;
                        const pf2_t sp = pf2_pack(icoef[c], icoef[c + 1]), tp = pf2_pack(icoef[CI + c], icoef[CI + c + 1]);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const pf2_t x = pf2_from_bf16x2(lds32(XBs + (h ? o1 : o0)));
                            float y0, y1, q0, q1;
                            pf2_unpack(pf2_fma(sp, x, tp), y0, y1);
                            pf2_unpack(pf2_from_bf16x2(h ? u1 : u0), q0, q1);
                            const bool v = h ? v1 : v0;
                            const pf2_t q = pf2_pack(v && y0 > 0.f ? q0 : 0.f, v && y1 > 0.f ? q1 : 0.f);
                            rs[nt][0] = pf2_fma(q, one, rs[nt][0]);
                            rs[nt][1] = pf2_fma(q, x, rs[nt][1]);
                        }
                    }
                }
            }
            // ---- backward-weight: dW[CO x CI] += dZ^T[CO x 16] A_x[16 x CI] ----
            if (do_wgrad) {
                uint32_t bx[NCI][2];
#pragma unroll
                for (int nt = 0; nt < NCI; nt += 2) {
                    if (nt + 1 < NCI) ldsm4t(xrow + b_wg + nt * 16, bx[nt][0], bx[nt][1], bx[nt + 1][0], bx[nt + 1][1]);
                    else ldsm2t(xrow + b_wg + nt * 16, bx[nt][0], bx[nt][1]);
                }
                // rows of the tile beyond M are zero in dZ (never transformed) -> no masking of A_x needed
                if (act_in) {
#pragma unroll
                    for (int nt = 0; nt < NCI; ++nt) {
                        const float s = icoef[8 * nt + g], sh = icoef[CI + 8 * nt + g];
                        const pf2_t s2 = pf2_pack(s, s), t2 = pf2_pack(sh, sh);
                        bx[nt][0] = pf2_relu_bf16x2(pf2_fma(pf2_from_bf16x2(bx[nt][0]), s2, t2));
                        bx[nt][1] = pf2_relu_bf16x2(pf2_fma(pf2_from_bf16x2(bx[nt][1]), s2, t2));
                    }
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    uint32_t a0, a1, a2, a3;
                    ldsm4t(grow + a_wg + mt * 32, a0, a1, a2, a3);
#pragma unroll
                    for (int nt = 0; nt < NCI; ++nt) mma16816(wacc[mt][nt], a0, a1, a2, a3, bx[nt][0], bx[nt][1]);
                }
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) tma_store4(&tm_dx, ci0, (int)row0, 0, 0, OUT);
        b ^= 1;
    }
    if (tid == 0) tma_store_wait_read();
    // ---- flush: weight gradient and the producer's reductions ----
    if (do_wgrad) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NCI; ++nt) {
                const int co = 16 * mt + g, ci = 8 * nt + 2 * t;
                atomicAdd(&dwacc[co * CI + ci], wacc[mt][nt][0]);
                atomicAdd(&dwacc[co * CI + ci + 1], wacc[mt][nt][1]);
                atomicAdd(&dwacc[(co + 8) * CI + ci], wacc[mt][nt][2]);
                atomicAdd(&dwacc[(co + 8) * CI + ci + 1], wacc[mt][nt][3]);
            }
    }
    if (do_red) {
#pragma unroll
        for (int nt = 0; nt < NCI; ++nt) {
            float v[4];
            pf2_unpack(rs[nt][0], v[0], v[1]);
            pf2_unpack(rs[nt][1], v[2], v[3]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[q] += __shfl_xor_sync(0xffffffffu, v[q], 4);
                v[q] += __shfl_xor_sync(0xffffffffu, v[q], 8);
                v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);
            }
            if (g == 0) {
                const int c = 8 * nt + 2 * t;
                atomicAdd(&red[c], v[0]);
                atomicAdd(&red[c + 1], v[1]);
                atomicAdd(&red[CI + c], v[2]);
                atomicAdd(&red[CI + c + 1], v[3]);
            }
        }
    }
    __syncthreads();
    if (do_wgrad)
        for (int i = tid; i < CO * CI; i += THREADS) atomicAdd(&p.dw[(i / CI) * CIT + ci0 + i % CI], dwacc[i]);
    if (do_red)
        for (int i = tid; i < CI; i += THREADS) {
            atomicAdd(&p.nsums[ci0 + i], (double)red[i]);
            atomicAdd(&p.nsums[CIT + ci0 + i], (double)red[CI + i]);
        }
}

template <int CO, int CI>
static int launch_pwb(const PwbP& p0, const void* G, const void* Z, const void* X, const void* add, void* dX, int Cin,
                      cudaStream_t st) {
    using Cfg = PwbCfg<CO, CI>;
    PwbP p = p0;
    p.cin_total = Cin;
    const int slices = Cin / CI;
    const char* name = "pw_bwd_fused";
    if (p.M >= (1ll << 31) - Cfg::R) { set_error("%s: too many rows", name); return MNB_ERR_UNSUPPORTED; }
    const int M = (int)p.M;
    CUtensorMap tm_g, tm_z, tm_x, tm_add, tm_dx;
    if (int e = dwm_tensor_map(&tm_g, G, 1, 1, M, CO, CO, Cfg::R, 1)) return e;
    if (int e = dwm_tensor_map(&tm_z, Z, 1, 1, M, CO, CO, Cfg::R, 1)) return e;
    if (int e = dwm_tensor_map(&tm_x, X, 1, 1, M, Cin, CI, Cfg::R, 1)) return e;
    if (int e = dwm_tensor_map(&tm_add, add ? add : X, 1, 1, M, Cin, CI, Cfg::R, 1)) return e;
    if (int e = dwm_tensor_map(&tm_dx, dX, 1, 1, M, Cin, CI, Cfg::R, 1)) return e;
    static int occ = -1;
    if (occ < 0) {
        cudaError_t e = cudaFuncSetAttribute(pw_bwd_fused_k<CO, CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return (int)e; }
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, pw_bwd_fused_k<CO, CI>, Cfg::THREADS, Cfg::SMEM);
        if (o < 1) { set_error("%s: kernel does not fit on an SM", name); return MNB_ERR_UNSUPPORTED; }
        occ = o;
    }
    const long long ntiles = (p.M + Cfg::R - 1) / Cfg::R;
    long long grid = std::max(1LL, (long long)num_sms() * occ / slices);      // resident CTAs only (no second wave)
    if (grid > ntiles) grid = ntiles;
    pw_bwd_fused_k<CO, CI><<<dim3((unsigned)grid, (unsigned)slices), Cfg::THREADS, Cfg::SMEM, st>>>(tm_g, tm_z, tm_x, tm_add,
                                                                                                  tm_dx, p);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

int pw_bwd_fused(const void* G, const void* Z, const float* scale, const float* shift, const double* sums, const float* mean,
                 const float* invstd, double m, float* dgamma, float* dbeta, float* dbias, const void* X,
                 const float* in_scale, const float* in_shift, const float* w, const void* add, void* dX, float* dw,
                 double* nsums, long long M, int Cin, int Cout, cudaStream_t st) {
    PwbP p = {};
    p.scale = scale; p.shift = shift; p.sums = sums; p.mean = mean; p.invstd = invstd; p.m = m;
    p.dgamma = dgamma; p.dbeta = dbeta; p.dbias = dbias; p.in_scale = in_scale; p.in_shift = in_shift; p.w = w; p.dw = dw;
    p.nsums = nsums; p.has_add = add != nullptr; p.M = M;
#define MNB_PWB(CO_, CI_) if (Cout == CO_ && Cin == CI_) return launch_pwb<CO_, CI_>(p, G, Z, X, add, dX, Cin, st)
    MNB_PWB(48, 16);
    MNB_PWB(16, 48);
    MNB_PWB(16, 32);
    MNB_PWB(72, 24);
    MNB_PWB(24, 72);
#undef MNB_PWB
    // wide inputs: CTA columns (blockIdx.y) own slices of CIS input channels; every slice repeats the dZ transform of the
    // narrow G / Z tile and the 210-255 registers per thread leave one CTA per SM: 174 / 169 us (slices of 48 / 80) against
    // 180 us for the unfused chain at 28x28 240 -> 40 (scripts/exp_pw_bwd.py slice) -- not a win, so it stays behind the
    // "pwb_slice" option (default 0).  Wide layers need the Cout / Cin split over warps.
    if (Cout == 40 && Cin % 48 == 0 && option_get(OPT_PWB_SLICE) == 48) return launch_pwb<40, 48>(p, G, Z, X, add, dX, Cin, st);
    if (Cout == 40 && Cin % 80 == 0 && option_get(OPT_PWB_SLICE) == 80) return launch_pwb<40, 80>(p, G, Z, X, add, dX, Cin, st);
    set_error("pw_bwd_fused: shape %d -> %d not instantiated", Cin, Cout);
    return MNB_ERR_UNSUPPORTED;
}

}  // namespace mnb
