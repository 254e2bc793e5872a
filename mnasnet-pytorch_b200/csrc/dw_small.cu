// Depthwise k x k (3 / 5) stride-1 convolution in WHOLE TILES (written for the small maps, H <= 28), bf16 NHWC, on the tensor pipe: forward (+ the
// producing ConvBlock's BN-apply+ReLU on load, + BN batch statistics on store), backward-data and backward-weight.
// Same math and entry points as dw_mma.cu (nn.Conv2d(groups=C) at src/models/mnasnet.py:76-81,120-125); same MMA
// formulation (m16n8k16 with a diagonal B per 8-channel chunk, A straight from the NHWC tile through ldmatrix).
//
// Why a second kernel family: dw_mma.cu walks DOWN the image in blocks of 10 / 12 rows with a k-1 row prologue per
// column segment.  On a 14 x 14 map that is three pipeline steps (prologue, a full block, a 4-row remainder on the
// range-checked path) for 14 rows of work, on 28 x 28 the prologues and the remainder cost 1.5x the MMAs
// (profiles/r2_exp_dw_mma.json: 0.19-0.26 of HBM on these maps against 0.45-0.57 on 112 x 112).  Here a work item is a
// WHOLE tile of TH = 14 output rows: one TMA box brings the TH + k - 1 input rows (halo rows / columns outside the image
// are zero-filled by the hardware), the row loop is fully unrolled with the tile-edge taps dropped at compile time (no
// prologue, no range checks, nothing recomputed but the in-place transform of the halo rows), one TMA store writes the
// TH x TW output box.  14 x 14 maps are one item per (image, channel group), 28 x 28 two.  Measured afterwards: the same
// kernels also beat the row-streaming ones on the 56 x 56 maps, and the fused backward does on 112 x 112 (14.85 -> 14.25 ms
// per step), so the default policy (dw_small_covers / dw_small_covers_bwd) takes those too.
#include "dw_mma_dev.cuh"

namespace mnb {

__host__ __device__ constexpr int dws_al128(int b) { return (b + 127) / 128 * 128; }

template <int K, int CG, int TWS, int TH>
struct DwsCfg {
    using F = DwmCfg<K, CG, TWS>;
    static constexpr int P = K / 2, TR = TH + K - 1, NCH = F::NCH, TW = F::TW, HC = F::HC, PITCH = F::PITCH;
    static constexpr int ROWB = F::ROWB, OROWB = F::OROWB, THREADS = F::THREADS;
    static constexpr int XB_BYTES = dws_al128(TR * ROWB);          // input tile: TR rows x HC columns
    static constexpr int OUT_BYTES = dws_al128(TH * OROWB);        // output tile: TH rows x TW columns
    static constexpr int SMEM = 2 * XB_BYTES + OUT_BYTES + 4 * CG * 4 + 16;
    // backward-weight: dZ tile TH x HC (+ halo columns), X tile TR x TW
    static constexpr int GZ_BYTES = dws_al128(TH * ROWB), XW_BYTES = dws_al128(TR * OROWB);
    static constexpr int SMEM_WG = 2 * (GZ_BYTES + XW_BYTES) + CG * K * K * 4 + 16;
    // resident CTAs per SM the compiler must allow: by registers (104 / 80 per thread), capped by what shared memory admits
    static constexpr int BY_REGS = 65536 / (THREADS * (K == 5 ? 104 : 80));
    static constexpr int BY_SMEM = 232448 / (SMEM + 1024), BY_SMEM_WG = 232448 / (SMEM_WG + 1024);
    static constexpr int MINB = BY_REGS < 1 ? 1 : (BY_REGS < BY_SMEM ? BY_REGS : BY_SMEM);
    static constexpr int MINB_WG = BY_REGS < 1 ? 1 : (BY_REGS < BY_SMEM_WG ? BY_REGS : BY_SMEM_WG);
};

struct DwsP {
    const float* in_scale;
    const float* in_shift;
    const float* w;             // [C][K][K] fp32
    double* stats;              // fwd: [2][C]
    float* dw;                  // wgrad
    int N, H, W, C;
    int tiles_w, tiles_h, cblocks, items;       // items = N * tiles_w * tiles_h
};

// Output row O of the tile is complete: stage it for the TMA store; STATS: sum / sum of squares of the stored (bf16)
// values, masked by k0 / k1 (column validity of the lane's two pixels x row validity).
template <int K, int CG, int TWS, bool STATS>
struct DwsEmit {
    uint32_t out_lane;
    f2_t mk0, mk1;
    f2_t st[2];
    template <int O>
    __device__ __forceinline__ void emit(float (&a)[4], bool row_ok) {
        using F = DwmCfg<K, CG, TWS>;
        const uint32_t u0 = pack_bf16x2(a[0], a[1]), u1 = pack_bf16x2(a[2], a[3]);
        sts32(out_lane + O * F::OROWB, u0);
        sts32(out_lane + O * F::OROWB + 8 * F::PITCH, u1);
        if constexpr (STATS) {
            const f2_t one = f2_pack(1.f, 1.f), zero = f2_pack(0.f, 0.f);
            const f2_t q0 = f2_fma(f2_from_bf16x2(u0), row_ok ? mk0 : zero, zero);
            const f2_t q1 = f2_fma(f2_from_bf16x2(u1), row_ok ? mk1 : zero, zero);
            st[0] = f2_fma(q0, one, st[0]);
            st[0] = f2_fma(q1, one, st[0]);
            st[1] = f2_fma(q0, q0, st[1]);
            st[1] = f2_fma(q1, q1, st[1]);
        }
    }
};

// Tile row I feeds the outputs I - kh that lie inside the tile (taps outside are dropped at compile time); output
// I - (K-1) is complete after it.  Straight-line code: rows above / below the image are zero-filled boxes rows and simply
// add zeros.  The last tap column is paired ACROSS rows as in dw_mma.cu -- tap (kh, K-1) of row I-1 and tap (kh+1, K-1) of
// row I feed the same output -- 13 instead of 15 MMAs per row for 5x5, 5 instead of 6 for 3x3.
template <int K, int CG, int TWS, int TH, int I, class EM>
struct DwsRows {
    static __device__ __forceinline__ void run(float (&acc)[K][4], const uint32_t (&bd)[K][K], uint32_t x4, uint32_t x2, int o0,
                                               int H, EM& em, uint32_t p0, uint32_t p1) {
        using F = DwmCfg<K, CG, TWS>;
        constexpr int TR = TH + K - 1;
        uint32_t m[K][2];
        dwm_load_row<K, F::PITCH>(m, x4 + I * F::ROWB, x2 + I * F::ROWB);
#pragma unroll
        for (int kw = 0; kw + 1 < K; kw += 2)
#pragma unroll
            for (int kh = 0; kh < K; ++kh)
                if (I - kh >= 0 && I - kh < TH)
                    mma16816(acc[(I - kh + K) % K], m[kw][0], m[kw][1], m[kw + 1][0], m[kw + 1][1], bd[kh][kw], bd[kh][kw + 1]);
#pragma unroll
        for (int kh = 0; kh + 1 < K; kh += 2)       // (row I-1, kh) + (row I, kh+1) -> output I-1-kh
            if (I - 1 - kh >= 0 && I - 1 - kh < TH)
                mma16816(acc[(I - 1 - kh + K) % K], p0, p1, m[K - 1][0], m[K - 1][1], bd[kh][K - 1], bd[kh + 1][K - 1]);
        if constexpr (I >= K - 1 && I - (K - 1) < TH) {
            constexpr int O = I - (K - 1);
            float (&a)[4] = acc[O % K];
            mma1688(a, m[K - 1][0], m[K - 1][1], bd[K - 1][K - 1]);
            em.template emit<O>(a, o0 + O < H);
            a[0] = a[1] = a[2] = a[3] = 0.f;
        }
        if constexpr (I + 1 < TR) DwsRows<K, CG, TWS, TH, I + 1, EM>::run(acc, bd, x4, x2, o0, H, em, m[K - 1][0], m[K - 1][1]);
    }
};

// FLIP: correlate with the 180-degree rotated kernel (backward-data of a stride-1 'same' depthwise conv)
template <int K, int CG, int TWS, int TH, bool FLIP>
__global__ void __launch_bounds__(DwsCfg<K, CG, TWS, TH>::THREADS, DwsCfg<K, CG, TWS, TH>::MINB)
    dws_fwd_k(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_z, const DwsP p) {
    using Cfg = DwsCfg<K, CG, TWS, TH>;
    constexpr int P = Cfg::P, NCH = Cfg::NCH, TW = Cfg::TW, HC = Cfg::HC, PITCH = Cfg::PITCH, THREADS = Cfg::THREADS, TR = Cfg::TR;
    extern __shared__ __align__(128) unsigned char dsm[];
    const uint32_t XB0 = smem_u32(dsm);
    const uint32_t OUT = XB0 + 2 * Cfg::XB_BYTES;
    float* red = reinterpret_cast<float*>(dsm + 2 * Cfg::XB_BYTES + Cfg::OUT_BYTES);      // [4][CG]
    const uint32_t bar0 = OUT + Cfg::OUT_BYTES + 4 * CG * 4;                             // two mbarriers
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int chunk = warp % NCH, strip = warp / NCH;
    const int cb = blockIdx.x % p.cblocks, slot = blockIdx.x / p.cblocks, nslots = gridDim.x / p.cblocks;
    const int cbase = cb * CG;
    const bool chunk_live = cbase + chunk * 8 < p.C;

    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    for (int i = tid; i < 4 * CG; i += THREADS) red[i] = 0.f;
    uint32_t bd[K][K];
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
            const int ih = FLIP ? K - 1 - kh : kh, iw = FLIP ? K - 1 - kw : kw;
            const float wv = chunk_live ? p.w[(size_t)(cbase + chunk * 8 + g) * K * K + ih * K + iw] : 0.f;
            bd[kh][kw] = dwm_diag(wv, g, t);
        }
    const int tchunk = tid % NCH, pix0 = tid / NCH;
    constexpr int PSTEP = THREADS / NCH;
    const bool xf = p.in_scale != nullptr && (cbase + tchunk * 8 < p.C);
    f2_t xs[4], xt[4];
    if (xf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 a = *reinterpret_cast<const float2*>(p.in_scale + cbase + tchunk * 8 + 2 * i);
            const float2 b = *reinterpret_cast<const float2*>(p.in_shift + cbase + tchunk * 8 + 2 * i);
            xs[i] = f2_pack(a.x, a.y);
            xt[i] = f2_pack(b.x, b.y);
        }
    }
    const int mi = lane >> 3, r8 = lane & 7;
    const uint32_t off4 = (uint32_t)(((mi >> 1) + (mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    const uint32_t off2 = (uint32_t)(((K - 1) + (mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    typedef DwsEmit<K, CG, TWS, !FLIP> EM;      // statistics belong to the forward; backward-data stores only
    EM em;
    em.out_lane = OUT + (uint32_t)((strip * 16 + g) * PITCH + chunk * 16 + t * 4);
    em.st[0] = em.st[1] = f2_pack(0.f, 0.f);
    const int per_img = p.tiles_w * p.tiles_h;
    __syncthreads();

    auto issue = [&](int item, int b) {
        const int n = item / per_img, rem = item - n * per_img, wt = rem / p.tiles_h, rt = rem - wt * p.tiles_h;
        mbar_expect_tx(bar0 + 8 * b, (uint32_t)(TR * Cfg::ROWB));
        tma_load4(XB0 + b * Cfg::XB_BYTES, &tm_x, cbase, wt * TW - P, rt * TH - P, n, bar0 + 8 * b);
    };
    int item = slot, b = 0;
    uint32_t ph = 0;
    if (tid == 0 && item < p.items) issue(item, 0);
    for (; item < p.items; item += nslots, b ^= 1) {
        const int n = item / per_img, rem = item - n * per_img, wt = rem / p.tiles_h, rt = rem - wt * p.tiles_h;
        const int w0 = wt * TW, o0 = rt * TH, r0 = o0 - P;
        if (tid == 0) {
            tma_store_wait_read();                      // the previous tile's store has finished reading OUT
            if (item + nslots < p.items) issue(item + nslots, b ^ 1);
        }
        const uint32_t XB = XB0 + b * Cfg::XB_BYTES;
        mbar_wait(bar0 + 8 * b, (ph >> b) & 1);
        ph ^= 1u << b;
        if (xf) {       // relu(scale * x + shift) in place; zero-filled padding stays zero
            for (int pix = pix0; pix < TR * HC; pix += PSTEP) {
                const int rr = pix / HC, cc = pix - rr * HC;
                if ((unsigned)(r0 + rr) < (unsigned)p.H && (unsigned)(w0 - P + cc) < (unsigned)p.W) {
                    const uint32_t a = XB + (uint32_t)(pix * PITCH + tchunk * 16);
                    uint4 u = lds128(a);
                    u.x = f2_relu_bf16x2(f2_fma(f2_from_bf16x2(u.x), xs[0], xt[0]));
                    u.y = f2_relu_bf16x2(f2_fma(f2_from_bf16x2(u.y), xs[1], xt[1]));
                    u.z = f2_relu_bf16x2(f2_fma(f2_from_bf16x2(u.z), xs[2], xt[2]));
                    u.w = f2_relu_bf16x2(f2_fma(f2_from_bf16x2(u.w), xs[3], xt[3]));
                    sts128(a, u);
                }
            }
        }
        __syncthreads();
        const int sw = w0 + strip * 16;
        if (chunk_live && sw < p.W) {
            const float m0 = sw + g < p.W ? 1.f : 0.f, m1 = sw + g + 8 < p.W ? 1.f : 0.f;
            em.mk0 = f2_pack(m0, m0);
            em.mk1 = f2_pack(m1, m1);
            float acc[K][4];
#pragma unroll
            for (int i = 0; i < K; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            DwsRows<K, CG, TWS, TH, 0, EM>::run(acc, bd, XB + off4, XB + off2, o0, p.H, em, 0u, 0u);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) tma_store4(&tm_z, cbase, w0, o0, n, OUT);
    }
    if (tid == 0) tma_store_wait_read();
    if (!FLIP && p.stats != nullptr) {
        float v[4];
        f2_unpack(em.st[0], v[0], v[1]);
        f2_unpack(em.st[1], v[2], v[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[q] += __shfl_xor_sync(0xffffffffu, v[q], 4);
            v[q] += __shfl_xor_sync(0xffffffffu, v[q], 8);
            v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);
        }
        if (g == 0 && chunk_live) {
            const int c = chunk * 8 + 2 * t;
            atomicAdd(&red[c], v[0]);
            atomicAdd(&red[c + 1], v[1]);
            atomicAdd(&red[CG + c], v[2]);
            atomicAdd(&red[CG + c + 1], v[3]);
        }
        __syncthreads();
        for (int i = tid; i < CG; i += THREADS) {
            if (cbase + i < p.C) {
                atomicAdd(&p.stats[cbase + i], (double)red[i]);
                atomicAdd(&p.stats[p.C + cbase + i], (double)red[CG + i]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward-weight: D[(tap column, c')][c] += sum over 16 pixels of dZ^T (ldmatrix.trans of the dZ row shifted by the tap
// column) x A (ldmatrix.trans of the X row of tap row kh, activated in registers); the diagonal c' == c of the
// accumulators is the weight gradient.  dZ tile row o pairs with X tile rows o .. o+K-1.
// ---------------------------------------------------------------------------------------------------------------
template <int K, int CG, int TWS, int TH>
__global__ void __launch_bounds__(DwsCfg<K, CG, TWS, TH>::THREADS, DwsCfg<K, CG, TWS, TH>::MINB_WG)
    dws_wgrad_k(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_x, const DwsP p) {
    using Cfg = DwsCfg<K, CG, TWS, TH>;
    constexpr int P = Cfg::P, NCH = Cfg::NCH, TW = Cfg::TW, PITCH = Cfg::PITCH, THREADS = Cfg::THREADS, TR = Cfg::TR;
    constexpr int KK = K * K, NPR = (K + 1) / 2;
    extern __shared__ __align__(128) unsigned char dsm[];
    const uint32_t GB0 = smem_u32(dsm);
    const uint32_t XW0 = GB0 + 2 * Cfg::GZ_BYTES;
    float* dwacc = reinterpret_cast<float*>(dsm + 2 * (Cfg::GZ_BYTES + Cfg::XW_BYTES));   // [CG][K][K]
    const uint32_t bar0 = XW0 + 2 * Cfg::XW_BYTES + CG * KK * 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int chunk = warp % NCH, strip = warp / NCH;
    const int cb = blockIdx.x % p.cblocks, slot = blockIdx.x / p.cblocks, nslots = gridDim.x / p.cblocks;
    const int cbase = cb * CG;
    const bool chunk_live = cbase + chunk * 8 < p.C;

    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    for (int i = tid; i < CG * KK; i += THREADS) dwacc[i] = 0.f;
    const int mi = lane >> 3, r8 = lane & 7;
    const uint32_t off4 = (uint32_t)(((mi >> 1) + (mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    const uint32_t off2 = (uint32_t)(((K - 1) + (mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    const uint32_t offx = (uint32_t)(((mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    const bool act = p.in_scale != nullptr;
    f2_t ag = f2_pack(1.f, 1.f), at = f2_pack(0.f, 0.f);
    if (act && chunk_live) {
        const float s = p.in_scale[cbase + chunk * 8 + g], sh = p.in_shift[cbase + chunk * 8 + g];
        ag = f2_pack(s, s);
        at = f2_pack(sh, sh);
    }
    float wacc[K][NPR][4];
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
        for (int pr = 0; pr < NPR; ++pr) wacc[kh][pr][0] = wacc[kh][pr][1] = wacc[kh][pr][2] = wacc[kh][pr][3] = 0.f;
    const int per_img = p.tiles_w * p.tiles_h;
    __syncthreads();

    auto issue = [&](int item, int b) {
        const int n = item / per_img, rem = item - n * per_img, wt = rem / p.tiles_h, rt = rem - wt * p.tiles_h;
        mbar_expect_tx(bar0 + 8 * b, (uint32_t)(TH * Cfg::ROWB + TR * Cfg::OROWB));
        tma_load4(GB0 + b * Cfg::GZ_BYTES, &tm_g, cbase, wt * TW - P, rt * TH, n, bar0 + 8 * b);
        tma_load4(XW0 + b * Cfg::XW_BYTES, &tm_x, cbase, wt * TW, rt * TH - P, n, bar0 + 8 * b);
    };
    int item = slot, b = 0;
    uint32_t ph = 0;
    if (tid == 0 && item < p.items) issue(item, 0);
    for (; item < p.items; item += nslots, b ^= 1) {
        const int n = item / per_img, rem = item - n * per_img, wt = rem / p.tiles_h, rt = rem - wt * p.tiles_h;
        const int w0 = wt * TW, o0 = rt * TH, r0 = o0 - P;
        if (tid == 0 && item + nslots < p.items) issue(item + nslots, b ^ 1);
        mbar_wait(bar0 + 8 * b, (ph >> b) & 1);
        ph ^= 1u << b;
        const int sw = w0 + strip * 16;
        if (chunk_live && sw < p.W) {
            const uint32_t gz4 = GB0 + b * Cfg::GZ_BYTES + off4, gz2 = GB0 + b * Cfg::GZ_BYTES + off2;
            const uint32_t xrow = XW0 + b * Cfg::XW_BYTES + offx;
            // X is zero-filled outside the image but relu(shift) is not zero: mask the lane's pixels (dZ is exactly zero there)
            const int c0 = sw + 2 * t;
            const uint32_t cm0 = (c0 < p.W ? 0x0000ffffu : 0u) | (c0 + 1 < p.W ? 0xffff0000u : 0u);
            const uint32_t cm1 = (c0 + 8 < p.W ? 0x0000ffffu : 0u) | (c0 + 9 < p.W ? 0xffff0000u : 0u);
            auto load_x = [&](int xr, uint32_t (&dst)[2]) {       // straight-line: rows outside the image are masked, not skipped
                const bool rv = (unsigned)(r0 + xr) < (unsigned)p.H;
                uint32_t u0, u1;
                ldsm2t(xrow + (uint32_t)(xr * Cfg::OROWB), u0, u1);
                if (act) {
                    u0 = f2_relu_bf16x2(f2_fma(f2_from_bf16x2(u0), ag, at));
                    u1 = f2_relu_bf16x2(f2_fma(f2_from_bf16x2(u1), ag, at));
                }
                dst[0] = rv ? u0 & cm0 : 0u;
                dst[1] = rv ? u1 & cm1 : 0u;
            };
            uint32_t bf[K][2];
#pragma unroll
            for (int kh = 0; kh + 1 < K; ++kh) load_x(kh, bf[kh]);
#pragma unroll
            for (int o = 0; o < TH; ++o) {
                load_x(o + K - 1, bf[(o + K - 1) % K]);
                {       // dZ rows below the image are zero-filled by the TMA load: they add nothing
                    uint32_t tt[K][2];
                    const uint32_t a4 = gz4 + (uint32_t)(o * Cfg::ROWB), a2 = gz2 + (uint32_t)(o * Cfg::ROWB);
                    ldsm4t(a4, tt[0][0], tt[0][1], tt[1][0], tt[1][1]);
                    if constexpr (K == 5) ldsm4t(a4 + 2 * PITCH, tt[2][0], tt[2][1], tt[3][0], tt[3][1]);
                    ldsm2t(a2, tt[K - 1][0], tt[K - 1][1]);
#pragma unroll
                    for (int kh = 0; kh < K; ++kh)
#pragma unroll
                        for (int pr = 0; pr < NPR; ++pr) {
                            const int s0 = 2 * pr, s1 = (2 * pr + 1 < K) ? 2 * pr + 1 : 2 * pr;   // odd K: last pair duplicates
                            mma16816(wacc[kh][pr], tt[s0][0], tt[s1][0], tt[s0][1], tt[s1][1], bf[(o + kh) % K][0], bf[(o + kh) % K][1]);
                        }
                }
            }
        }
        __syncthreads();            // buffer b is free for the load issued at the top of the next iteration
    }
    // the diagonal c' == c lives on the lanes with t == g >> 1: slot A in d[g & 1], slot B in d[2 + (g & 1)];
    // shift s of the dZ operand is tap column K-1-s, kh is the tap row
    if (chunk_live && (g >> 1) == t) {
        float* dst = dwacc + (chunk * 8 + g) * KK;
#pragma unroll
        for (int kh = 0; kh < K; ++kh)
#pragma unroll
            for (int pr = 0; pr < NPR; ++pr) {
                atomicAdd(dst + kh * K + (K - 1 - 2 * pr), (g & 1) ? wacc[kh][pr][1] : wacc[kh][pr][0]);
                if (2 * pr + 1 < K) atomicAdd(dst + kh * K + (K - 2 - 2 * pr), (g & 1) ? wacc[kh][pr][3] : wacc[kh][pr][2]);
            }
    }
    __syncthreads();
    for (int i = tid; i < CG * KK; i += THREADS)
        if (cbase + i / KK < p.C) atomicAdd(&p.dw[(size_t)cbase * KK + i], dwacc[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// Fused depthwise ConvBlock backward on a whole tile (same contract as dw_mma.cu's dw_mma_bwd_k / mnb_dw_bwd_fused):
//     dZ = a * G * [scale*Z + shift > 0] + b * Z + c  (in place over the G tile, halo included),
//     dX = dZ (*) rot180(w)  -> staged over the Z tile, TMA store;  sum(dX'), sum(dX' x) of the producing block;
//     dW += dZ^T A(X)  on the tile's own rows.
// G, Z: TR x HC boxes (halo rows / columns: zero outside the image), X: TR x TW box.  Single-buffered: 2-4 CTAs share an SM.
// ---------------------------------------------------------------------------------------------------------------
struct DwsBP {
    const float* scale;         // this block's BN scale / shift (ReLU mask)
    const float* shift;
    const double* sums;         // [2][C] sum(G'), sum(G' z) of this block (already reduced)
    const float* mean;
    const float* invstd;
    double m;
    float* dgamma;              // += (NULL = frozen)
    float* dbeta;
    float* dbias;
    const float* in_scale;      // producing block's BN scale / shift
    const float* in_shift;
    const float* w;
    float* dw;                  // += (NULL = frozen)
    double* nsums;              // [2][C] sum(dX'), sum(dX' x) for the producing block (NULL = not wanted)
    int N, H, W, C;
    int tiles_w, tiles_h, cblocks, items;
};

template <int K, int CG, int TWS, int TH>
struct DwsBCfg {
    using F = DwsCfg<K, CG, TWS, TH>;
    static constexpr int KK = K * K;
    static constexpr int GB_BYTES = F::XB_BYTES, XW_BYTES = F::XW_BYTES;
    static constexpr int WS_BYTES = (CG * KK * 2 + 15) / 16 * 16, DW_BYTES = CG * KK * 4, CO_BYTES = 7 * CG * 4;
    static constexpr int SMEM = 2 * GB_BYTES + XW_BYTES + WS_BYTES + DW_BYTES + CO_BYTES + 2 * CG * 4 + 16;
    static constexpr int BY_REGS = 65536 / (F::THREADS * 128) < 1 ? 1 : 65536 / (F::THREADS * 128);
    static constexpr int BY_SMEM = 232448 / (SMEM + 1024);
    static constexpr int MINB = BY_REGS < BY_SMEM ? BY_REGS : (BY_SMEM < 1 ? 1 : BY_SMEM);
};

// dX row O of the tile: stage for the TMA store; reduce dX' = dX * [s_in x + t_in > 0] and dX' * x against the raw X
template <int K, int CG, int TWS>
struct DwsEmitReduce {
    uint32_t out_lane, x_lane;
    bool do_red;
    f2_t sp, tp, mk0, mk1;
    f2_t rs[2];
    template <int O>
    __device__ __forceinline__ void emit(float (&a)[4], bool row_ok) {
        using F = DwmCfg<K, CG, TWS>;
        const uint32_t u0 = pack_bf16x2(a[0], a[1]), u1 = pack_bf16x2(a[2], a[3]);
        sts32(out_lane + O * F::OROWB, u0);
        sts32(out_lane + O * F::OROWB + 8 * F::PITCH, u1);
        if (do_red) {
            const f2_t zero = f2_pack(0.f, 0.f), one = f2_pack(1.f, 1.f);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const f2_t x = f2_from_bf16x2(lds32(x_lane + O * F::OROWB + h * 8 * F::PITCH));
                float y0, y1, q0, q1;
                f2_unpack(f2_fma(sp, x, tp), y0, y1);
                f2_unpack(f2_from_bf16x2(h ? u1 : u0), q0, q1);
                f2_t q = f2_pack(y0 > 0.f ? q0 : 0.f, y1 > 0.f ? q1 : 0.f);
                q = f2_fma(q, row_ok ? (h ? mk1 : mk0) : zero, zero);
                rs[0] = f2_fma(q, one, rs[0]);
                rs[1] = f2_fma(q, x, rs[1]);
            }
        }
    }
};

template <int K, int CG, int TWS, int TH>
__global__ void __launch_bounds__(DwsCfg<K, CG, TWS, TH>::THREADS, DwsBCfg<K, CG, TWS, TH>::MINB)
    dws_bwd_k(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_x,
              const __grid_constant__ CUtensorMap tm_dx, const DwsBP p) {
    using Cfg = DwsCfg<K, CG, TWS, TH>;
    using B = DwsBCfg<K, CG, TWS, TH>;
    constexpr int P = Cfg::P, NCH = Cfg::NCH, TW = Cfg::TW, HC = Cfg::HC, PITCH = Cfg::PITCH, THREADS = Cfg::THREADS, TR = Cfg::TR;
    constexpr int KK = K * K, NPR = (K + 1) / 2;
    extern __shared__ __align__(128) unsigned char dsm[];
    const uint32_t GB = smem_u32(dsm);                      // G tile -> dZ (in place)
    const uint32_t ZB = GB + B::GB_BYTES;                   // Z tile; after the transform: dX staging (TH x TW)
    const uint32_t XW = ZB + B::GB_BYTES;                   // raw X tile
    unsigned char* after = dsm + 2 * B::GB_BYTES + B::XW_BYTES;
    __nv_bfloat16* wsm = reinterpret_cast<__nv_bfloat16*>(after);                           // [CG][KK] bf16
    float* dwacc = reinterpret_cast<float*>(after + B::WS_BYTES);                           // [CG][KK]
    float* coef = reinterpret_cast<float*>(after + B::WS_BYTES + B::DW_BYTES);              // [7][CG]
    float* red = coef + 7 * CG;                                                             // [2][CG]
    const uint32_t bar = smem_u32(red + 2 * CG);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int chunk = warp % NCH, strip = warp / NCH;
    const int cb = blockIdx.x % p.cblocks, slot = blockIdx.x / p.cblocks, nslots = gridDim.x / p.cblocks;
    const int cbase = cb * CG;
    const bool chunk_live = cbase + chunk * 8 < p.C;
    const bool do_wgrad = p.dw != nullptr, do_red = p.nsums != nullptr && p.in_scale != nullptr;

    if (tid == 0) mbar_init(bar, 1);
    for (int i = tid; i < CG * KK; i += THREADS) {
        const int c = cbase + i / KK;
        wsm[i] = __float2bfloat16_rn(c < p.C ? p.w[(size_t)c * KK + i % KK] : 0.f);
        dwacc[i] = 0.f;
    }
    for (int i = tid; i < CG; i += THREADS) {
        // the per-channel work of BatchNorm backward (what bn_bwd_finalize does), redundantly per CTA
        const int c = cbase + i;
        float a = 0.f, b = 0.f, c3 = 0.f, sc = 0.f, sh = 0.f, isc = 1.f, ish = 0.f;
        if (c < p.C) {
            const double sg = p.sums[c], sgz = p.sums[p.C + c];
            const double mean = p.mean[c], invstd = p.invstd[c], s = p.scale[c];
            const double dga = invstd * (sgz - mean * sg);
            const double bb = -s * invstd * dga / p.m;
            const double cc = -s * sg / p.m - bb * mean;
            a = (float)s; b = (float)bb; c3 = (float)cc; sc = p.scale[c]; sh = p.shift[c];
            if (p.in_scale) { isc = p.in_scale[c]; ish = p.in_shift[c]; }
            if (slot == 0) {
                if (p.dgamma) p.dgamma[c] += (float)dga;
                if (p.dbeta) p.dbeta[c] += (float)sg;
                if (p.dbias) p.dbias[c] += (float)(s * sg + bb * mean * p.m + cc * p.m);   // analytically 0
            }
        }
        coef[0 * CG + i] = a; coef[1 * CG + i] = b; coef[2 * CG + i] = c3; coef[3 * CG + i] = sc; coef[4 * CG + i] = sh;
        coef[5 * CG + i] = isc; coef[6 * CG + i] = ish;
        red[i] = 0.f; red[CG + i] = 0.f;
    }
    const int tchunk = tid % NCH, pix0 = tid / NCH;
    constexpr int PSTEP = THREADS / NCH;
    const bool tlive = cbase + tchunk * 8 < p.C;
    const int mi = lane >> 3, r8 = lane & 7;
    const uint32_t off4 = (uint32_t)(((mi >> 1) + (mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    const uint32_t off2 = (uint32_t)(((K - 1) + (mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    const uint32_t offx = (uint32_t)(((mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    typedef DwsEmitReduce<K, CG, TWS> EM;
    EM em;
    em.out_lane = ZB + (uint32_t)((strip * 16 + g) * PITCH + chunk * 16 + t * 4);
    em.x_lane = XW + (uint32_t)(P * Cfg::OROWB + (strip * 16 + g) * PITCH + chunk * 16 + t * 4);
    em.do_red = do_red;
    em.rs[0] = em.rs[1] = f2_pack(0.f, 0.f);
    uint32_t ph = 0;
    __syncthreads();
    {
        const int c = chunk * 8 + 2 * t;
        em.sp = f2_pack(coef[5 * CG + c], coef[5 * CG + c + 1]);
        em.tp = f2_pack(coef[6 * CG + c], coef[6 * CG + c + 1]);
    }
    const f2_t ag = f2_pack(coef[5 * CG + chunk * 8 + g], coef[5 * CG + chunk * 8 + g]);
    const f2_t at = f2_pack(coef[6 * CG + chunk * 8 + g], coef[6 * CG + chunk * 8 + g]);
    const bool act = p.in_scale != nullptr;
    const int per_img = p.tiles_w * p.tiles_h;

    for (int item = slot; item < p.items; item += nslots) {
        const int n = item / per_img, rem = item - n * per_img, wt = rem / p.tiles_h, rt = rem - wt * p.tiles_h;
        const int w0 = wt * TW, o0 = rt * TH, r0 = o0 - P;
        if (tid == 0) {
            tma_store_wait_read();                      // the previous tile's dX store has finished reading ZB
            mbar_expect_tx(bar, (uint32_t)(2 * TR * Cfg::ROWB + TR * Cfg::OROWB));
            tma_load4(GB, &tm_g, cbase, w0 - P, r0, n, bar);
            tma_load4(ZB, &tm_z, cbase, w0 - P, r0, n, bar);
            tma_load4(XW, &tm_x, cbase, w0, r0, n, bar);
        }
        mbar_wait(bar, ph);
        ph ^= 1;
        if (tlive) {
            // dZ = a * G * [scale*Z + shift > 0] + b * Z + c in place over G (zero outside the image)
            f2_t ca[4], cb2[4], cc[4], cs[4], ct[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = tchunk * 8 + 2 * i;
                ca[i] = f2_pack(coef[c], coef[c + 1]);
                cb2[i] = f2_pack(coef[CG + c], coef[CG + c + 1]);
                cc[i] = f2_pack(coef[2 * CG + c], coef[2 * CG + c + 1]);
                cs[i] = f2_pack(coef[3 * CG + c], coef[3 * CG + c + 1]);
                ct[i] = f2_pack(coef[4 * CG + c], coef[4 * CG + c + 1]);
            }
            for (int pix = pix0; pix < TR * HC; pix += PSTEP) {
                const int rr = pix / HC, cc_ = pix - rr * HC;
                if ((unsigned)(r0 + rr) < (unsigned)p.H && (unsigned)(w0 - P + cc_) < (unsigned)p.W) {
                    const uint32_t o = (uint32_t)(pix * PITCH + tchunk * 16);
                    const uint4 ug = lds128(GB + o), uz = lds128(ZB + o);
                    const uint32_t gg[4] = {ug.x, ug.y, ug.z, ug.w}, zz[4] = {uz.x, uz.y, uz.z, uz.w};
                    uint32_t r[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const f2_t z2 = f2_from_bf16x2(zz[i]);
                        float y0, y1, g0, g1, d0, d1;
                        f2_unpack(f2_fma(cs[i], z2, ct[i]), y0, y1);
                        f2_unpack(f2_from_bf16x2(gg[i]), g0, g1);
                        const f2_t gm = f2_pack(y0 > 0.f ? g0 : 0.f, y1 > 0.f ? g1 : 0.f);
                        f2_unpack(f2_fma(ca[i], gm, f2_fma(cb2[i], z2, cc[i])), d0, d1);
                        r[i] = pack_bf16x2(d0, d1);
                    }
                    sts128(GB + o, make_uint4(r[0], r[1], r[2], r[3]));
                }
            }
        }
        __syncthreads();
        const int sw = w0 + strip * 16;
        if (chunk_live && sw < p.W) {
            const float m0 = sw + g < p.W ? 1.f : 0.f, m1 = sw + g + 8 < p.W ? 1.f : 0.f;
            em.mk0 = f2_pack(m0, m0);
            em.mk1 = f2_pack(m1, m1);
            {   // (a) backward-data: rotated diagonal weight fragments, rebuilt per tile (dead during (b))
                uint32_t bd[K][K];
#pragma unroll
                for (int kh = 0; kh < K; ++kh)
#pragma unroll
                    for (int kw = 0; kw < K; ++kw) {
                        const uint32_t hb = (uint32_t)__bfloat16_as_ushort(wsm[(chunk * 8 + g) * KK + (K - 1 - kh) * K + (K - 1 - kw)]);
                        bd[kh][kw] = (g >> 1) == t ? ((g & 1) ? (hb << 16) : hb) : 0u;
                    }
                float acc[K][4];
#pragma unroll
                for (int i = 0; i < K; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
                DwsRows<K, CG, TWS, TH, 0, EM>::run(acc, bd, GB + off4, GB + off2, o0, p.H, em, 0u, 0u);
            }
            if (do_wgrad) {
                // (b) backward-weight: dZ tile row P + o (image row o0 + o) pairs with X tile rows o .. o + K - 1
                const uint32_t gz4 = GB + P * Cfg::ROWB + off4, gz2 = GB + P * Cfg::ROWB + off2, xrow = XW + offx;
                const int c0 = sw + 2 * t;
                const uint32_t cm0 = (c0 < p.W ? 0x0000ffffu : 0u) | (c0 + 1 < p.W ? 0xffff0000u : 0u);
                const uint32_t cm1 = (c0 + 8 < p.W ? 0x0000ffffu : 0u) | (c0 + 9 < p.W ? 0xffff0000u : 0u);
                auto load_x = [&](int xr, uint32_t (&dst)[2]) {
                    const bool rv = (unsigned)(r0 + xr) < (unsigned)p.H;
                    uint32_t u0, u1;
                    ldsm2t(xrow + (uint32_t)(xr * Cfg::OROWB), u0, u1);
                    if (act) {
                        u0 = f2_relu_bf16x2(f2_fma(f2_from_bf16x2(u0), ag, at));
                        u1 = f2_relu_bf16x2(f2_fma(f2_from_bf16x2(u1), ag, at));
                    }
                    dst[0] = rv ? u0 & cm0 : 0u;
                    dst[1] = rv ? u1 & cm1 : 0u;
                };
                float wacc[K][NPR][4];
#pragma unroll
                for (int kh = 0; kh < K; ++kh)
#pragma unroll
                    for (int pr = 0; pr < NPR; ++pr) wacc[kh][pr][0] = wacc[kh][pr][1] = wacc[kh][pr][2] = wacc[kh][pr][3] = 0.f;
                uint32_t bf[K][2];
#pragma unroll
                for (int kh = 0; kh + 1 < K; ++kh) load_x(kh, bf[kh]);
#pragma unroll
                for (int o = 0; o < TH; ++o) {
                    load_x(o + K - 1, bf[(o + K - 1) % K]);
                    uint32_t tt[K][2];
                    const uint32_t a4 = gz4 + (uint32_t)(o * Cfg::ROWB), a2 = gz2 + (uint32_t)(o * Cfg::ROWB);
                    ldsm4t(a4, tt[0][0], tt[0][1], tt[1][0], tt[1][1]);
                    if constexpr (K == 5) ldsm4t(a4 + 2 * PITCH, tt[2][0], tt[2][1], tt[3][0], tt[3][1]);
                    ldsm2t(a2, tt[K - 1][0], tt[K - 1][1]);
#pragma unroll
                    for (int kh = 0; kh < K; ++kh)
#pragma unroll
                        for (int pr = 0; pr < NPR; ++pr) {
                            const int s0 = 2 * pr, s1 = (2 * pr + 1 < K) ? 2 * pr + 1 : 2 * pr;
                            mma16816(wacc[kh][pr], tt[s0][0], tt[s1][0], tt[s0][1], tt[s1][1], bf[(o + kh) % K][0], bf[(o + kh) % K][1]);
                        }
                }
                if ((g >> 1) == t) {
                    float* dst = dwacc + (chunk * 8 + g) * KK;
#pragma unroll
                    for (int kh = 0; kh < K; ++kh)
#pragma unroll
                        for (int pr = 0; pr < NPR; ++pr) {
                            atomicAdd(dst + kh * K + (K - 1 - 2 * pr), (g & 1) ? wacc[kh][pr][1] : wacc[kh][pr][0]);
                            if (2 * pr + 1 < K) atomicAdd(dst + kh * K + (K - 2 - 2 * pr), (g & 1) ? wacc[kh][pr][3] : wacc[kh][pr][2]);
                        }
                }
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) tma_store4(&tm_dx, cbase, w0, o0, n, ZB);
    }
    if (tid == 0) tma_store_wait_read();
    if (do_red) {
        float v[4];
        f2_unpack(em.rs[0], v[0], v[1]);
        f2_unpack(em.rs[1], v[2], v[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[q] += __shfl_xor_sync(0xffffffffu, v[q], 4);
            v[q] += __shfl_xor_sync(0xffffffffu, v[q], 8);
            v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);
        }
        if (g == 0 && chunk_live) {
            const int c = chunk * 8 + 2 * t;
            atomicAdd(&red[c], v[0]);
            atomicAdd(&red[c + 1], v[1]);
            atomicAdd(&red[CG + c], v[2]);
            atomicAdd(&red[CG + c + 1], v[3]);
        }
    }
    __syncthreads();
    if (do_red) {
        for (int i = tid; i < CG; i += THREADS)
            if (cbase + i < p.C) {
                atomicAdd(&p.nsums[cbase + i], (double)red[i]);
                atomicAdd(&p.nsums[p.C + cbase + i], (double)red[CG + i]);
            }
    }
    if (do_wgrad) {
        for (int i = tid; i < CG * KK; i += THREADS)
            if (cbase + i / KK < p.C) atomicAdd(&p.dw[(size_t)cbase * KK + i], dwacc[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
static bool dws_fill(DwsP& p, const DwmGeom& g, int N, int H, int W, int C, int TH) {
    p.N = N; p.H = H; p.W = W; p.C = C;
    p.tiles_w = g.tiles_w; p.tiles_h = (H + TH - 1) / TH; p.cblocks = g.cblocks;
    const long long items = (long long)N * p.tiles_w * p.tiles_h;
    if (items > (1 << 24)) return false;
    p.items = (int)items;
    return true;
}

template <class Kern>
static int dws_grid(Kern kern, int threads, int smem, const DwsP& p, int& occ, const char* name) {
    if (occ < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return -1; }
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, smem);
        if (o < 1) { set_error("%s: kernel does not fit on an SM", name); return -1; }
        occ = o;
    }
    // resident CTAs only: surplus CTAs of a persistent grid would run as a second wave
    long long slots = (long long)num_sms() * occ / p.cblocks;
    if (slots < 1) slots = 1;
    if (slots > p.items) slots = p.items;
    return (int)(slots * p.cblocks);
}

template <int K, int CG, int TWS, int TH, bool FLIP>
static int dws_launch_fwd(const DwmGeom& g, const void* x, const float* s, const float* t, const float* w, void* z, double* stats,
                          int N, int H, int W, int C, cudaStream_t st, const char* name) {
    using Cfg = DwsCfg<K, CG, TWS, TH>;
    CUtensorMap tm_x, tm_z;
    if (int e = dwm_tensor_map(&tm_x, x, N, H, W, C, CG, Cfg::HC, Cfg::TR)) return e;
    if (int e = dwm_tensor_map(&tm_z, z, N, H, W, C, CG, Cfg::TW, TH)) return e;
    DwsP p = {};
    p.in_scale = s; p.in_shift = t; p.w = w; p.stats = stats;
    if (!dws_fill(p, g, N, H, W, C, TH)) return MNB_ERR_UNSUPPORTED;
    static int occ = -1;
    const int grid = dws_grid(dws_fwd_k<K, CG, TWS, TH, FLIP>, Cfg::THREADS, Cfg::SMEM, p, occ, name);
    if (grid < 0) return MNB_ERR_UNSUPPORTED;
    dws_fwd_k<K, CG, TWS, TH, FLIP><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(tm_x, tm_z, p);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

template <int K, int CG, int TWS, int TH>
static int dws_launch_wgrad(const DwmGeom& g, const void* x, const float* s, const float* t, const void* dz, float* dw, int N,
                            int H, int W, int C, cudaStream_t st, const char* name) {
    using Cfg = DwsCfg<K, CG, TWS, TH>;
    CUtensorMap tm_g, tm_x;
    if (int e = dwm_tensor_map(&tm_g, dz, N, H, W, C, CG, Cfg::HC, TH)) return e;
    if (int e = dwm_tensor_map(&tm_x, x, N, H, W, C, CG, Cfg::TW, Cfg::TR)) return e;
    DwsP p = {};
    p.in_scale = s; p.in_shift = t; p.dw = dw;
    if (!dws_fill(p, g, N, H, W, C, TH)) return MNB_ERR_UNSUPPORTED;
    static int occ = -1;
    const int grid = dws_grid(dws_wgrad_k<K, CG, TWS, TH>, Cfg::THREADS, Cfg::SMEM_WG, p, occ, name);
    if (grid < 0) return MNB_ERR_UNSUPPORTED;
    dws_wgrad_k<K, CG, TWS, TH><<<grid, Cfg::THREADS, Cfg::SMEM_WG, st>>>(tm_g, tm_x, p);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

// maps of at most 28 rows; the channel-group / strip geometry is dw_mma's (24 | 40 channels, 1 | 2 strips of 16 columns)
// option "dw_small": 1 = where the whole-tile kernels measured faster than the row-streaming ones -- forward / backward-data /
// backward-weight on maps of 12..64 rows, the fused backward on every map of >= 12 rows (5x5 up to 64 rows); 2 = every
// shape (tests); 0 = never
bool dw_small_covers(int H, int W, int C, int k) {
    const int o = option_get(OPT_DW_SMALL);
    if (!o || (k != 3 && k != 5) || C % 8 != 0) return false;
    return o == 2 || (H >= 12 && H <= 64 && W >= 12);
}
bool dw_small_covers_bwd(int H, int W, int C, int k) {
    const int o = option_get(OPT_DW_SMALL);
    if (!o || (k != 3 && k != 5) || C % 8 != 0) return false;
    return o == 2 || (H >= 12 && W >= 12 && (k == 3 || H <= 64));
}

#define DWS_DISPATCH(CALL)                                                    \
    if (g.CG == 24 && g.TWS == 1) { CALL(24, 1) }                             \
    if (g.CG == 24 && g.TWS == 2) { CALL(24, 2) }                             \
    if (g.CG == 40 && g.TWS == 1) { CALL(40, 1) }                             \
    { CALL(40, 2) }

template <int K, bool FLIP>
static int dws_fwd_any(const void* x, const float* s, const float* t, const float* w, void* z, double* stats, int N, int H, int W,
                       int C, cudaStream_t st, const char* name) {
    const DwmGeom g = dwm_geometry(C, W, K);
    const bool th7 = H <= 7;
#define CALL(CGG, TT)                                                                                                 \
    return th7 ? dws_launch_fwd<K, CGG, TT, 7, FLIP>(g, x, s, t, w, z, stats, N, H, W, C, st, name)                   \
               : dws_launch_fwd<K, CGG, TT, 14, FLIP>(g, x, s, t, w, z, stats, N, H, W, C, st, name);
    DWS_DISPATCH(CALL)
#undef CALL
}

template <int K>
static int dws_wgrad_any(const void* x, const float* s, const float* t, const void* dz, float* dw, int N, int H, int W, int C,
                         cudaStream_t st, const char* name) {
    const DwmGeom g = dwm_geometry(C, W, K);
    const bool th7 = H <= 7;
#define CALL(CGG, TT)                                                                                                 \
    return th7 ? dws_launch_wgrad<K, CGG, TT, 7>(g, x, s, t, dz, dw, N, H, W, C, st, name)                            \
               : dws_launch_wgrad<K, CGG, TT, 14>(g, x, s, t, dz, dw, N, H, W, C, st, name);
    DWS_DISPATCH(CALL)
#undef CALL
}

int dw_fwd_small(const void* x, const float* s, const float* t, const float* w, void* z, double* stats, int N, int H, int W,
                 int C, int k, cudaStream_t st) {
    if (k == 3) return dws_fwd_any<3, false>(x, s, t, w, z, stats, N, H, W, C, st, "dw_fwd(small)");
    return dws_fwd_any<5, false>(x, s, t, w, z, stats, N, H, W, C, st, "dw_fwd(small)");
}
int dw_dgrad_small(const void* dz, const float* w, void* dx, int N, int H, int W, int C, int k, cudaStream_t st) {
    if (k == 3) return dws_fwd_any<3, true>(dz, nullptr, nullptr, w, dx, nullptr, N, H, W, C, st, "dw_dgrad(small)");
    return dws_fwd_any<5, true>(dz, nullptr, nullptr, w, dx, nullptr, N, H, W, C, st, "dw_dgrad(small)");
}
int dw_wgrad_small(const void* x, const float* s, const float* t, const void* dz, float* dw, int N, int H, int W, int C, int k,
                   cudaStream_t st) {
    if (k == 3) return dws_wgrad_any<3>(x, s, t, dz, dw, N, H, W, C, st, "dw_wgrad(small)");
    return dws_wgrad_any<5>(x, s, t, dz, dw, N, H, W, C, st, "dw_wgrad(small)");
}

template <int K, int CG, int TWS, int TH>
static int dws_launch_bwd(const DwmGeom& g, DwsBP p, const void* G, const void* Z, const void* X, void* dX, cudaStream_t st,
                          const char* name) {
    using Cfg = DwsCfg<K, CG, TWS, TH>;
    using B = DwsBCfg<K, CG, TWS, TH>;
    const int N = p.N, H = p.H, W = p.W, C = p.C;
    CUtensorMap tm_g, tm_z, tm_x, tm_dx;
    if (int e = dwm_tensor_map(&tm_g, G, N, H, W, C, CG, Cfg::HC, Cfg::TR)) return e;
    if (int e = dwm_tensor_map(&tm_z, Z, N, H, W, C, CG, Cfg::HC, Cfg::TR)) return e;
    if (int e = dwm_tensor_map(&tm_x, X, N, H, W, C, CG, Cfg::TW, Cfg::TR)) return e;
    if (int e = dwm_tensor_map(&tm_dx, dX, N, H, W, C, CG, Cfg::TW, TH)) return e;
    p.tiles_w = g.tiles_w; p.tiles_h = (H + TH - 1) / TH; p.cblocks = g.cblocks;
    const long long items = (long long)N * p.tiles_w * p.tiles_h;
    if (items > (1 << 24)) return MNB_ERR_UNSUPPORTED;
    p.items = (int)items;
    static int occ = -1;
    if (occ < 0) {
        cudaError_t e = cudaFuncSetAttribute(dws_bwd_k<K, CG, TWS, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, B::SMEM);
        if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return (int)e; }
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, dws_bwd_k<K, CG, TWS, TH>, Cfg::THREADS, B::SMEM);
        if (o < 1) { set_error("%s: kernel does not fit on an SM", name); return MNB_ERR_UNSUPPORTED; }
        occ = o;
    }
    long long slots = (long long)num_sms() * occ / p.cblocks;
    if (slots < 1) slots = 1;
    if (slots > p.items) slots = p.items;
    dws_bwd_k<K, CG, TWS, TH><<<(unsigned)(slots * p.cblocks), Cfg::THREADS, B::SMEM, st>>>(tm_g, tm_z, tm_x, tm_dx, p);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

int dw_bwd_small(const void* G, const void* Z, const float* scale, const float* shift, const double* sums, const float* mean,
                 const float* invstd, double m, float* dgamma, float* dbeta, float* dbias, const void* X,
                 const float* in_scale, const float* in_shift, const float* w, void* dX, float* dw, double* nsums, int N, int H,
                 int W, int C, int k, cudaStream_t st) {
    DwsBP p = {};
    p.scale = scale; p.shift = shift; p.sums = sums; p.mean = mean; p.invstd = invstd; p.m = m;
    p.dgamma = dgamma; p.dbeta = dbeta; p.dbias = dbias; p.in_scale = in_scale; p.in_shift = in_shift; p.w = w; p.dw = dw;
    p.nsums = nsums; p.N = N; p.H = H; p.W = W; p.C = C;
    const DwmGeom g = dwm_geometry(C, W, k);
    const char* name = "dw_bwd(small)";
    const bool th7 = H <= 7;
#define CALL3(CGG, TT) return th7 ? dws_launch_bwd<3, CGG, TT, 7>(g, p, G, Z, X, dX, st, name) : dws_launch_bwd<3, CGG, TT, 14>(g, p, G, Z, X, dX, st, name);
#define CALL5(CGG, TT) return th7 ? dws_launch_bwd<5, CGG, TT, 7>(g, p, G, Z, X, dX, st, name) : dws_launch_bwd<5, CGG, TT, 14>(g, p, G, Z, X, dX, st, name);
    if (k == 3) { DWS_DISPATCH(CALL3) }
    if (k == 5) { DWS_DISPATCH(CALL5) }
#undef CALL3
#undef CALL5
    return MNB_ERR_UNSUPPORTED;
}

}  // namespace mnb
