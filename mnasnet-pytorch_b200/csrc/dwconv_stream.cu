// Depthwise k x k (3 / 5) stride-1 convolution, bf16 NHWC, as register-resident ROW STREAMS (forward, dgrad, wgrad).
// Replaces nn.Conv2d(groups=C) (src/models/mnasnet.py:76-81,120-125) with the producing ConvBlock's BN-apply+ReLU
// fused into the load and this ConvBlock's BN statistics fused into the store, like dwconv_tile.cu -- but with no
// shared memory, no barriers and no vertical halo:
//
//   * a lane owns one channel PAIR (k*k packed fp32x2 weights in registers) and a strip of TW output columns; it
//     walks DOWN the image: every input row of the strip (TW+k-1 four-byte loads, the next row already in flight)
//     is transformed once and scattered into the k pending output rows held in a register ring, the completed
//     output row is rounded, stored and folded into the statistics.  Each input row is therefore loaded and
//     transformed exactly once per strip (the tile kernel reloads (TH+k-1)/TH of them and pays cp.async address
//     generation, a shared-memory round trip and two barriers per tile: ~1500 instructions per thread per tile of
//     which 350 are FFMA2 for 5x5).
//   * a warp packs G strips x PL channel pairs (PL*G <= 32, PL >= 8 so that a lane group reads whole 32-byte
//     sectors): all the MNASNet widths except 72 fill 32 lanes (72 channels: 3 strips x 9 pairs = 27 lanes).
//   * a warp keeps ONE channel block for its whole life (its statistics / weight gradients stay in registers and
//     are flushed once) and strides over the spatial tasks (image, row segment, strip group).
//
// The per-lane program is plain C++ over global pointers (no warp collectives), so the SAME source runs on the
// host: tests/test_dw_stream_cpu.py builds this file with -DMNB_DW_STREAM_EMUL and checks the lane program, the task
// decomposition and the padding logic against torch on the CPU before any GPU time is spent.
#include "common.cuh"

#if defined(__CUDA_ARCH__)
#define MNB_DEVICE_CODE 1
#else
#define MNB_DEVICE_CODE 0
#endif
#define MNB_HD __host__ __device__ __forceinline__

namespace mnb {

// ---- packed fp32 pair: one FFMA2 on sm_100, two fmaf on the host --------------------------------------------------
struct F2 {
#if MNB_DEVICE_CODE
    unsigned long long v;
#else
    float lo, hi;
#endif
};
MNB_HD F2 f2_make(float lo, float hi) {
    F2 r;
#if MNB_DEVICE_CODE
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
#else
    r.lo = lo; r.hi = hi;
#endif
    return r;
}
MNB_HD void f2_get(const F2& a, float& lo, float& hi) {
#if MNB_DEVICE_CODE
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
#else
    lo = a.lo; hi = a.hi;
#endif
}
MNB_HD F2 f2_fma(const F2& a, const F2& b, const F2& c) {
    F2 r;
#if MNB_DEVICE_CODE
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
#else
    r.lo = fmaf(a.lo, b.lo, c.lo); r.hi = fmaf(a.hi, b.hi, c.hi);
#endif
    return r;
}
MNB_HD float bits_f(uint32_t u) {
#if MNB_DEVICE_CODE
    return __uint_as_float(u);
#else
    union { uint32_t i; float f; } c; c.i = u; return c.f;
#endif
}
MNB_HD uint32_t f_bits(float f) {
#if MNB_DEVICE_CODE
    return __float_as_uint(f);
#else
    union { uint32_t i; float f; } c; c.f = f; return c.i;
#endif
}
MNB_HD F2 f2_from_bf16x2(uint32_t u) { return f2_make(bits_f(u << 16), bits_f(u & 0xffff0000u)); }
// round-to-nearest-even fp32 -> bf16 (finite inputs; NaN is not produced by this path), packed pair
MNB_HD uint32_t bf16_rn_bits(float f) {
    uint32_t u = f_bits(f);
    u += 0x7fffu + ((u >> 16) & 1u);
    return u >> 16;
}
MNB_HD uint32_t pack2_rn(float lo, float hi) {
#if MNB_DEVICE_CODE
    return pack_bf16x2(lo, hi);
#else
    return bf16_rn_bits(lo) | (bf16_rn_bits(hi) << 16);
#endif
}
MNB_HD uint32_t ld32(const bf16* p) {
#if MNB_DEVICE_CODE
    return __ldg(reinterpret_cast<const uint32_t*>(p));
#else
    return *reinterpret_cast<const uint32_t*>(p);
#endif
}

enum { DWS_FWD = 0, DWS_DGRAD = 1, DWS_WGRAD = 2 };

struct DwSP {
    const bf16* x;          // fwd / wgrad: input activations (raw conv output of the producer); dgrad: dz
    const float* in_scale;  // fwd / wgrad: producer BN scale / shift (NULL = raw)
    const float* in_shift;
    const float* w;         // [C][k][k]
    const float* bias;      // fwd, may be NULL
    const bf16* dz;         // wgrad
    bf16* out;              // fwd: z ; dgrad: dx
    float* dw;              // wgrad: [C][k][k], accumulated into
    double* stats;          // fwd: [2C] sum / sum of squares of the stored values, may be NULL
    int N, H, W, C;
    int PL, G;              // channel pairs per lane group, strips per warp (PL * G <= 32)
    int NB;                 // channel blocks = (C / 2) / PL
    int HS;                 // output rows per task
    int nws, nhs;           // strip groups per row = ceil(W / (TW * G)), row segments = ceil(H / HS)
    int spatial_tasks;      // N * nhs * nws
    int warps_per_cb;       // warps that share one channel block (task stride)
};

// what a lane hands to the warp-level flush
template <int K>
struct DwLaneOut {
    float st[4];            // fwd: sum(ch0), sum(ch1), sumsq(ch0), sumsq(ch1)
    F2 wg[K][K];            // wgrad: dW of the channel pair
};

// ---- one (image, row segment, strip) of forward / backward-data for one channel pair ------------------------------
// All addressing is 32-bit element offsets from the tensor base (the host checks N*H*W*C < 2^31): one running row
// offset plus the per-lane constants jC[j] = j*C.  EDGE = the warp's strips touch the left / right image border (then
// every column access is predicated by a bit mask); interior warps take the unpredicated variant: only the row
// validity is tested, and that is uniform over the warp.  PD = input rows in flight ahead of the one being consumed.
template <int K, int TW, int MODE, int PD, bool EDGE>
MNB_HD void dws_conv_task(const DwSP& p, int n, int h0, int h1, int c0, int ch, const int (&jC)[TW + K - 1],
                          const F2 (&wr)[K][K], bool xf, const F2& sc, const F2& sh, float b0, float b1, F2& ssum,
                          F2& ssq) {
    constexpr int P = K / 2, NI = TW + K - 1;
    const int H = p.H, W = p.W;
    const int rs = W * p.C;                             // row stride in elements
    unsigned cmask = (1u << NI) - 1u, omask = (1u << TW) - 1u;
    if (EDGE) {
        cmask = 0u; omask = 0u;
#pragma unroll
        for (int j = 0; j < NI; ++j) { const int col = c0 - P + j; if (col >= 0 && col < W) cmask |= 1u << j; }
#pragma unroll
        for (int j = 0; j < TW; ++j) if (c0 + j < W) omask |= 1u << j;
    }
    const F2 zero2 = f2_make(0.f, 0.f), one2 = f2_make(1.f, 1.f);
    F2 acc[K][TW];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < TW; ++j) acc[i][j] = zero2;

    const int first_ir = h0 - P, last_ir = h1 - 1 + P;
    // offsets of (row, column c0-P) of the input and (row, column c0) of the output; positions outside the image
    // are never dereferenced
    int off_next = ((n * H + first_ir) * W + (c0 - P)) * p.C + ch;
    int ooff = ((n * H + h0) * W + c0) * p.C + ch;
    uint32_t raw[PD][NI];
    auto load_row = [&](int rr, int roff, uint32_t (&dst)[NI]) {
        const bool rok = rr >= 0 && rr < H && rr <= last_ir;
        if (EDGE) {
#pragma unroll
            for (int j = 0; j < NI; ++j) dst[j] = (rok && ((cmask >> j) & 1u)) ? ld32(p.x + (roff + jC[j])) : 0u;
        } else if (rok) {
#pragma unroll
            for (int j = 0; j < NI; ++j) dst[j] = ld32(p.x + (roff + jC[j]));
        } else {
#pragma unroll
            for (int j = 0; j < NI; ++j) dst[j] = 0u;
        }
    };
#pragma unroll
    for (int d = 0; d < PD; ++d) { load_row(first_ir + d, off_next, raw[d]); off_next += rs; }
    for (int ir = first_ir; ir <= last_ir; ir += K) {
#pragma unroll
        for (int u = 0; u < K; ++u) {
            const int r = ir + u;                      // input row (uniform over the warp)
            if (r <= last_ir) {
                const bool rok = r >= 0 && r < H;
                F2 in[NI];
                if (MODE == DWS_FWD && xf) {
                    // zero padding is applied AFTER the activation (the conv pads the activated tensor)
                    if (EDGE) {
#pragma unroll
                        for (int j = 0; j < NI; ++j) {
                            float a, b;
                            f2_get(f2_fma(sc, f2_from_bf16x2(raw[0][j]), sh), a, b);
                            const bool ok = rok && ((cmask >> j) & 1u);
                            in[j] = f2_make(ok ? fmaxf(a, 0.f) : 0.f, ok ? fmaxf(b, 0.f) : 0.f);
                        }
                    } else if (rok) {
#pragma unroll
                        for (int j = 0; j < NI; ++j) {
                            float a, b;
                            f2_get(f2_fma(sc, f2_from_bf16x2(raw[0][j]), sh), a, b);
                            in[j] = f2_make(fmaxf(a, 0.f), fmaxf(b, 0.f));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < NI; ++j) in[j] = zero2;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NI; ++j) in[j] = f2_from_bf16x2(raw[0][j]);       // invalid positions hold 0
                }
                // raw[0] is consumed: shift the ring and put row r+PD in flight
#pragma unroll
                for (int d = 0; d + 1 < PD; ++d)
#pragma unroll
                    for (int j = 0; j < NI; ++j) raw[d][j] = raw[d + 1][j];
                load_row(r + PD, off_next, raw[PD - 1]);
                off_next += rs;
                // This input row feeds output row o = r + P - kh through kernel row kh; slot (o - h0) % K = (u - kh) % K is
                // static after unrolling.  Kernel row 0 is always the FIRST contribution to its output row, so it
                // overwrites the slot: no zeroing, and rows outside [h0, h1) need no test -- what they leave in a slot
                // is overwritten before the slot's next real row starts, or never stored.
#pragma unroll
                for (int kh = 0; kh < K; ++kh) {
                    const int slot = (u - kh + 2 * K) % K;
#pragma unroll
                    for (int tc = 0; tc < TW; ++tc)
#pragma unroll
                        for (int kw = 0; kw < K; ++kw)
                            acc[slot][tc] = f2_fma(in[tc + kw], wr[kh][kw], (kh == 0 && kw == 0) ? zero2 : acc[slot][tc]);
                }
                const int od = r - P;                  // output row completed by this input row (kernel row K-1)
                if (od >= h0 && od < h1) {
                    const int slot = (u + 1) % K;      // (u - (K-1) + 2K) % K
#pragma unroll
                    for (int tc = 0; tc < TW; ++tc) {
                        float v0, v1;
                        f2_get(acc[slot][tc], v0, v1);
                        if (!EDGE || ((omask >> tc) & 1u)) {
                            const uint32_t pk = pack2_rn(v0 + b0, v1 + b1);
                            *reinterpret_cast<uint32_t*>(p.out + (ooff + jC[tc])) = pk;
                            if (MODE == DWS_FWD) {     // statistics of the stored (rounded) values
                                const F2 q = f2_from_bf16x2(pk);
                                ssum = f2_fma(q, one2, ssum);
                                ssq = f2_fma(q, q, ssq);
                            }
                        }
                    }
                    ooff += rs;
                }
            }
        }
    }
}

// ---- one (image, row segment, strip) of backward-weight for one channel pair ---------------------------------------
template <int K, int TW, int PD, bool EDGE>
MNB_HD void dws_wgrad_task(const DwSP& p, int n, int h0, int h1, int c0, int ch, const int (&jC)[TW + K - 1],
                           F2 (&wg)[K][K], bool xf, const F2& sc, const F2& sh) {
    constexpr int P = K / 2, NI = TW + K - 1;
    const int H = p.H, W = p.W;
    const int rs = W * p.C;
    unsigned cmask = (1u << NI) - 1u, omask = (1u << TW) - 1u;
    if (EDGE) {
        cmask = 0u; omask = 0u;
#pragma unroll
        for (int j = 0; j < NI; ++j) { const int col = c0 - P + j; if (col >= 0 && col < W) cmask |= 1u << j; }
#pragma unroll
        for (int j = 0; j < TW; ++j) if (c0 + j < W) omask |= 1u << j;
    }
    const F2 zero2 = f2_make(0.f, 0.f);
    F2 g[K][TW];                                       // ring of dz rows, slot = (row - h0) % K
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < TW; ++j) g[i][j] = zero2;

    const int first_ir = h0 - P, last_ir = h1 - 1 + P;
    int xoff_next = ((n * H + first_ir) * W + (c0 - P)) * p.C + ch;      // input row rr, column c0-P
    int goff_next = ((n * H + first_ir + P) * W + c0) * p.C + ch;        // dz row rr+P, column c0
    uint32_t rawx[PD][NI], rawg[PD][TW];
    auto load_rows = [&](int rr, int xoff, int goff, uint32_t (&dx)[NI], uint32_t (&dg)[TW]) {
        const bool rok = rr >= 0 && rr < H && rr <= last_ir;
        const int o = rr + P;
        const bool ook = o >= h0 && o < h1 && rr <= last_ir;
        if (EDGE) {
#pragma unroll
            for (int j = 0; j < NI; ++j) dx[j] = (rok && ((cmask >> j) & 1u)) ? ld32(p.x + (xoff + jC[j])) : 0u;
#pragma unroll
            for (int tc = 0; tc < TW; ++tc) dg[tc] = (ook && ((omask >> tc) & 1u)) ? ld32(p.dz + (goff + jC[tc])) : 0u;
        } else {
            if (rok) {
#pragma unroll
                for (int j = 0; j < NI; ++j) dx[j] = ld32(p.x + (xoff + jC[j]));
            } else {
#pragma unroll
                for (int j = 0; j < NI; ++j) dx[j] = 0u;
            }
            if (ook) {
#pragma unroll
                for (int tc = 0; tc < TW; ++tc) dg[tc] = ld32(p.dz + (goff + jC[tc]));
            } else {
#pragma unroll
                for (int tc = 0; tc < TW; ++tc) dg[tc] = 0u;
            }
        }
    };
#pragma unroll
    for (int d = 0; d < PD; ++d) {
        load_rows(first_ir + d, xoff_next, goff_next, rawx[d], rawg[d]);
        xoff_next += rs; goff_next += rs;
    }
    for (int ir = first_ir; ir <= last_ir; ir += K) {
#pragma unroll
        for (int u = 0; u < K; ++u) {
            const int r = ir + u;                      // input row
            if (r <= last_ir) {
                const bool rok = r >= 0 && r < H;
                // dz row r + P enters the ring (kernel row 0 pairs it with this input row): slot u
#pragma unroll
                for (int tc = 0; tc < TW; ++tc) g[u][tc] = f2_from_bf16x2(rawg[0][tc]);
                F2 in[NI];
                if (xf) {
                    if (EDGE) {
#pragma unroll
                        for (int j = 0; j < NI; ++j) {
                            float a, b;
                            f2_get(f2_fma(sc, f2_from_bf16x2(rawx[0][j]), sh), a, b);
                            const bool ok = rok && ((cmask >> j) & 1u);
                            in[j] = f2_make(ok ? fmaxf(a, 0.f) : 0.f, ok ? fmaxf(b, 0.f) : 0.f);
                        }
                    } else if (rok) {
#pragma unroll
                        for (int j = 0; j < NI; ++j) {
                            float a, b;
                            f2_get(f2_fma(sc, f2_from_bf16x2(rawx[0][j]), sh), a, b);
                            in[j] = f2_make(fmaxf(a, 0.f), fmaxf(b, 0.f));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < NI; ++j) in[j] = zero2;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NI; ++j) in[j] = f2_from_bf16x2(rawx[0][j]);
                }
#pragma unroll
                for (int d = 0; d + 1 < PD; ++d) {
#pragma unroll
                    for (int j = 0; j < NI; ++j) rawx[d][j] = rawx[d + 1][j];
#pragma unroll
                    for (int tc = 0; tc < TW; ++tc) rawg[d][tc] = rawg[d + 1][tc];
                }
                load_rows(r + PD, xoff_next, goff_next, rawx[PD - 1], rawg[PD - 1]);
                xoff_next += rs; goff_next += rs;
                // dz row o = r + P - kh pairs with this input row through kernel row kh; the ring holds zeros for rows
                // outside [h0, h1) (initial state / masked loads), so no range test is needed
#pragma unroll
                for (int kh = 0; kh < K; ++kh) {
                    const int slot = (u - kh + 2 * K) % K;
#pragma unroll
                    for (int tc = 0; tc < TW; ++tc)
#pragma unroll
                        for (int kw = 0; kw < K; ++kw) wg[kh][kw] = f2_fma(in[tc + kw], g[slot][tc], wg[kh][kw]);
                }
            }
        }
    }
}

// ---- the whole life of one lane ------------------------------------------------------------------------------------
template <int K, int TW, int MODE, int PD>
MNB_HD void dws_lane(const DwSP& p, int warp_global, int lane, DwLaneOut<K>& out) {
    constexpr int P = K / 2, NI = TW + K - 1;
    const F2 zero2 = f2_make(0.f, 0.f);
    out.st[0] = out.st[1] = out.st[2] = out.st[3] = 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) out.wg[i][j] = zero2;
    const int cb = warp_global % p.NB, widx = warp_global / p.NB;
    const int gi = lane / p.PL, pl = lane - gi * p.PL;
    if (gi >= p.G || widx >= p.warps_per_cb) return;
    const int ch = (cb * p.PL + pl) * 2;
    const bool xf = (MODE != DWS_DGRAD) && p.in_scale != nullptr;
    F2 sc = f2_make(1.f, 1.f), sh = zero2;
    if (xf) { sc = f2_make(p.in_scale[ch], p.in_scale[ch + 1]); sh = f2_make(p.in_shift[ch], p.in_shift[ch + 1]); }
    float b0 = 0.f, b1 = 0.f;
    if (MODE == DWS_FWD && p.bias) { b0 = p.bias[ch]; b1 = p.bias[ch + 1]; }
    F2 wr[K][K];
    if (MODE != DWS_WGRAD) {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) {
                // backward-data correlates dz with the 180-degree rotated kernel
                const int ii = MODE == DWS_DGRAD ? K - 1 - i : i, jj = MODE == DWS_DGRAD ? K - 1 - j : j;
                wr[i][j] = f2_make(p.w[(ch + 0) * K * K + ii * K + jj], p.w[(ch + 1) * K * K + ii * K + jj]);
            }
    }
    int jC[NI];
#pragma unroll
    for (int j = 0; j < NI; ++j) jC[j] = j * p.C;
    F2 ssum = zero2, ssq = zero2;
    const int per_img = p.nhs * p.nws;
    for (int task = widx; task < p.spatial_tasks; task += p.warps_per_cb) {
        const int n = task / per_img, rem = task - n * per_img;
        const int hs = rem / p.nws, ws = rem - hs * p.nws;
        const int c0 = (ws * p.G + gi) * TW;
        const int h0 = hs * p.HS;
        int h1 = h0 + p.HS;
        if (h1 > p.H) h1 = p.H;
        // uniform over the warp: do this warp's strips (columns [ws*G*TW, (ws+1)*G*TW) plus the halo) touch a border?
        const bool edge = ws == 0 || (ws + 1) * p.G * TW + P > p.W;
        if (edge) {
            if (c0 >= p.W) continue;
            if (MODE == DWS_WGRAD) dws_wgrad_task<K, TW, PD, true>(p, n, h0, h1, c0, ch, jC, out.wg, xf, sc, sh);
            else dws_conv_task<K, TW, MODE, PD, true>(p, n, h0, h1, c0, ch, jC, wr, xf, sc, sh, b0, b1, ssum, ssq);
        } else {
            if (MODE == DWS_WGRAD) dws_wgrad_task<K, TW, PD, false>(p, n, h0, h1, c0, ch, jC, out.wg, xf, sc, sh);
            else dws_conv_task<K, TW, MODE, PD, false>(p, n, h0, h1, c0, ch, jC, wr, xf, sc, sh, b0, b1, ssum, ssq);
        }
    }
    f2_get(ssum, out.st[0], out.st[1]);
    f2_get(ssq, out.st[2], out.st[3]);
}

// ---- device kernel ---------------------------------------------------------------------------------------------------
template <int K, int TW, int MODE, int PD, int MINB>
__global__ void __launch_bounds__(128, MINB) dw_stream_k(const DwSP p) {
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    DwLaneOut<K> o;
    dws_lane<K, TW, MODE, PD>(p, warp_global, lane, o);
    // sum the G strips of a channel pair into lane group 0 (every lane takes part in the shuffles)
    const int pl = lane % p.PL;
    if (MODE == DWS_FWD) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v = o.st[q], s = v;
            for (int gi = 1; gi < p.G; ++gi) s += __shfl_sync(0xffffffffu, v, pl + gi * p.PL);
            o.st[q] = s;
        }
    }
    if (MODE == DWS_WGRAD) {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) {
                float a, b;
                f2_get(o.wg[i][j], a, b);
                float sa = a, sb = b;
                for (int gi = 1; gi < p.G; ++gi) {
                    sa += __shfl_sync(0xffffffffu, a, pl + gi * p.PL);
                    sb += __shfl_sync(0xffffffffu, b, pl + gi * p.PL);
                }
                o.wg[i][j] = f2_make(sa, sb);
            }
    }
    const int cb = warp_global % p.NB;
    if (lane < p.PL && warp_global / p.NB < p.warps_per_cb) {
        const int ch = (cb * p.PL + lane) * 2;
        if (MODE == DWS_FWD && p.stats) {
            atomicAdd(&p.stats[ch], (double)o.st[0]);
            atomicAdd(&p.stats[ch + 1], (double)o.st[1]);
            atomicAdd(&p.stats[p.C + ch], (double)o.st[2]);
            atomicAdd(&p.stats[p.C + ch + 1], (double)o.st[3]);
        }
        if (MODE == DWS_WGRAD) {
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    float a, b;
                    f2_get(o.wg[i][j], a, b);
                    atomicAdd(&p.dw[(ch + 0) * K * K + i * K + j], a);
                    atomicAdd(&p.dw[(ch + 1) * K * K + i * K + j], b);
                }
        }
    }
}

// ---- host side: geometry ---------------------------------------------------------------------------------------------
// lanes: the largest PL * G <= 32 with PL | C/2 and PL >= 8 (whole 32-byte sectors per lane group); ties -> fewer
// wasted columns at the right edge, then larger PL
static bool dws_geometry(DwSP& p, int TW, int total_warps) {
    const int P2 = p.C / 2;
    int best_pl = 0, best_g = 0;
    double best_cost = 1e30;
    for (int pl = 8; pl <= 32; ++pl) {
        if (P2 % pl) continue;
        for (int g = 1; g * pl <= 32 && g <= 4; ++g) {
            const int cols = TW * g;
            const double cover = (double)((p.W + cols - 1) / cols) * cols / p.W;      // column over-coverage
            const double cost = cover * 32.0 / (pl * g) - 1e-3 * pl;
            if (cost < best_cost) { best_cost = cost; best_pl = pl; best_g = g; }
        }
    }
    if (!best_pl) return false;
    p.PL = best_pl; p.G = best_g; p.NB = P2 / best_pl;
    p.nws = (p.W + TW * p.G - 1) / (TW * p.G);
    p.warps_per_cb = total_warps / p.NB;
    if (p.warps_per_cb < 1) return false;
    // row segments: as long as possible (no vertical halo) while every warp still gets >= ~6 tasks
    int hs = p.H;
    while (hs > 14) {
        const long long tasks = (long long)p.N * ((p.H + hs - 1) / hs) * p.nws;
        if (tasks >= 6ll * p.warps_per_cb) break;
        hs = (hs + 1) / 2;
    }
    p.HS = hs;
    p.nhs = (p.H + hs - 1) / hs;
    const long long st = (long long)p.N * p.nhs * p.nws;
    // 32-bit element offsets inside the kernel (one extra row of slack for the prefetch offsets)
    if (st >= (1ll << 31) || ((long long)p.N * p.H + 8) * p.W * p.C >= (1ll << 31)) return false;
    p.spatial_tasks = (int)st;
    return true;
}

template <int K, int TW, int MODE, int MINB>
static int dws_launch(DwSP p, cudaStream_t st, const char* name) {
    const int blocks = num_sms() * MINB;               // 128-thread CTAs: MINB = 4 -> 128 registers per lane, 3 -> 168
    if (!dws_geometry(p, TW, blocks * 4)) { set_error("%s: shape not covered by the row-stream kernel", name); return MNB_ERR_UNSUPPORTED; }
    const int pd = option_get(OPT_DW_STREAM_PD);       // rows of prefetch (1..3), see include/mnb200.h
    if (pd >= 3) dw_stream_k<K, TW, MODE, 3, MINB><<<blocks, 128, 0, st>>>(p);
    else if (pd == 2) dw_stream_k<K, TW, MODE, 2, MINB><<<blocks, 128, 0, st>>>(p);
    else dw_stream_k<K, TW, MODE, 1, MINB><<<blocks, 128, 0, st>>>(p);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

static bool dws_shape_ok(int C, int k) { return (k == 3 || k == 5) && C % 2 == 0; }   // lane packing: dws_geometry

int dw_fwd_stream(const void* x, const float* s, const float* t, const float* w, const float* bias, void* z,
                  double* stats, int N, int H, int W, int C, int k, cudaStream_t st) {
    if (!dws_shape_ok(C, k)) { set_error("dw_fwd(stream): C=%d k=%d not covered", C, k); return MNB_ERR_UNSUPPORTED; }
    DwSP p = {};
    p.x = (const bf16*)x; p.in_scale = s; p.in_shift = t; p.w = w; p.bias = bias; p.out = (bf16*)z; p.stats = stats;
    p.N = N; p.H = H; p.W = W; p.C = C;
    if (k == 3 && option_get(OPT_DW_STREAM_TW8)) return dws_launch<3, 8, DWS_FWD, 3>(p, st, "dw_fwd(stream, 8 columns)");
    if (k == 3) return dws_launch<3, 4, DWS_FWD, 4>(p, st, "dw_fwd(stream)");
    return dws_launch<5, 4, DWS_FWD, 3>(p, st, "dw_fwd(stream)");
}
int dw_dgrad_stream(const void* dz, const float* w, void* dx, int N, int H, int W, int C, int k, cudaStream_t st) {
    if (!dws_shape_ok(C, k)) { set_error("dw_dgrad(stream): C=%d k=%d not covered", C, k); return MNB_ERR_UNSUPPORTED; }
    DwSP p = {};
    p.x = (const bf16*)dz; p.w = w; p.out = (bf16*)dx;
    p.N = N; p.H = H; p.W = W; p.C = C;
    if (k == 3 && option_get(OPT_DW_STREAM_TW8)) return dws_launch<3, 8, DWS_DGRAD, 3>(p, st, "dw_dgrad(stream, 8 columns)");
    if (k == 3) return dws_launch<3, 4, DWS_DGRAD, 4>(p, st, "dw_dgrad(stream)");
    return dws_launch<5, 4, DWS_DGRAD, 3>(p, st, "dw_dgrad(stream)");
}
int dw_wgrad_stream(const void* x, const float* s, const float* t, const void* dz, float* dw, int N, int H, int W,
                    int C, int k, cudaStream_t st) {
    if (!dws_shape_ok(C, k)) { set_error("dw_wgrad(stream): C=%d k=%d not covered", C, k); return MNB_ERR_UNSUPPORTED; }
    DwSP p = {};
    p.x = (const bf16*)x; p.in_scale = s; p.in_shift = t; p.dz = (const bf16*)dz; p.dw = dw;
    p.N = N; p.H = H; p.W = W; p.C = C;
    if (k == 3) return dws_launch<3, 4, DWS_WGRAD, 4>(p, st, "dw_wgrad(stream)");
    return dws_launch<5, 2, DWS_WGRAD, 3>(p, st, "dw_wgrad(stream)");
}

}  // namespace mnb

// ---- host emulation of the lane program (test infrastructure; compiled only with -DMNB_DW_STREAM_EMUL) -------------------
#if defined(MNB_DW_STREAM_EMUL) && !defined(__CUDA_ARCH__)
namespace mnb {
void set_error(const char*, ...) {}
int option_get(int) { return 1; }
// flush of one lane's results after lane group 0 collected its G strips (mirrors the tail of dw_stream_k)
template <int K, int MODE>
static void dws_flush_lane(const DwSP& p, int warp_global, int lane, const DwLaneOut<K>& o,
                           void (*add_f64)(double*, double), void (*add_f32)(float*, float)) {
    const int cb = warp_global % p.NB;
    if (lane >= p.PL || warp_global / p.NB >= p.warps_per_cb) return;
    const int ch = (cb * p.PL + lane) * 2;
    if (MODE == DWS_FWD && p.stats) {
        add_f64(&p.stats[ch], (double)o.st[0]);
        add_f64(&p.stats[ch + 1], (double)o.st[1]);
        add_f64(&p.stats[p.C + ch], (double)o.st[2]);
        add_f64(&p.stats[p.C + ch + 1], (double)o.st[3]);
    }
    if (MODE == DWS_WGRAD) {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) {
                float a, b;
                f2_get(o.wg[i][j], a, b);
                add_f32(&p.dw[(ch + 0) * K * K + i * K + j], a);
                add_f32(&p.dw[(ch + 1) * K * K + i * K + j], b);
            }
    }
}

static void emul_add_f64(double* p, double v) { *p += v; }
static void emul_add_f32(float* p, float v) { *p += v; }

template <int K, int TW, int MODE, int PD>
static int emul_run_pd(DwSP p, int total_warps) {
    if (!dws_geometry(p, TW, total_warps)) return MNB_ERR_UNSUPPORTED;
    for (int wgl = 0; wgl < total_warps; ++wgl) {
        DwLaneOut<K> lanes[32];
        for (int lane = 0; lane < 32; ++lane) dws_lane<K, TW, MODE, PD>(p, wgl, lane, lanes[lane]);
        // what the shuffles do: lane group 0 collects the strips of its channel pair
        for (int pl = 0; pl < p.PL; ++pl)
            for (int gi = 1; gi < p.G; ++gi) {
                const DwLaneOut<K>& s = lanes[pl + gi * p.PL];
                for (int q = 0; q < 4; ++q) lanes[pl].st[q] += s.st[q];
                for (int i = 0; i < K; ++i)
                    for (int j = 0; j < K; ++j) {
                        lanes[pl].wg[i][j].lo += s.wg[i][j].lo;
                        lanes[pl].wg[i][j].hi += s.wg[i][j].hi;
                    }
            }
        for (int lane = 0; lane < 32; ++lane) dws_flush_lane<K, MODE>(p, wgl, lane, lanes[lane], emul_add_f64, emul_add_f32);
    }
    return 0;
}
}  // namespace mnb

namespace mnb {
static int g_emul_pd = 1, g_emul_tw8 = 0;
template <int K, int TW, int MODE>
static int emul_run(const DwSP& p, int total_warps) {
    if (g_emul_pd >= 3) return emul_run_pd<K, TW, MODE, 3>(p, total_warps);
    if (g_emul_pd == 2) return emul_run_pd<K, TW, MODE, 2>(p, total_warps);
    return emul_run_pd<K, TW, MODE, 1>(p, total_warps);
}
}  // namespace mnb
extern "C" void mnb_emul_dw_stream_set_pd(int pd) { mnb::g_emul_pd = pd; }
extern "C" void mnb_emul_dw_stream_set_tw8(int on) { mnb::g_emul_tw8 = on; }

// all pointers are HOST pointers; mode 0 fwd / 1 dgrad / 2 wgrad; geometry[6] returns PL, G, NB, HS, nws, nhs
extern "C" int mnb_emul_dw_stream(int mode, const void* x, const float* s, const float* t, const float* w,
                                  const float* bias, const void* dz, void* out, float* dw, double* stats, int N, int H,
                                  int W, int C, int k, int total_warps, int* geometry) {
    using namespace mnb;
    DwSP p = {};
    p.x = (const bf16*)x; p.in_scale = s; p.in_shift = t; p.w = w; p.bias = bias; p.dz = (const bf16*)dz;
    p.out = (bf16*)out; p.dw = dw; p.stats = stats; p.N = N; p.H = H; p.W = W; p.C = C;
    if (geometry) {
        DwSP q = p;
        const int tw = (mode == 2 && k == 5) ? 2 : ((k == 3 && mode != 2 && g_emul_tw8) ? 8 : 4);
        if (!dws_geometry(q, tw, total_warps)) return MNB_ERR_UNSUPPORTED;
        geometry[0] = q.PL; geometry[1] = q.G; geometry[2] = q.NB; geometry[3] = q.HS; geometry[4] = q.nws; geometry[5] = q.nhs;
    }
    if (k == 3) {
        if (mode == 0 && g_emul_tw8) return emul_run<3, 8, DWS_FWD>(p, total_warps);
        if (mode == 1 && g_emul_tw8) return emul_run<3, 8, DWS_DGRAD>(p, total_warps);
        if (mode == 0) return emul_run<3, 4, DWS_FWD>(p, total_warps);
        if (mode == 1) return emul_run<3, 4, DWS_DGRAD>(p, total_warps);
        return emul_run<3, 4, DWS_WGRAD>(p, total_warps);
    }
    if (mode == 0) return emul_run<5, 4, DWS_FWD>(p, total_warps);
    if (mode == 1) return emul_run<5, 4, DWS_DGRAD>(p, total_warps);
    return emul_run<5, 2, DWS_WGRAD>(p, total_warps);
}
#endif
