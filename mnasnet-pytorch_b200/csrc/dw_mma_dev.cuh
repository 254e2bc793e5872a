// Device-side building blocks shared by the TMA + mma.sync depthwise kernels: packed-fp32 helpers, the channel-group /
// strip configuration, diagonal weight fragments, the per-row ldmatrix / MMA steps and the output-row emitter with
// BatchNorm statistics.  Included by dw_mma.cu (row-streaming kernels) and dw_small.cu (whole-tile kernels for small maps).
#pragma once
#include "dw_mma.cuh"

namespace mnb {

// packed fp32 pairs (Blackwell FFMA2 / F2FP): one instruction per channel pair
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
    f2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
    f2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f2_t f2_from_bf16x2(uint32_t u) { return f2_pack(bf_lo(u), bf_hi(u)); }
// relu(v) rounded to bf16x2 (lo -> bits 0..15)
__device__ __forceinline__ uint32_t f2_relu_bf16x2(f2_t v) {
    float lo, hi;
    f2_unpack(v, lo, hi);
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

template <int K, int CG, int TWS>
struct DwmCfg {
    static constexpr int P = K / 2;
    static constexpr int RB = K == 3 ? 12 : 10;     // rows per block: a multiple of K so ring slots are static
    static constexpr int NCH = CG / 8, TW = 16 * TWS, HC = TW + K - 1, PITCH = CG * 2;
    static constexpr int ROWB = HC * PITCH;         // bytes of one input row in shared memory
    static constexpr int OROWB = TW * PITCH;        // bytes of one output row
    static constexpr int XB_BYTES = (RB * ROWB + 127) / 128 * 128;
    static constexpr int OUT_BYTES = (RB * OROWB + 127) / 128 * 128;
    static constexpr int THREADS = 32 * NCH * TWS;
    static constexpr int SMEM = 2 * XB_BYTES + OUT_BYTES + 4 * CG * 4 + 16;
    // register budget: 104 (5x5) / 80 (3x3) per thread -> resident CTAs per SM the compiler must allow
    static constexpr int MINB = 65536 / (THREADS * (K == 5 ? 104 : 80));
};

// Per-lane diagonal B fragment of tap value wv (already the lane's channel g): B[k = c'][n = g] = wv iff c' == g.
// Fragment register = rows k = 2t, 2t+1 of column n = g  ->  non-zero only on the lanes with g>>1 == t.
__device__ __forceinline__ uint32_t dwm_diag(float wv, int g, int t) {
    const uint32_t hb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(wv));
    return (g >> 1) == t ? ((g & 1) ? (hb << 16) : hb) : 0u;
}

// A work step: block j of a column segment (j = jb0 - 1 is the segment's prologue).
struct DwmStep {
    int item, j, jb0, jb1, n, w0;
};
template <int TW>
__device__ __forceinline__ void dwm_step_init(DwmStep& s, int item, int items, int nseg, int tiles_w, int seg, int nblocks) {
    s.item = item;
    if (item < items) {
        const int col = item / nseg, sg = item - col * nseg;
        s.n = col / tiles_w;
        s.w0 = (col - s.n * tiles_w) * TW;
        s.jb0 = sg * seg;
        s.jb1 = min(s.jb0 + seg, nblocks);
        s.j = s.jb0 - 1;
    }
}

// ldmatrix of the 2K 8x8 blocks of one input row (K shifts x two 8-pixel halves)
template <int K, int PITCH>
__device__ __forceinline__ void dwm_load_row(uint32_t (&m)[K][2], uint32_t a4, uint32_t a2) {
    ldsm4(a4, m[0][0], m[0][1], m[1][0], m[1][1]);
    if constexpr (K == 5) ldsm4(a4 + 2 * PITCH, m[2][0], m[2][1], m[3][0], m[3][1]);
    ldsm2(a2, m[K - 1][0], m[K - 1][1]);
}

// taps kw = 0 .. K-2 of every tap row: two taps per m16n8k16 (a = [shift kw | shift kw+1], b = their diagonals)
template <int K, int I>
__device__ __forceinline__ void dwm_mma_pairs(float (&acc)[K][4], const uint32_t (&bd)[K][K], const uint32_t (&m)[K][2]) {
#pragma unroll
    for (int kw = 0; kw + 1 < K; kw += 2)
#pragma unroll
        for (int kh = 0; kh < K; ++kh)
            mma16816(acc[(I + K - 1 - kh) % K], m[kw][0], m[kw][1], m[kw + 1][0], m[kw + 1][1], bd[kh][kw], bd[kh][kw + 1]);
}

// Output row completed by block row I: pack, stage for the TMA store, statistics of the stored (bf16) values.
template <int K, int CG, int TWS>
struct DwmEmitStats {
    uint32_t out_lane;
    bool do_stats;
    f2_t mk0, mk1;          // column validity of this lane's two pixels (1 / 0), used by the MASKED variants
    f2_t st[2];             // (sum, sum of squares) of the lane's channel pair
    template <int I, bool MASKED>
    __device__ __forceinline__ void emit(float (&a)[4]) {
        using Cfg = DwmCfg<K, CG, TWS>;
        const uint32_t u0 = pack_bf16x2(a[0], a[1]), u1 = pack_bf16x2(a[2], a[3]);
        sts32(out_lane + I * Cfg::OROWB, u0);
        sts32(out_lane + I * Cfg::OROWB + 8 * Cfg::PITCH, u1);
        if (do_stats) {
            const f2_t one = f2_pack(1.f, 1.f), zero = f2_pack(0.f, 0.f);
            f2_t q0 = f2_from_bf16x2(u0), q1 = f2_from_bf16x2(u1);
            if (MASKED) { q0 = f2_fma(q0, mk0, zero); q1 = f2_fma(q1, mk1, zero); }
            st[0] = f2_fma(q0, one, st[0]);
            st[0] = f2_fma(q1, one, st[0]);
            st[1] = f2_fma(q0, q0, st[1]);
            st[1] = f2_fma(q1, q1, st[1]);
        }
    }
};

}  // namespace mnb
