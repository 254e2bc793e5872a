// Stem convolution: 3x3, stride 2, pad 1, Cin = 3 (NCHW fp32 network input, src/train.py:427) -> Cout (32)
// NHWC.  Replaces features.0's nn.Conv2d (src/models/mnasnet.py:179) forward and backward-weight (there is no
// backward-data: the input needs no gradient).  K = 27 is too small for the tensor pipe to matter and the
// input layout is planar fp32, so this is a direct CUDA-core kernel: HBM-bound (154 MB fp32 in, 205 MB bf16 out
// at N=256).
//   forward: thread == output pixel, all Cout channels; the 27x32 weights are broadcast from shared memory as
//            float4; lanes run along the output row so the NCHW reads are (stride-2) coalesced and the NHWC
//            stores are 64 contiguous bytes per thread / 2 KB per warp; BN statistics in registers across a
//            persistent pixel loop, one smem + fp64-atomic flush per CTA.
//   wgrad  : thread == (pixel slot, 4-channel group): 27 taps x 4 channels of partial sums in registers over a
//            persistent pixel loop; smem + fp32-atomic flush per CTA.
#include "common.cuh"

namespace mnb {

constexpr int STEM_CIN = 3, STEM_K = 3, STEM_TAPS = 27;

template <typename T, int COUT>
__global__ void __launch_bounds__(128) stem_fwd_k(const float* __restrict__ x, const float* __restrict__ w,
                                                  const float* __restrict__ bias, T* __restrict__ z, double* stats,
                                                  int N, int H, int W, int Ho, int Wo) {
    __shared__ __align__(16) float sw[STEM_TAPS][COUT];      // [tap = ci*9 + kh*3 + kw][co]
    __shared__ float sb[COUT];
    __shared__ float sred[2][COUT];
    for (int i = threadIdx.x; i < STEM_TAPS * COUT; i += blockDim.x) {
        const int co = i % COUT, tap = i / COUT;
        sw[tap][co] = w[co * STEM_TAPS + tap];                // torch layout [co][ci][kh][kw]
    }
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) { sb[i] = bias ? bias[i] : 0.f; sred[0][i] = sred[1][i] = 0.f; }
    __syncthreads();
    float ssum[COUT], ssq[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) ssum[c] = ssq[c] = 0.f;
    const long long total = (long long)N * Ho * Wo;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
         pix += (long long)gridDim.x * blockDim.x) {
        const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
        float in[STEM_TAPS];
#pragma unroll
        for (int ci = 0; ci < STEM_CIN; ++ci)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int ih = ho * 2 - 1 + kh, iw = wo * 2 - 1 + kw;
                    const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
                    in[ci * 9 + kh * 3 + kw] = ok ? x[(((long long)n * STEM_CIN + ci) * H + ih) * W + iw] : 0.f;
                }
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = sb[c];
#pragma unroll
        for (int t = 0; t < STEM_TAPS; ++t) {
#pragma unroll
            for (int c4 = 0; c4 < COUT / 4; ++c4) {
                const float4 wv = *reinterpret_cast<const float4*>(&sw[t][c4 * 4]);
                acc[c4 * 4 + 0] = fmaf(in[t], wv.x, acc[c4 * 4 + 0]);
                acc[c4 * 4 + 1] = fmaf(in[t], wv.y, acc[c4 * 4 + 1]);
                acc[c4 * 4 + 2] = fmaf(in[t], wv.z, acc[c4 * 4 + 2]);
                acc[c4 * 4 + 3] = fmaf(in[t], wv.w, acc[c4 * 4 + 3]);
            }
        }
        T* zp = z + pix * COUT;
#pragma unroll
        for (int c8 = 0; c8 < COUT / 8; ++c8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = acc[c8 * 8 + j];
            store8(zp + c8 * 8, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float q = to_f(from_f<T>(v[j]));        // statistics of the stored (rounded) value
                ssum[c8 * 8 + j] += q;
                ssq[c8 * 8 + j] = fmaf(q, q, ssq[c8 * 8 + j]);
            }
        }
    }
    if (stats) {
        // warp-level reduction first (all lanes hold the same channels), then one smem atomic per warp
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
            float a = ssum[c], b = ssq[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if ((threadIdx.x & 31) == 0) { atomicAdd(&sred[0][c], a); atomicAdd(&sred[1][c], b); }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
            atomicAdd(&stats[i], (double)sred[0][i]);
            atomicAdd(&stats[COUT + i], (double)sred[1][i]);
        }
    }
}

// dw[co][tap] += sum_pix dz[pix][co] * xpatch[pix][tap]
template <typename T, int COUT>
__global__ void __launch_bounds__(256) stem_wgrad_k(const float* __restrict__ x, const T* __restrict__ dz,
                                                    float* dw, int N, int H, int W, int Ho, int Wo) {
    constexpr int CG = COUT / 4;                       // channel groups of 4 (8 for Cout = 32)
    __shared__ float sred[COUT * STEM_TAPS];
    for (int i = threadIdx.x; i < COUT * STEM_TAPS; i += blockDim.x) sred[i] = 0.f;
    __syncthreads();
    const int cg = threadIdx.x % CG, slot = threadIdx.x / CG;
    const int slots = blockDim.x / CG;
    float acc[STEM_TAPS][4];
#pragma unroll
    for (int t = 0; t < STEM_TAPS; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
    const long long total = (long long)N * Ho * Wo;
    for (long long pix = (long long)blockIdx.x * slots + slot; pix < total; pix += (long long)gridDim.x * slots) {
        const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
        const T* dp = dz + pix * COUT + cg * 4;
        const float g0 = to_f(dp[0]), g1 = to_f(dp[1]), g2 = to_f(dp[2]), g3 = to_f(dp[3]);
#pragma unroll
        for (int ci = 0; ci < STEM_CIN; ++ci)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int ih = ho * 2 - 1 + kh, iw = wo * 2 - 1 + kw;
                    const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
                    const float v = ok ? x[(((long long)n * STEM_CIN + ci) * H + ih) * W + iw] : 0.f;
                    const int t = ci * 9 + kh * 3 + kw;
                    acc[t][0] = fmaf(v, g0, acc[t][0]);
                    acc[t][1] = fmaf(v, g1, acc[t][1]);
                    acc[t][2] = fmaf(v, g2, acc[t][2]);
                    acc[t][3] = fmaf(v, g3, acc[t][3]);
                }
    }
#pragma unroll
    for (int t = 0; t < STEM_TAPS; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&sred[(cg * 4 + j) * STEM_TAPS + t], acc[t][j]);
    __syncthreads();
    for (int i = threadIdx.x; i < COUT * STEM_TAPS; i += blockDim.x) atomicAdd(&dw[i], sred[i]);
}

bool stem_supported(int Cin, int Cout, int k, int stride, int pad, int nchw_in) {
    return nchw_in && Cin == 3 && Cout == 32 && k == 3 && stride == 2 && pad == 1;
}

int stem_fwd(const float* x, const float* w, const float* bias, void* z, double* stats, int N, int H, int W, int dtype,
             cudaStream_t st) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * Ho * Wo;
    long long blocks = cdiv(total, 128);
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (dtype == MNB_F32) stem_fwd_k<float, 32><<<(unsigned)blocks, 128, 0, st>>>(x, w, bias, (float*)z, stats, N, H, W, Ho, Wo);
    else stem_fwd_k<bf16, 32><<<(unsigned)blocks, 128, 0, st>>>(x, w, bias, (bf16*)z, stats, N, H, W, Ho, Wo);
    MNB_LAUNCH_CHECK("stem_fwd");
    return 0;
}

int stem_wgrad(const float* x, const void* dz, float* dw, int N, int H, int W, int dtype, cudaStream_t st) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * Ho * Wo;
    long long blocks = cdiv(total, 32);
    const long long cap = (long long)num_sms() * 4;
    if (blocks > cap) blocks = cap;
    if (dtype == MNB_F32) stem_wgrad_k<float, 32><<<(unsigned)blocks, 256, 0, st>>>(x, (const float*)dz, dw, N, H, W, Ho, Wo);
    else stem_wgrad_k<bf16, 32><<<(unsigned)blocks, 256, 0, st>>>(x, (const bf16*)dz, dw, N, H, W, Ho, Wo);
    MNB_LAUNCH_CHECK("stem_wgrad");
    return 0;
}

}  // namespace mnb
