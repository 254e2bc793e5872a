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
#include <stdlib.h>

#include "common.cuh"

namespace mnb {

constexpr int STEM_CIN = 3, STEM_K = 3, STEM_TAPS = 27;

// Network input fetch.  U8 = false: N x 3 x H x W fp32 (src/train.py:427 hands over the normalised tensor).
// U8 = true: N x H x W x 3 uint8 straight from the decoder -- ToTensor + Normalize(mean, std) of the reference's
// pipeline (src/utils/datasets.py:456-462, src/models/classifiers.py:91-92) happen HERE, through a 3 x 256 table built
// per CTA with exactly torch's fp32 operations ((u / 255 - mean) / std), so the result is bit-identical to normalising
// first and the host-to-device copy carries one byte per value (38.5 MB instead of 154 MB at N = 256).
template <bool U8>
__device__ __forceinline__ float stem_in(const void* x, const float* lut, long long n, int ci, int ih, int iw, int H, int W) {
    if (U8) return lut[ci * 256 + static_cast<const unsigned char*>(x)[((n * H + ih) * W + iw) * STEM_CIN + ci]];
    return static_cast<const float*>(x)[((n * STEM_CIN + ci) * H + ih) * W + iw];
}
__device__ __forceinline__ void stem_build_lut(float* lut, const float* mean, const float* stdv) {
    for (int i = threadIdx.x; i < STEM_CIN * 256; i += blockDim.x) {
        const int c = i >> 8;
        const float v = (float)(i & 255) / 255.f;
        lut[i] = (v - mean[c]) / stdv[c];
    }
}

template <typename T, int COUT, bool U8>
__global__ void __launch_bounds__(128) stem_fwd_k(const void* __restrict__ x, const float* __restrict__ mean,
                                                  const float* __restrict__ stdv, const float* __restrict__ w,
                                                  const float* __restrict__ bias, T* __restrict__ z, double* stats,
                                                  int N, int H, int W, int Ho, int Wo) {
    __shared__ __align__(16) float sw[STEM_TAPS][COUT];      // [tap = ci*9 + kh*3 + kw][co]
    __shared__ float sb[COUT];
    __shared__ float sred[2][COUT];
    __shared__ float lut[U8 ? STEM_CIN * 256 : 1];
    if (U8) stem_build_lut(lut, mean, stdv);
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
                    in[ci * 9 + kh * 3 + kw] = ok ? stem_in<U8>(x, lut, n, ci, ih, iw, H, W) : 0.f;
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
template <typename T, int COUT, bool U8>
__global__ void __launch_bounds__(256) stem_wgrad_k(const void* __restrict__ x, const float* __restrict__ mean,
                                                    const float* __restrict__ stdv, const T* __restrict__ dz,
                                                    float* dw, int N, int H, int W, int Ho, int Wo) {
    constexpr int CG = COUT / 4;                       // channel groups of 4 (8 for Cout = 32)
    __shared__ float sred[COUT * STEM_TAPS];
    __shared__ float lut[U8 ? STEM_CIN * 256 : 1];
    if (U8) stem_build_lut(lut, mean, stdv);
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
                    const float v = ok ? stem_in<U8>(x, lut, n, ci, ih, iw, H, W) : 0.f;
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

// ---- backward-weight on the tensor pipe (bf16 activations) ---------------------------------------------------------
// dw[co][tap] = sum_pix dz[pix][co] * patch[pix][tap] as an implicit GEMM with M = 32 output channels, N = 27 (-> 32)
// taps and the output pixels as the reduction index.  A warp takes 16 consecutive output pixels per step: it stages
// their dz rows and their bf16-rounded input patches in a private shared buffer (the next step's global loads are
// already in flight), reads both back transposed with ldmatrix.trans and issues 2 x 4 mma.sync.m16n8k16.  No
// cross-warp synchronisation until the final reduction; replaces 108 fp32 accumulators per lane (254 registers, 8
// warps per SM, latency-bound at 0.3 TB/s) of stem_wgrad_k.
__device__ __forceinline__ void stem_mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void stem_ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}

template <bool U8>
__global__ void __launch_bounds__(128, 4) stem_wgrad_mma_k(const void* __restrict__ x, const float* __restrict__ mean,
                                                           const float* __restrict__ stdv, const bf16* __restrict__ dz,
                                                           float* dw, int N, int H, int W, int Ho, int Wo) {
    constexpr int WARPS = 4, PITCH = 80;                 // 64 bytes of payload per staged row, pitch = 5 x 16 bytes
    constexpr int SEGS = 5;                              // (ci,kh) segments of 3 taps per lane: lanes 0-15 take 0-4, 16-31 take 5-8
    __shared__ __align__(16) unsigned char s_z[WARPS][16 * PITCH];
    __shared__ __align__(16) unsigned char s_p[WARPS][16 * PITCH];
    __shared__ float s_dw[32 * 32];
    __shared__ float lut[U8 ? STEM_CIN * 256 : 1];
    if (U8) stem_build_lut(lut, mean, stdv);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) s_dw[i] = 0.f;
    for (int i = threadIdx.x; i < WARPS * 16 * PITCH / 4; i += blockDim.x) {
        reinterpret_cast<uint32_t*>(&s_z[0][0])[i] = 0u;
        reinterpret_cast<uint32_t*>(&s_p[0][0])[i] = 0u;  // taps 27..31 stay zero for the whole kernel
    }
    __syncthreads();

    const int HWo = Ho * Wo;
    const int total = N * HWo;                           // host guarantees < 2^31
    const int nsteps = (total + 15) / 16;
    const int wstride = gridDim.x * WARPS;
    int step = blockIdx.x * WARPS + warp;
    const int pix = lane & 15, half = lane >> 4;

    float xr[SEGS][3];
    uint4 zr[2];
    auto load_step = [&](int st) {
        const int p0 = st * 16;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = lane + 32 * i, r = c >> 2, cc = c & 3;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (p0 + r < total) v = __ldg(reinterpret_cast<const uint4*>(dz + (long long)(p0 + r) * 32 + cc * 8));
            zr[i] = v;
        }
        const int p = p0 + pix;
        const bool pok = p < total;
        const int n = p / HWo, rem = p - n * HWo, ho = rem / Wo, wo = rem - ho * Wo;
        const int iw0 = 2 * wo - 1;
#pragma unroll
        for (int i = 0; i < SEGS; ++i) {
            const int seg = half * SEGS + i;             // (ci,kh) = (seg / 3, seg % 3); seg 9 does not exist
            const int ci = seg / 3, kh = seg - 3 * ci;
            const int ih = 2 * ho - 1 + kh;
            const bool rok = pok && seg < 9 && ih >= 0 && ih < H;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int iw = iw0 + kw;
                xr[i][kw] = (rok && iw >= 0 && iw < W) ? stem_in<U8>(x, lut, n, ci, ih, iw, H, W) : 0.f;
            }
        }
    };
    if (step < nsteps) load_step(step);

    unsigned char* wz = s_z[warp];
    unsigned char* wp = s_p[warp];
    const int lm = lane >> 3, lr = lane & 7;
    const unsigned char* a_addr = wz + (lr + ((lm & 2) ? 8 : 0)) * PITCH + ((lm & 1) ? 16 : 0);   // + 32*m
    const unsigned char* b_addr = wp + (lr + ((lm & 1) ? 8 : 0)) * PITCH + ((lm & 2) ? 16 : 0);   // + 32*(n/2)

    float acc[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;

    for (; step < nsteps; step += wstride) {
        __syncwarp();                                    // the previous step's ldmatrix reads are done
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = lane + 32 * i, r = c >> 2, cc = c & 3;
            *reinterpret_cast<uint4*>(wz + r * PITCH + cc * 16) = zr[i];
        }
#pragma unroll
        for (int i = 0; i < SEGS; ++i) {
            const int seg = half * SEGS + i;
            if (seg < 9) {
                bf16* pp = reinterpret_cast<bf16*>(wp + pix * PITCH) + 3 * seg;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) pp[kw] = __float2bfloat16_rn(xr[i][kw]);
            }
        }
        __syncwarp();
        if (step + wstride < nsteps) load_step(step + wstride);
        uint32_t b[4][2];
#pragma unroll
        for (int n = 0; n < 4; n += 2) {
            uint32_t r4[4];
            stem_ldmatrix_x4_trans(r4, b_addr + 16 * n);
            b[n][0] = r4[0]; b[n][1] = r4[1]; b[n + 1][0] = r4[2]; b[n + 1][1] = r4[3];
        }
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            uint32_t a[4];
            stem_ldmatrix_x4_trans(a, a_addr + 32 * m);
#pragma unroll
            for (int n = 0; n < 4; ++n) stem_mma16816(acc[m][n], a, b[n][0], b[n][1]);
        }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int co = 16 * m + g, tap = 8 * n + 2 * t;
            atomicAdd(&s_dw[co * 32 + tap], acc[m][n][0]);
            atomicAdd(&s_dw[co * 32 + tap + 1], acc[m][n][1]);
            atomicAdd(&s_dw[(co + 8) * 32 + tap], acc[m][n][2]);
            atomicAdd(&s_dw[(co + 8) * 32 + tap + 1], acc[m][n][3]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * STEM_TAPS; i += blockDim.x) {
        const int co = i / STEM_TAPS, tap = i - co * STEM_TAPS;
        atomicAdd(&dw[i], s_dw[co * 32 + tap]);
    }
}

int stem_fwd_mma(const void* x, int x_u8, const float* mean, const float* stdv, const float* w, const float* bias, void* z,
                 double* stats, int N, int H, int W, cudaStream_t st);

bool stem_supported(int Cin, int Cout, int k, int stride, int pad, int nchw_in) {
    return nchw_in && Cin == 3 && Cout == 32 && k == 3 && stride == 2 && pad == 1;
}

// x_u8: the input is N x H x W x 3 uint8 and mean / stdv (3 floats each) are the Normalize constants
int stem_fwd(const void* x, int x_u8, const float* mean, const float* stdv, const float* w, const float* bias, void* z,
             double* stats, int N, int H, int W, int dtype, cudaStream_t st) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * Ho * Wo;
    long long blocks = cdiv(total, 128);
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    // bf16: tensor-pipe kernel (stem_mma.cu) unless the "stem_mma" option / MNB_STEM_MMA is 0 or the rows cannot be bulk-copied
    if (dtype == MNB_BF16 && option_get(OPT_STEM_MMA)) {
        int r = stem_fwd_mma(x, x_u8, mean, stdv, w, bias, z, stats, N, H, W, st);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    const unsigned b = (unsigned)blocks;
    if (dtype == MNB_F32) {
        if (x_u8) stem_fwd_k<float, 32, true><<<b, 128, 0, st>>>(x, mean, stdv, w, bias, (float*)z, stats, N, H, W, Ho, Wo);
        else stem_fwd_k<float, 32, false><<<b, 128, 0, st>>>(x, mean, stdv, w, bias, (float*)z, stats, N, H, W, Ho, Wo);
    } else {
        if (x_u8) stem_fwd_k<bf16, 32, true><<<b, 128, 0, st>>>(x, mean, stdv, w, bias, (bf16*)z, stats, N, H, W, Ho, Wo);
        else stem_fwd_k<bf16, 32, false><<<b, 128, 0, st>>>(x, mean, stdv, w, bias, (bf16*)z, stats, N, H, W, Ho, Wo);
    }
    MNB_LAUNCH_CHECK("stem_fwd");
    return 0;
}

int stem_wgrad(const void* x, int x_u8, const float* mean, const float* stdv, const void* dz, float* dw, int N, int H, int W,
               int dtype, int impl, cudaStream_t st) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * Ho * Wo;
    long long blocks = cdiv(total, 32);
    const long long cap = (long long)num_sms() * 4;
    if (blocks > cap) blocks = cap;
    // bf16: tensor-pipe kernel (11x faster, profiles/r1_exp_stream.json) unless impl 2 asks for the fp32-input SIMT
    // kernel or the "stem_mma" option / MNB_STEM_MMA is 0
    const int use_mma = option_get(OPT_STEM_MMA);
    if (dtype == MNB_BF16 && (impl == 3 || (impl == 0 && use_mma)) && total < (1ll << 31) - 64) {
        if (x_u8) stem_wgrad_mma_k<true><<<num_sms() * 4, 128, 0, st>>>(x, mean, stdv, (const bf16*)dz, dw, N, H, W, Ho, Wo);
        else stem_wgrad_mma_k<false><<<num_sms() * 4, 128, 0, st>>>(x, mean, stdv, (const bf16*)dz, dw, N, H, W, Ho, Wo);
        MNB_LAUNCH_CHECK("stem_wgrad(mma)");
        return 0;
    }
    const unsigned b = (unsigned)blocks;
    if (dtype == MNB_F32) {
        if (x_u8) stem_wgrad_k<float, 32, true><<<b, 256, 0, st>>>(x, mean, stdv, (const float*)dz, dw, N, H, W, Ho, Wo);
        else stem_wgrad_k<float, 32, false><<<b, 256, 0, st>>>(x, mean, stdv, (const float*)dz, dw, N, H, W, Ho, Wo);
    } else {
        if (x_u8) stem_wgrad_k<bf16, 32, true><<<b, 256, 0, st>>>(x, mean, stdv, (const bf16*)dz, dw, N, H, W, Ho, Wo);
        else stem_wgrad_k<bf16, 32, false><<<b, 256, 0, st>>>(x, mean, stdv, (const bf16*)dz, dw, N, H, W, Ho, Wo);
    }
    MNB_LAUNCH_CHECK("stem_wgrad");
    return 0;
}

}  // namespace mnb
