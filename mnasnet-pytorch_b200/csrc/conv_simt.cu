// Dense convolution (1x1 and 3x3, stride 1/2) as an implicit GEMM on CUDA cores, fp32 accumulate.
// This is the fp32-mode path (1e-4 parity needs full fp32 products, which the bf16/tf32 tensor pipe cannot
// give) and the checker for the tcgen05 path in gemm_tc.cu; in bf16 mode the tcgen05 kernels are used.
// Replaces nn.Conv2d(groups=1) forward / backward-data / backward-weight (src/models/mnasnet.py:48-54,
// called from :82-85,92-95,116-119,126-129,157-161,179).
//
// One kernel template, three modes.  Tile 64 x 64 x 16, 256 threads, 4x4 outputs per thread.  Every thread
// stages exactly one 8-element chunk of A or B per K-step (threads 0..127 -> A, 128..255 -> B):
//   FWD   C[m=(n,ho,wo)][co]      = sum_kk a(m,kk) w(co,kk)        kk=(kh,kw,ci)  A chunk: 8 ci  (vector load)
//   DGRAD C[m=(n,h,w)][ci]        = sum_kk dz(m,kk) w(ci,kk)       kk=(kh,kw,co)  A chunk: 8 co  (vector load)
//   WGRAD C[co][kk=(kh,kw,ci)]    = sum_pos dz(pos,co) a(pos,kk)   K = positions  chunks run along M / N
#include "common.cuh"
#include <stdlib.h>

#include "conv_params.cuh"

namespace mnb {

// a(n, ih, iw, ci..ci+7) with transform; zero outside the image (zero padding is applied AFTER BN+ReLU)
template <typename T>
__device__ __forceinline__ void load_act8(const ConvP& p, int n, int ih, int iw, int ci, int cvalid, float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) return;
    if (p.nchw_in) {
        const float* xp = (const float*)p.x;
        for (int i = 0; i < cvalid; ++i) v[i] = xp[(((long long)n * p.Cin + ci + i) * p.H + ih) * p.W + iw];
    } else {
        const T* xp = (const T*)p.x + (((long long)n * p.H + ih) * p.W + iw) * p.Cin + ci;
        if (cvalid == 8) load8(xp, v);
        else for (int i = 0; i < cvalid; ++i) v[i] = to_f(xp[i]);
    }
    if (p.in_scale) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < cvalid) v[i] = fmaxf(fmaf(p.in_scale[ci + i], v[i], p.in_shift[ci + i]), 0.f);
    }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256) conv_simt_k(ConvP p) {
    __shared__ __align__(16) float As[16][64 + 4];
    __shared__ __align__(16) float Bs[16][64 + 4];
    __shared__ float sred[2][64];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int kk2 = p.k * p.k;
    long long Mtot;
    int Ntot;
    long long Ktot;
    if (MODE == MODE_FWD) { Mtot = (long long)p.N * p.Ho * p.Wo; Ntot = p.Cout; Ktot = (long long)kk2 * p.Cin; }
    else if (MODE == MODE_DGRAD) { Mtot = (long long)p.N * p.H * p.W; Ntot = p.Cin; Ktot = (long long)kk2 * p.Cout; }
    else { Mtot = p.Cout; Ntot = kk2 * p.Cin; Ktot = (long long)p.N * p.Ho * p.Wo; }
    const long long m0 = (long long)blockIdx.y * 64;
    const int n0 = blockIdx.x * 64;
    long long kbeg = 0, kend = Ktot;
    if (MODE == MODE_WGRAD) {
        kbeg = (long long)blockIdx.z * p.kchunk;
        kend = kbeg + p.kchunk < Ktot ? kbeg + p.kchunk : Ktot;
    }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // ---- per-thread staging role ----
    const bool isA = tid < 128;
    const int lt = tid & 127;
    // FWD/DGRAD: chunk = (row lt/2, k-offset (lt%2)*8) ; WGRAD: chunk = (8 rows/cols (lt%8)*8.., kpos lt/8)
    int rowA = lt / 2, koff = (lt % 2) * 8;
    int an = 0, ah = 0, aw = 0;   // decoded position of this thread's A row (FWD: output pos; DGRAD: input pos)
    bool arow_ok = false;
    if (MODE != MODE_WGRAD && isA) {
        long long m = m0 + rowA;
        arow_ok = m < Mtot;
        if (arow_ok) {
            int Wd = MODE == MODE_FWD ? p.Wo : p.W, Hd = MODE == MODE_FWD ? p.Ho : p.H;
            aw = (int)(m % Wd);
            ah = (int)((m / Wd) % Hd);
            an = (int)(m / ((long long)Wd * Hd));
        }
    }

    for (long long k0 = kbeg; k0 < kend; k0 += 16) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (MODE == MODE_FWD) {
            if (isA) {
                long long kk = k0 + koff;
                if (arow_ok && kk < Ktot) {
                    if (p.Cin % 8 == 0) {
                        int tap = (int)(kk / p.Cin), ci = (int)(kk % p.Cin);
                        int kh = tap / p.k, kw = tap % p.k;
                        load_act8<T>(p, an, ah * p.stride - p.pad + kh, aw * p.stride - p.pad + kw, ci, 8, v);
                    } else {
                        for (int i = 0; i < 8 && kk + i < Ktot; ++i) {
                            int tap = (int)((kk + i) / p.Cin), ci = (int)((kk + i) % p.Cin);
                            int kh = tap / p.k, kw = tap % p.k;
                            float one[8];
                            load_act8<T>(p, an, ah * p.stride - p.pad + kh, aw * p.stride - p.pad + kw, ci, 1, one);
                            v[i] = one[0];
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) As[koff + i][rowA] = v[i];
            } else {
                int co = n0 + rowA;
                long long kk = k0 + koff;
                if (co < Ntot) {
                    for (int i = 0; i < 8 && kk + i < Ktot; ++i) {
                        int tap = (int)((kk + i) / p.Cin), ci = (int)((kk + i) % p.Cin);
                        v[i] = p.w[((long long)co * p.Cin + ci) * kk2 + tap];
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) Bs[koff + i][rowA] = v[i];
            }
        } else if (MODE == MODE_DGRAD) {
            if (isA) {
                long long kk = k0 + koff;
                if (arow_ok && kk < Ktot) {      // Cout % 8 == 0 always
                    int tap = (int)(kk / p.Cout), co = (int)(kk % p.Cout);
                    int kh = tap / p.k, kw = tap % p.k;
                    int hn = ah + p.pad - kh, wn = aw + p.pad - kw;
                    if (hn >= 0 && wn >= 0 && hn % p.stride == 0 && wn % p.stride == 0) {
                        int ho = hn / p.stride, wo = wn / p.stride;
                        if (ho < p.Ho && wo < p.Wo)
                            load8((const T*)p.dz + (((long long)an * p.Ho + ho) * p.Wo + wo) * p.Cout + co, v);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) As[koff + i][rowA] = v[i];
            } else {
                int ci = n0 + rowA;
                long long kk = k0 + koff;
                if (ci < Ntot) {
                    for (int i = 0; i < 8 && kk + i < Ktot; ++i) {
                        int tap = (int)((kk + i) / p.Cout), co = (int)((kk + i) % p.Cout);
                        v[i] = p.w[((long long)co * p.Cin + ci) * kk2 + tap];
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) Bs[koff + i][rowA] = v[i];
            }
        } else {  // WGRAD
            const int kp = lt / 8, r8 = (lt % 8) * 8;
            long long pos = k0 + kp;
            if (pos < kend) {
                int wo = (int)(pos % p.Wo), ho = (int)((pos / p.Wo) % p.Ho), n = (int)(pos / ((long long)p.Wo * p.Ho));
                if (isA) {
                    long long co = m0 + r8;
                    if (co < Mtot) load8((const T*)p.dz + pos * p.Cout + co, v);
                } else {
                    int kk = n0 + r8;
                    if (kk < Ntot) {
                        if (p.Cin % 8 == 0) {
                            int tap = kk / p.Cin, ci = kk % p.Cin;
                            int kh = tap / p.k, kw = tap % p.k;
                            load_act8<T>(p, n, ho * p.stride - p.pad + kh, wo * p.stride - p.pad + kw, ci, 8, v);
                        } else {
                            for (int i = 0; i < 8 && kk + i < Ntot; ++i) {
                                int tap = (kk + i) / p.Cin, ci = (kk + i) % p.Cin;
                                int kh = tap / p.k, kw = tap % p.k;
                                float one[8];
                                load_act8<T>(p, n, ho * p.stride - p.pad + kh, wo * p.stride - p.pad + kw, ci, 1, one);
                                v[i] = one[0];
                            }
                        }
                    }
                }
            }
            float* dst = isA ? &As[kp][r8] : &Bs[kp][r8];
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue ----
    if (MODE == MODE_FWD) {
        T* z = (T*)p.out;
        float cs[4] = {0, 0, 0, 0}, cq[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            long long m = m0 + ty * 4 + i;
            if (m >= Mtot) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int co = n0 + tx * 4 + j;
                if (co >= Ntot) continue;
                float v = acc[i][j] + (p.bias ? p.bias[co] : 0.f);
                T q = from_f<T>(v);
                z[m * p.Cout + co] = q;
                float qf = to_f(q);
                cs[j] += qf;
                cq[j] = fmaf(qf, qf, cq[j]);
            }
        }
        if (p.stats) {
            if (tid < 128) (&sred[0][0])[tid] = 0.f;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(&sred[0][tx * 4 + j], cs[j]);
                atomicAdd(&sred[1][tx * 4 + j], cq[j]);
            }
            __syncthreads();
            if (tid < 64 && n0 + tid < Ntot) {
                atomicAdd(&p.stats[n0 + tid], (double)sred[0][tid]);
                atomicAdd(&p.stats[p.Cout + n0 + tid], (double)sred[1][tid]);
            }
        }
    } else if (MODE == MODE_DGRAD) {
        T* dx = (T*)p.out;
        const T* add = (const T*)p.add;
        const T* bz = (const T*)p.bn_z;
        float cs[4] = {0, 0, 0, 0}, cq[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            long long m = m0 + ty * 4 + i;
            if (m >= Mtot) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int ci = n0 + tx * 4 + j;
                if (ci >= Ntot) continue;
                float v = acc[i][j];
                if (add) v += to_f(add[m * p.Cin + ci]);
                T q = from_f<T>(v);
                dx[m * p.Cin + ci] = q;
                if (bz) {       // fused BN-backward reduction of the producer block
                    const float zz = to_f(bz[m * p.Cin + ci]);
                    const float g = fmaf(p.bn_scale[ci], zz, p.bn_shift[ci]) > 0.f ? to_f(q) : 0.f;
                    cs[j] += g;
                    cq[j] = fmaf(g, zz, cq[j]);
                }
            }
        }
        if (bz && p.stats) {
            if (tid < 128) (&sred[0][0])[tid] = 0.f;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(&sred[0][tx * 4 + j], cs[j]);
                atomicAdd(&sred[1][tx * 4 + j], cq[j]);
            }
            __syncthreads();
            if (tid < 64 && n0 + tid < Ntot) {
                atomicAdd(&p.stats[n0 + tid], (double)sred[0][tid]);
                atomicAdd(&p.stats[p.Cin + n0 + tid], (double)sred[1][tid]);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            long long co = m0 + ty * 4 + i;
            if (co >= Mtot) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int kk = n0 + tx * 4 + j;
                if (kk >= Ntot) continue;
                int tap = kk / p.Cin, ci = kk % p.Cin;
                atomicAdd(&p.dw[((long long)co * p.Cin + ci) * kk2 + tap], acc[i][j]);
            }
        }
    }
}

// dedicated stem kernels (stem.cu)
bool stem_supported(int Cin, int Cout, int k, int stride, int pad, int nchw_in);
int stem_fwd(const void* x, int x_u8, const float* mean, const float* stdv, const float* w, const float* bias, void* z,
             double* stats, int N, int H, int W, int dtype,
             cudaStream_t st);
int stem_wgrad(const void* x, int x_u8, const float* mean, const float* stdv, const void* dz, float* dw, int N, int H, int W,
               int dtype, int impl, cudaStream_t st);

__global__ void pack_weights_k(const float* __restrict__ w, bf16* __restrict__ pf, bf16* __restrict__ pd, int Cout,
                               int Cin, int kk2) {
    const int total = Cout * Cin * kk2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tap = i % kk2, ci = (i / kk2) % Cin, co = i / (kk2 * Cin);      // torch layout [co][ci][tap]
        const bf16 v = __float2bfloat16_rn(w[i]);
        if (pf) pf[(long long)co * kk2 * Cin + tap * Cin + ci] = v;
        if (pd) pd[(long long)ci * kk2 * Cout + tap * Cout + co] = v;
    }
}

static int check_conv(const char* name, int N, int H, int W, int Cin, int Cout, int k, int stride, int pad, int dtype,
                      int x_layout) {
    MNB_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0, "%s: bad N/H/W/Cin", name);
    MNB_REQUIRE(Cout > 0 && Cout % 8 == 0, "%s: Cout=%d must be a positive multiple of 8", name, Cout);
    MNB_REQUIRE(x_layout == MNB_LAYOUT_NHWC || x_layout == MNB_LAYOUT_NCHW_F32 || x_layout == MNB_LAYOUT_NHWC_U8,
                "%s: bad x_layout %d", name, x_layout);
    MNB_REQUIRE(x_layout != MNB_LAYOUT_NHWC || Cin % 8 == 0, "%s: Cin=%d must be a multiple of 8 for NHWC input",
                name, Cin);
    MNB_REQUIRE(x_layout != MNB_LAYOUT_NHWC_U8 || (Cin == 3 && Cout == 32 && k == 3 && stride == 2 && pad == 1),
                "%s: uint8 input is lowered for the stem only (3 -> 32, k3 s2 p1)", name);
    MNB_REQUIRE((k == 1 && pad == 0) || (k == 3 && pad == 1), "%s: unsupported k=%d pad=%d", name, k, pad);
    MNB_REQUIRE(stride == 1 || stride == 2, "%s: unsupported stride %d", name, stride);
    MNB_REQUIRE(dtype == MNB_F32 || dtype == MNB_BF16, "%s: bad dtype %d", name, dtype);
    return 0;
}

int conv_fwd_simt(const ConvP& p, int dtype, cudaStream_t st) {
    long long M = (long long)p.N * p.Ho * p.Wo;
    dim3 grid((unsigned)cdiv(p.Cout, 64), (unsigned)cdiv(M, 64));
    if (dtype == MNB_F32) conv_simt_k<float, MODE_FWD><<<grid, 256, 0, st>>>(p);
    else conv_simt_k<bf16, MODE_FWD><<<grid, 256, 0, st>>>(p);
    MNB_LAUNCH_CHECK("conv_fwd(simt)");
    return 0;
}
int conv_dgrad_simt(const ConvP& p, int dtype, cudaStream_t st) {
    long long M = (long long)p.N * p.H * p.W;
    dim3 grid((unsigned)cdiv(p.Cin, 64), (unsigned)cdiv(M, 64));
    if (dtype == MNB_F32) conv_simt_k<float, MODE_DGRAD><<<grid, 256, 0, st>>>(p);
    else conv_simt_k<bf16, MODE_DGRAD><<<grid, 256, 0, st>>>(p);
    MNB_LAUNCH_CHECK("conv_dgrad(simt)");
    return 0;
}
int conv_wgrad_simt(ConvP p, int dtype, cudaStream_t st) {
    long long K = (long long)p.N * p.Ho * p.Wo;
    int gx = (int)cdiv((long long)p.k * p.k * p.Cin, 64), gy = (int)cdiv(p.Cout, 64);
    long long want = (long long)num_sms() * 4 / ((long long)gx * gy);
    if (want < 1) want = 1;
    long long chunk = cdiv(cdiv(K, want), 16) * 16;
    if (chunk < 16) chunk = 16;
    p.kchunk = chunk;
    dim3 grid(gx, gy, (unsigned)cdiv(K, chunk));
    if (dtype == MNB_F32) conv_simt_k<float, MODE_WGRAD><<<grid, 256, 0, st>>>(p);
    else conv_simt_k<bf16, MODE_WGRAD><<<grid, 256, 0, st>>>(p);
    MNB_LAUNCH_CHECK("conv_wgrad(simt)");
    return 0;
}

}  // namespace mnb

using namespace mnb;

// The warp-streaming kernels of pw_stream.cu take the 1x1 bf16 layers they are instantiated for (measured 2-3x faster
// than the tcgen05 pipeline on the 112x112 / 56x56 stages, profiles/r1_exp_stream.json) under impl 0 (auto) and impl 3;
// impl 2 forces tcgen05, impl 1 SIMT.  mnb_set_option("pw_stream", 0) / MNB_PW_STREAM=0 switch them off for auto.
static bool prefer_stream(int impl) { return impl == 3 || (impl == 0 && option_get(OPT_PW_STREAM)); }

extern "C" {

int mnb_pack_weights(const float* w, void* wpk_fwd, void* wpk_dgrad, int Cout, int Cin, int k, void* stream) {
    MNB_REQUIRE(Cout > 0 && Cin > 0 && (k == 1 || k == 3), "pack_weights: bad shape");
    const int total = Cout * Cin * k * k;
    int blocks = (total + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    pack_weights_k<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (bf16*)wpk_fwd, (bf16*)wpk_dgrad, Cout, Cin, k * k);
    MNB_LAUNCH_CHECK("pack_weights");
    return 0;
}

int mnb_conv_fwd(const void* x, const float* in_scale, const float* in_shift, const float* w, const float* bias,
                 void* z, double* stats, int N, int H, int W, int Cin, int Cout, int k, int stride, int pad, int dtype,
                 int x_layout, int impl, void* stream) {
    return mnb_conv_fwd_packed(x, in_scale, in_shift, w, nullptr, bias, z, stats, N, H, W, Cin, Cout, k, stride, pad,
                               dtype, x_layout, impl, stream);
}

int mnb_conv_dgrad(const void* dz, const float* w, const void* add, void* dx, const void* bn_z, const float* bn_scale,
                   const float* bn_shift, double* bn_sums, int N, int H, int W, int Cin, int Cout, int k, int stride,
                   int pad, int dtype, int impl, void* stream) {
    return mnb_conv_dgrad_packed(dz, w, nullptr, add, dx, bn_z, bn_scale, bn_shift, bn_sums, N, H, W, Cin, Cout, k,
                                 stride, pad, dtype, impl, stream);
}

int mnb_conv_fwd_packed(const void* x, const float* in_scale, const float* in_shift, const float* w, const void* wpk,
                        const float* bias, void* z, double* stats, int N, int H, int W, int Cin, int Cout, int k,
                        int stride, int pad, int dtype, int x_layout, int impl, void* stream) {
    if (int e = check_conv("conv_fwd", N, H, W, Cin, Cout, k, stride, pad, dtype, x_layout)) return e;
    ConvP p = {};
    p.wpk = wpk;
    p.x = x; p.in_scale = in_scale; p.in_shift = in_shift; p.w = w; p.bias = bias; p.out = z; p.stats = stats;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.k = k; p.stride = stride; p.pad = pad;
    p.Ho = (H + 2 * pad - k) / stride + 1; p.Wo = (W + 2 * pad - k) / stride + 1;
    p.nchw_in = x_layout == MNB_LAYOUT_NCHW_F32;
    if (x_layout == MNB_LAYOUT_NHWC_U8) {       // in_scale / in_shift carry Normalize's mean / std (3 floats each)
        MNB_REQUIRE(in_scale && in_shift, "conv_fwd: uint8 input needs mean (in_scale) and std (in_shift)");
        return stem_fwd(x, 1, in_scale, in_shift, w, bias, z, stats, N, H, W, dtype, (cudaStream_t)stream);
    }
    if (impl != 1 && in_scale == nullptr && stem_supported(Cin, Cout, k, stride, pad, p.nchw_in))
        return stem_fwd(x, 0, nullptr, nullptr, w, bias, z, stats, N, H, W, dtype, (cudaStream_t)stream);
    if (dtype == MNB_BF16 && prefer_stream(impl)) {
        int r = conv_fwd_stream(p, (cudaStream_t)stream);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_BF16 && (impl == 0 || impl == 3)) {
        int r = conv_fwd_c3(p, (cudaStream_t)stream);
        if (r != MNB_ERR_UNSUPPORTED) return r;
        r = conv_fwd_pwide(p, (cudaStream_t)stream);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (impl != 1 && dtype == MNB_BF16) {
        int r = conv_fwd_tc(p, (cudaStream_t)stream);
        if (r != MNB_ERR_UNSUPPORTED) return r;
        MNB_REQUIRE(impl != 2, "conv_fwd: tcgen05 path does not cover this shape (%s)", mnb_last_error());
    } else MNB_REQUIRE(impl != 2, "conv_fwd: tcgen05 path needs bf16");
    return conv_fwd_simt(p, dtype, (cudaStream_t)stream);
}

int mnb_conv_dgrad_packed(const void* dz, const float* w, const void* wpk, const void* add, void* dx, const void* bn_z,
                          const float* bn_scale, const float* bn_shift, double* bn_sums, int N, int H, int W, int Cin,
                          int Cout, int k, int stride, int pad, int dtype, int impl, void* stream) {
    if (int e = check_conv("conv_dgrad", N, H, W, Cin, Cout, k, stride, pad, dtype, MNB_LAYOUT_NHWC)) return e;
    MNB_REQUIRE(!bn_z || (bn_scale && bn_shift && bn_sums), "conv_dgrad: bn_z needs bn_scale/bn_shift/bn_sums");
    ConvP p = {};
    p.wpk = wpk;
    p.dz = dz; p.w = w; p.add = add; p.out = dx;
    p.bn_z = bn_z; p.bn_scale = bn_scale; p.bn_shift = bn_shift; p.stats = bn_z ? bn_sums : nullptr;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.k = k; p.stride = stride; p.pad = pad;
    p.Ho = (H + 2 * pad - k) / stride + 1; p.Wo = (W + 2 * pad - k) / stride + 1;
    if (dtype == MNB_BF16 && prefer_stream(impl)) {
        int r = conv_dgrad_stream(p, (cudaStream_t)stream);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_BF16 && (impl == 0 || impl == 3)) {
        int r = conv_dgrad_c3(p, (cudaStream_t)stream);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (impl != 1 && dtype == MNB_BF16) {
        int r = conv_dgrad_tc(p, (cudaStream_t)stream);
        if (r != MNB_ERR_UNSUPPORTED) return r;
        MNB_REQUIRE(impl != 2, "conv_dgrad: tcgen05 path does not cover this shape (%s)", mnb_last_error());
    } else MNB_REQUIRE(impl != 2, "conv_dgrad: tcgen05 path needs bf16");
    return conv_dgrad_simt(p, dtype, (cudaStream_t)stream);
}

int mnb_fc_fwd_tc(const void* xb, const float* w, const void* wpk, const float* b, float* y, int relu_out, int N, int K,
                  int O, void* stream) {
    MNB_REQUIRE(N > 0 && K > 0 && O > 0 && K % 8 == 0, "fc_fwd_tc: K=%d must be a positive multiple of 8", K);
    ConvP p = {};
    p.x = xb; p.w = w; p.wpk = wpk; p.bias = b; p.out = y; p.out_f32 = 1; p.relu_out = relu_out;
    p.N = N; p.H = 1; p.W = 1; p.Cin = K; p.Cout = O; p.k = 1; p.stride = 1; p.pad = 0; p.Ho = 1; p.Wo = 1;
    int r = conv_fwd_tc(p, (cudaStream_t)stream);
    MNB_REQUIRE(r != MNB_ERR_UNSUPPORTED, "fc_fwd_tc: %s", mnb_last_error());
    return r;
}

int mnb_fc_dgrad_tc(const void* dyb, const float* w, const void* wpk, float* dx, int N, int K, int O, void* stream) {
    MNB_REQUIRE(N > 0 && K % 8 == 0 && O % 8 == 0 && K > 0 && O > 0, "fc_dgrad_tc: K=%d, O=%d must be multiples of 8", K, O);
    ConvP p = {};
    p.dz = dyb; p.w = w; p.wpk = wpk; p.out = dx; p.out_f32 = 1;
    p.N = N; p.H = 1; p.W = 1; p.Cin = K; p.Cout = O; p.k = 1; p.stride = 1; p.pad = 0; p.Ho = 1; p.Wo = 1;
    int r = conv_dgrad_tc(p, (cudaStream_t)stream);
    MNB_REQUIRE(r != MNB_ERR_UNSUPPORTED, "fc_dgrad_tc: %s", mnb_last_error());
    return r;
}

int mnb_conv_wgrad(const void* x, const float* in_scale, const float* in_shift, const void* dz, float* dw, int N,
                   int H, int W, int Cin, int Cout, int k, int stride, int pad, int dtype, int x_layout, int impl,
                   void* stream) {
    if (int e = check_conv("conv_wgrad", N, H, W, Cin, Cout, k, stride, pad, dtype, x_layout)) return e;
    ConvP p = {};
    p.x = x; p.in_scale = in_scale; p.in_shift = in_shift; p.dz = dz; p.dw = dw;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.k = k; p.stride = stride; p.pad = pad;
    p.Ho = (H + 2 * pad - k) / stride + 1; p.Wo = (W + 2 * pad - k) / stride + 1;
    p.nchw_in = x_layout == MNB_LAYOUT_NCHW_F32;
    if (x_layout == MNB_LAYOUT_NHWC_U8) {
        MNB_REQUIRE(in_scale && in_shift, "conv_wgrad: uint8 input needs mean (in_scale) and std (in_shift)");
        return stem_wgrad(x, 1, in_scale, in_shift, dz, dw, N, H, W, dtype, impl == 1 ? 2 : impl, (cudaStream_t)stream);
    }
    if (impl != 1 && in_scale == nullptr && stem_supported(Cin, Cout, k, stride, pad, p.nchw_in))
        return stem_wgrad(x, 0, nullptr, nullptr, dz, dw, N, H, W, dtype, impl, (cudaStream_t)stream);
    if (dtype == MNB_BF16 && prefer_stream(impl)) {
        int r = conv_wgrad_stream(p, (cudaStream_t)stream);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_BF16 && (impl == 0 || impl == 3)) {
        int r = conv_wgrad_c3(p, (cudaStream_t)stream);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (impl != 1 && dtype == MNB_BF16) {
        int r = conv_wgrad_tc(p, (cudaStream_t)stream);
        if (r != MNB_ERR_UNSUPPORTED) return r;
        MNB_REQUIRE(impl != 2, "conv_wgrad: tcgen05 path does not cover this shape (%s)", mnb_last_error());
    } else MNB_REQUIRE(impl != 2, "conv_wgrad: tcgen05 path needs bf16");
    return conv_wgrad_simt(p, dtype, (cudaStream_t)stream);
}

}  // extern "C"
