// Depthwise k x k (k in {3,5}), stride 1, pad k/2 convolution on NHWC: forward, dgrad, wgrad.
// Replaces nn.Conv2d(groups=C) at src/models/mnasnet.py:76-81,120-125 (+ the BN-apply/ReLU of the producing
// ConvBlock fused into the load, + this ConvBlock's BN batch statistics fused into the store).
//
// Mapping (round-1 kernel): a thread owns ONE channel pair (weights live in registers: 2*k*k floats) and a
// TH x TW output strip; warp lanes run along C so every warp-level access is a contiguous 128-B (bf16) /
// 256-B (fp32) segment of an NHWC row.  Input rows slide through registers (each input row is loaded once
// per strip and feeds k output rows), so the FMA:load ratio is k*k*TW*2 : (TW+k-1).  CTAs are persistent
// (grid-stride over strips) and keep per-channel partial sums in registers; one fp64 atomic per channel per
// CTA at the end.  Arithmetic is fp32; the 5x5 layers need 25 MAC per 4 bytes, i.e. they sit at the fp32-FMA
// roof rather than the HBM roof.
#include "common.cuh"

namespace mnb {

struct DwGeom {
    dim3 grid, block;
    int txc;
};

template <int TH, int TW>
static DwGeom dw_geom(int N, int H, int W, int C, int ctas_per_sm) {
    int cp = C / 2;
    int tx = largest_divisor_le(cp, 32);
    int ty = 128 / tx;
    if (ty < 1) ty = 1;
    long long strips = (long long)N * cdiv(H, TH) * cdiv(W, TW);
    long long gy = cdiv(strips, ty);
    long long cap = (long long)num_sms() * ctas_per_sm / (cp / tx);
    if (cap < 1) cap = 1;
    if (gy > cap) gy = cap;
    DwGeom g;
    g.grid = dim3(cp / tx, (unsigned)gy);
    g.block = dim3(tx, ty);
    g.txc = tx;
    return g;
}

// XF: apply a = max(s*x+t,0) on load.  FLIP: correlate with the 180-degree rotated kernel (dgrad).
template <typename T, int K, int TH, int TW, bool XF, bool FLIP, bool STATS>
__global__ void __launch_bounds__(128) dw_fwd_k(const T* __restrict__ x, const float* __restrict__ in_scale,
                                                const float* __restrict__ in_shift, const float* __restrict__ w,
                                                const float* __restrict__ bias, T* __restrict__ z, double* stats,
                                                const T* __restrict__ bnz, const float* __restrict__ bns,
                                                const float* __restrict__ bnt, int N, int H, int W, int C) {
    constexpr int P = K / 2;
    constexpr int IW = TW + K - 1;
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    float gs0 = 0.f, gs1 = 0.f, gt0 = 0.f, gt1 = 0.f;       // dgrad: producer block's BN scale/shift
    if (FLIP && STATS) { gs0 = bns[c0]; gs1 = bns[c0 + 1]; gt0 = bnt[c0]; gt1 = bnt[c0 + 1]; }
    float wr[K][K][2];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            int ii = FLIP ? K - 1 - i : i, jj = FLIP ? K - 1 - j : j;
            wr[i][j][0] = w[(c0 + 0) * K * K + ii * K + jj];
            wr[i][j][1] = w[(c0 + 1) * K * K + ii * K + jj];
        }
    float s0 = 1.f, s1 = 1.f, t0 = 0.f, t1 = 0.f, b0 = 0.f, b1 = 0.f;
    if (XF) { s0 = in_scale[c0]; s1 = in_scale[c0 + 1]; t0 = in_shift[c0]; t1 = in_shift[c0 + 1]; }
    if (bias) { b0 = bias[c0]; b1 = bias[c0 + 1]; }
    float st[4] = {0.f, 0.f, 0.f, 0.f};

    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    const long long strips = (long long)N * tiles_h * tiles_w;
    for (long long sidx = (long long)blockIdx.y * blockDim.y + threadIdx.y; sidx < strips;
         sidx += (long long)gridDim.y * blockDim.y) {
        const int tw_i = (int)(sidx % tiles_w);
        const int th_i = (int)((sidx / tiles_w) % tiles_h);
        const int n = (int)(sidx / ((long long)tiles_w * tiles_h));
        const int h0 = th_i * TH, w0 = tw_i * TW;
        const T* xn = x + (long long)n * H * W * C + c0;
        T* zn = z + (long long)n * H * W * C + c0;
        float acc[K][TW][2];
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < TW; ++j) acc[i][j][0] = acc[i][j][1] = 0.f;
#pragma unroll
        for (int r = 0; r < TH + K - 1; ++r) {
            const int ih = h0 - P + r;
            float in[IW][2];
            const bool rv = (ih >= 0) && (ih < H);
#pragma unroll
            for (int j = 0; j < IW; ++j) {
                const int iw = w0 - P + j;
                float2 v = make_float2(0.f, 0.f);
                if (rv && iw >= 0 && iw < W) {
                    v = load2(xn + ((long long)ih * W + iw) * C);
                    if (XF) { v.x = fmaxf(fmaf(s0, v.x, t0), 0.f); v.y = fmaxf(fmaf(s1, v.y, t1), 0.f); }
                }
                in[j][0] = v.x; in[j][1] = v.y;
            }
#pragma unroll
            for (int kh = 0; kh < K; ++kh) {
                const int o = r - kh;           // output row (relative) fed by this input row through tap kh
                if (o >= 0 && o < TH) {
#pragma unroll
                    for (int j = 0; j < TW; ++j)
#pragma unroll
                        for (int kw = 0; kw < K; ++kw) {
                            acc[o % K][j][0] = fmaf(in[j + kw][0], wr[kh][kw][0], acc[o % K][j][0]);
                            acc[o % K][j][1] = fmaf(in[j + kw][1], wr[kh][kw][1], acc[o % K][j][1]);
                        }
                }
            }
            const int od = r - (K - 1);         // output row completed by this input row
            if (od >= 0) {
                const int oh = h0 + od;
#pragma unroll
                for (int j = 0; j < TW; ++j) {
                    float v0 = acc[od % K][j][0] + b0, v1 = acc[od % K][j][1] + b1;
                    acc[od % K][j][0] = 0.f; acc[od % K][j][1] = 0.f;
                    if (oh < H && w0 + j < W) {
                        store2(zn + ((long long)oh * W + w0 + j) * C, v0, v1);
                        if (STATS) {
                            // statistics of the values as stored (bf16-rounded in bf16 mode)
                            float q0 = to_f(from_f<T>(v0)), q1 = to_f(from_f<T>(v1));
                            if (!FLIP) {
                                st[0] += q0; st[1] += q1;
                                st[2] = fmaf(q0, q0, st[2]); st[3] = fmaf(q1, q1, st[3]);
                            } else {    // fused BN-backward reduction of the producer block
                                const float2 zz = load2(bnz + (long long)n * H * W * C + c0 + ((long long)oh * W + w0 + j) * C);
                                const float g0 = fmaf(gs0, zz.x, gt0) > 0.f ? q0 : 0.f;
                                const float g1 = fmaf(gs1, zz.y, gt1) > 0.f ? q1 : 0.f;
                                st[0] += g0; st[1] += g1;
                                st[2] = fmaf(g0, zz.x, st[2]); st[3] = fmaf(g1, zz.y, st[3]);
                            }
                        }
                    }
                }
            }
        }
    }
    if (STATS) {
        __shared__ float red[4][32];
        const int tid = threadIdx.y * blockDim.x + threadIdx.x;
        for (int i = tid; i < 4 * 32; i += blockDim.x * blockDim.y) (&red[0][0])[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) atomicAdd(&red[q][threadIdx.x], st[q]);
        __syncthreads();
        if (threadIdx.y == 0) {
            atomicAdd(&stats[c0], (double)red[0][threadIdx.x]);
            atomicAdd(&stats[c0 + 1], (double)red[1][threadIdx.x]);
            atomicAdd(&stats[C + c0], (double)red[2][threadIdx.x]);
            atomicAdd(&stats[C + c0 + 1], (double)red[3][threadIdx.x]);
        }
    }
}

// dw[c,kh,kw] += sum_{n,h,w} a(n,h+kh-P,w+kw-P,c) * dz(n,h,w,c)
template <typename T, int K, int TH, int TW, bool XF>
__global__ void __launch_bounds__(128) dw_wgrad_k(const T* __restrict__ x, const float* __restrict__ in_scale,
                                                  const float* __restrict__ in_shift, const T* __restrict__ dz,
                                                  float* dw, int N, int H, int W, int C) {
    constexpr int P = K / 2;
    constexpr int IW = TW + K - 1;
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    float wa[K][K][2];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) wa[i][j][0] = wa[i][j][1] = 0.f;
    float s0 = 1.f, s1 = 1.f, t0 = 0.f, t1 = 0.f;
    if (XF) { s0 = in_scale[c0]; s1 = in_scale[c0 + 1]; t0 = in_shift[c0]; t1 = in_shift[c0 + 1]; }

    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
    const long long strips = (long long)N * tiles_h * tiles_w;
    for (long long sidx = (long long)blockIdx.y * blockDim.y + threadIdx.y; sidx < strips;
         sidx += (long long)gridDim.y * blockDim.y) {
        const int tw_i = (int)(sidx % tiles_w);
        const int th_i = (int)((sidx / tiles_w) % tiles_h);
        const int n = (int)(sidx / ((long long)tiles_w * tiles_h));
        const int h0 = th_i * TH, w0 = tw_i * TW;
        const T* xn = x + (long long)n * H * W * C + c0;
        const T* dn = dz + (long long)n * H * W * C + c0;
        float g[K][TW][2];      // ring of the K live dz rows
#pragma unroll
        for (int r = 0; r < TH + K - 1; ++r) {
            const int ih = h0 - P + r;
            if (r < TH) {       // dz row o = r enters the ring
                const int oh = h0 + r;
#pragma unroll
                for (int j = 0; j < TW; ++j) {
                    float2 v = make_float2(0.f, 0.f);
                    if (oh < H && w0 + j < W) v = load2(dn + ((long long)oh * W + w0 + j) * C);
                    g[r % K][j][0] = v.x; g[r % K][j][1] = v.y;
                }
            }
            float in[IW][2];
            const bool rv = (ih >= 0) && (ih < H);
#pragma unroll
            for (int j = 0; j < IW; ++j) {
                const int iw = w0 - P + j;
                float2 v = make_float2(0.f, 0.f);
                if (rv && iw >= 0 && iw < W) {
                    v = load2(xn + ((long long)ih * W + iw) * C);
                    if (XF) { v.x = fmaxf(fmaf(s0, v.x, t0), 0.f); v.y = fmaxf(fmaf(s1, v.y, t1), 0.f); }
                }
                in[j][0] = v.x; in[j][1] = v.y;
            }
#pragma unroll
            for (int kh = 0; kh < K; ++kh) {
                const int o = r - kh;
                if (o >= 0 && o < TH) {
#pragma unroll
                    for (int j = 0; j < TW; ++j)
#pragma unroll
                        for (int kw = 0; kw < K; ++kw) {
                            wa[kh][kw][0] = fmaf(in[j + kw][0], g[o % K][j][0], wa[kh][kw][0]);
                            wa[kh][kw][1] = fmaf(in[j + kw][1], g[o % K][j][1], wa[kh][kw][1]);
                        }
                }
            }
        }
    }
    // block reduction over the threads that share a channel pair, then one fp32 atomic per tap per CTA
    __shared__ float red[32 * 2 * K * K];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nt = blockDim.x * blockDim.y;
    for (int i = tid; i < 32 * 2 * K * K; i += nt) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            atomicAdd(&red[(threadIdx.x * 2 + 0) * K * K + i * K + j], wa[i][j][0]);
            atomicAdd(&red[(threadIdx.x * 2 + 1) * K * K + i * K + j], wa[i][j][1]);
        }
    __syncthreads();
    const int cbase = blockIdx.x * blockDim.x * 2;
    for (int i = tid; i < (int)blockDim.x * 2 * K * K; i += nt) atomicAdd(&dw[(long long)cbase * K * K + i], red[i]);
}

template <typename T, int K, bool FLIP>
static int launch_dw_fwd(const T* x, const float* s, const float* t, const float* w, const float* bias, T* z,
                         double* stats, const T* bnz, const float* bns, const float* bnt, int N, int H, int W, int C,
                         cudaStream_t st) {
    constexpr int TH = 8, TW = 4;
    DwGeom g = dw_geom<TH, TW>(N, H, W, C, 8);
    if (s && stats) dw_fwd_k<T, K, TH, TW, true, FLIP, true><<<g.grid, g.block, 0, st>>>(x, s, t, w, bias, z, stats, bnz, bns, bnt, N, H, W, C);
    else if (s) dw_fwd_k<T, K, TH, TW, true, FLIP, false><<<g.grid, g.block, 0, st>>>(x, s, t, w, bias, z, stats, bnz, bns, bnt, N, H, W, C);
    else if (stats) dw_fwd_k<T, K, TH, TW, false, FLIP, true><<<g.grid, g.block, 0, st>>>(x, s, t, w, bias, z, stats, bnz, bns, bnt, N, H, W, C);
    else dw_fwd_k<T, K, TH, TW, false, FLIP, false><<<g.grid, g.block, 0, st>>>(x, s, t, w, bias, z, stats, bnz, bns, bnt, N, H, W, C);
    return 0;
}

template <typename T, int K>
static int launch_dw_wgrad(const T* x, const float* s, const float* t, const T* dz, float* dw, int N, int H, int W,
                           int C, cudaStream_t st) {
    constexpr int TH = 8, TW = 4;
    DwGeom g = dw_geom<TH, TW>(N, H, W, C, 4);
    if (s) dw_wgrad_k<T, K, TH, TW, true><<<g.grid, g.block, 0, st>>>(x, s, t, dz, dw, N, H, W, C);
    else dw_wgrad_k<T, K, TH, TW, false><<<g.grid, g.block, 0, st>>>(x, s, t, dz, dw, N, H, W, C);
    return 0;
}

// shared-memory halo-tile kernels (dwconv_tile.cu), bf16 only
int dw_fwd_tile(const void* x, const float* s, const float* t, const float* w, const float* bias, void* z, double* stats,
                int N, int H, int W, int C, int k, cudaStream_t st);
int dw_dgrad_tile(const void* dz, const float* w, void* dx, const void* bn_z, const float* bn_scale,
                  const float* bn_shift, double* bn_sums, int N, int H, int W, int C, int k, cudaStream_t st);
int dw_wgrad_tile(const void* x, const float* s, const float* t, const void* dz, float* dw, int N, int H, int W, int C,
                  int k, cudaStream_t st);

static int check_dw(const char* name, int N, int H, int W, int C, int k, int dtype) {
    MNB_REQUIRE(N > 0 && H > 0 && W > 0, "%s: bad N/H/W", name);
    MNB_REQUIRE(C > 0 && C % 8 == 0, "%s: C=%d must be a positive multiple of 8", name, C);
    MNB_REQUIRE(k == 3 || k == 5, "%s: k=%d unsupported (3 or 5)", name, k);
    MNB_REQUIRE(dtype == MNB_F32 || dtype == MNB_BF16, "%s: bad dtype %d", name, dtype);
    return 0;
}

}  // namespace mnb

using namespace mnb;

// TMA + mma.sync kernels (dw_mma.cu), option "dw_mma" (default 1): bf16 forward / stand-alone backward-data
namespace mnb {
int dw_fwd_mma(const void* x, const float* s, const float* t, const float* w, const float* bias, void* z, double* stats,
               int N, int H, int W, int C, int k, cudaStream_t st);
int dw_dgrad_mma(const void* dz, const float* w, void* dx, int N, int H, int W, int C, int k, cudaStream_t st);
int dw_wgrad_mma(const void* x, const float* s, const float* t, const void* dz, float* dw, int N, int H, int W, int C, int k,
                 cudaStream_t st);
// whole-tile variants for maps of at most 28 rows (dw_small.cu), option "dw_small"
bool dw_small_covers(int H, int W, int C, int k);
bool dw_small_covers_bwd(int H, int W, int C, int k);
int dw_fwd_small(const void* x, const float* s, const float* t, const float* w, void* z, double* stats, int N, int H, int W,
                 int C, int k, cudaStream_t st);
int dw_dgrad_small(const void* dz, const float* w, void* dx, int N, int H, int W, int C, int k, cudaStream_t st);
int dw_wgrad_small(const void* x, const float* s, const float* t, const void* dz, float* dw, int N, int H, int W, int C, int k,
                   cudaStream_t st);
int dw_bwd_small(const void* G, const void* Z, const float* scale, const float* shift, const double* sums, const float* mean,
                 const float* invstd, double m, float* dgamma, float* dbeta, float* dbias, const void* X,
                 const float* in_scale, const float* in_shift, const float* w, void* dX, float* dw, double* nsums, int N, int H,
                 int W, int C, int k, cudaStream_t st);
int pw_proj_bwd(const void* dz, const void* x, const float* in_scale, const float* in_shift, const float* w, const void* add,
                void* dx, float* dw, double* nsums, long long M, int Cin, int Cout, cudaStream_t st);
int pw_bwd_fused(const void* G, const void* Z, const float* scale, const float* shift, const double* sums, const float* mean,
                 const float* invstd, double m, float* dgamma, float* dbeta, float* dbias, const void* X,
                 const float* in_scale, const float* in_shift, const float* w, const void* add, void* dX, float* dw,
                 double* nsums, long long M, int Cin, int Cout, cudaStream_t st);
int dw_bwd_mma(const void* G, const void* Z, const float* scale, const float* shift, const double* sums, const float* mean,
               const float* invstd, double m, float* dgamma, float* dbeta, float* dbias, const void* X,
               const float* in_scale, const float* in_shift, const float* w, void* dX, float* dw, double* nsums, int N, int H,
               int W, int C, int k, cudaStream_t st);
}  // namespace mnb

extern "C" {

int mnb_dw_fwd(const void* x, const float* in_scale, const float* in_shift, const float* w, const float* bias,
               void* z, double* stats, int N, int H, int W, int C, int k, int dtype, void* stream) {
    if (int e = check_dw("dw_fwd", N, H, W, C, k, dtype)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    // tensor-pipe kernel from 14-row maps up (7x7 / 4x4 maps fill 16-pixel strips too poorly: tile kernel); dw_mma = 2 forces it
    if (dtype == MNB_BF16 && !bias && dw_small_covers(H, W, C, k)) {
        int r = dw_fwd_small(x, in_scale, in_shift, w, z, stats, N, H, W, C, k, st);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_BF16 && !bias &&
        (option_get(OPT_DW_MMA) == 2 || (option_get(OPT_DW_MMA) == 1 && H >= 12 && W >= 12))) {
        int r = dw_fwd_mma(x, in_scale, in_shift, w, bias, z, stats, N, H, W, C, k, st);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_BF16) {
        int r = dw_fwd_tile(x, in_scale, in_shift, w, bias, z, stats, N, H, W, C, k, st);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_F32) {
        if (k == 3) launch_dw_fwd<float, 3, false>((const float*)x, in_scale, in_shift, w, bias, (float*)z, stats, nullptr, nullptr, nullptr, N, H, W, C, st);
        else launch_dw_fwd<float, 5, false>((const float*)x, in_scale, in_shift, w, bias, (float*)z, stats, nullptr, nullptr, nullptr, N, H, W, C, st);
    } else {
        if (k == 3) launch_dw_fwd<bf16, 3, false>((const bf16*)x, in_scale, in_shift, w, bias, (bf16*)z, stats, nullptr, nullptr, nullptr, N, H, W, C, st);
        else launch_dw_fwd<bf16, 5, false>((const bf16*)x, in_scale, in_shift, w, bias, (bf16*)z, stats, nullptr, nullptr, nullptr, N, H, W, C, st);
    }
    MNB_LAUNCH_CHECK("dw_fwd");
    return 0;
}

int mnb_dw_dgrad(const void* dz, const float* w, void* dx, const void* bn_z, const float* bn_scale,
                 const float* bn_shift, double* bn_sums, int N, int H, int W, int C, int k, int dtype, void* stream) {
    if (int e = check_dw("dw_dgrad", N, H, W, C, k, dtype)) return e;
    MNB_REQUIRE(!bn_z || (bn_scale && bn_shift && bn_sums), "dw_dgrad: bn_z needs bn_scale/bn_shift/bn_sums");
    if (!bn_z) bn_sums = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MNB_BF16 && !bn_z && dw_small_covers(H, W, C, k)) {
        int r = dw_dgrad_small(dz, w, dx, N, H, W, C, k, st);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_BF16 && !bn_z &&
        (option_get(OPT_DW_MMA) == 2 || (option_get(OPT_DW_MMA) == 1 && H >= 12 && W >= 12))) {
        int r = dw_dgrad_mma(dz, w, dx, N, H, W, C, k, st);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_BF16) {
        int r = dw_dgrad_tile(dz, w, dx, bn_z, bn_scale, bn_shift, bn_sums, N, H, W, C, k, st);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_F32) {
        if (k == 3) launch_dw_fwd<float, 3, true>((const float*)dz, nullptr, nullptr, w, nullptr, (float*)dx, bn_sums, (const float*)bn_z, bn_scale, bn_shift, N, H, W, C, st);
        else launch_dw_fwd<float, 5, true>((const float*)dz, nullptr, nullptr, w, nullptr, (float*)dx, bn_sums, (const float*)bn_z, bn_scale, bn_shift, N, H, W, C, st);
    } else {
        if (k == 3) launch_dw_fwd<bf16, 3, true>((const bf16*)dz, nullptr, nullptr, w, nullptr, (bf16*)dx, bn_sums, (const bf16*)bn_z, bn_scale, bn_shift, N, H, W, C, st);
        else launch_dw_fwd<bf16, 5, true>((const bf16*)dz, nullptr, nullptr, w, nullptr, (bf16*)dx, bn_sums, (const bf16*)bn_z, bn_scale, bn_shift, N, H, W, C, st);
    }
    MNB_LAUNCH_CHECK("dw_dgrad");
    return 0;
}

int mnb_dw_wgrad(const void* x, const float* in_scale, const float* in_shift, const void* dz, float* dw, int N,
                 int H, int W, int C, int k, int dtype, void* stream) {
    if (int e = check_dw("dw_wgrad", N, H, W, C, k, dtype)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    // tensor-pipe backward-weight where it measured faster than the tile kernel (profiles/r2_exp_dw_mma.json): maps of
    // >= 12 rows whose channel count the 24-channel geometry tiles exactly (not the 32-channel stem block)
    if (dtype == MNB_BF16 && dw_small_covers(H, W, C, k)) {
        int r = dw_wgrad_small(x, in_scale, in_shift, dz, dw, N, H, W, C, k, st);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_BF16 &&
        (option_get(OPT_DW_MMA) == 2 || (option_get(OPT_DW_MMA) == 1 && H >= 12 && W >= 12 && C % 24 == 0))) {
        int r = dw_wgrad_mma(x, in_scale, in_shift, dz, dw, N, H, W, C, k, st);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_BF16) {
        int r = dw_wgrad_tile(x, in_scale, in_shift, dz, dw, N, H, W, C, k, st);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    if (dtype == MNB_F32) {
        if (k == 3) launch_dw_wgrad<float, 3>((const float*)x, in_scale, in_shift, (const float*)dz, dw, N, H, W, C, st);
        else launch_dw_wgrad<float, 5>((const float*)x, in_scale, in_shift, (const float*)dz, dw, N, H, W, C, st);
    } else {
        if (k == 3) launch_dw_wgrad<bf16, 3>((const bf16*)x, in_scale, in_shift, (const bf16*)dz, dw, N, H, W, C, st);
        else launch_dw_wgrad<bf16, 5>((const bf16*)x, in_scale, in_shift, (const bf16*)dz, dw, N, H, W, C, st);
    }
    MNB_LAUNCH_CHECK("dw_wgrad");
    return 0;
}

int mnb_dw_bwd_fused(const void* dA, const void* z, const float* scale, const float* shift, const double* sums,
                     const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, float* dbias,
                     const void* x, const float* in_scale, const float* in_shift, const float* w, void* dx, float* dw,
                     double* in_sums, int N, int H, int W, int C, int k, double m, int dtype, void* stream) {
    if (int e = check_dw("dw_bwd_fused", N, H, W, C, k, dtype)) return e;
    MNB_REQUIRE(dA && z && scale && shift && sums && save_mean && save_invstd && x && w && dx, "dw_bwd_fused: NULL operand");
    MNB_REQUIRE(!in_sums || in_scale, "dw_bwd_fused: in_sums needs in_scale / in_shift");
    MNB_REQUIRE(!in_scale || in_shift, "dw_bwd_fused: in_scale without in_shift");
    MNB_REQUIRE(m > 0, "dw_bwd_fused: bad element count");
    if (dtype != MNB_BF16) { set_error("dw_bwd_fused: bf16 only"); return MNB_ERR_UNSUPPORTED; }
    if (dw_small_covers_bwd(H, W, C, k)) {
        int r = dw_bwd_small(dA, z, scale, shift, sums, save_mean, save_invstd, m, dgamma, dbeta, dbias, x, in_scale, in_shift,
                             w, dx, dw, in_sums, N, H, W, C, k, (cudaStream_t)stream);
        if (r != MNB_ERR_UNSUPPORTED) return r;
    }
    return dw_bwd_mma(dA, z, scale, shift, sums, save_mean, save_invstd, m, dgamma, dbeta, dbias, x, in_scale, in_shift, w,
                      dx, dw, in_sums, N, H, W, C, k, (cudaStream_t)stream);
}

int mnb_pw_bwd_fused(const void* dA, const void* z, const float* scale, const float* shift, const double* sums,
                     const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, float* dbias,
                     const void* x, const float* in_scale, const float* in_shift, const float* w, const void* add,
                     void* dx, float* dw, double* in_sums, long long M, int Cin, int Cout, double m, int dtype,
                     void* stream) {
    MNB_REQUIRE(M > 0 && Cin > 0 && Cout > 0 && Cin % 8 == 0 && Cout % 8 == 0, "pw_bwd_fused: bad M / Cin / Cout");
    MNB_REQUIRE(dA && z && scale && shift && sums && save_mean && save_invstd && x && w && dx, "pw_bwd_fused: NULL operand");
    MNB_REQUIRE(!in_sums || in_scale, "pw_bwd_fused: in_sums needs in_scale / in_shift");
    MNB_REQUIRE(!in_scale || in_shift, "pw_bwd_fused: in_scale without in_shift");
    MNB_REQUIRE(m > 0, "pw_bwd_fused: bad element count");
    if (dtype != MNB_BF16) { set_error("pw_bwd_fused: bf16 only"); return MNB_ERR_UNSUPPORTED; }
    return pw_bwd_fused(dA, z, scale, shift, sums, save_mean, save_invstd, m, dgamma, dbeta, dbias, x, in_scale, in_shift, w,
                        add, dx, dw, in_sums, M, Cin, Cout, (cudaStream_t)stream);
}

int mnb_pw_proj_bwd(const void* dz, const void* x, const float* in_scale, const float* in_shift, const float* w,
                    const void* add, void* dx, float* dw, double* in_sums, long long M, int Cin, int Cout, int dtype,
                    void* stream) {
    MNB_REQUIRE(M > 0 && Cin > 0 && Cout > 0 && Cin % 8 == 0 && Cout % 8 == 0, "pw_proj_bwd: bad M / Cin / Cout");
    MNB_REQUIRE(dz && x && w && dx, "pw_proj_bwd: NULL operand");
    MNB_REQUIRE(!in_sums || in_scale, "pw_proj_bwd: in_sums needs in_scale / in_shift");
    MNB_REQUIRE(!in_scale || in_shift, "pw_proj_bwd: in_scale without in_shift");
    if (dtype != MNB_BF16) { set_error("pw_proj_bwd: bf16 only"); return MNB_ERR_UNSUPPORTED; }
    return pw_proj_bwd(dz, x, in_scale, in_shift, w, add, dx, dw, in_sums, M, Cin, Cout, (cudaStream_t)stream);
}

}  // extern "C"
