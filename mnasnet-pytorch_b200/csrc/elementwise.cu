// BatchNorm finalize / apply / backward, head (GAP, dropout, FC), loss, Adam, layout helpers.
// All HBM-bound streaming kernels: 128-bit vector IO on NHWC rows, channel-owner threads for the
// per-channel reductions (warp lanes run along C so every warp-level access is one contiguous segment),
// per-CTA shared-memory partials flushed with one fp64 atomic per channel per CTA.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace mnb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// kernel-selection switches: -1 = not set yet (the environment variable is consulted once, then the built-in default)
static int g_opt[OPT_COUNT] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
static const char* const g_opt_name[OPT_COUNT] = {"pw_stream", "stem_mma", "dw_mma", "dw_mma_cg", "dw_mma_tws", "dw_mma_seg",
                                                   "bn_ctas", "c3_mma", "dw_small", "pwb_slice", "pw_wide"};
static const char* const g_opt_env[OPT_COUNT] = {"MNB_PW_STREAM", "MNB_STEM_MMA", "MNB_DW_MMA", "MNB_DW_MMA_CG", "MNB_DW_MMA_TWS",
                                                  "MNB_DW_MMA_SEG", "MNB_BN_CTAS", "MNB_C3_MMA", "MNB_DW_SMALL", "MNB_PWB_SLICE",
                                                  "MNB_PW_WIDE"};
static const int g_opt_default[OPT_COUNT] = {1, 1, 1, 0, 0, 0, 0, 1, 1, 0, 0};
int option_get(int id) {
    if (g_opt[id] < 0) {
        const char* e = getenv(g_opt_env[id]);
        g_opt[id] = e ? atoi(e) : g_opt_default[id];
    }
    return g_opt[id];
}

// ---------------------------------------------------------------------------------------------------
// channel-owner launch geometry: blockDim = (TX, TY); thread x owns channel vector blockIdx.x*TX + x
// ---------------------------------------------------------------------------------------------------
struct ColGeom {
    dim3 grid, block;
};
// min_rows: rows every thread should get at least (amortises per-thread constants and the per-CTA flush on small tensors)
static ColGeom col_geom(long long M, int C, int ctas_per_sm = 8, int min_rows = 1) {
    int cv = C / 8;
    int tx = largest_divisor_le(cv, 32);
    int ty = 256 / tx;
    long long gy = cdiv(M, (long long)ty * min_rows);
    long long cap = (long long)num_sms() * ctas_per_sm / (cv / tx);
    if (cap < 1) cap = 1;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    ColGeom g;
    g.grid = dim3(cv / tx, (unsigned)gy);
    g.block = dim3(tx, ty);
    return g;
}

// block-level reduction of per-thread 8-channel partials (threads sharing threadIdx.x): shared-memory tree over
// threadIdx.y (no atomics: with narrow tensors 128 threads share one channel vector), then one fp64 atomic per
// channel per CTA
template <int NQ>
__device__ __forceinline__ void flush_channel_partials(float (&acc)[NQ][8], double* out, int C, int c0) {
    __shared__ float part[256 * NQ * 8];
    const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
    float* mine = part + (ty * TX + tx) * (NQ * 8);
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int i = 0; i < 8; ++i) mine[q * 8 + i] = acc[q][i];
    int s = 1;
    while (s < TY) s <<= 1;
    for (s >>= 1; s > 0; s >>= 1) {
        __syncthreads();
        if (ty < s && ty + s < TY) {
            const float* other = part + ((ty + s) * TX + tx) * (NQ * 8);
#pragma unroll
            for (int k = 0; k < NQ * 8; ++k) mine[k] += other[k];
        }
    }
    if (ty == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(&out[q * C + c0 + i], (double)mine[q * 8 + i]);
    }
}

// ---------------------------------------------------------------------------------------------------
// BN finalize / eval coefficients
// ---------------------------------------------------------------------------------------------------
__global__ void bn_finalize_k(const double* __restrict__ stats, const float* __restrict__ gamma,
                              const float* __restrict__ beta, const float* __restrict__ conv_bias,
                              float* running_mean, float* running_var,
                              long long* nbt, float* scale, float* shift, float* save_mean, float* save_invstd,
                              int C, double m, float eps, float momentum) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt) *nbt += 1;
    if (c >= C) return;
    double mean = stats[c] / m;
    double var = stats[C + c] / m - mean * mean;
    if (var < 0) var = 0;
    double invstd = 1.0 / sqrt(var + (double)eps);
    double s = (double)gamma[c] * invstd;
    scale[c] = (float)s;
    shift[c] = (float)((double)beta[c] - mean * s);
    if (save_mean) save_mean[c] = (float)mean;
    if (save_invstd) save_invstd[c] = (float)invstd;
    if (running_mean) {
        double unb = m > 1 ? var * m / (m - 1) : var;
        const double mb = mean + (conv_bias ? (double)conv_bias[c] : 0.0);      // folded conv bias
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mb);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
    }
}

__global__ void bn_eval_coeffs_k(const float* gamma, const float* beta, const float* conv_bias, const float* rm,
                                 const float* rv, float* scale, float* shift, int C, float eps) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = gamma[c] / sqrtf(rv[c] + eps);
    scale[c] = s;
    shift[c] = beta[c] - (rm[c] - (conv_bias ? conv_bias[c] : 0.f)) * s;
}

// ---------------------------------------------------------------------------------------------------
// y = [residual +] relu(scale*z + shift)
// ---------------------------------------------------------------------------------------------------
template <typename T, bool RES>
__global__ void __launch_bounds__(256) bn_relu_apply_k(const T* __restrict__ z, const float* __restrict__ scale,
                                                       const float* __restrict__ shift,
                                                       const T* __restrict__ res, T* __restrict__ y,
                                                       long long M, int C) {
    // channel-owner threads: the per-channel constants stay in registers for the whole row loop
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    float s[8], t[8];
    load8(scale + c0, s);
    load8(shift + c0, t);
    const long long stride = (long long)gridDim.y * blockDim.y;
    for (long long r = (long long)blockIdx.y * blockDim.y + threadIdx.y; r < M; r += stride) {
        float a[8], q[8];
        load8(z + r * C + c0, a);
        if (RES) load8(res + r * C + c0, q);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float o = fmaxf(fmaf(s[i], a[i], t[i]), 0.f);
            a[i] = RES ? o + q[i] : o;
        }
        store8(y + r * C + c0, a);
    }
}

// ---------------------------------------------------------------------------------------------------
// BN backward
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_k(const T* __restrict__ dA, const T* __restrict__ z,
                                                       const float* __restrict__ scale,
                                                       const float* __restrict__ shift, double* sums,
                                                       long long M, int C) {
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    float s[8], t[8], acc[2][8];
    load8(scale + c0, s);
    load8(shift + c0, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
    const long long stride = (long long)gridDim.y * blockDim.y;
    long long r = (long long)blockIdx.y * blockDim.y + threadIdx.y;
    // 4 rows (8 x 16-byte loads) in flight per thread: the kernel runs with few, fat CTAs (see the launcher)
    for (; r + 3 * stride < M; r += 4 * stride) {
        float a[4][8], zz[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            load8(dA + (r + u * stride) * C + c0, a[u]);
            load8(z + (r + u * stride) * C + c0, zz[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float g = fmaf(s[i], zz[u][i], t[i]) > 0.f ? a[u][i] : 0.f;
                acc[0][i] += g;
                acc[1][i] = fmaf(g, zz[u][i], acc[1][i]);
            }
    }
    for (; r < M; r += stride) {
        float a0[8], z0[8];
        load8(dA + r * C + c0, a0);
        load8(z + r * C + c0, z0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float g0 = fmaf(s[i], z0[i], t[i]) > 0.f ? a0[i] : 0.f;
            acc[0][i] += g0;
            acc[1][i] = fmaf(g0, z0[i], acc[1][i]);
        }
    }
    flush_channel_partials<2>(acc, sums, C, c0);
}

__global__ void bn_bwd_finalize_k(const double* __restrict__ sums, const float* __restrict__ scale,
                                  const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                                  float* dgamma, float* dbeta, float* dbias, float* coef, int C, double m) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double sg = sums[c], sgz = sums[C + c];
    double mean = save_mean[c], invstd = save_invstd[c], s = scale[c];
    double dga = invstd * (sgz - mean * sg);
    double a = s;
    double b = -s * invstd * dga / m;
    double cc = -s * sg / m - b * mean;
    coef[c] = (float)a;
    coef[C + c] = (float)b;
    coef[2 * C + c] = (float)cc;
    if (dgamma) dgamma[c] += (float)dga;
    if (dbeta) dbeta[c] += (float)sg;
    if (dbias) dbias[c] += (float)(a * sg + b * mean * m + cc * m);   // analytically 0 (bias cancels in BN)
}

template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_k(const T* __restrict__ dA, const T* __restrict__ z,
                                                      const float* __restrict__ scale,
                                                      const float* __restrict__ shift,
                                                      const float* __restrict__ coef, T* __restrict__ dZ,
                                                      long long M, int C) {
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    float s[8], t[8], ca[8], cb[8], cc[8];
    load8(scale + c0, s);
    load8(shift + c0, t);
    load8(coef + c0, ca);
    load8(coef + C + c0, cb);
    load8(coef + 2 * C + c0, cc);
    const long long stride = (long long)gridDim.y * blockDim.y;
    long long r = (long long)blockIdx.y * blockDim.y + threadIdx.y;
    for (; r + stride < M; r += 2 * stride) {
        float a0[8], z0[8], a1[8], z1[8];
        load8(dA + r * C + c0, a0);
        load8(z + r * C + c0, z0);
        load8(dA + (r + stride) * C + c0, a1);
        load8(z + (r + stride) * C + c0, z1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float g0 = fmaf(s[i], z0[i], t[i]) > 0.f ? a0[i] : 0.f;
            const float g1 = fmaf(s[i], z1[i], t[i]) > 0.f ? a1[i] : 0.f;
            a0[i] = fmaf(ca[i], g0, fmaf(cb[i], z0[i], cc[i]));
            a1[i] = fmaf(ca[i], g1, fmaf(cb[i], z1[i], cc[i]));
        }
        store8(dZ + r * C + c0, a0);
        store8(dZ + (r + stride) * C + c0, a1);
    }
    for (; r < M; r += stride) {
        float a0[8], z0[8];
        load8(dA + r * C + c0, a0);
        load8(z + r * C + c0, z0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float g0 = fmaf(s[i], z0[i], t[i]) > 0.f ? a0[i] : 0.f;
            a0[i] = fmaf(ca[i], g0, fmaf(cb[i], z0[i], cc[i]));
        }
        store8(dZ + r * C + c0, a0);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_fused_k(const T* __restrict__ dA, const T* __restrict__ z,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift,
                                                            const double* __restrict__ sums,
                                                            const float* __restrict__ save_mean,
                                                            const float* __restrict__ save_invstd, float* dgamma,
                                                            float* dbeta, float* dbias, T* __restrict__ dZ, long long M,
                                                            int C, double m) {
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    float s[8], t[8], ca[8], cb[8], cc[8];
    load8(scale + c0, s);
    load8(shift + c0, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) {       // the per-channel work of bn_bwd_finalize, redundantly per CTA
        const int c = c0 + i;
        const double sg = sums[c], sgz = sums[C + c];
        const double mean = save_mean[c], invstd = save_invstd[c], sc = s[i];
        const double dga = invstd * (sgz - mean * sg);
        const double b = -sc * invstd * dga / m;
        const double c3 = -sc * sg / m - b * mean;
        ca[i] = (float)sc; cb[i] = (float)b; cc[i] = (float)c3;
        if (blockIdx.y == 0 && threadIdx.y == 0) {
            if (dgamma) dgamma[c] += (float)dga;
            if (dbeta) dbeta[c] += (float)sg;
            if (dbias) dbias[c] += (float)(sc * sg + b * mean * m + c3 * m);
        }
    }
    const long long stride = (long long)gridDim.y * blockDim.y;
    long long r = (long long)blockIdx.y * blockDim.y + threadIdx.y;
    for (; r + stride < M; r += 2 * stride) {
        float a0[8], z0[8], a1[8], z1[8];
        load8(dA + r * C + c0, a0);
        load8(z + r * C + c0, z0);
        load8(dA + (r + stride) * C + c0, a1);
        load8(z + (r + stride) * C + c0, z1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float g0 = fmaf(s[i], z0[i], t[i]) > 0.f ? a0[i] : 0.f;
            const float g1 = fmaf(s[i], z1[i], t[i]) > 0.f ? a1[i] : 0.f;
            a0[i] = fmaf(ca[i], g0, fmaf(cb[i], z0[i], cc[i]));
            a1[i] = fmaf(ca[i], g1, fmaf(cb[i], z1[i], cc[i]));
        }
        store8(dZ + r * C + c0, a0);
        store8(dZ + (r + stride) * C + c0, a1);
    }
    for (; r < M; r += stride) {
        float a0[8], z0[8];
        load8(dA + r * C + c0, a0);
        load8(z + r * C + c0, z0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float g0 = fmaf(s[i], z0[i], t[i]) > 0.f ? a0[i] : 0.f;
            a0[i] = fmaf(ca[i], g0, fmaf(cb[i], z0[i], cc[i]));
        }
        store8(dZ + r * C + c0, a0);
    }
}

// ---------------------------------------------------------------------------------------------------
// head
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void gap_fwd_k(const T* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                          float* __restrict__ f, const unsigned char* __restrict__ mask, float ms, bf16* __restrict__ xb,
                          int N, int HW, int cv) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * cv) return;
    int n = idx / cv, c0 = (idx % cv) * 8;
    const int C = cv * 8;
    float s[8], t[8], acc[8];
    load8(scale + c0, s);
    load8(shift + c0, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const T* p = z + (long long)n * HW * C + c0;
    for (int q = 0; q < HW; ++q) {
        float a[8];
        load8(p + (long long)q * C, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += fmaxf(fmaf(s[i], a[i], t[i]), 0.f);
    }
    float inv = 1.f / (float)HW;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= inv;
    if (f) store8(f + (long long)n * C + c0, acc);
    if (xb) {       // the first Linear's bf16 A operand, dropout applied: pooling feeds the FC GEMM directly
        if (mask) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = mask[(long long)n * C + c0 + i] ? acc[i] * ms : 0.f;
        }
        store8(xb + (long long)n * C + c0, acc);
    }
}

template <typename T>
__global__ void gap_bwd_k(const float* __restrict__ df, T* __restrict__ dA, int N, int HW, int cv) {
    long long nvec = (long long)N * HW * cv;
    const int C = cv * 8;
    float inv = 1.f / (float)HW;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec;
         v += (long long)gridDim.x * blockDim.x) {
        int c0 = (int)(v % cv) * 8;
        int n = (int)(v / ((long long)HW * cv));
        float a[8];
        load8(df + (long long)n * C + c0, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] *= inv;
        store8(dA + v * 8, a);
    }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void dropout_mask_k(unsigned char* mask, long long n, float p, unsigned long long seed,
                               unsigned long long offset, const long long* dev_step) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (dev_step) offset += (unsigned long long)(*dev_step) * (unsigned long long)n;
    unsigned long long h = mix64(mix64(seed) ^ (offset + (unsigned long long)i));
    float u = (float)(h >> 40) * (1.0f / 16777216.0f);
    mask[i] = u >= p ? 1 : 0;
}

// Small tiled SGEMM for the classifier: C[M,N] = sum_k A(m,k) * B(n,k), 64x64 tile, BK 16, 4x4 per thread.
// The head is 0.2 % of the step's FLOPs; this keeps fp32 accuracy for the logits.
template <class AF, class BF, class EF>
__global__ void __launch_bounds__(256) small_gemm_k(AF af, BF bf, EF ef, int M, int N, int K) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
        for (int e = tid; e < 64 * 16; e += 256) {
            int r, kk;
            if (AF::k_contig) { r = e / 16; kk = e % 16; } else { r = e % 64; kk = e / 64; }
            As[kk][r] = (m0 + r < M && k0 + kk < K) ? af(m0 + r, k0 + kk) : 0.f;
            if (BF::k_contig) { r = e / 16; kk = e % 16; } else { r = e % 64; kk = e / 64; }
            Bs[kk][r] = (n0 + r < N && k0 + kk < K) ? bf(n0 + r, k0 + kk) : 0.f;
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
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) ef(m, n, acc[i][j]);
        }
}

struct XMasked {   // x[n,k]*mask*scale, k contiguous
    static const bool k_contig = true;
    const float* x; const unsigned char* mask; float ms; int K;
    __device__ float operator()(int n, int k) const {
        float v = x[(long long)n * K + k];
        if (mask) v = mask[(long long)n * K + k] ? v * ms : 0.f;
        return v;
    }
};
struct XMaskedT {  // as B/A operand with reduction over n: element (k, n)
    static const bool k_contig = false;
    const float* x; const unsigned char* mask; float ms; int K;
    __device__ float operator()(int k, int n) const {
        float v = x[(long long)n * K + k];
        if (mask) v = mask[(long long)n * K + k] ? v * ms : 0.f;
        return v;
    }
};
struct RowMajor {  // m[r, k] with leading dim ld, k contiguous
    static const bool k_contig = true;
    const float* p; int ld;
    __device__ float operator()(int r, int k) const { return p[(long long)r * ld + k]; }
};
struct ColMajor {  // element (r, k) = p[k*ld + r], r contiguous
    static const bool k_contig = false;
    const float* p; int ld;
    __device__ float operator()(int r, int k) const { return p[(long long)k * ld + r]; }
};
struct FcFwdEpi {
    float* y; const float* b; int O; int relu;
    __device__ void operator()(int n, int o, float v) const {
        v += b ? b[o] : 0.f;
        y[(long long)n * O + o] = relu ? fmaxf(v, 0.f) : v;
    }
};
struct FcDgradEpi {
    float* dx; const unsigned char* mask; float ms; const float* relu_ref; int K;
    __device__ void operator()(int n, int k, float v) const {
        long long i = (long long)n * K + k;
        if (mask) v = mask[i] ? v * ms : 0.f;
        if (relu_ref) v = relu_ref[i] > 0.f ? v : 0.f;
        dx[i] = v;
    }
};
struct AccumEpi {
    float* dw; int ld;
    __device__ void operator()(int o, int k, float v) const { dw[(long long)o * ld + k] += v; }
};

__global__ void fc_prep_bf16_k(const float* __restrict__ x, const unsigned char* __restrict__ mask, float ms,
                               bf16* __restrict__ xb, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = x[i];
        if (mask) v = mask[i] ? v * ms : 0.f;
        xb[i] = __float2bfloat16_rn(v);
    }
}
__global__ void fc_gate_k(float* __restrict__ dx, const unsigned char* __restrict__ mask, float ms,
                          const float* __restrict__ relu_ref, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = dx[i];
        if (mask) v = mask[i] ? v * ms : 0.f;
        if (relu_ref) v = relu_ref[i] > 0.f ? v : 0.f;
        dx[i] = v;
    }
}

__global__ void colsum_accum_k(const float* __restrict__ dy, float* db, int N, int O) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += dy[(long long)n * O + o];
    db[o] += s;
}

// ---------------------------------------------------------------------------------------------------
// softmax cross-entropy (mean) fused forward + backward: one CTA (128 threads) per row
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
    for (int o = 16; o > 0; o >>= 1) {
        float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, u) : v + u;
    }
    int w = threadIdx.x / 32, l = threadIdx.x % 32;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    int nw = blockDim.x / 32;
    v = sh[0];
    for (int i = 1; i < nw; ++i) v = is_max ? fmaxf(v, sh[i]) : v + sh[i];
    return v;
}
__global__ void __launch_bounds__(128) xent_k(const float* __restrict__ logits, const long long* __restrict__ target,
                                              float* loss, float* dlogits, int N, int O, float gscale) {
    __shared__ float sh[4];
    int n = blockIdx.x;
    const float* row = logits + (long long)n * O;
    float mx = -INFINITY;
    for (int o = threadIdx.x; o < O; o += blockDim.x) mx = fmaxf(mx, row[o]);
    mx = block_reduce(mx, true, sh);
    float se = 0.f;
    for (int o = threadIdx.x; o < O; o += blockDim.x) se += expf(row[o] - mx);
    se = block_reduce(se, false, sh);
    float lse = mx + logf(se);
    int t = (int)target[n];
    if (threadIdx.x == 0) atomicAdd(loss, (lse - row[t]) / (float)N);
    if (dlogits) {
        float inv = gscale / (float)N;
        for (int o = threadIdx.x; o < O; o += blockDim.x) {
            float p = expf(row[o] - lse);
            dlogits[(long long)n * O + o] = (p - (o == t ? 1.f : 0.f)) * inv;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, single-tensor form) over a flat buffer
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_k(float* __restrict__ p, const float* __restrict__ g,
                                              float* __restrict__ m, float* __restrict__ v, long long n,
                                              float lr, float b1, float b2, float eps, int step, float gscale,
                                              const float* dev_lr, const long long* dev_step) {
    // torch computes the bias corrections and step size in double python floats
    if (dev_lr) lr = *dev_lr;
    if (dev_step) step = (int)*dev_step;
    const double bc1 = 1.0 - pow((double)b1, (double)step);
    const double bc2 = 1.0 - pow((double)b2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * gscale;
        float mi = m[i] + (1.f - b1) * (gi - m[i]);      // lerp, as torch does
        float vi = v[i] * b2 + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= step_size * (mi / denom);
    }
}

// torch.optim.SGD (momentum 0) and torch.optim.RMSprop (alpha 0.99, not centered, momentum 0) over a flat buffer:
// the other two optimizers src/train.py:222-229 can select
__global__ void __launch_bounds__(256) sgd_k(float* __restrict__ p, const float* __restrict__ g, long long n, float lr,
                                             float gscale, const float* dev_lr) {
    if (dev_lr) lr = *dev_lr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] -= lr * (g[i] * gscale);
}
__global__ void __launch_bounds__(256) rmsprop_k(float* __restrict__ p, const float* __restrict__ g,
                                                 float* __restrict__ sq, long long n, float lr, float alpha, float eps,
                                                 float gscale, const float* dev_lr) {
    if (dev_lr) lr = *dev_lr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float s = sq[i] * alpha + (1.f - alpha) * (gi * gi);      // square_avg.mul_(alpha).addcmul_(g, g, 1-alpha)
        sq[i] = s;
        p[i] -= lr * (gi / (sqrtf(s) + eps));                           // param.addcdiv_(g, sqrt(square_avg)+eps, -lr)
    }
}

// ---------------------------------------------------------------------------------------------------
// layout helpers (module-boundary only; not on the fused step's path)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nhwc_to_nchw_k(const T* __restrict__ x, float* __restrict__ y, int N, int HW, int C) {
    long long total = (long long)N * HW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int p = (int)(i % HW);
        int c = (int)((i / HW) % C);
        long long n = i / ((long long)HW * C);
        y[i] = to_f(x[(n * HW + p) * C + c]);
    }
}
template <typename T>
__global__ void nchw_to_nhwc_k(const float* __restrict__ x, T* __restrict__ y, int N, int HW, int C) {
    long long total = (long long)N * HW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int p = (int)((i / C) % HW);
        long long n = i / ((long long)HW * C);
        y[i] = from_f<T>(x[(n * C + c) * HW + p]);
    }
}

// transforms.ToTensor() + transforms.Normalize(mean, std) (src/utils/datasets.py:460-462) on the device: uint8 HWC in,
// the fp32 NCHW tensor train.py:427 hands to the model out -- the host then ships 1 byte per value instead of 4
__global__ void u8hwc_to_nchw_norm_k(const unsigned char* __restrict__ x, const float* __restrict__ mean,
                                     const float* __restrict__ stdv, float* __restrict__ y, int N, int HW, int C) {
    const long long total = (long long)N * HW * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // i runs over the OUTPUT (coalesced fp32 stores; the byte gathers hit in L1)
        const int p = (int)(i % HW);
        const int c = (int)((i / HW) % C);
        const long long n = i / ((long long)HW * C);
        const float v = (float)x[(n * HW + p) * C + c] / 255.f;
        y[i] = (v - mean[c]) / stdv[c];
    }
}

static inline unsigned flat_grid(long long nvec, int ctas_per_sm = 8) {
    long long b = cdiv(nvec, 256);
    long long cap = (long long)num_sms() * ctas_per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace mnb

using namespace mnb;

extern "C" {

int mnb_version(void) { return 100; }
int mnb_set_option(const char* name, int value) {
    for (int i = 0; i < OPT_COUNT; ++i)
        if (!strcmp(name, g_opt_name[i])) { g_opt[i] = value < 0 ? 0 : value; return 0; }
    set_error("set_option: unknown option '%s'", name);
    return MNB_ERR_ARG;
}
int mnb_get_option(const char* name) {
    for (int i = 0; i < OPT_COUNT; ++i)
        if (!strcmp(name, g_opt_name[i])) return option_get(i);
    set_error("get_option: unknown option '%s'", name);
    return MNB_ERR_ARG;
}
const char* mnb_last_error(void) { return g_err; }
int mnb_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    return major == 10;
}

int mnb_bn_finalize_fb(const double* stats, const float* gamma, const float* beta, const float* conv_bias,
                       float* running_mean, float* running_var, long long* nbt, float* scale, float* shift,
                       float* save_mean, float* save_invstd, int C, double m, float eps, float momentum, void* stream) {
    MNB_REQUIRE(C > 0 && m > 0, "bn_finalize: bad C/m");
    bn_finalize_k<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, gamma, beta, conv_bias, running_mean,
                                                                      running_var, nbt, scale, shift, save_mean,
                                                                      save_invstd, C, m, eps, momentum);
    MNB_LAUNCH_CHECK("bn_finalize");
    return 0;
}

int mnb_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, long long* nbt, float* scale, float* shift, float* save_mean,
                    float* save_invstd, int C, double m, float eps, float momentum, void* stream) {
    return mnb_bn_finalize_fb(stats, gamma, beta, nullptr, running_mean, running_var, nbt, scale, shift, save_mean,
                              save_invstd, C, m, eps, momentum, stream);
}

int mnb_bn_eval_coeffs_fb(const float* gamma, const float* beta, const float* conv_bias, const float* rm,
                          const float* rv, float* scale, float* shift, int C, float eps, void* stream) {
    MNB_REQUIRE(C > 0, "bn_eval_coeffs: bad C");
    bn_eval_coeffs_k<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, conv_bias, rm, rv, scale, shift, C,
                                                                         eps);
    MNB_LAUNCH_CHECK("bn_eval_coeffs");
    return 0;
}

int mnb_bn_eval_coeffs(const float* gamma, const float* beta, const float* rm, const float* rv, float* scale,
                       float* shift, int C, float eps, void* stream) {
    return mnb_bn_eval_coeffs_fb(gamma, beta, nullptr, rm, rv, scale, shift, C, eps, stream);
}

int mnb_bn_bwd_apply_fused(const void* dA, const void* z, const float* scale, const float* shift, const double* sums,
                           const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, float* dbias,
                           void* dZ, long long M, int C, double m, int dtype, void* stream) {
    MNB_REQUIRE(C % 8 == 0 && M > 0 && m > 0, "bn_bwd_apply_fused: bad shape");
    ColGeom g = col_geom(M, C, 16, 16);     // >= 16 rows per thread: each thread derives its channels' coefficients first
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MNB_F32) bn_bwd_apply_fused_k<float><<<g.grid, g.block, 0, st>>>((const float*)dA, (const float*)z, scale, shift, sums, save_mean, save_invstd, dgamma, dbeta, dbias, (float*)dZ, M, C, m);
    else if (dtype == MNB_BF16) bn_bwd_apply_fused_k<bf16><<<g.grid, g.block, 0, st>>>((const bf16*)dA, (const bf16*)z, scale, shift, sums, save_mean, save_invstd, dgamma, dbeta, dbias, (bf16*)dZ, M, C, m);
    else MNB_REQUIRE(false, "bn_bwd_apply_fused: bad dtype");
    MNB_LAUNCH_CHECK("bn_bwd_apply_fused");
    return 0;
}

int mnb_bn_relu_apply(const void* z, const float* scale, const float* shift, const void* residual, void* y,
                      long long M, int C, int dtype, void* stream) {
    MNB_REQUIRE(C % 8 == 0 && M > 0, "bn_relu_apply: C %% 8 != 0 or M <= 0");
    ColGeom g = col_geom(M, C, 16, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MNB_F32) {
        if (residual) bn_relu_apply_k<float, true><<<g.grid, g.block, 0, st>>>((const float*)z, scale, shift, (const float*)residual, (float*)y, M, C);
        else bn_relu_apply_k<float, false><<<g.grid, g.block, 0, st>>>((const float*)z, scale, shift, nullptr, (float*)y, M, C);
    } else if (dtype == MNB_BF16) {
        if (residual) bn_relu_apply_k<bf16, true><<<g.grid, g.block, 0, st>>>((const bf16*)z, scale, shift, (const bf16*)residual, (bf16*)y, M, C);
        else bn_relu_apply_k<bf16, false><<<g.grid, g.block, 0, st>>>((const bf16*)z, scale, shift, nullptr, (bf16*)y, M, C);
    } else MNB_REQUIRE(false, "bn_relu_apply: bad dtype");
    MNB_LAUNCH_CHECK("bn_relu_apply");
    return 0;
}

int mnb_bn_bwd_reduce(const void* dA, const void* z, const float* scale, const float* shift, double* sums,
                      long long M, int C, int dtype, void* stream) {
    MNB_REQUIRE(C % 8 == 0 && M > 0, "bn_bwd_reduce: C %% 8 != 0 or M <= 0");
    // few, fat CTAs: every CTA ends with one fp64 atomic per channel and same-address atomics serialise in L2
    // Every CTA ends with one fp64 atomic per channel, and same-address atomics serialise in L2 (~8 us per CTA-per-SM
    // on B200, scripts/exp_bn.py): 3 CTAs per SM for the >= 200 MB passes, fewer for the small tensors.
    int ctas = option_get(OPT_BN_CTAS);
    if (ctas <= 0) {
        const double bytes = 2.0 * (double)M * C * (dtype == MNB_F32 ? 4 : 2);
        ctas = (int)(bytes / (num_sms() * 500.0e3) + 0.5);
        ctas = ctas < 1 ? 1 : (ctas > 3 ? 3 : ctas);
    }
    ColGeom g = col_geom(M, C, ctas);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MNB_F32) bn_bwd_reduce_k<float><<<g.grid, g.block, 0, st>>>((const float*)dA, (const float*)z, scale, shift, sums, M, C);
    else if (dtype == MNB_BF16) bn_bwd_reduce_k<bf16><<<g.grid, g.block, 0, st>>>((const bf16*)dA, (const bf16*)z, scale, shift, sums, M, C);
    else MNB_REQUIRE(false, "bn_bwd_reduce: bad dtype");
    MNB_LAUNCH_CHECK("bn_bwd_reduce");
    return 0;
}

int mnb_bn_bwd_finalize(const double* sums, const float* scale, const float* save_mean, const float* save_invstd,
                        float* dgamma, float* dbeta, float* dbias, float* coef, int C, double m, void* stream) {
    MNB_REQUIRE(C > 0 && m > 0, "bn_bwd_finalize: bad C/m");
    bn_bwd_finalize_k<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, scale, save_mean, save_invstd, dgamma,
                                                                          dbeta, dbias, coef, C, m);
    MNB_LAUNCH_CHECK("bn_bwd_finalize");
    return 0;
}

int mnb_bn_bwd_apply(const void* dA, const void* z, const float* scale, const float* shift, const float* coef,
                     void* dZ, long long M, int C, int dtype, void* stream) {
    MNB_REQUIRE(C % 8 == 0 && M > 0, "bn_bwd_apply: C %% 8 != 0 or M <= 0");
    ColGeom g = col_geom(M, C, 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MNB_F32) bn_bwd_apply_k<float><<<g.grid, g.block, 0, st>>>((const float*)dA, (const float*)z, scale, shift, coef, (float*)dZ, M, C);
    else if (dtype == MNB_BF16) bn_bwd_apply_k<bf16><<<g.grid, g.block, 0, st>>>((const bf16*)dA, (const bf16*)z, scale, shift, coef, (bf16*)dZ, M, C);
    else MNB_REQUIRE(false, "bn_bwd_apply: bad dtype");
    MNB_LAUNCH_CHECK("bn_bwd_apply");
    return 0;
}

int mnb_gap_fwd(const void* z, const float* scale, const float* shift, float* f, int N, int HW, int C, int dtype,
                void* stream) {
    MNB_REQUIRE(C % 8 == 0 && N > 0 && HW > 0, "gap_fwd: bad shape");
    int total = N * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MNB_F32) gap_fwd_k<float><<<(total + 127) / 128, 128, 0, st>>>((const float*)z, scale, shift, f, nullptr, 1.f, nullptr, N, HW, C / 8);
    else if (dtype == MNB_BF16) gap_fwd_k<bf16><<<(total + 127) / 128, 128, 0, st>>>((const bf16*)z, scale, shift, f, nullptr, 1.f, nullptr, N, HW, C / 8);
    else MNB_REQUIRE(false, "gap_fwd: bad dtype");
    MNB_LAUNCH_CHECK("gap_fwd");
    return 0;
}

int mnb_gap_fc_prep(const void* z, const float* scale, const float* shift, float* f, const unsigned char* mask,
                    float mask_scale, void* xb, int N, int HW, int C, int dtype, void* stream) {
    MNB_REQUIRE(C % 8 == 0 && N > 0 && HW > 0 && xb, "gap_fc_prep: bad shape / NULL operand");
    int total = N * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MNB_F32) gap_fwd_k<float><<<(total + 127) / 128, 128, 0, st>>>((const float*)z, scale, shift, f, mask, mask_scale, (bf16*)xb, N, HW, C / 8);
    else if (dtype == MNB_BF16) gap_fwd_k<bf16><<<(total + 127) / 128, 128, 0, st>>>((const bf16*)z, scale, shift, f, mask, mask_scale, (bf16*)xb, N, HW, C / 8);
    else MNB_REQUIRE(false, "gap_fc_prep: bad dtype");
    MNB_LAUNCH_CHECK("gap_fc_prep");
    return 0;
}

int mnb_gap_bwd(const float* df, void* dA, int N, int HW, int C, int dtype, void* stream) {
    MNB_REQUIRE(C % 8 == 0 && N > 0 && HW > 0, "gap_bwd: bad shape");
    long long nvec = (long long)N * HW * (C / 8);
    unsigned g = flat_grid(nvec);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MNB_F32) gap_bwd_k<float><<<g, 256, 0, st>>>(df, (float*)dA, N, HW, C / 8);
    else if (dtype == MNB_BF16) gap_bwd_k<bf16><<<g, 256, 0, st>>>(df, (bf16*)dA, N, HW, C / 8);
    else MNB_REQUIRE(false, "gap_bwd: bad dtype");
    MNB_LAUNCH_CHECK("gap_bwd");
    return 0;
}

int mnb_dropout_mask(unsigned char* mask, long long n, float p, unsigned long long seed, unsigned long long offset,
                     const long long* dev_step, void* stream) {
    MNB_REQUIRE(n > 0 && p >= 0.f && p < 1.f, "dropout_mask: bad n/p");
    dropout_mask_k<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(mask, n, p, seed, offset, dev_step);
    MNB_LAUNCH_CHECK("dropout_mask");
    return 0;
}

int mnb_fc_fwd(const float* x, const unsigned char* mask, float mask_scale, const float* w, const float* b, float* y,
               int relu_out, int N, int K, int O, void* stream) {
    MNB_REQUIRE(N > 0 && K > 0 && O > 0, "fc_fwd: bad shape");
    XMasked af{x, mask, mask_scale, K};
    RowMajor bf_{w, K};
    FcFwdEpi ef{y, b, O, relu_out};
    dim3 grid((O + 63) / 64, (N + 63) / 64);
    small_gemm_k<<<grid, 256, 0, (cudaStream_t)stream>>>(af, bf_, ef, N, O, K);
    MNB_LAUNCH_CHECK("fc_fwd");
    return 0;
}

int mnb_fc_dgrad(const float* dy, const float* w, const unsigned char* mask, float mask_scale, const float* relu_ref,
                 float* dx, int N, int K, int O, void* stream) {
    MNB_REQUIRE(N > 0 && K > 0 && O > 0, "fc_dgrad: bad shape");
    RowMajor af{dy, O};        // A(n, o)
    ColMajor bf_{w, K};        // B(k, o) = w[o*K + k]
    FcDgradEpi ef{dx, mask, mask_scale, relu_ref, K};
    dim3 grid((K + 63) / 64, (N + 63) / 64);
    small_gemm_k<<<grid, 256, 0, (cudaStream_t)stream>>>(af, bf_, ef, N, K, O);
    MNB_LAUNCH_CHECK("fc_dgrad");
    return 0;
}

int mnb_fc_wgrad(const float* x, const unsigned char* mask, float mask_scale, const float* dy, float* dw, float* db,
                 int N, int K, int O, void* stream) {
    MNB_REQUIRE(N > 0 && K > 0 && O > 0, "fc_wgrad: bad shape");
    ColMajor af{dy, O};                      // A(o, n) = dy[n*O + o]
    XMaskedT bf_{x, mask, mask_scale, K};    // B(k, n) = xm[n*K + k]
    AccumEpi ef{dw, K};
    dim3 grid((K + 63) / 64, (O + 63) / 64);
    small_gemm_k<<<grid, 256, 0, (cudaStream_t)stream>>>(af, bf_, ef, O, K, N);
    MNB_LAUNCH_CHECK("fc_wgrad");
    if (db) {
        colsum_accum_k<<<(O + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dy, db, N, O);
        MNB_LAUNCH_CHECK("fc_wgrad(bias)");
    }
    return 0;
}

int mnb_fc_prep_bf16(const float* x, const unsigned char* mask, float mask_scale, void* xb, long long n, void* stream) {
    MNB_REQUIRE(n > 0, "fc_prep_bf16: empty");
    fc_prep_bf16_k<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(x, mask, mask_scale, (bf16*)xb, n);
    MNB_LAUNCH_CHECK("fc_prep_bf16");
    return 0;
}

int mnb_fc_gate(float* dx, const unsigned char* mask, float mask_scale, const float* relu_ref, long long n,
                void* stream) {
    MNB_REQUIRE(n > 0, "fc_gate: empty");
    fc_gate_k<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(dx, mask, mask_scale, relu_ref, n);
    MNB_LAUNCH_CHECK("fc_gate");
    return 0;
}

int mnb_fc_bias_grad(const float* dy, float* db, int N, int O, void* stream) {
    MNB_REQUIRE(N > 0 && O > 0, "fc_bias_grad: bad shape");
    colsum_accum_k<<<(O + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dy, db, N, O);
    MNB_LAUNCH_CHECK("fc_bias_grad");
    return 0;
}

int mnb_xent_fwd_bwd(const float* logits, const long long* target, float* loss, float* dlogits, int N, int O,
                     float grad_scale, void* stream) {
    MNB_REQUIRE(N > 0 && O > 0, "xent: bad shape");
    xent_k<<<N, 128, 0, (cudaStream_t)stream>>>(logits, target, loss, dlogits, N, O, grad_scale);
    MNB_LAUNCH_CHECK("xent");
    return 0;
}

int mnb_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, int step, float grad_scale, const float* dev_lr, const long long* dev_step,
                  void* stream) {
    MNB_REQUIRE(n > 0 && (step >= 1 || dev_step), "adam: bad n/step");
    adam_k<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, step, grad_scale,
                                                           dev_lr, dev_step);
    MNB_LAUNCH_CHECK("adam");
    return 0;
}

int mnb_sgd_step(float* p, const float* g, long long n, float lr, float grad_scale, const float* dev_lr, void* stream) {
    MNB_REQUIRE(n > 0, "sgd: bad n");
    sgd_k<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, n, lr, grad_scale, dev_lr);
    MNB_LAUNCH_CHECK("sgd");
    return 0;
}

int mnb_rmsprop_step(float* p, const float* g, float* square_avg, long long n, float lr, float alpha, float eps,
                     float grad_scale, const float* dev_lr, void* stream) {
    MNB_REQUIRE(n > 0, "rmsprop: bad n");
    rmsprop_k<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, square_avg, n, lr, alpha, eps, grad_scale, dev_lr);
    MNB_LAUNCH_CHECK("rmsprop");
    return 0;
}

__global__ void counter_inc_k(long long* c) { *c += 1; }
int mnb_counter_inc(long long* counter, void* stream) {
    counter_inc_k<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
    MNB_LAUNCH_CHECK("counter_inc");
    return 0;
}

int mnb_u8hwc_to_nchw_f32(const unsigned char* x, const float* mean, const float* stdv, float* y, int N, int H, int W,
                          int C, void* stream) {
    const long long total = (long long)N * H * W * C;
    MNB_REQUIRE(total > 0 && x && mean && stdv && y, "u8hwc_to_nchw: empty or NULL argument");
    u8hwc_to_nchw_norm_k<<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>(x, mean, stdv, y, N, H * W, C);
    MNB_LAUNCH_CHECK("u8hwc_to_nchw");
    return 0;
}

int mnb_nhwc_to_nchw_f32(const void* x, float* y, int N, int H, int W, int C, int dtype, void* stream) {
    long long total = (long long)N * H * W * C;
    MNB_REQUIRE(total > 0, "nhwc_to_nchw: empty");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MNB_F32) nhwc_to_nchw_k<float><<<flat_grid(total), 256, 0, st>>>((const float*)x, y, N, H * W, C);
    else if (dtype == MNB_BF16) nhwc_to_nchw_k<bf16><<<flat_grid(total), 256, 0, st>>>((const bf16*)x, y, N, H * W, C);
    else MNB_REQUIRE(false, "nhwc_to_nchw: bad dtype");
    MNB_LAUNCH_CHECK("nhwc_to_nchw");
    return 0;
}

int mnb_nchw_f32_to_nhwc(const float* x, void* y, int N, int H, int W, int C, int dtype, void* stream) {
    long long total = (long long)N * H * W * C;
    MNB_REQUIRE(total > 0, "nchw_to_nhwc: empty");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MNB_F32) nchw_to_nhwc_k<float><<<flat_grid(total), 256, 0, st>>>(x, (float*)y, N, H * W, C);
    else if (dtype == MNB_BF16) nchw_to_nhwc_k<bf16><<<flat_grid(total), 256, 0, st>>>(x, (bf16*)y, N, H * W, C);
    else MNB_REQUIRE(false, "nchw_to_nhwc: bad dtype");
    MNB_LAUNCH_CHECK("nchw_to_nhwc");
    return 0;
}

}  // extern "C"
