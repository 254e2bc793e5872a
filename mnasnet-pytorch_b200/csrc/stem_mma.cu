// Stem convolution forward on the tensor pipe, bf16 output: 3x3, stride 2, pad 1, Cin = 3 -> Cout = 32 (features.0's
// nn.Conv2d, src/models/mnasnet.py:179), network input either N x 3 x H x W fp32 (src/train.py:427) or N x H x W x 3 uint8
// with ToTensor + Normalize fused through the same 3 x 256 table as stem.cu (bit-identical normalised values).
//
// The CUDA-core kernel (stem.cu: thread == output pixel, 27 x 32 FMAs + 27 scattered byte / strided fp32 loads) takes
// 291 us at batch 256 against a ~45 us HBM floor (38 / 154 MB in, 205 MB out).  Here a CTA owns TH full output rows of
// one image: the 2*TH+1 input rows are contiguous in global memory (one 1-D bulk copy for uint8 NHWC, one per colour
// plane for fp32 NCHW), each thread builds ONE row of the im2col tile in shared memory (27 normalised values rounded to
// bf16, k = ci*9 + kh*3 + kw as in the torch weight layout, padded to 32), and the GEMM [pixels][32] x [32][32] runs as
// mma.sync.m16n8k16 with the weights in registers; the output tile is contiguous in global memory and leaves through
// one bulk store; BN statistics of the stored (bf16) values from the staged tile.  Operands are rounded to bf16 like
// every tensor-pipe operand of the bf16 path (the backward-weight kernel stem_wgrad_mma_k rounds them the same way).
#include <algorithm>

#include "dw_mma.cuh"

namespace mnb {

struct StemP {
    const void* x;
    const float* mean;
    const float* stdv;
    const float* w;             // [32][3][3][3] fp32
    const float* bias;
    bf16* z;
    double* stats;
    int N, H, W, Ho, Wo;
    int TH, PT, nblk, items;    // tile: TH output rows x Wo columns = PT pixels; nblk tiles per image
    int BH;                     // input rows per tile (2 TH + 1)
    int xb_bytes;               // bytes of one input-tile buffer (128-aligned)
    uint32_t magic_wo;
};

constexpr int STM_THREADS = 224, STM_WARPS = 7, STM_MTW = 2, STM_AP = 80, STM_OP = 64, STM_COUT = 32;

__host__ __device__ constexpr int stm_al128(int b) { return (b + 127) / 128 * 128; }

template <bool U8>
__global__ void __launch_bounds__(STM_THREADS, 3) stem_fwd_mma_k(const StemP p) {
    constexpr int THREADS = STM_THREADS, WARPS = STM_WARPS, MTW = STM_MTW, AP = STM_AP, OP = STM_OP, COUT = STM_COUT;
    extern __shared__ __align__(128) unsigned char dsm[];
    // [weights 32 x 80 B][table 3 x 256 fp32][bias 32][statistics 64][2 mbarriers + pad][2 x input tile][im2col tile][output tile]
    const uint32_t WS = smem_u32(dsm);
    float* lut = reinterpret_cast<float*>(dsm + COUT * AP);
    float* s_bias = lut + 3 * 256;
    float* red = s_bias + COUT;
    const uint32_t bar0 = WS + COUT * AP + (3 * 256 + COUT + 2 * COUT) * 4;
    constexpr int HEAD = stm_al128(COUT * AP + (3 * 256 + COUT + 2 * COUT) * 4 + 16);
    unsigned char* xb0 = dsm + HEAD;
    const uint32_t XB0 = WS + HEAD;
    const uint32_t AS = XB0 + 2 * p.xb_bytes;
    const uint32_t OUT = AS + stm_al128(16 * MTW * WARPS * AP);      // the MMAs read all 14 m-tiles of the im2col tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, mi = lane >> 3, r8 = lane & 7;

    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    if (U8)
        for (int i = tid; i < 3 * 256; i += THREADS) {
            const int c = i >> 8;
            const float v = (float)(i & 255) / 255.f;
            lut[i] = (v - p.mean[c]) / p.stdv[c];
        }
    for (int i = tid; i < COUT; i += THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    for (int i = tid; i < 2 * COUT; i += THREADS) red[i] = 0.f;
    for (int i = tid; i < COUT * 32; i += THREADS) {        // [co][k = ci*9 + kh*3 + kw] bf16, k 27..31 zero
        const int co = i >> 5, k = i & 31;
        *reinterpret_cast<bf16*>(dsm + co * AP + k * 2) = __float2bfloat16_rn(k < 27 ? p.w[co * 27 + k] : 0.f);
    }
    __syncthreads();
    // weight fragments: 4 n-tiles x 2 k16 steps
    uint32_t bfr[4][2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int jp = 0; jp < 2; ++jp)
            ldsm4(WS + (uint32_t)(((2 * jp + (mi >> 1)) * 8 + r8) * AP + (mi & 1) * 16 + ks * 32), bfr[2 * jp][ks][0], bfr[2 * jp][ks][1],
                  bfr[2 * jp + 1][ks][0], bfr[2 * jp + 1][ks][1]);
    uint32_t aoff[MTW];
    int prow[MTW][2];
#pragma unroll
    for (int i = 0; i < MTW; ++i) {
        const int mt = i * WARPS + warp;
        aoff[i] = AS + (uint32_t)((mt * 16 + (mi & 1) * 8 + r8) * AP + (mi >> 1) * 16);
        prow[i][0] = mt * 16 + g;
        prow[i][1] = mt * 16 + g + 8;
    }
    // this thread's pixel of the im2col tile
    const bool plive = tid < p.PT;
    const int pty = (int)fastdiv((uint32_t)(plive ? tid : 0), p.magic_wo), ptx = (plive ? tid : 0) - pty * p.Wo;
    // statistics pass: thread = (channel pair, pixel phase)
    constexpr int CP = COUT / 2, PSTEP = THREADS / CP;
    const int scp = tid % CP, sp0 = tid / CP;
    float ssum0 = 0.f, ssum1 = 0.f, ssq0 = 0.f, ssq1 = 0.f;
    const int rowb = U8 ? p.W * 3 : p.W * 4;              // bytes of one input row (uint8 NHWC: all channels; fp32: one plane)

    auto issue = [&](int item, int b) {
        const int n = item / p.nblk, blk = item - n * p.nblk;
        const int iy0 = 2 * blk * p.TH - 1;
        const int r_lo = max(0, -iy0), r_hi = min(p.BH, p.H - iy0);
        const uint32_t bytes = (uint32_t)((r_hi - r_lo) * rowb);
        const uint32_t dst = XB0 + b * p.xb_bytes;
        if (U8) {
            mbar_expect_tx(bar0 + 8 * b, bytes);
            bulk_load1(dst + r_lo * rowb, static_cast<const unsigned char*>(p.x) + ((size_t)n * p.H + iy0 + r_lo) * rowb, bytes, bar0 + 8 * b);
        } else {
            mbar_expect_tx(bar0 + 8 * b, 3 * bytes);
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
                bulk_load1(dst + (ci * p.BH + r_lo) * rowb,
                           static_cast<const unsigned char*>(p.x) + (((size_t)n * 3 + ci) * p.H + iy0 + r_lo) * rowb, bytes, bar0 + 8 * b);
        }
    };
    int item = blockIdx.x, b = 0;
    uint32_t ph = 0;
    if (tid == 0 && item < p.items) issue(item, 0);
    for (; item < p.items; item += gridDim.x, b ^= 1) {
        const int n = item / p.nblk, blk = item - n * p.nblk;
        const int oy0 = blk * p.TH, iy0 = 2 * oy0 - 1;
        const int rows = min(p.TH, p.Ho - oy0);
        if (tid == 0 && item + (int)gridDim.x < p.items) issue(item + gridDim.x, b ^ 1);
        mbar_wait(bar0 + 8 * b, (ph >> b) & 1);
        ph ^= 1u << b;
        if (plive) {
            // one im2col row: 27 normalised inputs (zero outside the image) -> bf16, k = ci*9 + kh*3 + kw
            const unsigned char* xb = xb0 + b * p.xb_bytes;
            float v[32];
#pragma unroll
            for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int r = 2 * pty + kh;
                const bool rok = (unsigned)(iy0 + r) < (unsigned)p.H;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int iw = 2 * ptx - 1 + kw;
                    const bool ok = rok && (unsigned)iw < (unsigned)p.W;
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci) {
                        float val = 0.f;
                        if (ok) {
                            if (U8) val = lut[ci * 256 + xb[(r * p.W + iw) * 3 + ci]];
                            else val = reinterpret_cast<const float*>(xb)[(ci * p.BH + r) * p.W + iw];
                        }
                        v[ci * 9 + kh * 3 + kw] = val;
                    }
                }
            }
            const uint32_t dst = AS + (uint32_t)(tid * AP);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                sts128(dst + q * 16, make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                                pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
        }
        __syncthreads();
        float acc[MTW][4][4];
#pragma unroll
        for (int i = 0; i < MTW; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t a[MTW][4];
#pragma unroll
            for (int i = 0; i < MTW; ++i) ldsm4(aoff[i] + ks * 32, a[i][0], a[i][1], a[i][2], a[i][3]);
#pragma unroll
            for (int i = 0; i < MTW; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma16816(acc[i][j], a[i][0], a[i][1], a[i][2], a[i][3], bfr[j][ks][0], bfr[j][ks][1]);
        }
        if (tid == 0) tma_store_wait_read();            // the previous tile's store has finished reading OUT
        __syncthreads();
#pragma unroll
        for (int i = 0; i < MTW; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (prow[i][h] < p.PT) {
                    const uint32_t o = OUT + (uint32_t)(prow[i][h] * OP + t * 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        sts32(o + j * 16, pack_bf16x2(acc[i][j][2 * h] + s_bias[j * 8 + 2 * t], acc[i][j][2 * h + 1] + s_bias[j * 8 + 2 * t + 1]));
                }
        fence_proxy_async();
        __syncthreads();
        const int npix = rows * p.Wo;
        if (tid == 0) bulk_store1(p.z + ((size_t)(n * p.Ho + oy0) * p.Wo) * COUT, OUT, (uint32_t)(npix * OP));
        if (p.stats != nullptr) {
            for (int px = sp0; px < npix; px += PSTEP) {
                const uint32_t u = lds32(OUT + (uint32_t)(px * OP + scp * 4));
                const float v0 = bf_lo(u), v1 = bf_hi(u);
                ssum0 += v0; ssum1 += v1;
                ssq0 = fmaf(v0, v0, ssq0); ssq1 = fmaf(v1, v1, ssq1);
            }
        }
    }
    if (p.stats != nullptr) {
        atomicAdd(&red[2 * scp], ssum0); atomicAdd(&red[2 * scp + 1], ssum1);
        atomicAdd(&red[COUT + 2 * scp], ssq0); atomicAdd(&red[COUT + 2 * scp + 1], ssq1);
        __syncthreads();
        for (int i = tid; i < 2 * COUT; i += THREADS) atomicAdd(&p.stats[i], (double)red[i]);
    }
    if (tid == 0) tma_store_wait_read();
}

// MNB_ERR_UNSUPPORTED when the geometry does not allow whole-row bulk copies (the caller falls back to stem.cu)
int stem_fwd_mma(const void* x, int x_u8, const float* mean, const float* stdv, const float* w, const float* bias, void* z,
                 double* stats, int N, int H, int W, cudaStream_t st) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int rowb = x_u8 ? W * 3 : W * 4;
    if (rowb % 16 != 0 || Wo > 16 * STM_MTW * STM_WARPS || ((uintptr_t)x & 15) || ((uintptr_t)z & 15)) return MNB_ERR_UNSUPPORTED;
    StemP p = {};
    p.x = x; p.mean = mean; p.stdv = stdv; p.w = w; p.bias = bias; p.z = (bf16*)z; p.stats = stats;
    p.N = N; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo;
    const int thmax = std::min(16 * STM_MTW * STM_WARPS / Wo, Ho);
    const int nblk = (Ho + thmax - 1) / thmax;
    p.TH = (Ho + nblk - 1) / nblk;
    p.nblk = (Ho + p.TH - 1) / p.TH;
    p.PT = p.TH * Wo;
    p.BH = 2 * p.TH + 1;
    p.xb_bytes = stm_al128((x_u8 ? 1 : 3) * p.BH * rowb);
    if ((long long)N * p.nblk > (1 << 30) || p.PT >= 65536) return MNB_ERR_UNSUPPORTED;
    p.items = N * p.nblk;
    p.magic_wo = (uint32_t)((1ull << 32) / (unsigned)Wo) + 1u;
    const size_t head = stm_al128(STM_COUT * STM_AP + (3 * 256 + STM_COUT + 2 * STM_COUT) * 4 + 16);
    const size_t smem = head + 2 * p.xb_bytes + stm_al128(16 * STM_MTW * STM_WARPS * STM_AP) + stm_al128(p.PT * STM_OP);
    if (smem > 75 * 1024) return MNB_ERR_UNSUPPORTED;      // three CTAs per SM
    cudaError_t e = x_u8 ? cudaFuncSetAttribute(stem_fwd_mma_k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(stem_fwd_mma_k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("stem_fwd(mma): %s", cudaGetErrorString(e)); return (int)e; }
    const int grid = std::min(p.items, num_sms() * 3);
    if (x_u8) stem_fwd_mma_k<true><<<grid, STM_THREADS, smem, st>>>(p);
    else stem_fwd_mma_k<false><<<grid, STM_THREADS, smem, st>>>(p);
    MNB_LAUNCH_CHECK("stem_fwd(mma)");
    return 0;
}

}  // namespace mnb
