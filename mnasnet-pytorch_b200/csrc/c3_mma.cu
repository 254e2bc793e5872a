// Dense 3x3 (pad 1, stride 1 / 2) convolutions with few channels, bf16 NHWC: forward (+ the producing ConvBlock's
// BN-apply+ReLU on load, + this ConvBlock's BN batch statistics on store), backward-data (stride 2) and backward-weight.
// Replaces nn.Conv2d(k=3) inside ConvBlock at src/models/mnasnet.py:130-161 (the stride-2 stage transitions 16->24,
// 24->40, 40->80) for the bf16 path; called through mnb_conv_fwd / mnb_conv_dgrad / mnb_conv_wgrad (conv_simt.cu).
//
// These layers move 25..140 MB and do < 6 GFLOP each: HBM-bound by two orders of magnitude.  The generic tcgen05
// pipeline (gemm_tc.cu) gathers every 16-byte im2col vector with its own index arithmetic and is bound by instruction
// issue (177 / 404 / 238 us for fwd / dgrad / wgrad of 16->24 at 112x112 against a 26 us HBM floor).  Here the im2col is
// free: a CTA owns TH full output rows of one image; the (TH-1)*S+3 full input rows it needs are CONTIGUOUS in global
// memory, so ONE 1-D bulk copy (cp.async.bulk, mbarrier completion) brings them into shared memory exactly as they lie
// in HBM (no tensor maps; the first version of this file used 4-D TMA boxes with zero-filled halos and ran at the same
// speed, 73 vs 69 us for the 16->24 forward: the copy is not the limit).  relu(scale*x+shift) is applied in place, and
// ldmatrix reads the A fragments of every tap straight from that NHWC tile: row addresses are per lane, so the stride
// between output pixels costs nothing and a tap that falls into the left / right padding simply points at a block of
// zeros (rows above / below the image are zero-filled in shared memory).  mma.sync.m16n8k16 / m16n8k8 with the weights
// resident in shared memory; the output tile is contiguous in global memory too and leaves through one bulk store.
// What bounds it (ncu, 16->24 forward): shared-memory wavefronts -- the 8 rows of a stride-2 ldmatrix at a 32-byte pixel
// pitch fall into two bank groups (4-way conflict, l1tex 75 %); three CTAs per SM: 54 / 46 / 78 us fwd / dgrad / wgrad.
//   FWD    D[m = pixel][n = co]      += A[pixel + tap][k = ci] * W[co][tap][ci]
//   DGRAD  by input-pixel parity class (py, px): D[m = (a, b)][n = ci] += dZ[(a + dy, b + dx)][k = co] * W[co][ci][tap],
//          only the 1 / 2 / 2 / 4 taps with (py + 1 - ky), (px + 1 - kx) even contribute
//   WGRAD  one warp per tap: D[m = co][n = ci] += dZ^T[co][k = pixel] * A[pixel + tap][ci]  (ldmatrix.trans both operands)
#include <algorithm>

#include "conv_params.cuh"
#include "dw_mma.cuh"

namespace mnb {

struct C3P {
    const float* in_scale;      // BN scale / shift of the producing ConvBlock (NULL = plain input)
    const float* in_shift;
    const float* w;             // [Cout][Cin][3][3] fp32
    const bf16* wpk;            // optional bf16 packing: fwd [Cout][tap][Cin], dgrad [Cin][tap][Cout]
    double* stats;              // fwd: [2][Cout]
    const bf16* src;            // the tensor whose rows are tiled: fwd / wgrad x [N][H][W][Cin], dgrad dz [N][Ho][Wo][Cout]
    bf16* out;                  // fwd: z; dgrad: dx
    const bf16* dz;             // wgrad
    float* dw;                  // wgrad
    int N, H, W, Ho, Wo;
    int TH, TWO, PT, nblk, items;       // tile: TH output rows x TWO (= Wo) columns = PT pixels; nblk tiles per image
    int BH;                     // source rows per tile
    int RP;                     // bytes of one source row (shared memory == global memory layout)
    int xb_bytes;               // bytes of one source-tile buffer (128-aligned)
    uint32_t magic_two;
};

template <int CIN, int COUT>
struct C3Cfg {
    static constexpr int WARPS = 7, THREADS = 32 * WARPS, MTW = 2, PTMAX = 16 * MTW * WARPS;
    static constexpr int NT = COUT / 8, NCH8 = CIN / 8;
    static constexpr int PITCH = CIN * 2, OPITCH = COUT * 2;
    static constexpr int WP = c3_odd16(9 * CIN * 2);                 // bytes per weight row (one co, K = (tap, ci))
    static constexpr int W_BYTES = c3_al128(COUT * WP);
    static constexpr int ZB_BYTES = c3_al128(PITCH);
    static constexpr int NSTAT = THREADS / (COUT / 2) * (COUT / 2);
    static constexpr int MINB = (MTW * NT * 4 <= 48) ? 2 : 1;
};

// thread 0: bulk copy of the in-range rows of a source tile (rows y0 .. y0+BH-1 of one image with HS rows)
__device__ __forceinline__ void c3_issue_rows(uint32_t dst, const bf16* img, int y0, int BH, int HS, int RP, uint32_t bar,
                                              uint32_t extra_tx) {
    const int r_lo = max(0, -y0), r_hi = min(BH, HS - y0);
    const uint32_t bytes = (uint32_t)((r_hi - r_lo) * RP);
    mbar_expect_tx(bar, bytes + extra_tx);
    bulk_load1(dst + (uint32_t)(r_lo * RP), reinterpret_cast<const unsigned char*>(img) + (size_t)(y0 + r_lo) * RP, bytes, bar);
}

// zero the rows of a source tile that lie outside the image
template <int THREADS>
__device__ __forceinline__ void c3_zero_rows(uint32_t XB, int y0, int BH, int HS, int RP, int tid) {
    const int r_lo = max(0, -y0), r_hi = min(BH, HS - y0), rpv = RP >> 4;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int v = tid; v < r_lo * rpv; v += THREADS) sts128(XB + (uint32_t)v * 16, z);
    for (int v = r_hi * rpv + tid; v < BH * rpv; v += THREADS) sts128(XB + (uint32_t)v * 16, z);
}

// relu(scale * x + shift) in place on the in-image rows of a source tile
template <int CIN, int THREADS>
__device__ __forceinline__ void c3_transform(uint32_t XB, const float* s_sc, int y0, int BH, int HS, int RP, int tid) {
    constexpr int NCH8 = CIN / 8;
    const int r_lo = max(0, -y0), r_hi = min(BH, HS - y0), rpv = RP >> 4;
    for (int v = r_lo * rpv + tid; v < r_hi * rpv; v += THREADS) {
        const int ch = v % NCH8;
        const uint32_t a = XB + (uint32_t)v * 16;
        uint4 u = lds128(a);
        const float4 s0 = *reinterpret_cast<const float4*>(s_sc + ch * 8), s1 = *reinterpret_cast<const float4*>(s_sc + ch * 8 + 4);
        const float4 t0 = *reinterpret_cast<const float4*>(s_sc + CIN + ch * 8), t1 = *reinterpret_cast<const float4*>(s_sc + CIN + ch * 8 + 4);
        u.x = pack_bf16x2(fmaxf(fmaf(bf_lo(u.x), s0.x, t0.x), 0.f), fmaxf(fmaf(bf_hi(u.x), s0.y, t0.y), 0.f));
        u.y = pack_bf16x2(fmaxf(fmaf(bf_lo(u.y), s0.z, t0.z), 0.f), fmaxf(fmaf(bf_hi(u.y), s0.w, t0.w), 0.f));
        u.z = pack_bf16x2(fmaxf(fmaf(bf_lo(u.z), s1.x, t1.x), 0.f), fmaxf(fmaf(bf_hi(u.z), s1.y, t1.y), 0.f));
        u.w = pack_bf16x2(fmaxf(fmaf(bf_lo(u.w), s1.z, t1.z), 0.f), fmaxf(fmaf(bf_hi(u.w), s1.w, t1.w), 0.f));
        sts128(a, u);
    }
}

// weights -> shared memory rows [NR][pitch WP] with K = (tap, c) contiguous.  FWD: row = co, c = ci; DGRAD: row = ci, c = co.
template <int NR, int NC, int WP, bool DGRAD, int THREADS>
__device__ __forceinline__ void c3_load_weights(unsigned char* ws, const C3P& p, int tid) {
    if (p.wpk) {
        constexpr int VPR = 9 * NC / 8;           // 16-byte vectors per row
        for (int i = tid; i < NR * VPR; i += THREADS) {
            const int r = i / VPR, v = i - r * VPR;
            *reinterpret_cast<uint4*>(ws + r * WP + v * 16) = *reinterpret_cast<const uint4*>(p.wpk + (size_t)r * 9 * NC + v * 8);
        }
    } else {
        for (int i = tid; i < NR * 9 * NC; i += THREADS) {
            const int r = i / (9 * NC), k = i - r * 9 * NC, tap = k / NC, c = k - tap * NC;
            const int co = DGRAD ? c : r, ci = DGRAD ? r : c;
            const int cin = DGRAD ? NR : NC;
            *reinterpret_cast<bf16*>(ws + r * WP + k * 2) = __float2bfloat16_rn(p.w[((size_t)co * cin + ci) * 9 + tap]);
        }
    }
}

// one tap of the implicit GEMM: acc[i][j] += A_i (K = KC channels at abase[i]) * B_j (weight rows at byte offset kb0)
template <int KC, int MTW, int NT, int WP>
__device__ __forceinline__ void c3_tap(float (&acc)[MTW][NT][4], const uint32_t (&abase)[MTW], uint32_t khalf, uint32_t b4,
                                       uint32_t b2, uint32_t b8, int kb0) {
#pragma unroll
    for (int ks = 0; ks < KC / 16; ++ks) {
        uint32_t a[MTW][4], bf[NT][2];
#pragma unroll
        for (int i = 0; i < MTW; ++i) ldsm4(abase[i] + khalf + ks * 32, a[i][0], a[i][1], a[i][2], a[i][3]);
        c3_load_b16<NT, WP>(bf, b4, b2, kb0 + ks * 32);
#pragma unroll
        for (int i = 0; i < MTW; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) mma16816(acc[i][j], a[i][0], a[i][1], a[i][2], a[i][3], bf[j][0], bf[j][1]);
    }
    if constexpr (KC % 16 != 0) {
        uint32_t a[MTW][2], bf[NT][2];
#pragma unroll
        for (int i = 0; i < MTW; ++i) ldsm2(abase[i] + (KC / 16) * 32, a[i][0], a[i][1]);
        c3_load_b8<NT, WP>(bf, b8, kb0 + (KC / 16) * 32);
#pragma unroll
        for (int i = 0; i < MTW; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) mma1688(acc[i][j], a[i][0], a[i][1], bf[j][0]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
template <int CIN, int COUT, int S>
__global__ void __launch_bounds__(C3Cfg<CIN, COUT>::THREADS, C3Cfg<CIN, COUT>::MINB) c3_fwd_k(const C3P p) {
    using Cfg = C3Cfg<CIN, COUT>;
    constexpr int THREADS = Cfg::THREADS, WARPS = Cfg::WARPS, MTW = Cfg::MTW, NT = Cfg::NT, PITCH = Cfg::PITCH,
                  OPITCH = Cfg::OPITCH, WP = Cfg::WP;
    extern __shared__ __align__(128) unsigned char dsm[];
    // [weights][zero block][2 x source tile][output tile][scale, shift][statistics][2 mbarriers]
    unsigned char* ws = dsm;
    const uint32_t WS = smem_u32(dsm);
    const uint32_t ZB = WS + Cfg::W_BYTES;
    const uint32_t XB0 = ZB + Cfg::ZB_BYTES;
    const int out_bytes = c3_al128(p.PT * OPITCH);
    const uint32_t OUT = XB0 + 2 * p.xb_bytes;
    float* s_sc = reinterpret_cast<float*>(dsm + Cfg::W_BYTES + Cfg::ZB_BYTES + 2 * p.xb_bytes + out_bytes);      // [2][CIN]
    float* red = s_sc + 2 * CIN;                                                                                    // [2][COUT]
    const uint32_t bar0 = OUT + out_bytes + (2 * CIN + 2 * COUT) * 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, mi = lane >> 3, r8 = lane & 7;

    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    const bool xf = p.in_scale != nullptr;
    if (xf)
        for (int i = tid; i < CIN; i += THREADS) { s_sc[i] = p.in_scale[i]; s_sc[CIN + i] = p.in_shift[i]; }
    for (int i = tid; i < 2 * COUT; i += THREADS) red[i] = 0.f;
    for (int i = tid; i < Cfg::ZB_BYTES / 16; i += THREADS) sts128(ZB + i * 16, make_uint4(0, 0, 0, 0));
    c3_load_weights<COUT, CIN, WP, false, THREADS>(ws, p, tid);

    // lane constants: the A rows of this warp's m-tiles (ldmatrix; one offset per tap column, padding -> zero block)
    // and the D rows it owns (staging)
    uint32_t aoff[MTW][3], amask = 0;
    int prow[MTW][2];
#pragma unroll
    for (int i = 0; i < MTW; ++i) {
        const int mt = i * WARPS + warp;
        int pa = mt * 16 + (mi & 1) * 8 + r8;
        pa = pa < p.PT ? pa : p.PT - 1;
        const int ty = (int)fastdiv((uint32_t)pa, p.magic_two), tx = pa - ty * p.TWO;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = tx * S - 1 + kx;
            const bool ok = (unsigned)ix < (unsigned)p.W;
            aoff[i][kx] = ok ? (uint32_t)((ty * S * p.W + ix) * PITCH) : 0u;
            amask |= (ok ? 1u : 0u) << (i * 3 + kx);
        }
        prow[i][0] = mt * 16 + g;
        prow[i][1] = mt * 16 + g + 8;
    }
    const uint32_t khalf = (uint32_t)((mi >> 1) * 16);
    const uint32_t b4 = WS + (uint32_t)(((mi >> 1) * 8 + r8) * WP + (mi & 1) * 16);
    const uint32_t b2 = WS + (uint32_t)(((NT - 1) * 8 + r8) * WP + (mi & 1) * 16);
    const uint32_t b8 = WS + (uint32_t)((mi * 8 + r8) * WP);
    // statistics pass: thread = (channel pair, pixel phase)
    constexpr int CP = COUT / 2, PSTEP = Cfg::NSTAT / CP;
    const int scp = tid % CP, sp0 = tid / CP;
    float ssum0 = 0.f, ssum1 = 0.f, ssq0 = 0.f, ssq1 = 0.f;
    const bool do_stats = p.stats != nullptr && tid < Cfg::NSTAT;
    __syncthreads();

    auto issue = [&](int item, int b) {
        const int n = item / p.nblk, blk = item - n * p.nblk;
        c3_issue_rows(XB0 + b * p.xb_bytes, p.src + (size_t)n * p.H * p.W * CIN, blk * p.TH * S - 1, p.BH, p.H, p.RP, bar0 + 8 * b, 0);
    };
    int item = blockIdx.x, b = 0;
    uint32_t ph = 0;
    if (tid == 0 && item < p.items) issue(item, 0);
    for (; item < p.items; item += gridDim.x, b ^= 1) {
        const int n = item / p.nblk, blk = item - n * p.nblk;
        const int oy0 = blk * p.TH, iy0 = oy0 * S - 1;
        const int rows = min(p.TH, p.Ho - oy0);
        if (tid == 0 && item + (int)gridDim.x < p.items) issue(item + gridDim.x, b ^ 1);
        const uint32_t XB = XB0 + b * p.xb_bytes;
        c3_zero_rows<THREADS>(XB, iy0, p.BH, p.H, p.RP, tid);
        mbar_wait(bar0 + 8 * b, (ph >> b) & 1);
        ph ^= 1u << b;
        if (xf) c3_transform<CIN, THREADS>(XB, s_sc, iy0, p.BH, p.H, p.RP, tid);
        __syncthreads();

        float acc[MTW][NT][4];
#pragma unroll
        for (int i = 0; i < MTW; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap % 3;
            uint32_t abase[MTW];
#pragma unroll
            for (int i = 0; i < MTW; ++i) abase[i] = (amask >> (i * 3 + kx)) & 1u ? XB + (uint32_t)(ky * p.RP) + aoff[i][kx] : ZB;
            c3_tap<CIN, MTW, NT, WP>(acc, abase, khalf, b4, b2, b8, tap * CIN * 2);
        }
        if (tid == 0) tma_store_wait_read();            // the previous tile's store has finished reading OUT
        __syncthreads();
#pragma unroll
        for (int i = 0; i < MTW; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (prow[i][h] < p.PT) {
                    const uint32_t o = OUT + (uint32_t)(prow[i][h] * OPITCH + t * 4);
#pragma unroll
                    for (int j = 0; j < NT; ++j) sts32(o + j * 16, pack_bf16x2(acc[i][j][2 * h], acc[i][j][2 * h + 1]));
                }
        fence_proxy_async();
        __syncthreads();
        const int npix = rows * p.TWO;
        if (tid == 0)
            bulk_store1(p.out + ((size_t)(n * p.Ho + oy0) * p.Wo) * COUT, OUT, (uint32_t)(npix * OPITCH));
        if (do_stats) {
            for (int px = sp0; px < npix; px += PSTEP) {
                const uint32_t u = lds32(OUT + (uint32_t)(px * OPITCH + scp * 4));
                const float v0 = bf_lo(u), v1 = bf_hi(u);
                ssum0 += v0; ssum1 += v1;
                ssq0 = fmaf(v0, v0, ssq0); ssq1 = fmaf(v1, v1, ssq1);
            }
        }
    }
    if (p.stats != nullptr) {
        if (do_stats) {
            atomicAdd(&red[2 * scp], ssum0); atomicAdd(&red[2 * scp + 1], ssum1);
            atomicAdd(&red[COUT + 2 * scp], ssq0); atomicAdd(&red[COUT + 2 * scp + 1], ssq1);
        }
        __syncthreads();
        for (int i = tid; i < 2 * COUT; i += THREADS) atomicAdd(&p.stats[i], (double)red[i]);
    }
    if (tid == 0) tma_store_wait_read();
}

// ---------------------------------------------------------------------------------------------------------------
// backward-data, stride 2: one pass per input-pixel parity class
// ---------------------------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct C3DCfg {
    static constexpr int WARPS = 7, THREADS = 32 * WARPS, MTW = 2;
    static constexpr int NT = CIN / 8;                                  // n = input channels
    static constexpr int ZP = COUT * 2, SP = CIN * 2;                   // dz / dx pixel pitch
    static constexpr int WP = c3_odd16(9 * COUT * 2);                   // bytes per weight row (one ci, K = (tap, co))
    static constexpr int W_BYTES = c3_al128(CIN * WP);
    static constexpr int ZB_BYTES = c3_al128(ZP);
    static constexpr int MINB = (MTW * NT * 4 <= 48) ? 2 : 1;
};

template <int CIN, int COUT, int PY, int PX, class Cfg>
__device__ __forceinline__ void c3_dgrad_class(uint32_t XB, uint32_t ZB, int RP, const uint32_t (&aoff)[Cfg::MTW][2], uint32_t amask,
                                               uint32_t khalf, uint32_t b4, uint32_t b2, uint32_t b8, uint32_t OUT,
                                               const uint32_t (&ooff)[Cfg::MTW][2], const bool (&ook)[Cfg::MTW][2], int W2, int t) {
    constexpr int MTW = Cfg::MTW, NT = Cfg::NT, WP = Cfg::WP, SP = Cfg::SP;
    float acc[MTW][NT][4];
#pragma unroll
    for (int i = 0; i < MTW; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
#pragma unroll
    for (int iy = 0; iy < (PY ? 2 : 1); ++iy)
#pragma unroll
        for (int ix = 0; ix < (PX ? 2 : 1); ++ix) {
            // input row 2a+PY takes tap row ky from output row a + dy: PY = 0 -> (ky 1, dy 0); PY = 1 -> (0, 1), (2, 0)
            const int ky = PY ? 2 * iy : 1, dy = PY ? 1 - iy : 0;
            const int kx = PX ? 2 * ix : 1, dx = PX ? 1 - ix : 0;
            uint32_t abase[MTW];
#pragma unroll
            for (int i = 0; i < MTW; ++i) abase[i] = (amask >> (i * 2 + dx)) & 1u ? XB + (uint32_t)(dy * RP) + aoff[i][dx] : ZB;
            c3_tap<COUT, MTW, NT, WP>(acc, abase, khalf, b4, b2, b8, (ky * 3 + kx) * COUT * 2);
        }
    const uint32_t coff = (uint32_t)((PY * W2 + PX) * SP + t * 4);
#pragma unroll
    for (int i = 0; i < MTW; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (ook[i][h]) {
#pragma unroll
                for (int j = 0; j < NT; ++j)
                    sts32(OUT + ooff[i][h] + coff + j * 16, pack_bf16x2(acc[i][j][2 * h], acc[i][j][2 * h + 1]));
            }
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(C3DCfg<CIN, COUT>::THREADS, C3DCfg<CIN, COUT>::MINB) c3_dgrad_s2_k(const C3P p) {
    using Cfg = C3DCfg<CIN, COUT>;
    constexpr int THREADS = Cfg::THREADS, WARPS = Cfg::WARPS, MTW = Cfg::MTW, NT = Cfg::NT, ZP = Cfg::ZP, SP = Cfg::SP, WP = Cfg::WP;
    extern __shared__ __align__(128) unsigned char dsm[];
    // [weights][zero block][2 x dz tile][dx tile][2 mbarriers]
    unsigned char* ws = dsm;
    const uint32_t WS = smem_u32(dsm);
    const uint32_t ZB = WS + Cfg::W_BYTES;
    const uint32_t XB0 = ZB + Cfg::ZB_BYTES;
    const int W2 = 2 * p.TWO;
    const int out_bytes = c3_al128(2 * p.TH * W2 * SP);
    const uint32_t OUT = XB0 + 2 * p.xb_bytes;
    const uint32_t bar0 = OUT + out_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, mi = lane >> 3, r8 = lane & 7;

    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    for (int i = tid; i < Cfg::ZB_BYTES / 16; i += THREADS) sts128(ZB + i * 16, make_uint4(0, 0, 0, 0));
    c3_load_weights<CIN, COUT, WP, true, THREADS>(ws, p, tid);
    uint32_t aoff[MTW][2], ooff[MTW][2], amask = 0;
    bool ook[MTW][2];
#pragma unroll
    for (int i = 0; i < MTW; ++i) {
        const int mt = i * WARPS + warp;
        int pa = mt * 16 + (mi & 1) * 8 + r8;
        pa = pa < p.PT ? pa : p.PT - 1;
        const int ty = (int)fastdiv((uint32_t)pa, p.magic_two), tx = pa - ty * p.TWO;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const bool ok = tx + dx < p.Wo;
            aoff[i][dx] = ok ? (uint32_t)((ty * p.Wo + tx + dx) * ZP) : 0u;
            amask |= (ok ? 1u : 0u) << (i * 2 + dx);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pr = mt * 16 + g + 8 * h;
            ook[i][h] = pr < p.PT;
            const int prc = ook[i][h] ? pr : 0;
            const int oy = (int)fastdiv((uint32_t)prc, p.magic_two), ox = prc - oy * p.TWO;
            ooff[i][h] = (uint32_t)((2 * oy * W2 + 2 * ox) * SP);
        }
    }
    const uint32_t khalf = (uint32_t)((mi >> 1) * 16);
    const uint32_t b4 = WS + (uint32_t)(((mi >> 1) * 8 + r8) * WP + (mi & 1) * 16);
    const uint32_t b2 = WS + (uint32_t)(((NT - 1) * 8 + r8) * WP + (mi & 1) * 16);
    const uint32_t b8 = WS + (uint32_t)((mi * 8 + r8) * WP);
    __syncthreads();

    auto issue = [&](int item, int b) {
        const int n = item / p.nblk, blk = item - n * p.nblk;
        c3_issue_rows(XB0 + b * p.xb_bytes, p.src + (size_t)n * p.Ho * p.Wo * COUT, blk * p.TH, p.BH, p.Ho, p.RP, bar0 + 8 * b, 0);
    };
    int item = blockIdx.x, b = 0;
    uint32_t ph = 0;
    if (tid == 0 && item < p.items) issue(item, 0);
    for (; item < p.items; item += gridDim.x, b ^= 1) {
        const int n = item / p.nblk, blk = item - n * p.nblk;
        const int iy0 = 2 * blk * p.TH;
        const int rows = min(2 * p.TH, p.H - iy0);
        if (tid == 0) {
            tma_store_wait_read();                      // the previous tile's store has finished reading OUT
            if (item + (int)gridDim.x < p.items) issue(item + gridDim.x, b ^ 1);
        }
        const uint32_t XB = XB0 + b * p.xb_bytes;
        c3_zero_rows<THREADS>(XB, blk * p.TH, p.BH, p.Ho, p.RP, tid);
        mbar_wait(bar0 + 8 * b, (ph >> b) & 1);
        ph ^= 1u << b;
        __syncthreads();
        c3_dgrad_class<CIN, COUT, 0, 0, Cfg>(XB, ZB, p.RP, aoff, amask, khalf, b4, b2, b8, OUT, ooff, ook, W2, t);
        c3_dgrad_class<CIN, COUT, 0, 1, Cfg>(XB, ZB, p.RP, aoff, amask, khalf, b4, b2, b8, OUT, ooff, ook, W2, t);
        c3_dgrad_class<CIN, COUT, 1, 0, Cfg>(XB, ZB, p.RP, aoff, amask, khalf, b4, b2, b8, OUT, ooff, ook, W2, t);
        c3_dgrad_class<CIN, COUT, 1, 1, Cfg>(XB, ZB, p.RP, aoff, amask, khalf, b4, b2, b8, OUT, ooff, ook, W2, t);
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) bulk_store1(p.out + ((size_t)(n * p.H + iy0) * p.W) * CIN, OUT, (uint32_t)(rows * W2 * SP));
    }
    if (tid == 0) tma_store_wait_read();
}

// ---------------------------------------------------------------------------------------------------------------
// backward-weight: one warp per (tap, slice of the output channels)
// ---------------------------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct C3WCfg {
    static constexpr int MT = (COUT + 15) / 16, NTI = CIN / 8;
    static constexpr int MSPLIT = (MT * NTI * 4 > 64) ? 2 : 1;
    static constexpr int MTW = (MT + MSPLIT - 1) / MSPLIT;
    static constexpr int WARPS = 9 * MSPLIT, THREADS = 32 * WARPS;
    static constexpr int PITCH = CIN * 2, ZP = COUT * 2;
    static constexpr int ZB_BYTES = c3_al128(PITCH);
    static constexpr int MINB = THREADS <= 288 ? 2 : 1;
};

template <int CIN, int COUT, int S>
__global__ void __launch_bounds__(C3WCfg<CIN, COUT>::THREADS, C3WCfg<CIN, COUT>::MINB) c3_wgrad_k(const C3P p) {
    using Cfg = C3WCfg<CIN, COUT>;
    constexpr int THREADS = Cfg::THREADS, MTW = Cfg::MTW, NTI = Cfg::NTI, PITCH = Cfg::PITCH, ZP = Cfg::ZP;
    extern __shared__ __align__(128) unsigned char dsm[];
    // [zero block][2 x input tile][2 x dz tile (+ one m-tile of slack)][scale, shift][2 mbarriers]
    const uint32_t ZB = smem_u32(dsm);
    const uint32_t XB0 = ZB + Cfg::ZB_BYTES;
    const int dz_bytes = c3_al128((p.PT + 15) / 16 * 16 * ZP + 32);
    const uint32_t DZ0 = XB0 + 2 * p.xb_bytes;
    float* s_sc = reinterpret_cast<float*>(dsm + Cfg::ZB_BYTES + 2 * p.xb_bytes + 2 * dz_bytes);
    const uint32_t bar0 = DZ0 + 2 * dz_bytes + 2 * CIN * 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, mi = lane >> 3, r8 = lane & 7;
    const int tap = warp % 9, msl = warp / 9;
    const int mt0 = msl * MTW;
    const int ky = tap / 3, kx = tap - ky * 3;

    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    const bool xf = p.in_scale != nullptr;
    if (xf)
        for (int i = tid; i < CIN; i += THREADS) { s_sc[i] = p.in_scale[i]; s_sc[CIN + i] = p.in_shift[i]; }
    for (int i = tid; i < Cfg::ZB_BYTES / 16; i += THREADS) sts128(ZB + i * 16, make_uint4(0, 0, 0, 0));
    float acc[MTW][NTI][4];
#pragma unroll
    for (int i = 0; i < MTW; ++i)
#pragma unroll
        for (int j = 0; j < NTI; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
    const uint32_t nhalf = (uint32_t)((mi >> 1) * 16);
    const uint32_t alane = (uint32_t)(((mi >> 1) * 8 + r8) * ZP + ((mt0 * 2 + (mi & 1)) * 8) * 2);
    __syncthreads();

    auto issue = [&](int item, int b) {
        const int n = item / p.nblk, blk = item - n * p.nblk;
        const int oy0 = blk * p.TH;
        const uint32_t zbytes = (uint32_t)(min(p.TH, p.Ho - oy0) * p.TWO * ZP);
        c3_issue_rows(XB0 + b * p.xb_bytes, p.src + (size_t)n * p.H * p.W * CIN, oy0 * S - 1, p.BH, p.H, p.RP, bar0 + 8 * b, zbytes);
        bulk_load1(DZ0 + b * dz_bytes, p.dz + ((size_t)(n * p.Ho + oy0) * p.Wo) * COUT, zbytes, bar0 + 8 * b);
    };
    int item = blockIdx.x, b = 0;
    uint32_t ph = 0;
    if (tid == 0 && item < p.items) issue(item, 0);
    for (; item < p.items; item += gridDim.x, b ^= 1) {
        const int blk = item % p.nblk;
        const int oy0 = blk * p.TH, iy0 = oy0 * S - 1;
        const int npix = min(p.TH, p.Ho - oy0) * p.TWO;
        const int nk = (npix + 15) >> 4;
        if (tid == 0 && item + (int)gridDim.x < p.items) issue(item + gridDim.x, b ^ 1);
        const uint32_t XB = XB0 + b * p.xb_bytes, DZ = DZ0 + b * dz_bytes;
        // rows npix .. 16 nk - 1 of the dz tile are outside the bulk copy: zero them (they multiply real activations)
        for (int i = tid; i < (nk * 16 - npix) * (ZP / 16); i += THREADS) sts128(DZ + (uint32_t)(npix * ZP + i * 16), make_uint4(0, 0, 0, 0));
        c3_zero_rows<THREADS>(XB, iy0, p.BH, p.H, p.RP, tid);
        mbar_wait(bar0 + 8 * b, (ph >> b) & 1);
        ph ^= 1u << b;
        if (xf) c3_transform<CIN, THREADS>(XB, s_sc, iy0, p.BH, p.H, p.RP, tid);
        __syncthreads();
        // this lane's pixel of the k-step (B operand row): walks 16 pixels per step through the TH x TWO tile
        int pb0 = (mi & 1) * 8 + r8;
        int ty = (int)fastdiv((uint32_t)pb0, p.magic_two), tx = pb0 - ty * p.TWO;
        const int dty = (int)fastdiv(16u, p.magic_two), dtx = 16 - dty * p.TWO;
        for (int ks = 0; ks < nk; ++ks) {
            if (ty >= p.TH) { ty = p.TH - 1; tx = p.TWO - 1; }        // beyond the tile: any valid pixel (its dz rows are zero)
            const int ix = tx * S - 1 + kx;
            const uint32_t baddr = ((unsigned)ix < (unsigned)p.W ? XB + (uint32_t)(((ty * S + ky) * p.W + ix) * PITCH) : ZB);
            uint32_t bf[NTI][2], a[MTW][4];
#pragma unroll
            for (int jp = 0; jp < NTI / 2; ++jp) ldsm4t(baddr + nhalf + jp * 32, bf[2 * jp][0], bf[2 * jp][1], bf[2 * jp + 1][0], bf[2 * jp + 1][1]);
            if constexpr (NTI & 1) ldsm2t(baddr + (NTI - 1) * 16, bf[NTI - 1][0], bf[NTI - 1][1]);
            const uint32_t aaddr = DZ + (uint32_t)(ks * 16 * ZP) + alane;
#pragma unroll
            for (int i = 0; i < MTW; ++i) ldsm4t(aaddr + i * 32, a[i][0], a[i][1], a[i][2], a[i][3]);
#pragma unroll
            for (int i = 0; i < MTW; ++i)
#pragma unroll
                for (int j = 0; j < NTI; ++j) mma16816(acc[i][j], a[i][0], a[i][1], a[i][2], a[i][3], bf[j][0], bf[j][1]);
            ty += dty; tx += dtx;
            if (tx >= p.TWO) { tx -= p.TWO; ++ty; }
        }
        fence_proxy_async();        // the in-place transform / zero fill precede the next bulk write to this buffer
        __syncthreads();            // buffer b is free for the load issued at the top of the next iteration
    }
#pragma unroll
    for (int i = 0; i < MTW; ++i)
#pragma unroll
        for (int j = 0; j < NTI; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int co = (mt0 + i) * 16 + g + (e >> 1) * 8, ci = j * 8 + 2 * t + (e & 1);
                if (co < COUT) atomicAdd(&p.dw[((size_t)co * CIN + ci) * 9 + tap], acc[i][j][e]);
            }
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
static uint32_t c3_magic(int d) { return (uint32_t)((1ull << 32) / (unsigned)d) + 1u; }

// tile geometry: TH output rows x all Wo columns, at most ptmax pixels; prefer a footprint that lets two CTAs share an SM
template <class SmemFn>
static bool c3_tiles(C3P& p, int S, int ptmax, bool dgrad, int row_bytes, SmemFn smem_of) {
    if (p.Wo > ptmax || p.Wo < 1 || row_bytes % 16 != 0) return false;
    const int thmax = std::min(ptmax / p.Wo, p.Ho);
    auto set = [&](int nblk) {
        p.TH = (p.Ho + nblk - 1) / nblk;
        p.nblk = (p.Ho + p.TH - 1) / p.TH;
        p.TWO = p.Wo;
        p.PT = p.TH * p.TWO;
        p.BH = dgrad ? p.TH + 1 : (p.TH - 1) * S + 3;        // dgrad: dz rows + the halo row
        p.RP = row_bytes;
        p.xb_bytes = c3_al128(p.BH * p.RP);
    };
    int nblk = (p.Ho + thmax - 1) / thmax;
    set(nblk);
    const int first = nblk;
    while (smem_of(p) > 112 * 1024 && p.TH > 2) set(++nblk);
    if (smem_of(p) > 112 * 1024) {          // cannot share an SM: take the largest tile that fits one
        nblk = first;
        set(nblk);
        while (smem_of(p) > 200 * 1024 && p.TH > 1) set(++nblk);
        if (smem_of(p) > 200 * 1024) return false;
    }
    p.items = p.N * p.nblk;
    p.magic_two = c3_magic(p.TWO);
    return p.PT < 65536 && (long long)p.BH * p.RP < (1 << 20);
}

static void c3_fill(C3P& q, const ConvP& p) {
    q.in_scale = p.in_scale; q.in_shift = p.in_shift; q.w = p.w; q.wpk = (const bf16*)p.wpk; q.stats = p.stats;
    q.out = (bf16*)p.out; q.dz = (const bf16*)p.dz; q.dw = p.dw;
    q.N = p.N; q.H = p.H; q.W = p.W; q.Ho = p.Ho; q.Wo = p.Wo;
}

template <class K>
static int c3_launch(K kern, const C3P& q, int threads, size_t smem, cudaStream_t st, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return (int)e; }
    // persistent grid of resident CTAs only (registers and shared memory decide how many share an SM)
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (per_sm < 1) { set_error("%s: kernel does not fit on an SM", name); return MNB_ERR_UNSUPPORTED; }
    const int grid = std::min(q.items, num_sms() * std::min(per_sm, 4));
    kern<<<grid, threads, smem, st>>>(q);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

template <int CIN, int COUT, int S>
static int c3_fwd_launch(const ConvP& p, cudaStream_t st) {
    using Cfg = C3Cfg<CIN, COUT>;
    C3P q = {};
    c3_fill(q, p);
    q.src = (const bf16*)p.x;
    auto smem_of = [](const C3P& g) {
        return (size_t)Cfg::W_BYTES + Cfg::ZB_BYTES + 2 * g.xb_bytes + c3_al128(g.PT * Cfg::OPITCH) + (2 * CIN + 2 * COUT) * 4 + 16;
    };
    if (!c3_tiles(q, S, Cfg::PTMAX, false, p.W * Cfg::PITCH, smem_of)) return MNB_ERR_UNSUPPORTED;
    return c3_launch(c3_fwd_k<CIN, COUT, S>, q, Cfg::THREADS, smem_of(q), st, "conv_fwd(c3_mma)");
}

template <int CIN, int COUT>
static int c3_dgrad_launch(const ConvP& p, cudaStream_t st) {
    using Cfg = C3DCfg<CIN, COUT>;
    C3P q = {};
    c3_fill(q, p);
    q.src = (const bf16*)p.dz;
    auto smem_of = [](const C3P& g) {
        return (size_t)Cfg::W_BYTES + Cfg::ZB_BYTES + 2 * g.xb_bytes + c3_al128(2 * g.TH * 2 * g.TWO * Cfg::SP) + 16;
    };
    if (!c3_tiles(q, 2, 16 * Cfg::MTW * Cfg::WARPS, true, p.Wo * Cfg::ZP, smem_of)) return MNB_ERR_UNSUPPORTED;
    return c3_launch(c3_dgrad_s2_k<CIN, COUT>, q, Cfg::THREADS, smem_of(q), st, "conv_dgrad(c3_mma)");
}

template <int CIN, int COUT, int S>
static int c3_wgrad_launch(const ConvP& p, cudaStream_t st) {
    using Cfg = C3WCfg<CIN, COUT>;
    C3P q = {};
    c3_fill(q, p);
    q.src = (const bf16*)p.x;
    auto smem_of = [](const C3P& g) {
        return (size_t)Cfg::ZB_BYTES + 2 * g.xb_bytes + 2 * c3_al128((g.PT + 15) / 16 * 16 * Cfg::ZP + 32) + 2 * CIN * 4 + 16;
    };
    if (!c3_tiles(q, S, 224, false, p.W * Cfg::PITCH, smem_of)) return MNB_ERR_UNSUPPORTED;
    return c3_launch(c3_wgrad_k<CIN, COUT, S>, q, Cfg::THREADS, smem_of(q), st, "conv_wgrad(c3_mma)");
}

static bool c3_common(const ConvP& p) {
    return option_get(OPT_C3_MMA) && p.k == 3 && p.pad == 1 && !p.nchw_in && !p.out_f32 && ((uintptr_t)p.x & 15) == 0 &&
           ((uintptr_t)p.dz & 15) == 0 && ((uintptr_t)p.out & 15) == 0;
}

// measured at batch 256 (scripts/exp_c3.py): 40->80 forward / backward-weight stay on the tcgen05 path (58 KB of weights
// per CTA, 1 CTA per SM), its backward-data is 3x faster here
#define C3_SHAPES(X) X(16, 24, 2) X(24, 40, 2)
#define C3_SHAPES_DGRAD(X) X(16, 24, 2) X(24, 40, 2) X(40, 80, 2)

int conv_fwd_c3(const ConvP& p, cudaStream_t st) {
    if (!c3_common(p) || p.bias) return MNB_ERR_UNSUPPORTED;
#define X(CI, CO, S) if (p.Cin == CI && p.Cout == CO && p.stride == S) return c3_fwd_launch<CI, CO, S>(p, st);
    C3_SHAPES(X)
#undef X
    return MNB_ERR_UNSUPPORTED;
}
int conv_dgrad_c3(const ConvP& p, cudaStream_t st) {
    // full input rows leave through one contiguous store: the input width must be exactly twice the output's
    if (!c3_common(p) || p.add || p.bn_z || p.stride != 2 || p.W != 2 * p.Wo) return MNB_ERR_UNSUPPORTED;
#define X(CI, CO, S) if (p.Cin == CI && p.Cout == CO && S == 2) return c3_dgrad_launch<CI, CO>(p, st);
    C3_SHAPES_DGRAD(X)
#undef X
    return MNB_ERR_UNSUPPORTED;
}
int conv_wgrad_c3(const ConvP& p, cudaStream_t st) {
    if (!c3_common(p)) return MNB_ERR_UNSUPPORTED;
#define X(CI, CO, S) if (p.Cin == CI && p.Cout == CO && p.stride == S) return c3_wgrad_launch<CI, CO, S>(p, st);
    C3_SHAPES(X)
#undef X
    return MNB_ERR_UNSUPPORTED;
}

}  // namespace mnb
