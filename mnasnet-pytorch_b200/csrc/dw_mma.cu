// Depthwise k x k (3 / 5) stride-1 convolution, bf16 NHWC, on the tensor pipe: forward (+ the producing ConvBlock's
// BN-apply+ReLU on load, + this ConvBlock's BN batch statistics on store) and stand-alone backward-data.
// Replaces nn.Conv2d(groups=C) at src/models/mnasnet.py:76-81,120-125 (called from ConvBlock.forward :58-62).
//
// Why tensor cores for a per-channel stencil: the CUDA-core kernels (dwconv_tile.cu) are bound by instruction issue
// (25 MACs per 4 bytes for 5x5 = ~19 lane-instructions per output with packed FFMA2, against a budget of 23 at HBM
// speed).  With NHWC data the MMA that needs no transposition is, per group of 8 channels,
//     D[m = 16 pixels along W][n = 8 channels] += A[m][k = (tap j, channel c')] * B[(j, c')][n = c],
//     A = the input at 16 consecutive pixels shifted by tap j  -- 8x8 blocks that ldmatrix reads straight from the
//         NHWC tile (8 pixel rows of 16 bytes),            B = diag(w[tap j][c]) (zero off the diagonal),
// i.e. mma.sync.m16n8k16 with two taps per instruction.  Only 1/8 of the MACs are useful, but one warp instruction
// does 256 of them: 15 (5x5) / 6 (3x3) MMAs + 3 / 2 ldmatrix per 128 outputs instead of ~600 / ~250 FFMA2 + loads.
// The legacy tensor path issues 0.5 mma.sync per clock per SM on B200 (scripts/ubench/ub_mma.cu): 5x5 is bound by it
// at ~0.77 of the HBM roofline, 3x3 is HBM-bound.
//
// Structure: a CTA owns CG channels x TW = 16*TWS output columns and walks DOWN the image in blocks of RB rows
// (RB = 12 / 10, a multiple of k), one warp per (8-channel chunk, 16-column strip).  Per block:
//   1. one thread issues a TMA load of the RB x (TW+k-1) x CG input box (negative / overhanging coordinates are
//      zero-filled by the hardware = the convolution's padding) and everybody waits on its mbarrier;
//   2. every thread applies relu(scale*x+shift) in place to its share of 16-byte vectors (padding stays 0);
//   3. each warp feeds every input row ONCE through ldmatrix into k accumulators (the k output rows that row
//      contributes to) which rotate through registers across rows AND across blocks -- no vertical halo is ever
//      re-read or re-transformed; finished rows go to a staging tile (+ statistics of the bf16-rounded values);
//   4. one thread issues the TMA store of the RB x TW x CG output box (clipped at the image border).
// A work item is a column segment of `seg` blocks, preceded by a k/2-row prologue box that primes the accumulators.
#include "dw_mma.cuh"
#include "dw_mma_dev.cuh"

#include <algorithm>
#include <mutex>
#include <unordered_map>

namespace mnb {

// ---------------------------------------------------------------------------------------------------------------
// host: tensor maps + geometry
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
    }
    return fn;
}

struct TmKey {
    const void* ptr;
    int N, H, W, C, bc, bw, bh;
    bool operator==(const TmKey& o) const {
        return ptr == o.ptr && N == o.N && H == o.H && W == o.W && C == o.C && bc == o.bc && bw == o.bw && bh == o.bh;
    }
};
struct TmHash {
    size_t operator()(const TmKey& k) const {
        size_t h = (size_t)k.ptr;
        const int v[7] = {k.N, k.H, k.W, k.C, k.bc, k.bw, k.bh};
        for (int i = 0; i < 7; ++i) h = h * 1000003u ^ (size_t)v[i];
        return h;
    }
};

int dwm_tensor_map(CUtensorMap* out, const void* ptr, int N, int H, int W, int C, int bc, int bw, int bh) {
    static std::mutex mu;
    static std::unordered_map<TmKey, CUtensorMap, TmHash> cache;
    const TmKey key = {ptr, N, H, W, C, bc, bw, bh};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error("dw_mma: cuTensorMapEncodeTiled is not available"); return MNB_ERR_UNSUPPORTED; }
    if (((uintptr_t)ptr & 15) != 0 || bc > 256 || bw > 256 || bh > 256) {
        set_error("dw_mma: tensor map needs a 16-byte aligned pointer and box extents <= 256");
        return MNB_ERR_UNSUPPORTED;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("dw_mma: cuTensorMapEncodeTiled failed (%d)", (int)r); return MNB_ERR_UNSUPPORTED; }
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return 0;
}

DwmGeom dwm_geometry(int C, int W, int K) {
    // CG = 24 (3 chunks) or 40 (5 chunks): the smallest padded channel count wins, ties go to 24 (more CTAs per SM).
    DwmGeom g = {};
    int cg = option_get(OPT_DW_MMA_CG);
    if (cg != 24 && cg != 40) {
        const int pad24 = (C + 23) / 24 * 24, pad40 = (C + 39) / 40 * 40;
        cg = pad40 < pad24 ? 40 : 24;
    }
    // one 16-column strip per CTA: half the shared memory of two strips, so twice the CTAs share an SM -- measured 6-17 %
    // faster on the 112 x 112 maps (fused backward 513 -> 428 us) and neutral elsewhere; "dw_mma_tws" = 2 restores two strips
    int tws = option_get(OPT_DW_MMA_TWS);
    if (tws != 1 && tws != 2) tws = 1;
    if (W <= 16) tws = 1;
    g.CG = cg; g.NCH = cg / 8; g.TWS = tws; g.TW = 16 * tws; g.HC = g.TW + K - 1; g.PITCH = cg * 2;
    g.tiles_w = (W + g.TW - 1) / g.TW;
    g.cblocks = (C + cg - 1) / cg;
    g.threads = 32 * g.NCH * tws;
    return g;
}

// ---------------------------------------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------------------------------------
struct DwmP {
    const float* in_scale;      // forward: BN scale/shift of the producing ConvBlock (NULL = plain input)
    const float* in_shift;
    const float* w;             // [C][K][K] fp32
    double* stats;              // [2][C] sum / sum of squares of the stored outputs (NULL = none)
    int N, H, W, C;
    int tiles_w, cblocks;
    int items;                  // N * tiles_w * nseg
    int nseg, seg;              // column segments per column, blocks per segment
    int nblocks;                // ceil(H / RB)
};

// General path (image borders, segment prologues): every row and output is range-checked; each input row applies all
// its taps at once.
template <int K, int CG, int TWS, int I, class EM>
struct DwmRowsEdge {
    static __device__ __forceinline__ void run(float (&acc)[K][4], const uint32_t (&bd)[K][K], uint32_t x4, uint32_t x2,
                                               int rbase, int rlo, int H, int obase, EM& em) {
        using Cfg = DwmCfg<K, CG, TWS>;
        const int r = rbase + I;
        if (r >= rlo && r < H) {
            uint32_t m[K][2];
            dwm_load_row<K, Cfg::PITCH>(m, x4 + I * Cfg::ROWB, x2 + I * Cfg::ROWB);
            dwm_mma_pairs<K, I>(acc, bd, m);
#pragma unroll
            for (int kh = 0; kh < K; ++kh) mma1688(acc[(I + K - 1 - kh) % K], m[K - 1][0], m[K - 1][1], bd[kh][K - 1]);
        }
        const int o = obase + I;                       // output row completed by input row r (tap row K-1)
        float (&a)[4] = acc[I % K];
        if (o >= 0 && o < H) em.template emit<I, true>(a);
        a[0] = a[1] = a[2] = a[3] = 0.f;               // the slot now belongs to output row o + K
        if constexpr (I + 1 < Cfg::RB)
            DwmRowsEdge<K, CG, TWS, I + 1, EM>::run(acc, bd, x4, x2, rbase, rlo, H, obase, em);
    }
};

// Interior path (all RB input rows and outputs inside the image; MASKED: the strip overhangs the right border, its
// surplus columns are computed and clipped by the TMA store, only the statistics mask them): no range checks, and the last
// tap column is paired ACROSS rows -- tap (kh, K-1) of row r-1 and tap (kh+1, K-1) of row r feed the same output row,
// so they share one m16n8k16: 13 instead of 15 MMAs per row for 5x5, 5 instead of 6 for 3x3 (the tensor pipe issues
// one mma.sync per 2 clocks per SM whatever its k extent).
template <int K, int CG, int TWS, int I, bool MASKED, class EM>
struct DwmRowsFull {
    static __device__ __forceinline__ void run(float (&acc)[K][4], const uint32_t (&bd)[K][K], uint32_t x4, uint32_t x2,
                                               EM& em, uint32_t p0, uint32_t p1) {
        using Cfg = DwmCfg<K, CG, TWS>;
        uint32_t m[K][2];
        dwm_load_row<K, Cfg::PITCH>(m, x4 + I * Cfg::ROWB, x2 + I * Cfg::ROWB);
        dwm_mma_pairs<K, I>(acc, bd, m);
#pragma unroll
        for (int kh = 0; kh + 1 < K; kh += 2) {
            if constexpr (I > 0)    // (row I-1, kh) + (row I, kh+1) -> output slot of (I, kh+1)
                mma16816(acc[(I + K - 2 - kh) % K], p0, p1, m[K - 1][0], m[K - 1][1], bd[kh][K - 1], bd[kh + 1][K - 1]);
            else
                mma1688(acc[(I + K - 2 - kh) % K], m[K - 1][0], m[K - 1][1], bd[kh + 1][K - 1]);
        }
        mma1688(acc[I % K], m[K - 1][0], m[K - 1][1], bd[K - 1][K - 1]);
        if constexpr (I + 1 == Cfg::RB) {               // last row of the block: its even tap rows have no partner
#pragma unroll
            for (int kh = 0; kh + 1 < K; kh += 2) mma1688(acc[(I + K - 1 - kh) % K], m[K - 1][0], m[K - 1][1], bd[kh][K - 1]);
        }
        float (&a)[4] = acc[I % K];
        em.template emit<I, MASKED>(a);
        a[0] = a[1] = a[2] = a[3] = 0.f;
        if constexpr (I + 1 < Cfg::RB)
            DwmRowsFull<K, CG, TWS, I + 1, MASKED, EM>::run(acc, bd, x4, x2, em, m[K - 1][0], m[K - 1][1]);
    }
};

// FLIP: correlate with the 180-degree rotated kernel (backward-data of a stride-1 'same' depthwise conv)
template <int K, int CG, int TWS, bool FLIP>
__global__ void __launch_bounds__(DwmCfg<K, CG, TWS>::THREADS, DwmCfg<K, CG, TWS>::MINB)
    dw_mma_fwd_k(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_xp,
                 const __grid_constant__ CUtensorMap tm_z, const DwmP p) {
    using Cfg = DwmCfg<K, CG, TWS>;
    constexpr int P = Cfg::P, RB = Cfg::RB, NCH = Cfg::NCH, TW = Cfg::TW, HC = Cfg::HC, PITCH = Cfg::PITCH;
    constexpr int THREADS = Cfg::THREADS;
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

    // diagonal weight fragments of this warp's chunk (bf16, like every tensor-pipe operand in bf16 mode)
    uint32_t bd[K][K];
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
            const int ih = FLIP ? K - 1 - kh : kh, iw = FLIP ? K - 1 - kw : kw;
            const float wv = chunk_live ? p.w[(size_t)(cbase + chunk * 8 + g) * K * K + ih * K + iw] : 0.f;
            bd[kh][kw] = dwm_diag(wv, g, t);
        }
    // transform mapping: a thread keeps one chunk for life (THREADS is a multiple of NCH) -> constants in registers
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
    // ldmatrix lane addresses (relative to the first row of a block): x4 = shifts (2q, 2q+1) x halves, x2 = shift K-1
    const int mi = lane >> 3, r8 = lane & 7;
    const uint32_t off4 = (uint32_t)(((mi >> 1) + (mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    const uint32_t off2 = (uint32_t)(((K - 1) + (mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    const uint32_t out_lane = OUT + (uint32_t)((strip * 16 + g) * PITCH + chunk * 16 + t * 4);
    const bool do_stats = p.stats != nullptr;
    typedef DwmEmitStats<K, CG, TWS> EM;
    EM em;
    em.out_lane = out_lane;
    em.do_stats = do_stats;
    em.st[0] = em.st[1] = f2_pack(0.f, 0.f);
    float acc[K][4];
    uint32_t ph = 0;
    __syncthreads();

    // TMA load of a step's input box into buffer b (thread 0).  Block j holds input rows RB*j+P .. RB*j+P+RB-1 and
    // completes output rows RB*j .. RB*j+RB-1; the prologue of a segment holds the K-1 rows above its first block
    // (rows RB*jb0-P .. RB*jb0+P-1; negative rows are zero-filled) and completes nothing.
    auto issue = [&](const DwmStep& s, int b) {
        const bool pro = s.j < s.jb0;
        const int row_first = pro ? RB * s.jb0 - P : RB * s.j + P;
        if (row_first < p.H) {
            mbar_expect_tx(bar0 + 8 * b, (uint32_t)((pro ? K - 1 : RB) * Cfg::ROWB));
            tma_load4(XB0 + b * Cfg::XB_BYTES, pro ? &tm_xp : &tm_x, cbase, s.w0 - P, row_first, s.n, bar0 + 8 * b);
        }
    };

    DwmStep cur, nxt;
    dwm_step_init<TW>(cur, slot, p.items, p.nseg, p.tiles_w, p.seg, p.nblocks);
    if (tid == 0 && cur.item < p.items) issue(cur, 0);
    int b = 0;
    while (cur.item < p.items) {
        nxt = cur;
        if (++nxt.j >= nxt.jb1) dwm_step_init<TW>(nxt, nxt.item + nslots, p.items, p.nseg, p.tiles_w, p.seg, p.nblocks);
        if (tid == 0) {
            tma_store_wait_read();                      // the previous step's store has finished reading OUT
            if (nxt.item < p.items) issue(nxt, b ^ 1);  // buffer b^1 was released by the previous step's last barrier
        }
        const bool pro = cur.j < cur.jb0;
        const int rbase = RB * cur.j + P;               // input row of block row 0
        const int row_first = pro ? RB * cur.jb0 - P : rbase;
        const uint32_t XB = XB0 + b * Cfg::XB_BYTES;
        if (row_first < p.H) {
            mbar_wait(bar0 + 8 * b, (ph >> b) & 1);
            ph ^= 1u << b;
            if (xf) {       // relu(scale * x + shift) in place; zero-filled padding stays zero
                const int npix = (pro ? K - 1 : RB) * HC;
                for (int pix = pix0; pix < npix; pix += PSTEP) {
                    const int rr = pix / HC, cc = pix - rr * HC;
                    const int ih = row_first + rr, iw = cur.w0 - P + cc;
                    if ((unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W) {
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
        }
        __syncthreads();
        if (pro) {
#pragma unroll
            for (int i = 0; i < K; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        }
        const int sw = cur.w0 + strip * 16;
        if (chunk_live && sw < p.W) {
            const bool full = !pro && rbase + RB <= p.H;
            const float m0 = sw + g < p.W ? 1.f : 0.f, m1 = sw + g + 8 < p.W ? 1.f : 0.f;
            em.mk0 = f2_pack(m0, m0);
            em.mk1 = f2_pack(m1, m1);
            if (full && sw + 16 <= p.W) {
                DwmRowsFull<K, CG, TWS, 0, false, EM>::run(acc, bd, XB + off4, XB + off2, em, 0u, 0u);
            } else if (full) {
                DwmRowsFull<K, CG, TWS, 0, true, EM>::run(acc, bd, XB + off4, XB + off2, em, 0u, 0u);
            } else {
                // the prologue box sits at the start of the buffer but holds block rows RB-(K-1) .. RB-1
                const uint32_t base = pro ? XB - (uint32_t)((RB - (K - 1)) * Cfg::ROWB) : XB;
                DwmRowsEdge<K, CG, TWS, 0, EM>::run(acc, bd, base + off4, base + off2, rbase, max(row_first, 0), p.H,
                                                    pro ? -2 * RB : RB * cur.j, em);
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0 && !pro) tma_store4(&tm_z, cbase, cur.w0, RB * cur.j, cur.n, OUT);
        cur = nxt;
        b ^= 1;
    }
    if (tid == 0) tma_store_wait_read();
    if (do_stats) {
        // lanes with equal t hold the same channel pair: reduce over g, then one shared atomic per warp and value
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
// launchers
// ---------------------------------------------------------------------------------------------------------------
template <int K, int CG, int TWS, bool FLIP>
static int launch_fwd_cfg(const DwmGeom& g, const void* x, const float* s, const float* t, const float* w, void* z,
                          double* stats, int N, int H, int W, int C, cudaStream_t st, const char* name) {
    using Cfg = DwmCfg<K, CG, TWS>;
    constexpr int RB = Cfg::RB;
    CUtensorMap tm_x, tm_xp, tm_z;
    if (int e = dwm_tensor_map(&tm_x, x, N, H, W, C, CG, Cfg::HC, RB)) return e;
    if (int e = dwm_tensor_map(&tm_xp, x, N, H, W, C, CG, Cfg::HC, K - 1)) return e;
    if (int e = dwm_tensor_map(&tm_z, z, N, H, W, C, CG, Cfg::TW, RB)) return e;
    DwmP p = {};
    p.in_scale = s; p.in_shift = t; p.w = w; p.stats = stats;
    p.N = N; p.H = H; p.W = W; p.C = C;
    p.tiles_w = g.tiles_w; p.cblocks = g.cblocks;
    p.nblocks = (H + RB - 1) / RB;
    static int occ = -1;
    if (occ < 0) {
        cudaError_t e = cudaFuncSetAttribute(dw_mma_fwd_k<K, CG, TWS, FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM);
        if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return (int)e; }
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, dw_mma_fwd_k<K, CG, TWS, FLIP>, Cfg::THREADS, Cfg::SMEM);
        if (o < 1) { set_error("%s: kernel does not fit on an SM", name); return MNB_ERR_UNSUPPORTED; }
        occ = o;
    }
    const long long cols = (long long)N * g.tiles_w;
    if (cols > (1 << 24)) { set_error("%s: too many columns", name); return MNB_ERR_UNSUPPORTED; }
    // segments: enough items for >= 6 rounds over the resident CTAs when the map is tall enough, at least 2 blocks each
    // resident CTAs only: a grid that exceeds SMs x occupancy runs its surplus CTAs as a second wave (2x the time)
    const long long slots_max = std::max(1LL, (long long)num_sms() * occ / g.cblocks);
    int seg = option_get(OPT_DW_MMA_SEG);
    if (seg <= 0) {
        seg = p.nblocks;
        while (seg > 2 && cols * ((p.nblocks + seg - 1) / seg) < 6 * slots_max) seg = (seg + 1) / 2;
    }
    if (seg > p.nblocks) seg = p.nblocks;
    p.seg = seg;
    p.nseg = (p.nblocks + seg - 1) / seg;
    p.items = (int)(cols * p.nseg);
    long long slots = slots_max;
    if (slots > p.items) slots = p.items;
    dw_mma_fwd_k<K, CG, TWS, FLIP><<<(unsigned)(slots * g.cblocks), Cfg::THREADS, Cfg::SMEM, st>>>(tm_x, tm_xp, tm_z, p);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

template <int K, bool FLIP>
static int launch_fwd(const void* x, const float* s, const float* t, const float* w, void* z, double* stats, int N, int H,
                      int W, int C, cudaStream_t st, const char* name) {
    const DwmGeom g = dwm_geometry(C, W, K);
    if (g.CG == 24 && g.TWS == 1) return launch_fwd_cfg<K, 24, 1, FLIP>(g, x, s, t, w, z, stats, N, H, W, C, st, name);
    if (g.CG == 24 && g.TWS == 2) return launch_fwd_cfg<K, 24, 2, FLIP>(g, x, s, t, w, z, stats, N, H, W, C, st, name);
    if (g.CG == 40 && g.TWS == 1) return launch_fwd_cfg<K, 40, 1, FLIP>(g, x, s, t, w, z, stats, N, H, W, C, st, name);
    return launch_fwd_cfg<K, 40, 2, FLIP>(g, x, s, t, w, z, stats, N, H, W, C, st, name);
}

int dw_fwd_mma(const void* x, const float* s, const float* t, const float* w, const float* bias, void* z, double* stats,
               int N, int H, int W, int C, int k, cudaStream_t st) {
    if (bias) { set_error("dw_fwd(mma): the conv bias is folded into BatchNorm on this path"); return MNB_ERR_UNSUPPORTED; }
    if (k == 3) return launch_fwd<3, false>(x, s, t, w, z, stats, N, H, W, C, st, "dw_fwd(mma)");
    if (k == 5) return launch_fwd<5, false>(x, s, t, w, z, stats, N, H, W, C, st, "dw_fwd(mma)");
    return MNB_ERR_UNSUPPORTED;
}

int dw_dgrad_mma(const void* dz, const float* w, void* dx, int N, int H, int W, int C, int k, cudaStream_t st) {
    if (k == 3) return launch_fwd<3, true>(dz, nullptr, nullptr, w, dx, nullptr, N, H, W, C, st, "dw_dgrad(mma)");
    if (k == 5) return launch_fwd<5, true>(dz, nullptr, nullptr, w, dx, nullptr, N, H, W, C, st, "dw_dgrad(mma)");
    return MNB_ERR_UNSUPPORTED;
}


// ===============================================================================================================
// Fused depthwise ConvBlock backward: BatchNorm-backward elementwise pass + backward-data + backward-weight (+ the
// BatchNorm-backward reductions of the ConvBlock that produced this layer's input) in ONE pass over
//     G (gradient w.r.t. this block's ReLU output), Z (this block's raw conv output), X (raw output of the producing
//     block)  ->  dX (gradient w.r.t. the producing block's ReLU output), dW, sum(G'), sum(G' x).
// Replaces native_batch_norm_backward + convolution_backward of one depthwise ConvBlock (src/models/mnasnet.py:58-62
// under autograd); math contract SURVEY.md appendix F:
//     dZ = a * G * [scale*Z + shift > 0] + b * Z + c      (a, b, c per channel from sum(G'), sum(G' z), mean, invstd)
//     dX[p] = sum_taps dZ[p + P - tap] * w[tap]            dW[tap] = sum_p dZ[p] * A[p + tap - P],  A = relu(s_in*X + t_in)
// 4 tensor passes (3 reads, 1 write) instead of 9 (apply 3, backward-data 2, backward-weight 2, next reduction 2).
//
// Same walk as the forward kernel (CTA = CG channels x TW columns, blocks of RB rows down the image, one warp per
// chunk x strip).  Per block: TMA boxes of G and Z (with the horizontal halo, rows RB*j+P ..) and of X (no halo, rows
// RB*j .. RB*j+RB+2P-1); dZ is computed in place over G; then per warp
//   (a) backward-data = the forward row pipeline with the rotated kernel; a finished dX row is staged for the TMA
//       store and, masked by the producing block's ReLU, reduced against the raw X of the same pixels;
//   (b) backward-weight on the tensor pipe as well: D[(tap, c')][c] += sum over 16 pixels of dZ^T (ldmatrix.trans of
//       the shifted dZ blocks) x A (ldmatrix.trans of the X blocks, activated in registers); each dZ row is paired
//       with the K input rows around it, the diagonal c' == c of the accumulators is the weight gradient.
// ===============================================================================================================
struct DwbP {
    const float* scale;         // this block's BN scale / shift (ReLU mask)
    const float* shift;
    const double* sums;         // [2][C] sum(G'), sum(G' z) of this block (already reduced)
    const float* mean;
    const float* invstd;
    double m;                   // N*H*W
    float* dgamma;              // += (NULL = frozen)
    float* dbeta;
    float* dbias;
    const float* in_scale;      // producing block's BN scale / shift
    const float* in_shift;
    const float* w;             // [C][K][K]
    float* dw;                  // += (NULL = frozen: no backward-weight)
    double* nsums;              // [2][C] sum(dX'), sum(dX' x) for the producing block (NULL = not wanted)
    int N, H, W, C;
    int tiles_w, cblocks, items, nseg, seg, nblocks;
};

template <int K, int CG, int TWS>
struct DwbCfg {
    using F = DwmCfg<K, CG, TWS>;
    static constexpr int P = F::P, RB = F::RB, NCH = F::NCH, TW = F::TW, HC = F::HC, PITCH = F::PITCH;
    static constexpr int XROWS = RB + 2 * P;
    static constexpr int GB_BYTES = F::XB_BYTES;                                  // RB x HC pixels (G -> dZ; Z -> dX staging)
    static constexpr int XB_BYTES = (XROWS * F::OROWB + 127) / 128 * 128;         // XROWS x TW pixels
    static constexpr int WS_BYTES = (CG * K * K * 2 + 15) / 16 * 16;              // bf16 weights
    static constexpr int DW_BYTES = CG * K * K * 4;                               // fp32 weight-gradient accumulators
    static constexpr int CO_BYTES = 7 * CG * 4;                                   // a, b, c, scale, shift, in_scale, in_shift
    static constexpr int SMEM = 2 * GB_BYTES + XB_BYTES + WS_BYTES + DW_BYTES + CO_BYTES + 2 * CG * 4 + 16;
    static constexpr int THREADS = F::THREADS;
    static constexpr int MINB = 65536 / (THREADS * (K == 5 ? 128 : 104)) > 0 ? 65536 / (THREADS * (K == 5 ? 128 : 104)) : 1;
};

// dX row finished by block row I: stage for the TMA store; reduce dX' = dX * [s_in x + t_in > 0] and dX' * x.
template <int K, int CG, int TWS>
struct DwbEmitReduce {
    uint32_t out_lane;
    uint32_t x_lane;        // raw X of pixel (row 0 of the block, column of this lane), this lane's channel pair
    bool do_red;
    f2_t sp, tp;            // producing block's scale / shift of the lane's channel pair
    f2_t mk0, mk1;
    f2_t rs[2];             // (sum dX', sum dX' x)
    template <int I, bool MASKED>
    __device__ __forceinline__ void emit(float (&a)[4]) {
        using Cfg = DwmCfg<K, CG, TWS>;
        const uint32_t u0 = pack_bf16x2(a[0], a[1]), u1 = pack_bf16x2(a[2], a[3]);
        sts32(out_lane + I * Cfg::OROWB, u0);
        sts32(out_lane + I * Cfg::OROWB + 8 * Cfg::PITCH, u1);
        if (do_red) {
            const f2_t one = f2_pack(1.f, 1.f), zero = f2_pack(0.f, 0.f);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const f2_t x = f2_from_bf16x2(lds32(x_lane + I * Cfg::OROWB + h * 8 * Cfg::PITCH));
                float y0, y1, q0, q1;
                f2_unpack(f2_fma(sp, x, tp), y0, y1);
                f2_unpack(f2_from_bf16x2(h ? u1 : u0), q0, q1);
                f2_t q = f2_pack(y0 > 0.f ? q0 : 0.f, y1 > 0.f ? q1 : 0.f);
                if (MASKED) q = f2_fma(q, h ? mk1 : mk0, zero);
                rs[0] = f2_fma(q, one, rs[0]);
                rs[1] = f2_fma(q, x, rs[1]);
            }
        }
    }
};

// Backward-weight of one step for the tap rows KH0 .. KH1-1 (5x5 runs two passes, 3 + 2 tap rows, so that at most 36
// accumulator registers are live next to the backward-data ring).  dZ block row i (ldmatrix.trans of its shifted 8x8
// blocks, the A operand) pairs with the activated X block rows i + kh (ldmatrix.trans, the B operand; a ring of
// KH1-KH0 fragments, one new row per dZ row): wacc[kh][pair] += dZ^T(shift 2*pair | 2*pair+1) x A.
struct DwbWg {
    uint32_t gz4, gz2;      // this lane's ldmatrix addresses of dZ block row 0 (x4: shifts (2q,2q+1) x halves, x2: shift K-1)
    uint32_t xrow;          // this lane's ldmatrix address of X block row 0 (two 8-pixel halves)
    uint32_t cm0, cm1;      // column validity masks of the lane's pixel pairs (X is zero-filled outside the image but
                            // relu(shift) is not zero; dZ is exactly zero there already)
    f2_t ag, at;            // producing block's scale / shift of the lane's channel g
    bool act;
    int xrow_img0;          // image row of X block row 0
    int rbase, i0, H;
};

template <int K, int CG, int TWS>
__device__ __forceinline__ void dwb_load_x(const DwbWg& q, int xrow, uint32_t (&dst)[2]) {
    using Cfg = DwmCfg<K, CG, TWS>;
    const int r = q.xrow_img0 + xrow;
    dst[0] = dst[1] = 0u;
    if (r >= 0 && r < q.H) {
        uint32_t u0, u1;
        ldsm2t(q.xrow + (uint32_t)(xrow * Cfg::OROWB), u0, u1);
        if (q.act) {
            u0 = f2_relu_bf16x2(f2_fma(f2_from_bf16x2(u0), q.ag, q.at));
            u1 = f2_relu_bf16x2(f2_fma(f2_from_bf16x2(u1), q.ag, q.at));
        }
        dst[0] = u0 & q.cm0;
        dst[1] = u1 & q.cm1;
    }
}

template <int K, int CG, int TWS, int KH0, int KH1>
__device__ __forceinline__ void dwb_wgrad_rows(const DwbWg& q, float (&wacc)[KH1 - KH0][(K + 1) / 2][4]) {
    using Cfg = DwmCfg<K, CG, TWS>;
    constexpr int RB = Cfg::RB, PITCH = Cfg::PITCH, NPR = (K + 1) / 2, R = KH1 - KH0;
    uint32_t bf[R][2];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
        if (i < q.i0) continue;
        if (i == q.i0) {
#pragma unroll
            for (int kh = KH0; kh + 1 < KH1; ++kh) dwb_load_x<K, CG, TWS>(q, i + kh, bf[(i + kh) % R]);
        }
        dwb_load_x<K, CG, TWS>(q, i + KH1 - 1, bf[(i + KH1 - 1) % R]);
        const int r = q.rbase + i;
        if (r >= 0 && r < q.H) {
            uint32_t tt[K][2];
            const uint32_t a4 = q.gz4 + (uint32_t)(i * Cfg::ROWB), a2 = q.gz2 + (uint32_t)(i * Cfg::ROWB);
            ldsm4t(a4, tt[0][0], tt[0][1], tt[1][0], tt[1][1]);
            if constexpr (K == 5) ldsm4t(a4 + 2 * PITCH, tt[2][0], tt[2][1], tt[3][0], tt[3][1]);
            ldsm2t(a2, tt[K - 1][0], tt[K - 1][1]);
#pragma unroll
            for (int kh = KH0; kh < KH1; ++kh)
#pragma unroll
                for (int pr = 0; pr < NPR; ++pr) {
                    const int s0 = 2 * pr, s1 = (2 * pr + 1 < K) ? 2 * pr + 1 : 2 * pr;   // odd K: last pair duplicates
                    mma16816(wacc[kh - KH0][pr], tt[s0][0], tt[s1][0], tt[s0][1], tt[s1][1], bf[(i + kh) % R][0],
                             bf[(i + kh) % R][1]);
                }
        }
    }
}

// the diagonal c' == c lives on the lanes with t == g >> 1: slot A in d[g & 1], slot B in d[2 + (g & 1)];
// shift s of the dZ operand is tap column K-1-s, kh is the tap row
template <int K, int KH0, int KH1>
__device__ __forceinline__ void dwb_wgrad_flush(const float (&wacc)[KH1 - KH0][(K + 1) / 2][4], float* dst, int g, int t) {
    constexpr int NPR = (K + 1) / 2;
    if ((g >> 1) == t) {
#pragma unroll
        for (int kh = KH0; kh < KH1; ++kh)
#pragma unroll
            for (int pr = 0; pr < NPR; ++pr) {
                atomicAdd(dst + kh * K + (K - 1 - 2 * pr), (g & 1) ? wacc[kh - KH0][pr][1] : wacc[kh - KH0][pr][0]);
                if (2 * pr + 1 < K)
                    atomicAdd(dst + kh * K + (K - 2 - 2 * pr), (g & 1) ? wacc[kh - KH0][pr][3] : wacc[kh - KH0][pr][2]);
            }
    }
}

template <int K, int CG, int TWS, int KH0, int KH1>
__device__ __forceinline__ void dwb_wgrad_pass(const DwbWg& q, float* dst, int g, int t) {
    constexpr int NPR = (K + 1) / 2, R = KH1 - KH0;
    float wacc[R][NPR][4];
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
        for (int pr = 0; pr < NPR; ++pr) wacc[a][pr][0] = wacc[a][pr][1] = wacc[a][pr][2] = wacc[a][pr][3] = 0.f;
    dwb_wgrad_rows<K, CG, TWS, KH0, KH1>(q, wacc);
    dwb_wgrad_flush<K, KH0, KH1>(wacc, dst, g, t);
}

// Stand-alone backward-weight: dZ (with the horizontal halo) and X boxes double-buffered, one barrier per step, the
// accumulators stay in registers for the CTA's whole life (its channel group never changes) and are flushed once.
template <int K, int CG, int TWS>
__global__ void __launch_bounds__(DwmCfg<K, CG, TWS>::THREADS)
    dw_mma_wgrad_k(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_gp,
                   const __grid_constant__ CUtensorMap tm_x, const DwbP p) {
    using Cfg = DwmCfg<K, CG, TWS>;
    using B = DwbCfg<K, CG, TWS>;
    constexpr int P = Cfg::P, RB = Cfg::RB, NCH = Cfg::NCH, TW = Cfg::TW, PITCH = Cfg::PITCH;
    constexpr int THREADS = Cfg::THREADS, KK = K * K, NPR = (K + 1) / 2, STAGE = B::GB_BYTES + B::XB_BYTES;
    extern __shared__ __align__(128) unsigned char dsm[];
    const uint32_t S0 = smem_u32(dsm);
    float* dwacc = reinterpret_cast<float*>(dsm + 2 * STAGE);                    // [CG][KK]
    const uint32_t bar0 = S0 + 2 * STAGE + B::DW_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int chunk = warp % NCH, strip = warp / NCH;
    const int cb = blockIdx.x % p.cblocks, slot = blockIdx.x / p.cblocks, nslots = gridDim.x / p.cblocks;
    const int cbase = cb * CG;
    const int ch = cbase + chunk * 8 + g;
    const bool chunk_live = cbase + chunk * 8 < p.C;
    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    for (int i = tid; i < CG * KK; i += THREADS) dwacc[i] = 0.f;
    const bool act_in = p.in_scale != nullptr;
    const float sg = (act_in && chunk_live) ? p.in_scale[ch] : 1.f, tg = (act_in && chunk_live) ? p.in_shift[ch] : 0.f;
    const int mi = lane >> 3, r8 = lane & 7;
    const uint32_t off4 = (uint32_t)(((mi >> 1) + (mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    const uint32_t off2 = (uint32_t)(((K - 1) + (mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    const uint32_t offx = (uint32_t)(((mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    float wacc[K][NPR][4];
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
        for (int pr = 0; pr < NPR; ++pr) wacc[a][pr][0] = wacc[a][pr][1] = wacc[a][pr][2] = wacc[a][pr][3] = 0.f;
    uint32_t ph = 0;
    __syncthreads();

    // a segment's prologue rows belong to the segment above, except the K/2 rows above block 0
    auto first_j = [](const DwmStep& s) { return s.jb0 == 0 ? -1 : s.jb0; };
    auto issue = [&](const DwmStep& s, int b) {
        const bool pro = s.j < s.jb0;
        const int row_first = pro ? RB * s.jb0 - P : RB * s.j + P;
        if (row_first < p.H) {
            const uint32_t gb = S0 + b * STAGE;
            mbar_expect_tx(bar0 + 8 * b, (uint32_t)((pro ? K - 1 : RB) * Cfg::ROWB + B::XROWS * Cfg::OROWB));
            tma_load4(gb, pro ? &tm_gp : &tm_g, cbase, s.w0 - P, row_first, s.n, bar0 + 8 * b);
            tma_load4(gb + B::GB_BYTES, &tm_x, cbase, s.w0, RB * s.j, s.n, bar0 + 8 * b);
        }
    };
    DwmStep cur, nxt;
    dwm_step_init<TW>(cur, slot, p.items, p.nseg, p.tiles_w, p.seg, p.nblocks);
    if (cur.item < p.items) cur.j = first_j(cur);
    if (tid == 0 && cur.item < p.items) issue(cur, 0);
    int b = 0;
    while (cur.item < p.items) {
        nxt = cur;
        if (++nxt.j >= nxt.jb1) {
            dwm_step_init<TW>(nxt, nxt.item + nslots, p.items, p.nseg, p.tiles_w, p.seg, p.nblocks);
            if (nxt.item < p.items) nxt.j = first_j(nxt);
        }
        if (tid == 0 && nxt.item < p.items) issue(nxt, b ^ 1);      // released by the previous step's barrier
        const bool pro = cur.j < cur.jb0;
        const int rbase = RB * cur.j + P;
        const int row_first = pro ? RB * cur.jb0 - P : rbase;
        if (row_first < p.H) {
            mbar_wait(bar0 + 8 * b, (ph >> b) & 1);
            ph ^= 1u << b;
            const int sw = cur.w0 + strip * 16;
            if (chunk_live && sw < p.W) {
                const uint32_t gb = S0 + b * STAGE;
                const uint32_t base = pro ? gb - (uint32_t)((RB - (K - 1)) * Cfg::ROWB) : gb;
                DwbWg q;
                q.gz4 = base + off4; q.gz2 = base + off2; q.xrow = gb + B::GB_BYTES + offx;
                const int c0 = sw + 2 * t;
                q.cm0 = (c0 < p.W ? 0x0000ffffu : 0u) | (c0 + 1 < p.W ? 0xffff0000u : 0u);
                q.cm1 = (c0 + 8 < p.W ? 0x0000ffffu : 0u) | (c0 + 9 < p.W ? 0xffff0000u : 0u);
                q.ag = f2_pack(sg, sg); q.at = f2_pack(tg, tg); q.act = act_in;
                q.xrow_img0 = RB * cur.j; q.rbase = rbase; q.i0 = pro ? RB - (K - 1) : 0; q.H = p.H;
                dwb_wgrad_rows<K, CG, TWS, 0, K>(q, wacc);
            }
        }
        __syncthreads();
        cur = nxt;
        b ^= 1;
    }
    if (chunk_live) dwb_wgrad_flush<K, 0, K>(wacc, dwacc + (chunk * 8 + g) * KK, g, t);
    __syncthreads();
    for (int i = tid; i < CG * KK; i += THREADS)
        if (cbase + i / KK < p.C) atomicAdd(&p.dw[(size_t)cbase * KK + i], dwacc[i]);
}

// WGO: backward-weight only -- the G box already holds dZ (no Z box, no transform, no backward-data, no store)
template <int K, int CG, int TWS, bool WGO>
__global__ void __launch_bounds__(DwbCfg<K, CG, TWS>::THREADS, DwbCfg<K, CG, TWS>::MINB)
    dw_mma_bwd_k(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_gp,
                 const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_zp,
                 const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dx, const DwbP p) {
    using Cfg = DwmCfg<K, CG, TWS>;
    using B = DwbCfg<K, CG, TWS>;
    constexpr int P = Cfg::P, RB = Cfg::RB, NCH = Cfg::NCH, TW = Cfg::TW, HC = Cfg::HC, PITCH = Cfg::PITCH;
    constexpr int THREADS = Cfg::THREADS, KK = K * K;
    extern __shared__ __align__(128) unsigned char dsm[];
    const uint32_t GB = smem_u32(dsm);                      // G box -> dZ (in place)
    const uint32_t ZB = GB + B::GB_BYTES;                   // Z box; after the transform: dX staging tile
    constexpr int ZBB = WGO ? 0 : B::GB_BYTES;              // no Z box in backward-weight-only mode
    const uint32_t XB = ZB + ZBB;                           // raw X box
    unsigned char* after = dsm + B::GB_BYTES + ZBB + B::XB_BYTES;
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
    const bool do_wgrad = p.dw != nullptr, do_red = !WGO && p.nsums != nullptr && p.in_scale != nullptr;

    if (tid == 0) mbar_init(bar, 1);
    for (int i = tid; i < CG * KK; i += THREADS) {
        const int c = cbase + i / KK;
        wsm[i] = __float2bfloat16_rn(!WGO && c < p.C ? p.w[(size_t)c * KK + i % KK] : 0.f);
        dwacc[i] = 0.f;
    }
    for (int i = tid; i < CG; i += THREADS) {
        // the per-channel work of BatchNorm backward (what bn_bwd_finalize does), redundantly per CTA
        const int c = cbase + i;
        float a = 0.f, b = 0.f, c3 = 0.f, sc = 0.f, sh = 0.f, isc = 1.f, ish = 0.f;
        if (WGO) {
            if (c < p.C && p.in_scale) { isc = p.in_scale[c]; ish = p.in_shift[c]; }
        } else if (c < p.C) {
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
    // backward-weight operands: x4.trans = shifts (2q, 2q+1) x halves of a dZ row (same addresses as off4 / off2);
    // x2.trans of an X row = its two 8-pixel halves
    const uint32_t offx = (uint32_t)(((mi & 1) * 8 + r8 + strip * 16) * PITCH + chunk * 16);
    typedef DwbEmitReduce<K, CG, TWS> EM;
    EM em;
    em.out_lane = ZB + (uint32_t)((strip * 16 + g) * PITCH + chunk * 16 + t * 4);
    em.x_lane = XB + (uint32_t)((strip * 16 + g) * PITCH + chunk * 16 + t * 4);
    em.do_red = do_red;
    em.rs[0] = em.rs[1] = f2_pack(0.f, 0.f);
    float acc[K][4];
    uint32_t ph = 0;
    __syncthreads();
    {
        const int c = chunk * 8 + 2 * t;
        em.sp = f2_pack(coef[5 * CG + c], coef[5 * CG + c + 1]);
        em.tp = f2_pack(coef[6 * CG + c], coef[6 * CG + c + 1]);
    }
    // activation constants of the lane's channel g for the backward-weight B operand
    const f2_t ag = f2_pack(coef[5 * CG + chunk * 8 + g], coef[5 * CG + chunk * 8 + g]);
    const f2_t at = f2_pack(coef[6 * CG + chunk * 8 + g], coef[6 * CG + chunk * 8 + g]);
    const bool act_in = p.in_scale != nullptr;

    DwmStep cur;
    dwm_step_init<TW>(cur, slot, p.items, p.nseg, p.tiles_w, p.seg, p.nblocks);
    while (cur.item < p.items) {
        const bool pro = cur.j < cur.jb0;
        const int rbase = RB * cur.j + P;               // dZ row of block row 0
        const int row_first = pro ? RB * cur.jb0 - P : rbase;
        const bool has_rows = row_first < p.H;
        // backward-weight of a prologue's rows belongs to the segment above, except for the rows above block 0
        const bool wg_step = do_wgrad && (!pro || cur.jb0 == 0);
        // X rows RB*j ..: the backward-weight partners of this step's dZ rows, and the raw inputs of the dX rows this
        // step completes (which exist even when the block itself lies below the image: has_rows false)
        const bool need_x = (wg_step && has_rows) || (!pro && do_red && RB * cur.j < p.H);
        if (tid == 0) {
            tma_store_wait_read();                      // the previous step's dX store has finished reading ZB
            if (has_rows || need_x) {
                const int nr = pro ? K - 1 : RB;
                mbar_expect_tx(bar, (uint32_t)((has_rows ? (WGO ? 1 : 2) * nr * Cfg::ROWB : 0) +
                                               (need_x ? B::XROWS * Cfg::OROWB : 0)));
                if (has_rows) {
                    tma_load4(GB, pro ? &tm_gp : &tm_g, cbase, cur.w0 - P, row_first, cur.n, bar);
                    if (!WGO) tma_load4(ZB, pro ? &tm_zp : &tm_z, cbase, cur.w0 - P, row_first, cur.n, bar);
                }
                if (need_x) tma_load4(XB, &tm_x, cbase, cur.w0, RB * cur.j, cur.n, bar);
            }
        }
        if (has_rows || need_x) {
            mbar_wait(bar, ph);
            ph ^= 1;
            if (!WGO && tlive && has_rows) {
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
                const int npix = (pro ? K - 1 : RB) * HC;
                for (int pix = pix0; pix < npix; pix += PSTEP) {
                    const int rr = pix / HC, cc_ = pix - rr * HC;
                    const int ih = row_first + rr, iw = cur.w0 - P + cc_;
                    if ((unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W) {
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
        }
        __syncthreads();
        if (pro) {
#pragma unroll
            for (int i = 0; i < K; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        }
        const int sw = cur.w0 + strip * 16;
        if (chunk_live && sw < p.W) {
            const bool full = !pro && rbase + RB <= p.H;
            const float m0 = sw + g < p.W ? 1.f : 0.f, m1 = sw + g + 8 < p.W ? 1.f : 0.f;
            em.mk0 = f2_pack(m0, m0);
            em.mk1 = f2_pack(m1, m1);
            const uint32_t base = pro ? GB - (uint32_t)((RB - (K - 1)) * Cfg::ROWB) : GB;
            if constexpr (!WGO) {   // (a) backward-data: rotated diagonal weight fragments, rebuilt per step (dead during (b))
                uint32_t bd[K][K];
#pragma unroll
                for (int kh = 0; kh < K; ++kh)
#pragma unroll
                    for (int kw = 0; kw < K; ++kw) {
                        const uint32_t hb = (uint32_t)__bfloat16_as_ushort(wsm[(chunk * 8 + g) * KK + (K - 1 - kh) * K + (K - 1 - kw)]);
                        bd[kh][kw] = (g >> 1) == t ? ((g & 1) ? (hb << 16) : hb) : 0u;
                    }
                if (full && sw + 16 <= p.W) {
                    DwmRowsFull<K, CG, TWS, 0, false, EM>::run(acc, bd, base + off4, base + off2, em, 0u, 0u);
                } else if (full) {
                    DwmRowsFull<K, CG, TWS, 0, true, EM>::run(acc, bd, base + off4, base + off2, em, 0u, 0u);
                } else {
                    DwmRowsEdge<K, CG, TWS, 0, EM>::run(acc, bd, base + off4, base + off2, rbase, max(row_first, 0), p.H,
                                                        pro ? -2 * RB : RB * cur.j, em);
                }
            }
            if (wg_step && has_rows) {
                // (b) backward-weight.  dZ block row i (image row rbase + i) pairs with X block rows i .. i+K-1
                // (image rows RB*j + i + kh = dZ row + kh - P).
                DwbWg q;
                q.gz4 = base + off4; q.gz2 = base + off2; q.xrow = XB + offx;
                const int c0 = sw + 2 * t;      // this lane's two pixels of a half: columns 2t, 2t+1 (+8: second half)
                q.cm0 = (c0 < p.W ? 0x0000ffffu : 0u) | (c0 + 1 < p.W ? 0xffff0000u : 0u);
                q.cm1 = (c0 + 8 < p.W ? 0x0000ffffu : 0u) | (c0 + 9 < p.W ? 0xffff0000u : 0u);
                q.ag = ag; q.at = at; q.act = act_in;
                q.xrow_img0 = RB * cur.j; q.rbase = rbase; q.i0 = pro ? RB - (K - 1) : 0; q.H = p.H;
                float* dst = dwacc + (chunk * 8 + g) * KK;
                dwb_wgrad_pass<K, CG, TWS, 0, K>(q, dst, g, t);
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (!WGO && tid == 0 && !pro) tma_store4(&tm_dx, cbase, cur.w0, RB * cur.j, cur.n, ZB);
        if (++cur.j >= cur.jb1) dwm_step_init<TW>(cur, cur.item + nslots, p.items, p.nseg, p.tiles_w, p.seg, p.nblocks);
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

template <int K, int CG, int TWS, bool WGO>
static int launch_bwd_cfg(const DwmGeom& g, const DwbP& p0, const void* G, const void* Z, const void* X, void* dX,
                          cudaStream_t st, const char* name) {
    using Cfg = DwmCfg<K, CG, TWS>;
    using B = DwbCfg<K, CG, TWS>;
    constexpr int RB = Cfg::RB;
    DwbP p = p0;
    const int N = p.N, H = p.H, W = p.W, C = p.C;
    CUtensorMap tm_g, tm_gp, tm_z, tm_zp, tm_x, tm_dx;
    if (int e = dwm_tensor_map(&tm_g, G, N, H, W, C, CG, Cfg::HC, RB)) return e;
    if (int e = dwm_tensor_map(&tm_gp, G, N, H, W, C, CG, Cfg::HC, K - 1)) return e;
    if (int e = dwm_tensor_map(&tm_z, WGO ? G : Z, N, H, W, C, CG, Cfg::HC, RB)) return e;
    if (int e = dwm_tensor_map(&tm_zp, WGO ? G : Z, N, H, W, C, CG, Cfg::HC, K - 1)) return e;
    if (int e = dwm_tensor_map(&tm_x, X, N, H, W, C, CG, Cfg::TW, B::XROWS)) return e;
    if (int e = dwm_tensor_map(&tm_dx, WGO ? X : dX, N, H, W, C, CG, Cfg::TW, RB)) return e;
    p.tiles_w = g.tiles_w; p.cblocks = g.cblocks;
    p.nblocks = (H + RB - 1) / RB;
    constexpr int SMEM = B::SMEM - (WGO ? B::GB_BYTES : 0);
    static int occ = -1;
    if (occ < 0) {
        cudaError_t e = cudaFuncSetAttribute(dw_mma_bwd_k<K, CG, TWS, WGO>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return (int)e; }
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, dw_mma_bwd_k<K, CG, TWS, WGO>, B::THREADS, SMEM);
        if (o < 1) { set_error("%s: kernel does not fit on an SM", name); return MNB_ERR_UNSUPPORTED; }
        occ = o;
    }
    const long long cols = (long long)N * g.tiles_w;
    if (cols > (1 << 24)) { set_error("%s: too many columns", name); return MNB_ERR_UNSUPPORTED; }
    // resident CTAs only: a grid that exceeds SMs x occupancy runs its surplus CTAs as a second wave (2x the time)
    const long long slots_max = std::max(1LL, (long long)num_sms() * occ / g.cblocks);
    int seg = option_get(OPT_DW_MMA_SEG);
    if (seg <= 0) {
        seg = p.nblocks;
        while (seg > 2 && cols * ((p.nblocks + seg - 1) / seg) < 6 * slots_max) seg = (seg + 1) / 2;
    }
    if (seg > p.nblocks) seg = p.nblocks;
    p.seg = seg;
    p.nseg = (p.nblocks + seg - 1) / seg;
    p.items = (int)(cols * p.nseg);
    long long slots = slots_max;
    if (slots > p.items) slots = p.items;
    dw_mma_bwd_k<K, CG, TWS, WGO><<<(unsigned)(slots * g.cblocks), B::THREADS, SMEM, st>>>(tm_g, tm_gp, tm_z, tm_zp, tm_x, tm_dx, p);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

int dw_bwd_mma(const void* G, const void* Z, const float* scale, const float* shift, const double* sums, const float* mean,
               const float* invstd, double m, float* dgamma, float* dbeta, float* dbias, const void* X,
               const float* in_scale, const float* in_shift, const float* w, void* dX, float* dw, double* nsums, int N, int H,
               int W, int C, int k, cudaStream_t st) {
    DwbP p = {};
    p.scale = scale; p.shift = shift; p.sums = sums; p.mean = mean; p.invstd = invstd; p.m = m;
    p.dgamma = dgamma; p.dbeta = dbeta; p.dbias = dbias; p.in_scale = in_scale; p.in_shift = in_shift; p.w = w; p.dw = dw;
    p.nsums = nsums; p.N = N; p.H = H; p.W = W; p.C = C;
    const DwmGeom g = dwm_geometry(C, W, k);
    const char* name = "dw_bwd(mma)";
#define MNB_DWB(KK, CGG, TT) return launch_bwd_cfg<KK, CGG, TT, false>(g, p, G, Z, X, dX, st, name)
    if (k == 3) {
        if (g.CG == 24 && g.TWS == 1) MNB_DWB(3, 24, 1);
        if (g.CG == 24 && g.TWS == 2) MNB_DWB(3, 24, 2);
        if (g.CG == 40 && g.TWS == 1) MNB_DWB(3, 40, 1);
        MNB_DWB(3, 40, 2);
    } else if (k == 5) {
        if (g.CG == 24 && g.TWS == 1) MNB_DWB(5, 24, 1);
        if (g.CG == 24 && g.TWS == 2) MNB_DWB(5, 24, 2);
        if (g.CG == 40 && g.TWS == 1) MNB_DWB(5, 40, 1);
        MNB_DWB(5, 40, 2);
    }
#undef MNB_DWB
    return MNB_ERR_UNSUPPORTED;
}

template <int K, int CG, int TWS>
static int launch_wgrad_cfg(const DwmGeom& g, const DwbP& p0, const void* dZ, const void* X, cudaStream_t st,
                            const char* name) {
    using Cfg = DwmCfg<K, CG, TWS>;
    using B = DwbCfg<K, CG, TWS>;
    constexpr int RB = Cfg::RB;
    constexpr int SMEM = 2 * (B::GB_BYTES + B::XB_BYTES) + B::DW_BYTES + 16;
    DwbP p = p0;
    const int N = p.N, H = p.H, W = p.W, C = p.C;
    CUtensorMap tm_g, tm_gp, tm_x;
    if (int e = dwm_tensor_map(&tm_g, dZ, N, H, W, C, CG, Cfg::HC, RB)) return e;
    if (int e = dwm_tensor_map(&tm_gp, dZ, N, H, W, C, CG, Cfg::HC, K - 1)) return e;
    if (int e = dwm_tensor_map(&tm_x, X, N, H, W, C, CG, Cfg::TW, B::XROWS)) return e;
    p.tiles_w = g.tiles_w; p.cblocks = g.cblocks;
    p.nblocks = (H + RB - 1) / RB;
    static int occ = -1;
    if (occ < 0) {
        cudaError_t e = cudaFuncSetAttribute(dw_mma_wgrad_k<K, CG, TWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return (int)e; }
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, dw_mma_wgrad_k<K, CG, TWS>, B::THREADS, SMEM);
        if (o < 1) { set_error("%s: kernel does not fit on an SM", name); return MNB_ERR_UNSUPPORTED; }
        occ = o;
    }
    const long long cols = (long long)N * g.tiles_w;
    if (cols > (1 << 24)) { set_error("%s: too many columns", name); return MNB_ERR_UNSUPPORTED; }
    // resident CTAs only: a grid that exceeds SMs x occupancy runs its surplus CTAs as a second wave (2x the time)
    const long long slots_max = std::max(1LL, (long long)num_sms() * occ / g.cblocks);
    int seg = option_get(OPT_DW_MMA_SEG);
    if (seg <= 0) {
        seg = p.nblocks;
        while (seg > 2 && cols * ((p.nblocks + seg - 1) / seg) < 6 * slots_max) seg = (seg + 1) / 2;
    }
    if (seg > p.nblocks) seg = p.nblocks;
    p.seg = seg;
    p.nseg = (p.nblocks + seg - 1) / seg;
    p.items = (int)(cols * p.nseg);
    long long slots = slots_max;
    if (slots > p.items) slots = p.items;
    dw_mma_wgrad_k<K, CG, TWS><<<(unsigned)(slots * g.cblocks), B::THREADS, SMEM, st>>>(tm_g, tm_gp, tm_x, p);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

// backward-weight alone on the tensor pipe: dw[c,kh,kw] += sum dz * relu(s*x+t) shifted
int dw_wgrad_mma(const void* x, const float* s, const float* t, const void* dz, float* dw, int N, int H, int W, int C, int k,
                 cudaStream_t st) {
    DwbP p = {};
    p.in_scale = s; p.in_shift = t; p.dw = dw; p.N = N; p.H = H; p.W = W; p.C = C; p.m = 1.0;
    const DwmGeom g = dwm_geometry(C, W, k);
    const char* name = "dw_wgrad(mma)";
#define MNB_DWB(KK, CGG, TT) return launch_wgrad_cfg<KK, CGG, TT>(g, p, dz, x, st, name)
    if (k == 3) {
        if (g.CG == 24 && g.TWS == 1) MNB_DWB(3, 24, 1);
        if (g.CG == 24 && g.TWS == 2) MNB_DWB(3, 24, 2);
        if (g.CG == 40 && g.TWS == 1) MNB_DWB(3, 40, 1);
        MNB_DWB(3, 40, 2);
    } else if (k == 5) {
        if (g.CG == 24 && g.TWS == 1) MNB_DWB(5, 24, 1);
        if (g.CG == 24 && g.TWS == 2) MNB_DWB(5, 24, 2);
        if (g.CG == 40 && g.TWS == 1) MNB_DWB(5, 40, 1);
        MNB_DWB(5, 40, 2);
    }
#undef MNB_DWB
    return MNB_ERR_UNSUPPORTED;
}

}  // namespace mnb
