// tcgen05 / TMEM implicit-GEMM kernels for the dense convolutions (1x1 and 3x3, stride 1/2) in bf16:
// forward, backward-data and backward-weight.  sm_100a only (tcgen05.mma / tcgen05.ld / tcgen05.alloc).
//
// Why the operands are staged by threads and not by TMA: every A operand on this path needs an elementwise
// prologue between HBM and the tensor core -- relu(scale*z+shift) of the producing ConvBlock's BatchNorm
// (fwd, wgrad), an im2col gather with zero padding applied AFTER that transform (3x3), -- so the data has to
// pass through registers anyway.  Threads load 16-byte NHWC channel vectors (coalesced, 64-B runs per row),
// transform in fp32, and store 16-byte vectors straight into the canonical no-swizzle UMMA core-matrix
// layout (8 rows x 16 B = 128 contiguous bytes per core matrix), conflict-free.  One elected thread issues
// tcgen05.mma (M=128, N<=128, K=16 per instruction) with the fp32 accumulator in TMEM; completion is tracked
// with tcgen05.commit -> mbarrier.  The epilogue reads TMEM with tcgen05.ld (32x32b: thread == row), adds
// bias / the residual gradient, rounds to bf16, stages the tile in shared memory, writes it out with
// coalesced 16-byte stores and accumulates the BatchNorm sum / sum-of-squares of the stored values per column.
// Several CTAs are resident per SM (<= 85 KB smem, <= 128 TMEM columns each) so one CTA's loads overlap
// another's MMA / epilogue.  These GEMMs have K,N in 16..1152 and are HBM-bound.
//
//   FWD    D[m=(n,ho,wo)][co]   = sum_kk a(m,kk) w(co,kk)      A,B K-major     kk = (kh,kw,ci)
//   DGRAD  D[m=(n,h,w)][ci]     = sum_kk dz(m,kk) w(ci,kk)     A,B K-major     kk = (kh,kw,co)
//   WGRAD  D[co][kk=(kh,kw,ci)] = sum_pos dz(pos,co) a(pos,kk) A,B MN-major    K = output positions, split-K
#include "conv_params.cuh"

namespace mnb {

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, addr = smem_u32(bar);
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32, A=B=bf16, M=128, N=n, majors (0 = K-major, 1 = MN-major)
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

constexpr int KC = 64;            // K elements staged per chunk (4 MMAs of K=16)
constexpr int TC_THREADS = 128;   // 4 warps: warp w reads TMEM lanes 32w..32w+31

struct TcGeom {
    int BN;          // N tile (multiple of 16, <= 128)
    int n_tiles;     // tiles along N
    int tmem_cols;   // power of two >= 32
    long long m_tiles;
    int ksplit;      // wgrad: K splits
    long long kchunks_per_split;
};

// ---- activation gather: 8 channels of a(n, ih, iw, ci..ci+7) -> packed bf16x8, transform optional --------
__device__ __forceinline__ uint4 gather_act8(const ConvP& p, const float* s_scale, const float* s_shift, int n, int ih,
                                             int iw, int ci) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) return u;      // zero padding applies to the ACTIVATION
    u = *reinterpret_cast<const uint4*>((const bf16*)p.x + (((long long)n * p.H + ih) * p.W + iw) * p.Cin + ci);
    if (s_scale) {
        float v[8];
        unpack_bf16x8(u, v);
        float4 sa = *reinterpret_cast<const float4*>(s_scale + ci), sb = *reinterpret_cast<const float4*>(s_scale + ci + 4);
        float4 ta = *reinterpret_cast<const float4*>(s_shift + ci), tb = *reinterpret_cast<const float4*>(s_shift + ci + 4);
        v[0] = fmaxf(fmaf(sa.x, v[0], ta.x), 0.f); v[1] = fmaxf(fmaf(sa.y, v[1], ta.y), 0.f);
        v[2] = fmaxf(fmaf(sa.z, v[2], ta.z), 0.f); v[3] = fmaxf(fmaf(sa.w, v[3], ta.w), 0.f);
        v[4] = fmaxf(fmaf(sb.x, v[4], tb.x), 0.f); v[5] = fmaxf(fmaf(sb.y, v[5], tb.y), 0.f);
        v[6] = fmaxf(fmaf(sb.z, v[6], tb.z), 0.f); v[7] = fmaxf(fmaf(sb.w, v[7], tb.w), 0.f);
        u = pack_bf16x8(v);
    }
    return u;
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_k(ConvP p, TcGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = g.BN;
    const int kk2 = p.k * p.k;
    // ---- shared memory carve-up ----
    unsigned char* sA = smem_raw;                               // 128 x KC bf16 = 16 KB
    unsigned char* sB = sA + 128 * KC * 2;                      // BN x KC bf16
    unsigned char* sC = sB + BN * KC * 2;                       // epilogue stage: 128 rows x (BN*2+16) B
    const int c_pitch = BN * 2 + 16;
    float* s_scale = reinterpret_cast<float*>(sC + 128 * c_pitch);
    const int xch = (MODE == MODE_DGRAD) ? 0 : p.Cin;           // channels of the transformed operand
    float* s_shift = s_scale + xch;
    float* s_red = s_shift + xch;                               // [4][BN/2 * 2]... sized 2*BN floats
    uint64_t* mbar = reinterpret_cast<uint64_t*>(s_red + 2 * 256);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(mbar + 1);

    const bool xf = (MODE != MODE_DGRAD) && p.in_scale != nullptr;
    if (xf)
        for (int i = tid; i < p.Cin; i += TC_THREADS) { s_scale[i] = p.in_scale[i]; s_shift[i] = p.in_shift[i]; }
    if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tmem_holder, (uint32_t)g.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const float* xs = xf ? s_scale : nullptr;
    const float* xt = xf ? s_shift : nullptr;

    long long Mtot, Ktot;
    int Ntot;
    if (MODE == MODE_FWD) { Mtot = (long long)p.N * p.Ho * p.Wo; Ntot = p.Cout; Ktot = (long long)kk2 * p.Cin; }
    else if (MODE == MODE_DGRAD) { Mtot = (long long)p.N * p.H * p.W; Ntot = p.Cin; Ktot = (long long)kk2 * p.Cout; }
    else { Mtot = p.Cout; Ntot = kk2 * p.Cin; Ktot = (long long)p.N * p.Ho * p.Wo; }

    const uint32_t idesc = make_idesc(BN, MODE == MODE_WGRAD, MODE == MODE_WGRAD);
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    uint32_t phase = 0;
    float st[4] = {0.f, 0.f, 0.f, 0.f};                         // FWD: per-thread column-pair statistics

    const long long units = g.m_tiles * g.n_tiles * (MODE == MODE_WGRAD ? g.ksplit : 1);
    for (long long unit = blockIdx.x; unit < units; unit += gridDim.x) {
        long long u = unit;
        int ks = 0;
        if (MODE == MODE_WGRAD) { ks = (int)(u % g.ksplit); u /= g.ksplit; }
        const int nt = (int)(u % g.n_tiles);
        const long long mt = u / g.n_tiles;
        const long long m0 = mt * 128;
        const int n0 = nt * BN;
        long long kbeg = 0, kend = Ktot;
        if (MODE == MODE_WGRAD) {
            kbeg = (long long)ks * g.kchunks_per_split * KC;
            kend = kbeg + g.kchunks_per_split * KC;
            if (kend > Ktot) kend = Ktot;
            if (kbeg >= kend) continue;       // uniform across the CTA
        }
        bool first = true;
        for (long long k0 = kbeg; k0 < kend; k0 += KC) {
            const int kvalid = (int)((kend - k0) < KC ? (kend - k0) : KC);
            const int k16 = (kvalid + 15) / 16;                 // MMAs this chunk
            if (MODE != MODE_WGRAD) {
                // ---- K-major staging: vec(row, k8) -> ((k8*ROWS)+row)*16 ----
                const int nk8 = k16 * 2;
                // A: 16 row-blocks(8 rows) x 2 k8-blocks(4)
                for (int idx = warp; idx < 32; idx += 4) {
                    const int rb = idx >> 1, kb = idx & 1;
                    const int row = rb * 8 + (lane & 7), k8 = kb * 4 + (lane >> 3);
                    if (k8 >= nk8) continue;
                    uint4 v = make_uint4(0, 0, 0, 0);
                    const long long m = m0 + row;
                    const long long kk = k0 + k8 * 8;
                    if (m < Mtot && kk < kend) {
                        if (MODE == MODE_FWD) {
                            const int wo = (int)(m % p.Wo), ho = (int)((m / p.Wo) % p.Ho);
                            const int n = (int)(m / ((long long)p.Wo * p.Ho));
                            const int tap = (int)(kk / p.Cin), ci = (int)(kk % p.Cin);
                            const int kh = tap / p.k, kw = tap % p.k;
                            v = gather_act8(p, xs, xt, n, ho * p.stride - p.pad + kh, wo * p.stride - p.pad + kw, ci);
                        } else {
                            const int w_ = (int)(m % p.W), h_ = (int)((m / p.W) % p.H);
                            const int n = (int)(m / ((long long)p.W * p.H));
                            const int tap = (int)(kk / p.Cout), co = (int)(kk % p.Cout);
                            const int kh = tap / p.k, kw = tap % p.k;
                            const int hn = h_ + p.pad - kh, wn = w_ + p.pad - kw;
                            if (hn >= 0 && wn >= 0 && hn % p.stride == 0 && wn % p.stride == 0) {
                                const int ho = hn / p.stride, wo = wn / p.stride;
                                if (ho < p.Ho && wo < p.Wo)
                                    v = *reinterpret_cast<const uint4*>(
                                        (const bf16*)p.dz + (((long long)n * p.Ho + ho) * p.Wo + wo) * p.Cout + co);
                            }
                        }
                    }
                    *reinterpret_cast<uint4*>(sA + ((k8 * 128) + row) * 16) = v;
                }
                // B: (BN/8) col-blocks x 2 k8-blocks
                const int ncb = BN >> 3;
                for (int idx = warp; idx < ncb * 2; idx += 4) {
                    const int cb = idx >> 1, kb = idx & 1;
                    const int col = cb * 8 + (lane & 7), k8 = kb * 4 + (lane >> 3);
                    if (k8 >= nk8) continue;
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.f;
                    const int nn = n0 + col;
                    const long long kk = k0 + k8 * 8;
                    if (nn < Ntot && kk < kend) {
                        if (MODE == MODE_FWD) {
                            const int tap = (int)(kk / p.Cin), ci = (int)(kk % p.Cin);
                            const float* wp = p.w + ((long long)nn * p.Cin + ci) * kk2 + tap;
                            if (kk2 == 1) load8(wp, v);
                            else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = wp[i * kk2];
                            }
                        } else {
                            const int tap = (int)(kk / p.Cout), co = (int)(kk % p.Cout);
                            const float* wp = p.w + ((long long)co * p.Cin + nn) * kk2 + tap;
                            const long long cs = (long long)p.Cin * kk2;
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = wp[i * cs];
                        }
                    }
                    *reinterpret_cast<uint4*>(sB + ((k8 * BN) + col) * 16) = pack_bf16x8(v);
                }
            } else {
                // ---- MN-major staging: vec(mn8, p) -> ((mn8*KC)+p)*16 ; K (=positions) zero-filled past kvalid ----
                const int np8 = k16 * 2;                        // 8-position blocks that the MMAs will read
                // A = dz^T : 16 mn8 (co/8) x KC positions ; combos: 8 p-blocks x 4 mn8-blocks
                for (int idx = warp; idx < 32; idx += 4) {
                    const int pb = idx & 7, mb = idx >> 3;
                    if (pb >= np8) continue;
                    const int pp = pb * 8 + (lane & 7), mn8 = mb * 4 + (lane >> 3);
                    uint4 v = make_uint4(0, 0, 0, 0);
                    const long long pos = k0 + pp;
                    const long long co = m0 + mn8 * 8;
                    if (pos < kend && co < Mtot)
                        v = *reinterpret_cast<const uint4*>((const bf16*)p.dz + pos * p.Cout + co);
                    *reinterpret_cast<uint4*>(sA + ((mn8 * KC) + pp) * 16) = v;
                }
                // B = a^T : (BN/8) mn8 (kk/8) x KC positions
                const int nmb = ((BN >> 3) + 3) >> 2;
                for (int idx = warp; idx < 8 * nmb; idx += 4) {
                    const int pb = idx & 7, mb = idx >> 3;
                    if (pb >= np8) continue;
                    const int pp = pb * 8 + (lane & 7), mn8 = mb * 4 + (lane >> 3);
                    if (mn8 >= (BN >> 3)) continue;
                    uint4 v = make_uint4(0, 0, 0, 0);
                    const long long pos = k0 + pp;
                    const int kk = n0 + mn8 * 8;
                    if (pos < kend && kk < Ntot) {
                        const int wo = (int)(pos % p.Wo), ho = (int)((pos / p.Wo) % p.Ho);
                        const int n = (int)(pos / ((long long)p.Wo * p.Ho));
                        const int tap = kk / p.Cin, ci = kk % p.Cin;
                        const int kh = tap / p.k, kw = tap % p.k;
                        v = gather_act8(p, xs, xt, n, ho * p.stride - p.pad + kh, wo * p.stride - p.pad + kw, ci);
                    }
                    *reinterpret_cast<uint4*>(sB + ((mn8 * KC) + pp) * 16) = v;
                }
            }
            fence_async_proxy();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                for (int j = 0; j < k16; ++j) {
                    uint64_t ad, bd;
                    if (MODE != MODE_WGRAD) {
                        ad = make_desc(sA_u + j * 2 * (128 * 16), 128 * 16, 128);
                        bd = make_desc(sB_u + j * 2 * (BN * 16), BN * 16, 128);
                    } else {
                        ad = make_desc(sA_u + j * 256, 128, KC * 16);
                        bd = make_desc(sB_u + j * 256, 128, KC * 16);
                    }
                    umma_bf16(tmem_base, ad, bd, idesc, (first && j == 0) ? 0u : 1u);
                }
                umma_commit(mbar);
            }
            first = false;
            mbar_wait(mbar, phase);           // MMAs done: smem reusable, accumulator up to date
            phase ^= 1;
        }
        tc_fence_after();
        // ---- epilogue: thread == accumulator row ----
        const long long row_g = m0 + tid;
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        if (MODE == MODE_WGRAD) {
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(t_lane + c0, r);
                if (row_g < Mtot) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int kk = n0 + c0 + j;
                        if (kk < Ntot) {
                            const int tap = kk / p.Cin, ci = kk % p.Cin;
                            atomicAdd(&p.dw[((long long)row_g * p.Cin + ci) * kk2 + tap], __uint_as_float(r[j]));
                        }
                    }
                }
            }
        } else {
            const int ldo = (MODE == MODE_FWD) ? p.Cout : p.Cin;
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(t_lane + c0, r);
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                if (MODE == MODE_FWD) {
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += (n0 + c0 + j < Ntot) ? p.bias[n0 + c0 + j] : 0.f;
                    }
                } else if (p.add && row_g < Mtot) {
                    const bf16* ap = (const bf16*)p.add + row_g * ldo + n0 + c0;
                    if (n0 + c0 + 16 <= Ntot) {
                        float a0[8], a1[8];
                        load8(ap, a0);
                        load8(ap + 8, a1);
#pragma unroll
                        for (int j = 0; j < 8; ++j) { v[j] += a0[j]; v[8 + j] += a1[j]; }
                    } else {
                        for (int j = 0; j < 16; ++j)
                            if (n0 + c0 + j < Ntot) v[j] += to_f(ap[j]);
                    }
                }
                float lo[8], hi[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
                unsigned char* dst = sC + tid * c_pitch + c0 * 2;
                *reinterpret_cast<uint4*>(dst) = pack_bf16x8(lo);
                *reinterpret_cast<uint4*>(dst + 16) = pack_bf16x8(hi);
            }
            tc_fence_before();
            __syncthreads();
            // coalesced write-out of the staged tile
            const int vpr = BN >> 3;                            // 16-B vectors per row
            for (int vi = tid; vi < 128 * vpr; vi += TC_THREADS) {
                const int row = vi / vpr, c8 = vi % vpr;
                const long long m = m0 + row;
                const int col = n0 + c8 * 8;
                if (m < Mtot && col < Ntot) {                   // Ntot % 8 == 0
                    uint4 q = *reinterpret_cast<const uint4*>(sC + row * c_pitch + c8 * 16);
                    *reinterpret_cast<uint4*>((bf16*)p.out + m * ldo + col) = q;
                }
            }
            if (MODE == MODE_FWD && p.stats) {
                // column-pair owner threads: sum / sum of squares of the STORED (bf16-rounded) values
                const int ncp = BN >> 1;                        // <= 128 column pairs
                const int groups = TC_THREADS / ncp;            // row groups sharing a column pair
                const int cpi = tid % ncp, rg = tid / ncp;
                if (rg < groups) {
                    const int rows_here = (int)(Mtot - m0 < 128 ? Mtot - m0 : 128);
                    for (int row = rg; row < rows_here; row += groups) {
                        uint32_t w2 = *reinterpret_cast<const uint32_t*>(sC + row * c_pitch + cpi * 4);
                        float a = __uint_as_float(w2 << 16), b = __uint_as_float(w2 & 0xffff0000u);
                        st[0] += a; st[1] += b;
                        st[2] = fmaf(a, a, st[2]); st[3] = fmaf(b, b, st[3]);
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();      // TMEM + stage buffer free for the next unit
        tc_fence_after();
        if (MODE == MODE_FWD && p.stats && g.n_tiles > 1) {
            // the column assignment changes with the N tile: flush per unit
            const int ncp = BN >> 1;
            for (int i = tid; i < 2 * BN; i += TC_THREADS) s_red[i] = 0.f;
            __syncthreads();
            if (tid / ncp < TC_THREADS / ncp) {
                const int cpi = tid % ncp;
                atomicAdd(&s_red[cpi * 2], st[0]); atomicAdd(&s_red[cpi * 2 + 1], st[1]);
                atomicAdd(&s_red[BN + cpi * 2], st[2]); atomicAdd(&s_red[BN + cpi * 2 + 1], st[3]);
            }
            st[0] = st[1] = st[2] = st[3] = 0.f;
            __syncthreads();
            for (int i = tid; i < BN; i += TC_THREADS)
                if (n0 + i < Ntot) {
                    atomicAdd(&p.stats[n0 + i], (double)s_red[i]);
                    atomicAdd(&p.stats[p.Cout + n0 + i], (double)s_red[BN + i]);
                }
            __syncthreads();
        }
    }
    if (MODE == MODE_FWD && p.stats && g.n_tiles == 1) {
        const int ncp = BN >> 1;
        for (int i = tid; i < 2 * BN; i += TC_THREADS) s_red[i] = 0.f;
        __syncthreads();
        if (tid / ncp < TC_THREADS / ncp) {
            const int cpi = tid % ncp;
            atomicAdd(&s_red[cpi * 2], st[0]); atomicAdd(&s_red[cpi * 2 + 1], st[1]);
            atomicAdd(&s_red[BN + cpi * 2], st[2]); atomicAdd(&s_red[BN + cpi * 2 + 1], st[3]);
        }
        __syncthreads();
        for (int i = tid; i < BN; i += TC_THREADS)
            if (i < Ntot) {
                atomicAdd(&p.stats[i], (double)s_red[i]);
                atomicAdd(&p.stats[p.Cout + i], (double)s_red[BN + i]);
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

static int pow2_cols(int n) {
    int c = 32;
    while (c < n) c <<= 1;
    return c;
}

static bool tc_geom(int mode, const ConvP& p, TcGeom& g, size_t& smem) {
    const int kk2 = p.k * p.k;
    long long M;
    int Nn;
    if (mode == MODE_FWD) { M = (long long)p.N * p.Ho * p.Wo; Nn = p.Cout; }
    else if (mode == MODE_DGRAD) { M = (long long)p.N * p.H * p.W; Nn = p.Cin; }
    else { M = p.Cout; Nn = kk2 * p.Cin; }
    if (p.nchw_in || p.Cin % 8 != 0 || p.Cout % 8 != 0) return false;
    int nt = (Nn + 127) / 128;
    int BN = ((Nn + nt - 1) / nt + 15) / 16 * 16;
    g.BN = BN;
    g.n_tiles = (Nn + BN - 1) / BN;
    g.tmem_cols = pow2_cols(BN);
    g.m_tiles = (M + 127) / 128;
    g.ksplit = 1;
    g.kchunks_per_split = 0;
    const int xch = mode == MODE_DGRAD ? 0 : p.Cin;
    smem = (size_t)128 * KC * 2 + (size_t)BN * KC * 2 + (size_t)128 * (BN * 2 + 16) + (size_t)2 * xch * 4 +
           2 * 256 * 4 + 64;
    return true;
}

template <int MODE>
static int launch_tc(const ConvP& p, cudaStream_t st, const char* name) {
    if (!mnb_device_is_sm100()) { set_error("%s: device is not sm_100", name); return MNB_ERR_UNSUPPORTED; }
    TcGeom g;
    size_t smem;
    if (!tc_geom(MODE, p, g, smem)) { set_error("%s: shape not covered", name); return MNB_ERR_UNSUPPORTED; }
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    const int sms = num_sms();
    int ctas_per_sm = (int)(200 * 1024 / (smem + 1024));
    if (ctas_per_sm > 512 / g.tmem_cols) ctas_per_sm = 512 / g.tmem_cols;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    long long slots = (long long)sms * ctas_per_sm;
    long long tiles = g.m_tiles * g.n_tiles;
    if (MODE == MODE_WGRAD) {
        long long K = (long long)p.N * p.Ho * p.Wo;
        long long chunks = (K + KC - 1) / KC;
        long long want = slots / tiles;
        if (want < 1) want = 1;
        if (want > chunks) want = chunks;
        g.kchunks_per_split = (chunks + want - 1) / want;
        g.ksplit = (int)((chunks + g.kchunks_per_split - 1) / g.kchunks_per_split);
        tiles *= g.ksplit;
    }
    long long grid = tiles < slots ? tiles : slots;
    conv_tc_k<MODE><<<(unsigned)grid, TC_THREADS, smem, st>>>(p, g);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

int conv_fwd_tc(const ConvP& p, cudaStream_t st) { return launch_tc<MODE_FWD>(p, st, "conv_fwd(tcgen05)"); }
int conv_dgrad_tc(const ConvP& p, cudaStream_t st) { return launch_tc<MODE_DGRAD>(p, st, "conv_dgrad(tcgen05)"); }
int conv_wgrad_tc(const ConvP& p, cudaStream_t st) { return launch_tc<MODE_WGRAD>(p, st, "conv_wgrad(tcgen05)"); }

}  // namespace mnb
