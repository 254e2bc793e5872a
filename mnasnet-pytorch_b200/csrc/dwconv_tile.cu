// Depthwise k x k (3 / 5) stride-1 convolution, bf16 NHWC: shared-memory halo-tile kernels (forward, dgrad,
// wgrad).  Replaces nn.Conv2d(groups=C) (src/models/mnasnet.py:76-81,120-125) with the producing ConvBlock's
// BN-apply+ReLU fused into the tile load and this ConvBlock's BN statistics fused into the store.
//
// A CTA (256 threads) owns one channel block CB (64/32/24/16 channels) and walks spatial tiles of
// (CW*8) x 16 outputs persistently.  Per tile the (rows+k-1) x (16+k-1) x CB input halo is fetched with
// 16-byte cp.async (zero-fill outside the image = the conv's zero padding), DOUBLE-BUFFERED: the loads of
// tile i+1 are in flight while tile i is computed.  Each thread then applies relu(scale*x+shift) in place to the
// vectors it fetched itself (padding stays exactly 0), and after one __syncthreads the compute phase runs from
// shared memory: lanes run along channel pairs (conflict-free 4-byte ld.shared, 128-byte coalesced st.global),
// a thread owns 2 output columns x 8 rows of one channel pair, input rows slide through registers so every
// ld.shared feeds k*2 FMAs, weights (2*k*k floats) live in registers.  fp32 accumulate; statistics of the
// bf16-rounded outputs are kept in registers across tiles and flushed once per CTA (fp64 atomics).
#include "common.cuh"

namespace mnb {

__device__ __forceinline__ uint32_t dsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dcp_async16(void* dst_smem, const void* src, bool pred) {
    const uint32_t n = pred ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dsmem_u32(dst_smem)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void dcp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void dcp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Blackwell packed fp32: one FFMA2 does the FMA of BOTH channels of a pair (halves the FMA issue slots of the
// FMA-bound 5x5 layers: 25 MAC per 4 bytes of traffic)
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
__device__ __forceinline__ f2_t f2_from_bf16x2(uint32_t u) {
    return f2_pack(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

struct DwP {
    const bf16* x;          // fwd: input z of previous block; dgrad: dz; wgrad: input
    const float* in_scale;
    const float* in_shift;
    const float* w;         // [C][k][k]
    const float* bias;
    const bf16* dz;         // wgrad
    bf16* out;              // fwd: z ; dgrad: dx
    float* dw;              // wgrad
    double* stats;          // fwd: BN statistics ; dgrad: fused BN-backward sums of the producer block
    const bf16* bn_z;       // dgrad: producer block's raw conv output (NULL = no fused reduction)
    const float* bn_scale;
    const float* bn_shift;
    int N, H, W, C;
    int tiles_h, tiles_w;
    long long items;        // N * tiles_h * tiles_w
};

template <int K, int CB>
struct DwCfg {
    static constexpr int P = K / 2;
    static constexpr int NP = CB / 2;               // channel pairs per block
    static constexpr int CW = 32 / NP;              // output column groups per warp (each group = 2 columns)
    static constexpr int WX = 8 / CW;               // warps along W  (tile width = WX*CW*2 = 16)
    static constexpr int WY = CW;                   // warps along H
    static constexpr int TH = 7;                    // rows per thread (every MNASNet-224 map height is a multiple of 7)
    static constexpr int TW = 16;                   // tile width
    static constexpr int THT = WY * TH;             // tile height
    static constexpr int HR = THT + K - 1, HC = TW + K - 1;
    static constexpr int CV8 = CB / 8;              // 16-byte vectors per pixel
    // pixel pitch in shared memory, padded so that the CW column groups of a warp (2 pixels apart) fall into
    // disjoint banks: 2*PP mod 128 must clear the NP*4 bytes a column group reads
    static constexpr int PP = CB == 64 ? 128 : (CB == 32 ? 96 : (CB == 24 ? 96 : 48));
    static constexpr int NV = HR * HC * CV8;        // halo vectors per tile
    static constexpr int MAXV = (NV + 255) / 256;
    static constexpr int TILE_BYTES = HR * HC * PP;
    static constexpr int NVD = THT * TW * CV8;      // dz vectors per tile (wgrad)
    static constexpr int MAXVD = (NVD + 255) / 256;
    static constexpr int DZ_BYTES = THT * TW * PP;
};

__device__ __forceinline__ uint4 dw_xform8(uint4 u, const float* s, const float* t) {
    float v[8];
    unpack_bf16x8(u, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(fmaf(s[i], v[i], t[i]), 0.f);
    return pack_bf16x8(v);
}

// MODE 0: forward (XF optional, bias, stats)   MODE 1: dgrad (flipped kernel, plain input)   MODE 2: wgrad
// MODE 3: dgrad + fused BN-backward reduction of the producer block (second, halo-free tile holds its z)
template <int K, int CB, int MODE>
__global__ void __launch_bounds__(256, (K == 3 ? 3 : 2)) dw_tile_k(DwP p) {
    using Cfg = DwCfg<K, CB>;
    constexpr int P = Cfg::P, NP = Cfg::NP, CW = Cfg::CW, WX = Cfg::WX, TH = Cfg::TH, HC = Cfg::HC, HR = Cfg::HR;
    constexpr int CV8 = Cfg::CV8, NV = Cfg::NV, MAXV = Cfg::MAXV, THT = Cfg::THT, TW = Cfg::TW, PP = Cfg::PP;
    extern __shared__ __align__(128) unsigned char dsm[];
    constexpr bool SIDE = (MODE == 2 || MODE == 3);  // wgrad: dz tile; fused dgrad: bn_z tile
    constexpr bool DGRAD = (MODE == 1 || MODE == 3);
    constexpr int STAGE = Cfg::TILE_BYTES + (SIDE ? Cfg::DZ_BYTES : 0);
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    // 1-D grid, channel block fastest: the CTAs that own the different channel blocks of the SAME spatial tile
    // are co-scheduled and walk the tiles in the same order, so a pixel's bytes are fetched from DRAM once (as
    // whole bursts) and the sibling CTAs hit in L2
    const int cblocks = p.C / CB;
    const int cbase = (blockIdx.x % cblocks) * CB;
    const int slot = blockIdx.x / cblocks, nslots = gridDim.x / cblocks;
    const int cp = lane % NP, csub = lane / NP;
    const bool lane_on = csub < CW;
    const int wx = wid % WX, wy = wid / WX;
    const int c0 = (wx * CW + csub) * 2;            // first of the 2 tile-local output columns of this thread
    const int r0 = wy * TH;                         // first tile-local output row
    const int ch = cbase + cp * 2;                  // global channel of the pair
    const bool xf = !DGRAD && p.in_scale != nullptr;

    // per-thread constants
    f2_t wr[K][K];
    if (MODE != 2) {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int ii = DGRAD ? K - 1 - i : i, jj = DGRAD ? K - 1 - j : j;
                wr[i][j] = f2_pack(lane_on ? p.w[(ch + 0) * K * K + ii * K + jj] : 0.f,
                                   lane_on ? p.w[(ch + 1) * K * K + ii * K + jj] : 0.f);
            }
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) wr[i][j] = f2_pack(0.f, 0.f);
    }
    float b0 = 0.f, b1 = 0.f;
    if (MODE == 0 && p.bias && lane_on) { b0 = p.bias[ch]; b1 = p.bias[ch + 1]; }
    float st[4] = {0.f, 0.f, 0.f, 0.f};
    const bool bn_red = (MODE == 3);
    float gs0 = 0.f, gs1 = 0.f, gt0 = 0.f, gt1 = 0.f;
    if (bn_red && lane_on) { gs0 = p.bn_scale[ch]; gs1 = p.bn_scale[ch + 1]; gt0 = p.bn_shift[ch]; gt1 = p.bn_shift[ch + 1]; }

    // ---- tile loader.  The decomposition of this thread's vector slots v = tid + i*256 -> (halo row, halo col,
    //      8-channel group) does not depend on the tile: do the divisions once, outside the tile loop. ----
    int s_off[MAXV], s_rc[MAXV];                    // smem byte offset ; (r << 16) | (c << 8) | cv
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int v = tid + i * 256;
        const int cv = v % CV8, rc = v / CV8;
        const int c = rc % HC, r = rc / HC;
        s_off[i] = v < NV ? rc * PP + cv * 16 : -1;
        s_rc[i] = (r << 16) | (c << 8) | cv;
    }
    int d_off[Cfg::MAXVD], d_rc[Cfg::MAXVD];
    const bf16* side = MODE == 2 ? p.dz : p.bn_z;      // second (halo-free) tile: dz for wgrad, bn_z for fused dgrad
    if (SIDE) {
#pragma unroll
        for (int i = 0; i < Cfg::MAXVD; ++i) {
            const int v = tid + i * 256;
            const int cv = v % CV8, rc = v / CV8;
            const int c = rc % TW, r = rc / TW;
            d_off[i] = v < Cfg::NVD ? rc * PP + cv * 16 : -1;
            d_rc[i] = (r << 16) | (c << 8) | cv;
        }
    }
    constexpr bool CV_FIXED = (256 % CV8) == 0;     // then every slot of a thread has the same channel group
    float xs[8], xt[8];
    if (xf && CV_FIXED) {
        load8(p.in_scale + cbase + (tid % CV8) * 8, xs);
        load8(p.in_shift + cbase + (tid % CV8) * 8, xt);
    }
    auto issue = [&](long long item, int stage) -> unsigned {
        unsigned char* tile = dsm + stage * STAGE;
        const int tw_i = (int)(item % p.tiles_w);
        const int th_i = (int)((item / p.tiles_w) % p.tiles_h);
        const int n = (int)(item / ((long long)p.tiles_w * p.tiles_h));
        const int h0 = th_i * THT - P, w0 = tw_i * TW - P;
        const bf16* xn = p.x + (long long)n * p.H * p.W * p.C + cbase;
        unsigned mask = 0;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            if (s_off[i] >= 0) {
                const int ih = h0 + (s_rc[i] >> 16), iw = w0 + ((s_rc[i] >> 8) & 0xff), cv = s_rc[i] & 0xff;
                const bool pred = (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
                dcp_async16(tile + s_off[i], pred ? xn + (long long)(ih * p.W + iw) * p.C + cv * 8 : p.x, pred);
                mask |= (pred ? 1u : 0u) << i;
            }
        }
        if (SIDE) {
            unsigned char* dzt = tile + Cfg::TILE_BYTES;
            const bf16* dn = side + (long long)n * p.H * p.W * p.C + cbase;
#pragma unroll
            for (int i = 0; i < Cfg::MAXVD; ++i) {
                if (d_off[i] >= 0) {
                    const int oh = th_i * THT + (d_rc[i] >> 16), ow = tw_i * TW + ((d_rc[i] >> 8) & 0xff), cv = d_rc[i] & 0xff;
                    const bool pred = oh < p.H && ow < p.W;
                    dcp_async16(dzt + d_off[i], pred ? dn + (long long)(oh * p.W + ow) * p.C + cv * 8 : side, pred);
                }
            }
        }
        dcp_commit();
        return mask;
    };

    long long item = slot;
    int stage = 0;
    unsigned mask = 0, mask_next = 0;
    if (item < p.items) mask = issue(item, 0);
    for (; item < p.items; item += nslots) {
        const long long nxt = item + nslots;
        if (nxt < p.items) {
            mask_next = issue(nxt, stage ^ 1);
            dcp_wait<1>();
        } else {
            dcp_wait<0>();
        }
        unsigned char* tile = dsm + stage * STAGE;
        if (xf) {
            // BN-apply + ReLU in place on this thread's own vectors (zero padding stays zero)
#pragma unroll
            for (int i = 0; i < MAXV; ++i) {
                if (mask & (1u << i)) {
                    uint4* q = reinterpret_cast<uint4*>(tile + s_off[i]);
                    if (CV_FIXED) {
                        *q = dw_xform8(*q, xs, xt);
                    } else {
                        float s2[8], t2[8];
                        load8(p.in_scale + cbase + (s_rc[i] & 0xff) * 8, s2);
                        load8(p.in_shift + cbase + (s_rc[i] & 0xff) * 8, t2);
                        *q = dw_xform8(*q, s2, t2);
                    }
                }
            }
        }
        __syncthreads();
        const int tw_i = (int)(item % p.tiles_w);
        const int th_i = (int)((item / p.tiles_w) % p.tiles_h);
        const int n = (int)(item / ((long long)p.tiles_w * p.tiles_h));
        if (lane_on) {
            const unsigned char* tp = tile + (r0 * HC + c0) * PP + cp * 4;
            if (MODE != 2) {
                const int oh0 = th_i * THT + r0, ow0 = tw_i * TW + c0;
                const bool col_ok0 = ow0 < p.W, col_ok1 = ow0 + 1 < p.W;
                bf16* zrow = p.out + ((long long)n * p.H * p.W + (long long)oh0 * p.W + ow0) * p.C + ch;
                const long long row_step = (long long)p.W * p.C;
                f2_t acc[K][2];
                const f2_t zero2 = f2_pack(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < K; ++i) acc[i][0] = acc[i][1] = zero2;
#pragma unroll
                for (int r = 0; r < TH + K - 1; ++r) {
                    f2_t in[K + 1];
#pragma unroll
                    for (int j = 0; j < K + 1; ++j)
                        in[j] = f2_from_bf16x2(*reinterpret_cast<const uint32_t*>(tp + (r * HC + j) * PP));
#pragma unroll
                    for (int kh = 0; kh < K; ++kh) {
                        const int o = r - kh;
                        if (o >= 0 && o < TH) {
#pragma unroll
                            for (int tc = 0; tc < 2; ++tc)
#pragma unroll
                                for (int kw = 0; kw < K; ++kw)
                                    acc[o % K][tc] = f2_fma(in[tc + kw], wr[kh][kw], acc[o % K][tc]);
                        }
                    }
                    const int od = r - (K - 1);
                    if (od >= 0) {
                        const bool row_ok = oh0 + od < p.H;
#pragma unroll
                        for (int tc = 0; tc < 2; ++tc) {
                            float v0, v1;
                            f2_unpack(acc[od % K][tc], v0, v1);
                            v0 += b0; v1 += b1;
                            acc[od % K][tc] = zero2;
                            if (row_ok && (tc == 0 ? col_ok0 : col_ok1)) {
                                const uint32_t pk = pack_bf16x2(v0, v1);
                                *reinterpret_cast<uint32_t*>(zrow + tc * p.C) = pk;
                                if (MODE == 0) {
                                    const float q0 = __uint_as_float(pk << 16), q1 = __uint_as_float(pk & 0xffff0000u);
                                    st[0] += q0; st[1] += q1;
                                    st[2] = fmaf(q0, q0, st[2]); st[3] = fmaf(q1, q1, st[3]);
                                } else if (bn_red) {
                                    // fused BN-backward reduction of the producer block: G = dx*[s*z+t>0]
                                    const float q0 = __uint_as_float(pk << 16), q1 = __uint_as_float(pk & 0xffff0000u);
                                    const uint32_t zu = *reinterpret_cast<const uint32_t*>(
                                        tile + Cfg::TILE_BYTES + ((r0 + od) * TW + c0 + tc) * PP + cp * 4);
                                    const float z0 = __uint_as_float(zu << 16), z1 = __uint_as_float(zu & 0xffff0000u);
                                    const float g0 = fmaf(gs0, z0, gt0) > 0.f ? q0 : 0.f;
                                    const float g1 = fmaf(gs1, z1, gt1) > 0.f ? q1 : 0.f;
                                    st[0] += g0; st[1] += g1;
                                    st[2] = fmaf(g0, z0, st[2]); st[3] = fmaf(g1, z1, st[3]);
                                }
                            }
                        }
                        zrow += row_step;
                    }
                }
            } else {
                const unsigned char* dzp = tile + Cfg::TILE_BYTES + (r0 * TW + c0) * PP + cp * 4;
                f2_t g[K][2];
#pragma unroll
                for (int r = 0; r < TH + K - 1; ++r) {
                    if (r < TH) {
#pragma unroll
                        for (int tc = 0; tc < 2; ++tc)
                            g[r % K][tc] = f2_from_bf16x2(*reinterpret_cast<const uint32_t*>(dzp + (r * TW + tc) * PP));
                    }
                    f2_t in[K + 1];
#pragma unroll
                    for (int j = 0; j < K + 1; ++j)
                        in[j] = f2_from_bf16x2(*reinterpret_cast<const uint32_t*>(tp + (r * HC + j) * PP));
#pragma unroll
                    for (int kh = 0; kh < K; ++kh) {
                        const int o = r - kh;
                        if (o >= 0 && o < TH) {
#pragma unroll
                            for (int tc = 0; tc < 2; ++tc)
#pragma unroll
                                for (int kw = 0; kw < K; ++kw)
                                    wr[kh][kw] = f2_fma(in[tc + kw], g[o % K][tc], wr[kh][kw]);
                        }
                    }
                }
            }
        }
        __syncthreads();          // tile consumed: its buffer may be refilled two iterations from now
        stage ^= 1;
        mask = mask_next;
    }

    // ---- per-CTA flush ----
    float* red = reinterpret_cast<float*>(dsm);     // reuse the (now idle) tile memory
    if ((MODE == 0 && p.stats) || bn_red) {
        for (int i = tid; i < 4 * NP; i += 256) red[i] = 0.f;
        __syncthreads();
        if (lane_on) {
#pragma unroll
            for (int q = 0; q < 4; ++q) atomicAdd(&red[q * NP + cp], st[q]);
        }
        __syncthreads();
        if (tid < NP) {
            atomicAdd(&p.stats[cbase + tid * 2], (double)red[0 * NP + tid]);
            atomicAdd(&p.stats[cbase + tid * 2 + 1], (double)red[1 * NP + tid]);
            atomicAdd(&p.stats[p.C + cbase + tid * 2], (double)red[2 * NP + tid]);
            atomicAdd(&p.stats[p.C + cbase + tid * 2 + 1], (double)red[3 * NP + tid]);
        }
    }
    if (MODE == 2) {
        for (int i = tid; i < CB * K * K; i += 256) red[i] = 0.f;
        __syncthreads();
        if (lane_on) {
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    float a0, a1;
                    f2_unpack(wr[i][j], a0, a1);
                    atomicAdd(&red[(cp * 2 + 0) * K * K + i * K + j], a0);
                    atomicAdd(&red[(cp * 2 + 1) * K * K + i * K + j], a1);
                }
        }
        __syncthreads();
        for (int i = tid; i < CB * K * K; i += 256) atomicAdd(&p.dw[(long long)cbase * K * K + i], red[i]);
    }
}

template <int K, int CB, int MODE>
static int launch_dw_tile(DwP p, cudaStream_t st, const char* name) {
    using Cfg = DwCfg<K, CB>;
    constexpr int STAGE = Cfg::TILE_BYTES + ((MODE == 2 || MODE == 3) ? Cfg::DZ_BYTES : 0);
    const int smem = 2 * STAGE;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(dw_tile_k<K, CB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    p.tiles_h = (p.H + Cfg::THT - 1) / Cfg::THT;
    p.tiles_w = (p.W + Cfg::TW - 1) / Cfg::TW;
    p.items = (long long)p.N * p.tiles_h * p.tiles_w;
    const int cblocks = p.C / CB;
    int per_sm = (200 * 1024) / (smem + 1024);
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    long long gx = ((long long)num_sms() * per_sm + cblocks - 1) / cblocks;
    if (gx > p.items) gx = p.items;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)(gx * cblocks));
    dw_tile_k<K, CB, MODE><<<grid, 256, smem, st>>>(p);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

template <int MODE>
static int dispatch_dw_tile(const DwP& p, int k, cudaStream_t st, const char* name) {
    const int C = p.C;
#define MNB_DW_CASE(KK, CBB) return launch_dw_tile<KK, CBB, MODE>(p, st, name)
    if (k == 3) {
        if (C % 64 == 0) MNB_DW_CASE(3, 64);
        if (C % 32 == 0) MNB_DW_CASE(3, 32);
        if (C % 16 == 0) MNB_DW_CASE(3, 16);
        if (C % 24 == 0) MNB_DW_CASE(3, 24);
    } else if (k == 5) {
        if (C % 64 == 0) MNB_DW_CASE(5, 64);
        if (C % 32 == 0) MNB_DW_CASE(5, 32);
        if (C % 16 == 0) MNB_DW_CASE(5, 16);
        if (C % 24 == 0) MNB_DW_CASE(5, 24);
    }
#undef MNB_DW_CASE
    set_error("%s: channel count %d not covered by the tile kernel", name, C);
    return MNB_ERR_UNSUPPORTED;
}

int dw_fwd_tile(const void* x, const float* s, const float* t, const float* w, const float* bias, void* z, double* stats,
                int N, int H, int W, int C, int k, cudaStream_t st) {
    DwP p = {};
    p.x = (const bf16*)x; p.in_scale = s; p.in_shift = t; p.w = w; p.bias = bias; p.out = (bf16*)z; p.stats = stats;
    p.N = N; p.H = H; p.W = W; p.C = C;
    return dispatch_dw_tile<0>(p, k, st, "dw_fwd(tile)");
}
int dw_dgrad_tile(const void* dz, const float* w, void* dx, const void* bn_z, const float* bn_scale,
                  const float* bn_shift, double* bn_sums, int N, int H, int W, int C, int k, cudaStream_t st) {
    DwP p = {};
    p.x = (const bf16*)dz; p.w = w; p.out = (bf16*)dx;
    p.bn_z = (const bf16*)bn_z; p.bn_scale = bn_scale; p.bn_shift = bn_shift; p.stats = bn_z ? bn_sums : nullptr;
    p.N = N; p.H = H; p.W = W; p.C = C;
    if (p.bn_z && p.stats) return dispatch_dw_tile<3>(p, k, st, "dw_dgrad+bn(tile)");
    return dispatch_dw_tile<1>(p, k, st, "dw_dgrad(tile)");
}
int dw_wgrad_tile(const void* x, const float* s, const float* t, const void* dz, float* dw, int N, int H, int W, int C,
                  int k, cudaStream_t st) {
    DwP p = {};
    p.x = (const bf16*)x; p.in_scale = s; p.in_shift = t; p.dz = (const bf16*)dz; p.dw = dw;
    p.N = N; p.H = H; p.W = W; p.C = C;
    return dispatch_dw_tile<2>(p, k, st, "dw_wgrad(tile)");
}

}  // namespace mnb
