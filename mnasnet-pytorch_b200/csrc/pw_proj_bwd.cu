// Backward of a "project" 1x1 ConvBlock (wide input, narrow output: 240->40, 480->80, 576->96 of the 28x28 / 14x14
// stages; nn.Conv2d(k=1) inside ConvBlock, src/models/mnasnet.py:58-62,120-128) AFTER its BatchNorm-backward elementwise
// pass: one kernel does backward-data (+ the residual skip gradient), backward-weight and the BatchNorm-backward
// reductions of the block that produced the input, from the narrow dZ:
//     dX[M][Cin] = dZ[M][Cout] W[Cout][Cin] (+ skip),   dW[Cout][Cin] += dZ^T A(X),   A = relu(s_in X + t_in),
//     sum(dX'), sum(dX' x)  with  dX' = dX * [s_in x + t_in > 0]      (SURVEY.md appendix F).
// The unfused chain ran three kernels over these tensors (tcgen05 backward-data 35 us, backward-weight 64 us, the
// producer's reduction 38 us at 14x14 576->96) and read dX back once more.
//
// Everything is independent per input channel, so a CTA column (blockIdx.y) owns a slice of CIS input channels: X, the
// skip gradient, dX, dW and the reductions of that slice are touched by nobody else; only the narrow dZ tile is re-read
// by every slice (from L2).  A CTA of 6 warps walks 96-row tiles: cp.async brings dZ / X / skip rows into shared memory
// at a padded pitch (16 B x odd: conflict-free ldmatrix); backward-data gives every warp its own 16 rows (A = dZ rows,
// B = the weight slice resident in shared memory), the epilogue adds the skip gradient in place, reduces against the raw
// X and stages dX; X is then activated in place and backward-weight gives every warp one 16-channel m-tile of Cout over
// ALL 96 rows (ldmatrix.trans operands), accumulators in registers for the CTA's life.
#include <algorithm>

#include "dw_mma.cuh"

namespace mnb {

struct PpP {
    const bf16* dz;             // [M][CO]
    const bf16* x;              // [M][Cin] raw output of the producing block
    const float* in_scale;      // producing block's BN scale / shift (NULL = plain input)
    const float* in_shift;
    const float* w;             // [CO][Cin] fp32
    const bf16* add;            // [M][Cin] residual skip gradient (NULL = none)
    bf16* dx;                   // [M][Cin]
    float* dw;                  // [CO][Cin] += (NULL = frozen)
    double* nsums;              // [2][Cin] (NULL = none)
    long long M;
    int cin;
};

typedef unsigned long long ppf2_t;
__device__ __forceinline__ ppf2_t ppf2_pack(float lo, float hi) { ppf2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void ppf2_unpack(ppf2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ ppf2_t ppf2_fma(ppf2_t a, ppf2_t b, ppf2_t c) { ppf2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ void pp_cp16(uint32_t dst, const void* src, bool pred) {
    const uint32_t n = pred ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

template <int CO, int CIS>
struct PpCfg {
    static constexpr int WARPS = 6, THREADS = 32 * WARPS, R = 16 * WARPS;
    static constexpr int MT = (CO + 15) / 16, NCI = CIS / 8, KS16 = CO / 16, KS8 = (CO % 16) / 8;
    static constexpr int NSPLIT = WARPS / MT;               // warps per m-tile in backward-weight (n-tiles split)
    static constexpr int NW = NCI / NSPLIT;                 // n-tiles per warp in backward-weight
    static_assert(NSPLIT >= 1 && NCI % NSPLIT == 0 && CO % 8 == 0 && CIS % 8 == 0, "pw_proj_bwd: unsupported slice shape");
    static constexpr int GP = c3_odd16(MT * 16 * 2);        // dZ row pitch (padded to the m-tile extent so ldmatrix.trans of
                                                            // a partial last m-tile stays inside the row)
    static constexpr int XP = c3_odd16(CIS * 2);
    static constexpr int WP = c3_odd16(CO * 2);             // weight row (one ci, K = co)
    static constexpr int DZ_BYTES = c3_al128(R * GP), X_BYTES = c3_al128(R * XP), W_BYTES = c3_al128(CIS * WP);
    static constexpr int SMEM = DZ_BYTES + 2 * X_BYTES + W_BYTES + 4 * CIS * 4 + 16;
};

template <int CO, int CIS>
__global__ void __launch_bounds__(PpCfg<CO, CIS>::THREADS, 2) pw_proj_bwd_k(const PpP p) {
    using Cfg = PpCfg<CO, CIS>;
    constexpr int THREADS = Cfg::THREADS, R = Cfg::R, MT = Cfg::MT, NCI = Cfg::NCI, KS16 = Cfg::KS16, KS8 = Cfg::KS8;
    constexpr int NSPLIT = Cfg::NSPLIT, NW = Cfg::NW, GP = Cfg::GP, XP = Cfg::XP, WP = Cfg::WP;
    extern __shared__ __align__(128) unsigned char dsm[];
    const uint32_t DZ = smem_u32(dsm);
    const uint32_t XB = DZ + Cfg::DZ_BYTES;
    const uint32_t AB = XB + Cfg::X_BYTES;                  // skip gradient; then the dX staging (in place)
    const uint32_t WS = AB + Cfg::X_BYTES;
    unsigned char* ws = dsm + Cfg::DZ_BYTES + 2 * Cfg::X_BYTES;
    float* icoef = reinterpret_cast<float*>(ws + Cfg::W_BYTES);       // [2][CIS]
    float* red = icoef + 2 * CIS;                                     // [2][CIS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, mi = lane >> 3, r8 = lane & 7;
    const int ci0 = blockIdx.y * CIS, CIT = p.cin;
    const bool act_in = p.in_scale != nullptr, has_add = p.add != nullptr;
    const bool do_red = p.nsums != nullptr && act_in, do_wgrad = p.dw != nullptr;

    for (int c = tid; c < CIS; c += THREADS) {
        icoef[c] = act_in ? p.in_scale[ci0 + c] : 1.f;
        icoef[CIS + c] = act_in ? p.in_shift[ci0 + c] : 0.f;
        red[c] = 0.f; red[CIS + c] = 0.f;
    }
    // weight slice: rows n = ci, K = co contiguous (B operand of backward-data)
    for (int i = tid; i < CIS * CO; i += THREADS) {
        const int co = i / CIS, ci = i - co * CIS;
        *reinterpret_cast<bf16*>(ws + ci * WP + co * 2) = __float2bfloat16_rn(p.w[(size_t)co * CIT + ci0 + ci]);
    }
    // the padding columns of the dZ rows (channels CO .. 16 MT - 1) feed discarded accumulator rows only, but must be finite
    if (MT * 16 > CO)
        for (int i = tid; i < R; i += THREADS) sts128(DZ + (uint32_t)(i * GP + CO * 2), make_uint4(0, 0, 0, 0));

    const uint32_t a_dg = (uint32_t)(((mi & 1) * 8 + r8) * GP + (mi >> 1) * 16);          // backward-data A (rows = pixels)
    const uint32_t b4 = WS + (uint32_t)(((mi >> 1) * 8 + r8) * WP + (mi & 1) * 16);
    const uint32_t b2 = WS + (uint32_t)(((NCI - 1) * 8 + r8) * WP + (mi & 1) * 16);
    const uint32_t b8 = WS + (uint32_t)((mi * 8 + r8) * WP);
    // backward-weight: this warp's m-tile of Cout and its range of n-tiles
    const int wmt = warp % MT, wns = warp / MT;
    const bool wg_live = do_wgrad && warp < MT * NSPLIT;
    const uint32_t a_wg = (uint32_t)(((mi >> 1) * 8 + r8) * GP + (wmt * 2 + (mi & 1)) * 16);
    const uint32_t b_wg = (uint32_t)(((mi & 1) * 8 + r8) * XP + (wns * NW + (mi >> 1)) * 16);
    float wacc[NW][4];
#pragma unroll
    for (int j = 0; j < NW; ++j) wacc[j][0] = wacc[j][1] = wacc[j][2] = wacc[j][3] = 0.f;
    ppf2_t rs[NCI][2];          // per n-tile: (sum dX', sum dX' x) of the lane's channel pair
#pragma unroll
    for (int j = 0; j < NCI; ++j) rs[j][0] = rs[j][1] = ppf2_pack(0.f, 0.f);
    __syncthreads();

    const long long ntiles = (p.M + R - 1) / R;
    constexpr int GV = CO / 8, XV = CIS / 8;                // 16-byte vectors per row
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * R;
        // ---- loads: dZ rows, X slice, skip slice (rows beyond M zero-filled) ----
        for (int i = tid; i < R * GV; i += THREADS) {
            const int r = i / GV, v = i - r * GV;
            const bool ok = row0 + r < p.M;
            pp_cp16(DZ + (uint32_t)(r * GP + v * 16), p.dz + (ok ? (row0 + r) * CO + v * 8 : 0), ok);
        }
        for (int i = tid; i < R * XV; i += THREADS) {
            const int r = i / XV, v = i - r * XV;
            const bool ok = row0 + r < p.M;
            const long long off = ok ? (row0 + r) * CIT + ci0 + v * 8 : 0;
            pp_cp16(XB + (uint32_t)(r * XP + v * 16), p.x + off, ok);
            if (has_add) pp_cp16(AB + (uint32_t)(r * XP + v * 16), p.add + off, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- backward-data: this warp's 16 rows x CIS ----
        {
            float acc[NCI][4];
#pragma unroll
            for (int j = 0; j < NCI; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            const uint32_t grow = DZ + (uint32_t)(16 * warp * GP) + a_dg;
#pragma unroll
            for (int ks = 0; ks < KS16; ++ks) {
                uint32_t a0, a1, a2, a3, bf[NCI][2];
                ldsm4(grow + ks * 32, a0, a1, a2, a3);
                c3_load_b16<NCI, WP>(bf, b4, b2, ks * 32);
#pragma unroll
                for (int j = 0; j < NCI; ++j) mma16816(acc[j], a0, a1, a2, a3, bf[j][0], bf[j][1]);
            }
            if constexpr (KS8 == 1) {
                uint32_t a0, a1, bf[NCI][2];
                ldsm2(grow - (mi >> 1) * 16 + KS16 * 32, a0, a1);
                c3_load_b8<NCI, WP>(bf, b8, KS16 * 32);
#pragma unroll
                for (int j = 0; j < NCI; ++j) mma1688(acc[j], a0, a1, bf[j][0]);
            }
            // epilogue: + skip gradient (in place), stage dX, reduce against the raw X
            const bool v0 = row0 + 16 * warp + g < p.M, v1 = row0 + 16 * warp + g + 8 < p.M;
            const ppf2_t one = ppf2_pack(1.f, 1.f);
#pragma unroll
            for (int j = 0; j < NCI; ++j) {
                const uint32_t o0 = (uint32_t)((16 * warp + g) * XP + j * 16 + t * 4), o1 = o0 + 8 * XP;
                float d0 = acc[j][0], d1 = acc[j][1], d2 = acc[j][2], d3 = acc[j][3];
                if (has_add) {
                    const uint32_t s0 = lds32(AB + o0), s1 = lds32(AB + o1);
                    d0 += bf_lo(s0); d1 += bf_hi(s0); d2 += bf_lo(s1); d3 += bf_hi(s1);
                }
                const uint32_t u0 = pack_bf16x2(d0, d1), u1 = pack_bf16x2(d2, d3);
                sts32(AB + o0, u0);
                sts32(AB + o1, u1);
                if (do_red) {
                    const int c = 8 * j + 2 * t;
                    const ppf2_t sp = ppf2_pack(icoef[c], icoef[c + 1]), tp = ppf2_pack(icoef[CIS + c], icoef[CIS + c + 1]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t xr = lds32(XB + (h ? o1 : o0));
                        const ppf2_t x = ppf2_pack(bf_lo(xr), bf_hi(xr));
                        float y0, y1;
                        ppf2_unpack(ppf2_fma(sp, x, tp), y0, y1);
                        const uint32_t u = h ? u1 : u0;
                        const bool v = h ? v1 : v0;
                        const ppf2_t q = ppf2_pack(v && y0 > 0.f ? bf_lo(u) : 0.f, v && y1 > 0.f ? bf_hi(u) : 0.f);
                        rs[j][0] = ppf2_fma(q, one, rs[j][0]);
                        rs[j][1] = ppf2_fma(q, x, rs[j][1]);
                    }
                }
            }
        }
        __syncthreads();
        // ---- dX rows out (coalesced 16-byte stores); X activated in place for backward-weight ----
        for (int i = tid; i < R * XV; i += THREADS) {
            const int r = i / XV, v = i - r * XV;
            if (row0 + r < p.M)
                *reinterpret_cast<uint4*>(p.dx + (row0 + r) * CIT + ci0 + v * 8) = lds128(AB + (uint32_t)(r * XP + v * 16));
            if (do_wgrad && act_in) {
                const uint32_t a = XB + (uint32_t)(r * XP + v * 16);
                uint4 u = lds128(a);
                const float4 s0 = *reinterpret_cast<const float4*>(icoef + v * 8), s1 = *reinterpret_cast<const float4*>(icoef + v * 8 + 4);
                const float4 t0 = *reinterpret_cast<const float4*>(icoef + CIS + v * 8), t1 = *reinterpret_cast<const float4*>(icoef + CIS + v * 8 + 4);
                u.x = pack_bf16x2(fmaxf(fmaf(bf_lo(u.x), s0.x, t0.x), 0.f), fmaxf(fmaf(bf_hi(u.x), s0.y, t0.y), 0.f));
                u.y = pack_bf16x2(fmaxf(fmaf(bf_lo(u.y), s0.z, t0.z), 0.f), fmaxf(fmaf(bf_hi(u.y), s0.w, t0.w), 0.f));
                u.z = pack_bf16x2(fmaxf(fmaf(bf_lo(u.z), s1.x, t1.x), 0.f), fmaxf(fmaf(bf_hi(u.z), s1.y, t1.y), 0.f));
                u.w = pack_bf16x2(fmaxf(fmaf(bf_lo(u.w), s1.z, t1.z), 0.f), fmaxf(fmaf(bf_hi(u.w), s1.w, t1.w), 0.f));
                sts128(a, u);
            }
        }
        __syncthreads();
        // ---- backward-weight: m-tile wmt of Cout x this warp's n-tiles, K = the tile's 96 rows ----
        // (rows beyond M are zero in dZ, so the activated zero-filled X rows add nothing)
        if (wg_live) {
#pragma unroll
            for (int ks = 0; ks < R / 16; ++ks) {
                uint32_t a0, a1, a2, a3, bx[NW][2];
                ldsm4t(DZ + (uint32_t)(ks * 16 * GP) + a_wg, a0, a1, a2, a3);
#pragma unroll
                for (int j = 0; j < NW; j += 2) {
                    if (j + 1 < NW) ldsm4t(XB + (uint32_t)(ks * 16 * XP) + b_wg + j * 16, bx[j][0], bx[j][1], bx[j + 1][0], bx[j + 1][1]);
                    else ldsm2t(XB + (uint32_t)(ks * 16 * XP) + b_wg - (mi >> 1) * 16 + j * 16, bx[j][0], bx[j][1]);
                }
#pragma unroll
                for (int j = 0; j < NW; ++j) mma16816(wacc[j], a0, a1, a2, a3, bx[j][0], bx[j][1]);
            }
        }
        __syncthreads();            // the tile buffers are free for the next tile's loads
    }
    // ---- flush: weight gradient and the producer's reductions ----
    if (wg_live) {
#pragma unroll
        for (int j = 0; j < NW; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int co = wmt * 16 + g + (e >> 1) * 8, ci = (wns * NW + j) * 8 + 2 * t + (e & 1);
                if (co < CO) atomicAdd(&p.dw[(size_t)co * CIT + ci0 + ci], wacc[j][e]);
            }
    }
    if (do_red) {
#pragma unroll
        for (int j = 0; j < NCI; ++j) {
            float v[4];
            ppf2_unpack(rs[j][0], v[0], v[1]);
            ppf2_unpack(rs[j][1], v[2], v[3]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[q] += __shfl_xor_sync(0xffffffffu, v[q], 4);
                v[q] += __shfl_xor_sync(0xffffffffu, v[q], 8);
                v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);
            }
            if (g == 0) {
                const int c = 8 * j + 2 * t;
                atomicAdd(&red[c], v[0]);
                atomicAdd(&red[c + 1], v[1]);
                atomicAdd(&red[CIS + c], v[2]);
                atomicAdd(&red[CIS + c + 1], v[3]);
            }
        }
        __syncthreads();
        for (int i = tid; i < CIS; i += THREADS) {
            atomicAdd(&p.nsums[ci0 + i], (double)red[i]);
            atomicAdd(&p.nsums[CIT + ci0 + i], (double)red[CIS + i]);
        }
    }
}

template <int CO, int CIS>
static int launch_pp(const PpP& p, cudaStream_t st) {
    using Cfg = PpCfg<CO, CIS>;
    const char* name = "pw_proj_bwd";
    static int occ = -1;
    if (occ < 0) {
        cudaError_t e = cudaFuncSetAttribute(pw_proj_bwd_k<CO, CIS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return (int)e; }
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, pw_proj_bwd_k<CO, CIS>, Cfg::THREADS, Cfg::SMEM);
        if (o < 1) { set_error("%s: kernel does not fit on an SM", name); return MNB_ERR_UNSUPPORTED; }
        occ = o;
    }
    const int slices = p.cin / CIS;
    const long long ntiles = (p.M + Cfg::R - 1) / Cfg::R;
    long long grid = std::max(1LL, (long long)num_sms() * occ / slices);       // resident CTAs only
    if (grid > ntiles) grid = ntiles;
    pw_proj_bwd_k<CO, CIS><<<dim3((unsigned)grid, (unsigned)slices), Cfg::THREADS, Cfg::SMEM, st>>>(p);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

int pw_proj_bwd(const void* dz, const void* x, const float* in_scale, const float* in_shift, const float* w, const void* add,
                void* dx, float* dw, double* nsums, long long M, int Cin, int Cout, cudaStream_t st) {
    PpP p = {};
    p.dz = (const bf16*)dz; p.x = (const bf16*)x; p.in_scale = in_scale; p.in_shift = in_shift; p.w = w;
    p.add = (const bf16*)add; p.dx = (bf16*)dx; p.dw = dw; p.nsums = nsums; p.M = M; p.cin = Cin;
    if (Cout == 96 && Cin % 96 == 0) return launch_pp<96, 96>(p, st);
    if (Cout == 80 && Cin % 96 == 0) return launch_pp<80, 96>(p, st);
    if (Cout == 40 && Cin % 80 == 0) return launch_pp<40, 80>(p, st);
    set_error("pw_proj_bwd: shape %d -> %d not instantiated", Cin, Cout);
    return MNB_ERR_UNSUPPORTED;
}

}  // namespace mnb
