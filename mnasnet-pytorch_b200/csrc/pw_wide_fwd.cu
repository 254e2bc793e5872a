// Forward of the WIDE 1x1 ConvBlocks of the 28x28 / 14x14 stages, bf16 NHWC (nn.Conv2d(k=1) inside ConvBlock,
// src/models/mnasnet.py:58-62,116-128): expand blocks 40->240, 80->480, 96->576 and project blocks 240->40, 480->80,
// 576->96, with the producing block's BN-apply+ReLU on load and this block's BN batch statistics on store.
//
// These layers ran the generic tcgen05 pipeline (gemm_tc.cu) at 1.2-2.5 TB/s: its producers gather every 16-byte operand
// vector with their own index arithmetic and the kernel is bound by instruction issue (ncu: 6-7 % tensor-pipe active).
// A 1x1 convolution has no spatial structure -- the activations are an [M][Cin] matrix whose rows are contiguous -- so
// here a CTA walks 96-row tiles of one N-slice of the output channels (blockIdx.y): cp.async streams [96][KC] chunks of
// the rows into a two-stage ring at a padded pitch (16 B x odd: conflict-free ldmatrix), every thread applies
// relu(scale*x+shift) in place to the vectors it copied, the weight slice [NS][Cin] stays resident in shared memory for
// the CTA's life, mma.sync.m16n8k16 accumulates over the chunks (each warp 16 rows x NS columns), and the output tile is
// staged, written with coalesced 16-byte stores and reduced for the BN statistics of the stored (bf16) values.
// Expand shapes re-read (from L2) and re-transform the narrow A tile once per N-slice; project shapes have one slice.
//
// MEASURED (scripts/exp_pw_wide.py, batch 256): 65 / 72 / 52 us (576->96, 240->40, 480->80) and 56 / 62 / 45 us (96->576,
// 40->240, 80->480) against 54 / 63 / 46 and 44 / 46 / 34 us on the tcgen05 pipeline -- with six warps per CTA and one or
// two CTAs per SM the chunk loop is latency-bound.  The kernel is correct (tests/test_pw_wide_gpu.py) but NOT the default:
// option "pw_wide" = 1 selects it.
#include <algorithm>

#include "conv_params.cuh"
#include "dw_mma.cuh"

namespace mnb {

struct PfP {
    const bf16* x;              // [M][CIN] raw output of the producing block
    const float* in_scale;      // NULL = plain input
    const float* in_shift;
    const float* w;             // [Cout][CIN] fp32
    const bf16* wpk;            // optional bf16 packing [Cout][CIN]
    bf16* z;                    // [M][Cout]
    double* stats;              // [2][Cout] (NULL = none)
    long long M;
    int cout;
};

__device__ __forceinline__ void pf_cp16(uint32_t dst, const void* src, bool pred) {
    const uint32_t n = pred ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

template <int CIN, int KC, int NS>
struct PfCfg {
    static constexpr int WARPS = 6, THREADS = 32 * WARPS, R = 16 * WARPS;
    static constexpr int NCH = CIN / KC;                    // K chunks per tile
    static_assert(CIN % KC == 0 && KC % 8 == 0 && NS % 8 == 0, "pw_wide_fwd: unsupported shape");
    static constexpr int NT = NS / 8, KS16 = KC / 16, KS8 = (KC % 16) / 8, KV = KC / 8, NV = NS / 8;
    static constexpr int AP = c3_odd16(KC * 2), OP = c3_odd16(NS * 2), WP = c3_odd16(CIN * 2);
    static constexpr int A_BYTES = c3_al128(R * AP), O_BYTES = c3_al128(R * OP), W_BYTES = c3_al128(NS * WP);
    static constexpr int SMEM = W_BYTES + 2 * A_BYTES + O_BYTES + 2 * CIN * 4 + 2 * NS * 4 + 16;
    static constexpr int CP = NS / 2, NSTAT = THREADS / CP * CP;
    static constexpr int MINB = SMEM + 1024 <= 113 * 1024 ? 2 : 1;
};

template <int CIN, int KC, int NS>
__global__ void __launch_bounds__(PfCfg<CIN, KC, NS>::THREADS, PfCfg<CIN, KC, NS>::MINB) pw_wide_fwd_k(const PfP p) {
    using Cfg = PfCfg<CIN, KC, NS>;
    constexpr int THREADS = Cfg::THREADS, R = Cfg::R, NCH = Cfg::NCH, NT = Cfg::NT, KS16 = Cfg::KS16, KS8 = Cfg::KS8;
    constexpr int KV = Cfg::KV, NV = Cfg::NV, AP = Cfg::AP, OP = Cfg::OP, WP = Cfg::WP;
    extern __shared__ __align__(128) unsigned char dsm[];
    const uint32_t WS = smem_u32(dsm);
    const uint32_t A0 = WS + Cfg::W_BYTES;
    const uint32_t OUT = A0 + 2 * Cfg::A_BYTES;
    float* s_sc = reinterpret_cast<float*>(dsm + Cfg::W_BYTES + 2 * Cfg::A_BYTES + Cfg::O_BYTES);      // [2][CIN]
    float* red = s_sc + 2 * CIN;                                                                      // [2][NS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3, mi = lane >> 3, r8 = lane & 7;
    const int n0 = blockIdx.y * NS;
    const bool xf = p.in_scale != nullptr;

    if (xf)
        for (int i = tid; i < CIN; i += THREADS) { s_sc[i] = p.in_scale[i]; s_sc[CIN + i] = p.in_shift[i]; }
    for (int i = tid; i < 2 * NS; i += THREADS) red[i] = 0.f;
    // weight slice rows n = co - n0, K = ci contiguous
    if (p.wpk) {
        constexpr int VPR = CIN / 8;
        for (int i = tid; i < NS * VPR; i += THREADS) {
            const int r = i / VPR, v = i - r * VPR;
            *reinterpret_cast<uint4*>(dsm + r * WP + v * 16) = *reinterpret_cast<const uint4*>(p.wpk + (size_t)(n0 + r) * CIN + v * 8);
        }
    } else {
        for (int i = tid; i < NS * CIN; i += THREADS) {
            const int r = i / CIN, k = i - r * CIN;
            *reinterpret_cast<bf16*>(dsm + r * WP + k * 2) = __float2bfloat16_rn(p.w[(size_t)(n0 + r) * CIN + k]);
        }
    }
    const uint32_t a_ld = (uint32_t)((16 * warp + (mi & 1) * 8 + r8) * AP + (mi >> 1) * 16);
    const uint32_t b4 = WS + (uint32_t)(((mi >> 1) * 8 + r8) * WP + (mi & 1) * 16);
    const uint32_t b2 = WS + (uint32_t)(((NT - 1) * 8 + r8) * WP + (mi & 1) * 16);
    const uint32_t b8 = WS + (uint32_t)((mi * 8 + r8) * WP);
    constexpr int CP = Cfg::CP, PSTEP = Cfg::NSTAT / CP;
    const int scp = tid % CP, sp0 = tid / CP;
    float ssum0 = 0.f, ssum1 = 0.f, ssq0 = 0.f, ssq1 = 0.f;
    const bool do_stats = p.stats != nullptr && tid < Cfg::NSTAT;
    __syncthreads();

    const long long ntiles = (p.M + R - 1) / R;
    const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long nsteps = my_tiles * NCH;            // (tile, chunk) stream of this CTA
    auto issue = [&](long long s) {                     // cp.async of stream element s into stage s & 1
        const long long tile = blockIdx.x + (s / NCH) * gridDim.x;
        const int c = (int)(s % NCH);
        const long long row0 = tile * R;
        const uint32_t dst = A0 + (uint32_t)((s & 1) * Cfg::A_BYTES);
        for (int i = tid; i < R * KV; i += THREADS) {
            const int r = i / KV, v = i - r * KV;
            const bool ok = row0 + r < p.M;
            pf_cp16(dst + (uint32_t)(r * AP + v * 16), p.x + (ok ? (row0 + r) * CIN + c * KC + v * 8 : 0), ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    if (nsteps > 0) issue(0);
    for (long long s = 0; s < nsteps; ++s) {
        const int c = (int)(s % NCH);
        const long long tile = blockIdx.x + (s / NCH) * gridDim.x;
        const long long row0 = tile * R;
        const uint32_t AS = A0 + (uint32_t)((s & 1) * Cfg::A_BYTES);
        if (s + 1 < nsteps) {
            issue(s + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (xf) {       // relu(scale * x + shift) in place on the vectors this thread copied (visible to it after the wait)
            for (int i = tid; i < R * KV; i += THREADS) {
                const int r = i / KV, v = i - r * KV;
                const uint32_t a = AS + (uint32_t)(r * AP + v * 16);
                uint4 u = lds128(a);
                const float* sc = s_sc + c * KC + v * 8;
                const float4 s0 = *reinterpret_cast<const float4*>(sc), s1 = *reinterpret_cast<const float4*>(sc + 4);
                const float4 t0 = *reinterpret_cast<const float4*>(sc + CIN), t1 = *reinterpret_cast<const float4*>(sc + CIN + 4);
                u.x = pack_bf16x2(fmaxf(fmaf(bf_lo(u.x), s0.x, t0.x), 0.f), fmaxf(fmaf(bf_hi(u.x), s0.y, t0.y), 0.f));
                u.y = pack_bf16x2(fmaxf(fmaf(bf_lo(u.y), s0.z, t0.z), 0.f), fmaxf(fmaf(bf_hi(u.y), s0.w, t0.w), 0.f));
                u.z = pack_bf16x2(fmaxf(fmaf(bf_lo(u.z), s1.x, t1.x), 0.f), fmaxf(fmaf(bf_hi(u.z), s1.y, t1.y), 0.f));
                u.w = pack_bf16x2(fmaxf(fmaf(bf_lo(u.w), s1.z, t1.z), 0.f), fmaxf(fmaf(bf_hi(u.w), s1.w, t1.w), 0.f));
                sts128(a, u);
            }
        }
        __syncthreads();
        // ---- this warp's 16 rows x NS columns += A chunk x W[:, chunk] ----
#pragma unroll
        for (int ks = 0; ks < KS16; ++ks) {
            uint32_t a0, a1, a2, a3, bf[NT][2];
            ldsm4(AS + a_ld + ks * 32, a0, a1, a2, a3);
            c3_load_b16<NT, WP>(bf, b4, b2, (c * KC + ks * 16) * 2);
#pragma unroll
            for (int j = 0; j < NT; ++j) mma16816(acc[j], a0, a1, a2, a3, bf[j][0], bf[j][1]);
        }
        if constexpr (KS8 == 1) {
            uint32_t a0, a1, bf[NT][2];
            ldsm2(AS + a_ld - (mi >> 1) * 16 + KS16 * 32, a0, a1);
            c3_load_b8<NT, WP>(bf, b8, (c * KC + KS16 * 16) * 2);
#pragma unroll
            for (int j = 0; j < NT; ++j) mma1688(acc[j], a0, a1, bf[j][0]);
        }
        if (c == NCH - 1) {
            // ---- tile complete: stage, store, statistics ----
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const uint32_t o = OUT + (uint32_t)((16 * warp + g) * OP + j * 16 + t * 4);
                sts32(o, pack_bf16x2(acc[j][0], acc[j][1]));
                sts32(o + 8 * OP, pack_bf16x2(acc[j][2], acc[j][3]));
                acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            }
            __syncthreads();
            const int rows = p.M - row0 < R ? (int)(p.M - row0) : R;
            for (int i = tid; i < rows * NV; i += THREADS) {
                const int r = i / NV, v = i - r * NV;
                *reinterpret_cast<uint4*>(p.z + (row0 + r) * p.cout + n0 + v * 8) = lds128(OUT + (uint32_t)(r * OP + v * 16));
            }
            if (do_stats) {
                for (int r = sp0; r < rows; r += PSTEP) {
                    const uint32_t u = lds32(OUT + (uint32_t)(r * OP + scp * 4));
                    const float v0 = bf_lo(u), v1 = bf_hi(u);
                    ssum0 += v0; ssum1 += v1;
                    ssq0 = fmaf(v0, v0, ssq0); ssq1 = fmaf(v1, v1, ssq1);
                }
            }
        }
        __syncthreads();            // stage s & 1 (and OUT) are free for the copies issued next iteration
    }
    if (p.stats != nullptr) {
        if (do_stats) {
            atomicAdd(&red[2 * scp], ssum0); atomicAdd(&red[2 * scp + 1], ssum1);
            atomicAdd(&red[NS + 2 * scp], ssq0); atomicAdd(&red[NS + 2 * scp + 1], ssq1);
        }
        __syncthreads();
        for (int i = tid; i < NS; i += THREADS) {
            atomicAdd(&p.stats[n0 + i], (double)red[i]);
            atomicAdd(&p.stats[p.cout + n0 + i], (double)red[NS + i]);
        }
    }
}

template <int CIN, int KC, int NS>
static int launch_pf(const PfP& p, cudaStream_t st) {
    using Cfg = PfCfg<CIN, KC, NS>;
    const char* name = "conv_fwd(pw_wide)";
    static int occ = -1;
    if (occ < 0) {
        cudaError_t e = cudaFuncSetAttribute(pw_wide_fwd_k<CIN, KC, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return (int)e; }
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, pw_wide_fwd_k<CIN, KC, NS>, Cfg::THREADS, Cfg::SMEM);
        if (o < 1) { set_error("%s: kernel does not fit on an SM", name); return MNB_ERR_UNSUPPORTED; }
        occ = o;
    }
    const int slices = p.cout / NS;
    const long long ntiles = (p.M + Cfg::R - 1) / Cfg::R;
    long long grid = std::max(1LL, (long long)num_sms() * occ / slices);       // resident CTAs only
    if (grid > ntiles) grid = ntiles;
    pw_wide_fwd_k<CIN, KC, NS><<<dim3((unsigned)grid, (unsigned)slices), Cfg::THREADS, Cfg::SMEM, st>>>(p);
    MNB_LAUNCH_CHECK(name);
    return 0;
}

// the wide 1x1 layers of the 28x28 / 14x14 stages; MNB_ERR_UNSUPPORTED otherwise (the caller keeps the tcgen05 path)
int conv_fwd_pwide(const ConvP& c, cudaStream_t st) {
    if (!option_get(OPT_PW_WIDE) || c.k != 1 || c.stride != 1 || c.pad != 0 || c.nchw_in || c.out_f32 || c.bias) return MNB_ERR_UNSUPPORTED;
    if (((uintptr_t)c.x & 15) || ((uintptr_t)c.out & 15)) return MNB_ERR_UNSUPPORTED;
    PfP p = {};
    p.x = (const bf16*)c.x; p.in_scale = c.in_scale; p.in_shift = c.in_shift; p.w = c.w; p.wpk = (const bf16*)c.wpk;
    p.z = (bf16*)c.out; p.stats = c.stats; p.M = (long long)c.N * c.H * c.W; p.cout = c.Cout;
#define PF(CI, KC_, NS_, CO) if (c.Cin == CI && c.Cout == CO) return launch_pf<CI, KC_, NS_>(p, st);
    PF(576, 96, 96, 96) PF(480, 96, 80, 80) PF(240, 80, 40, 40)          // project
    PF(96, 96, 96, 576) PF(80, 80, 96, 480) PF(40, 40, 80, 240)          // expand
#undef PF
    return MNB_ERR_UNSUPPORTED;
}

}  // namespace mnb
