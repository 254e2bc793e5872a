// Pointwise (1x1) convolutions of the high-resolution stages as warp-streaming kernels.
//
// Replaces nn.Conv2d(k=1) forward / backward-data / backward-weight of the MBConv expand / project blocks
// (src/models/mnasnet.py:82-85, 92-95, 116-119, 126-129) at 112x112 and 56x56, where Cin, Cout <= 80: 0.2-0.8 KB per
// pixel row and 11-40 flop/B, i.e. HBM-bound by a factor > 5.  The tcgen05 pipeline of gemm_tc.cu pays ~1000 cycles
// of mbarrier hand-offs per 128-row tile on these shapes (profiles/README.md, knock-out experiment); here every warp
// is an independent streaming unit with no cross-warp synchronisation at all:
//
//   forward / dgrad : a warp owns 16 pixel rows at a time.  Each lane loads its slice of the A fragment straight
//       from global memory as 8-byte units (the K index of the MMA is permuted so that the four elements a lane
//       needs per k-step are four CONSECUTIVE channels), applies relu(scale*x+shift) in registers, issues
//       mma.sync.m16n8k16 (bf16 x bf16 -> fp32) against weight fragments that live in registers (or in a
//       lane-ordered shared array), and writes its accumulators back with the N index permuted the same way, so a
//       lane stores 16 / 8 / 4 consecutive bytes of the output row.  BatchNorm statistics of the stored (rounded)
//       values are accumulated per lane in registers and flushed once per CTA.  The next chunk's rows are
//       prefetched into registers while the current one is processed.
//   wgrad : dW[co][ci] = sum_rows dz[row][co] * a[row][ci]; both operands have the reduction index as the slow
//       one, so a warp stages 16 rows of each in its private shared buffer and reads them back transposed with
//       ldmatrix.trans; the activation transform is applied to the B fragments (channel == lane group).
//
// Everything here is bf16 activations / fp32 accumulate; shapes outside the instantiated set return
// MNB_ERR_UNSUPPORTED and the caller falls back to the tcgen05 path.
#include "conv_params.cuh"

namespace mnb {

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
// relu(s*x+t) on a packed bf16 pair, re-rounded to bf16 (the value the tensor pipe consumes)
__device__ __forceinline__ uint32_t bn_relu_pair(uint32_t u, float s0, float s1, float t0, float t1) {
    const float a = fmaxf(fmaf(bf_lo(u), s0, t0), 0.f);
    const float b = fmaxf(fmaf(bf_hi(u), s1, t1), 0.f);
    return pack_bf16x2(a, b);
}

// Accumulator element (n-tile j, lane-in-quad tq, e) <-> output channel.  N-tiles are taken in groups of four: inside
// a group of R n-tiles (R = 4, or NT % 4 for the last one) lane tq owns the 2R consecutive channels
// [32q + 2R*tq, 32q + 2R*(tq+1)), so a row is written as 16 / 8 / 4-byte pieces per lane.
template <int NT> struct OutMap {
    __host__ __device__ static constexpr int groups() { return (NT + 3) / 4; }
    __host__ __device__ static constexpr int rq(int q) { return (q < NT / 4) ? 4 : (NT % 4); }
    __host__ __device__ static constexpr int phys(int j, int tq, int e) {
        return 32 * (j >> 2) + 2 * rq(j >> 2) * tq + 2 * (j & 3) + e;
    }
};

struct PwP {
    const bf16* a;        // fwd: x [M][K]; dgrad: dz [M][K]
    const float* scale;   // fwd: per-input-channel BN scale / shift of the producer (NULL = raw input)
    const float* shift;
    const bf16* wpk;      // [NO][K] bf16, K-major (mnb_pack_weights: wpk_fwd for forward, wpk_dgrad for dgrad)
    const float* bias;    // fwd, may be NULL
    const bf16* add;      // dgrad residual, may be NULL
    bf16* out;            // [M][NO]
    double* stats;        // fwd: [2*NO] sum / sum of squares of the stored values, may be NULL
    long long M;
    int K;                // true channel count of A (<= 16*KS)
};

enum { PW_FWD = 0, PW_DGRAD = 1 };

// KS k-steps of 16 input channels, NT n-tiles of 8 output channels, MT 16-row tiles per prefetched chunk.
// 128-thread CTAs, MINB of them per SM (4 -> 128 registers per lane, 3 -> 168).
template <int KS, int NT, int MT, int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) pw_stream_k(const PwP p) {
    constexpr int NO = 8 * NT;
    constexpr bool BREG = KS * NT <= 12;          // weight fragments in registers
    typedef OutMap<NT> OM;
    __shared__ __align__(16) float s_scale[16 * KS];
    __shared__ __align__(16) float s_shift[16 * KS];
    __shared__ __align__(8) float s_bias[NO];
    __shared__ float s_red[2 * NO];
    __shared__ __align__(8) uint2 s_b[BREG ? 1 : KS * NT * 32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int K = p.K;
    const bool xf = (MODE == PW_FWD) && p.scale != nullptr;

    for (int i = threadIdx.x; i < 16 * KS; i += blockDim.x) {
        s_scale[i] = (xf && i < K) ? p.scale[i] : 0.f;
        s_shift[i] = (xf && i < K) ? p.shift[i] : 0.f;
    }
    for (int i = threadIdx.x; i < NO; i += blockDim.x) s_bias[i] = (MODE == PW_FWD && p.bias) ? p.bias[i] : 0.f;
    for (int i = threadIdx.x; i < 2 * NO; i += blockDim.x) s_red[i] = 0.f;
    if (!BREG) {
        for (int i = threadIdx.x; i < KS * NT * 32; i += blockDim.x) {
            const int l = i & 31, sj = i >> 5, s = sj / NT, j = sj % NT;
            const int gg = l >> 2, tt = l & 3;
            const int ch = 16 * s + 4 * tt, n = OM::phys(j, gg >> 1, gg & 1);
            uint2 v = make_uint2(0u, 0u);
            if (ch < K) v = *reinterpret_cast<const uint2*>(p.wpk + (long long)n * K + ch);
            s_b[i] = v;
        }
    }
    // weight fragments: lane (g,t), k-step s, n-tile j <- W[phys(j, g>>1, g&1)][16s+4t .. +3]
    uint32_t bfr[BREG ? KS : 1][BREG ? NT : 1][2];
    if (BREG) {
#pragma unroll
        for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int ch = 16 * s + 4 * t, n = OM::phys(j, g >> 1, g & 1);
                uint2 v = make_uint2(0u, 0u);
                if (ch < K) v = *reinterpret_cast<const uint2*>(p.wpk + (long long)n * K + ch);
                bfr[BREG ? s : 0][BREG ? j : 0][0] = v.x;
                bfr[BREG ? s : 0][BREG ? j : 0][1] = v.y;
            }
    }
    __syncthreads();

    float ssum[MODE == PW_FWD ? NT : 1][2], ssq[MODE == PW_FWD ? NT : 1][2];
#pragma unroll
    for (int j = 0; j < (MODE == PW_FWD ? NT : 1); ++j) ssum[j][0] = ssum[j][1] = ssq[j][0] = ssq[j][1] = 0.f;

    const long long rows_per_chunk = 16 * MT;
    const long long nchunks = (p.M + rows_per_chunk - 1) / rows_per_chunk;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    long long chunk = (long long)blockIdx.x * (blockDim.x >> 5) + warp;

    // raw A units of one chunk: [m-tile][k-step][row half] -> 4 consecutive channels (8 bytes)
    uint2 nxt[MT][KS][2];
    auto load_chunk = [&](long long c) {
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long row = c * rows_per_chunk + m * 16 + g + 8 * h;
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    const int ch = 16 * s + 4 * t;
                    uint2 v = make_uint2(0u, 0u);
                    if (row < p.M && ch < K) v = __ldg(reinterpret_cast<const uint2*>(p.a + row * K + ch));
                    nxt[m][s][h] = v;
                }
            }
    };
    if (chunk < nchunks) load_chunk(chunk);

    for (; chunk < nchunks; chunk += wstride) {
        uint2 cur[MT][KS][2];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int s = 0; s < KS; ++s) { cur[m][s][0] = nxt[m][s][0]; cur[m][s][1] = nxt[m][s][1]; }
        if (chunk + wstride < nchunks) load_chunk(chunk + wstride);

#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const long long r0 = chunk * rows_per_chunk + m * 16;
            if (r0 >= p.M) break;                               // warp-uniform
            const long long row_lo = r0 + g, row_hi = r0 + g + 8;
            const bool ok_lo = row_lo < p.M, ok_hi = row_hi < p.M;
            float d[NT][4];
            if (MODE == PW_FWD) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const float2 b = *reinterpret_cast<const float2*>(&s_bias[OM::phys(j, t, 0)]);
                    d[j][0] = d[j][2] = b.x;
                    d[j][1] = d[j][3] = b.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NT; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
                if (p.add) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const long long row = h ? row_hi : row_lo;
                        if (row < p.M) {
                            const bf16* ap = p.add + row * NO;
#pragma unroll
                            for (int q = 0; q < OM::groups(); ++q) {
                                const int R = OM::rq(q);
                                uint32_t u[4] = {0u, 0u, 0u, 0u};
                                const bf16* ptr = ap + 32 * q + 2 * R * t;
                                if (R == 4) {
                                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(ptr));
                                    u[0] = v.x; u[1] = v.y; u[2] = v.z; u[3] = v.w;
                                } else if (R == 2) {
                                    const uint2 v = __ldg(reinterpret_cast<const uint2*>(ptr));
                                    u[0] = v.x; u[1] = v.y;
                                } else {
#pragma unroll
                                    for (int jj = 0; jj < 3; ++jj)
                                        if (jj < R) u[jj] = __ldg(reinterpret_cast<const uint32_t*>(ptr) + jj);
                                }
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj)
                                    if (jj < R) {
                                        d[4 * q + jj][2 * h] = bf_lo(u[jj]);
                                        d[4 * q + jj][2 * h + 1] = bf_hi(u[jj]);
                                    }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                uint32_t a[4];
                if (xf) {
                    const float4 sc = *reinterpret_cast<const float4*>(&s_scale[16 * s + 4 * t]);
                    const float4 sh = *reinterpret_cast<const float4*>(&s_shift[16 * s + 4 * t]);
                    a[0] = bn_relu_pair(cur[m][s][0].x, sc.x, sc.y, sh.x, sh.y);     // row g,   k 2t,2t+1
                    a[1] = bn_relu_pair(cur[m][s][1].x, sc.x, sc.y, sh.x, sh.y);     // row g+8, k 2t,2t+1
                    a[2] = bn_relu_pair(cur[m][s][0].y, sc.z, sc.w, sh.z, sh.w);     // row g,   k 2t+8,2t+9
                    a[3] = bn_relu_pair(cur[m][s][1].y, sc.z, sc.w, sh.z, sh.w);     // row g+8, k 2t+8,2t+9
                } else {
                    a[0] = cur[m][s][0].x; a[1] = cur[m][s][1].x; a[2] = cur[m][s][0].y; a[3] = cur[m][s][1].y;
                }
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if (BREG) {
                        mma16816(d[j], a, bfr[BREG ? s : 0][BREG ? j : 0][0], bfr[BREG ? s : 0][BREG ? j : 0][1]);
                    } else {
                        const uint2 b = s_b[(s * NT + j) * 32 + lane];
                        mma16816(d[j], a, b.x, b.y);
                    }
                }
            }
            // write-out: row halves h = 0 (row g) / 1 (row g+8)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool ok = h ? ok_hi : ok_lo;
                bf16* op = p.out + (h ? row_hi : row_lo) * NO;
#pragma unroll
                for (int q = 0; q < OM::groups(); ++q) {
                    const int R = OM::rq(q);
                    uint32_t u[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        if (jj < R) u[jj] = ok ? pack_bf16x2(d[4 * q + jj][2 * h], d[4 * q + jj][2 * h + 1]) : 0u;
                    if (ok) {
                        bf16* ptr = op + 32 * q + 2 * R * t;
                        if (R == 4) *reinterpret_cast<uint4*>(ptr) = make_uint4(u[0], u[1], u[2], u[3]);
                        else if (R == 2) *reinterpret_cast<uint2*>(ptr) = make_uint2(u[0], u[1]);
                        else {
#pragma unroll
                            for (int jj = 0; jj < 3; ++jj)
                                if (jj < R) reinterpret_cast<uint32_t*>(ptr)[jj] = u[jj];
                        }
                    }
                    if (MODE == PW_FWD) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (jj < R) {
                                const float v0 = bf_lo(u[jj]), v1 = bf_hi(u[jj]);   // statistics of the stored value
                                const int j = 4 * q + jj;
                                ssum[MODE == PW_FWD ? j : 0][0] += v0;
                                ssum[MODE == PW_FWD ? j : 0][1] += v1;
                                ssq[MODE == PW_FWD ? j : 0][0] = fmaf(v0, v0, ssq[MODE == PW_FWD ? j : 0][0]);
                                ssq[MODE == PW_FWD ? j : 0][1] = fmaf(v1, v1, ssq[MODE == PW_FWD ? j : 0][1]);
                            }
                    }
                }
            }
        }
    }

    if (MODE == PW_FWD && p.stats) {
#pragma unroll
        for (int j = 0; j < (MODE == PW_FWD ? NT : 1); ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float a = ssum[j][e], b = ssq[j][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {            // lanes with the same t hold the same channels
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                }
                if (g == 0) {
                    atomicAdd(&s_red[OM::phys(j, t, e)], a);
                    atomicAdd(&s_red[NO + OM::phys(j, t, e)], b);
                }
            }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * NO; i += blockDim.x) atomicAdd(&p.stats[i], (double)s_red[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct PwWgP {
    const bf16* x;        // [M][Cin]
    const float* scale;   // per Cin, NULL = raw
    const float* shift;
    const bf16* dz;       // [M][Cout]
    float* dw;            // [Cout][Cin] fp32, accumulated into
    long long M;
    int Cin, Cout;
};

// MC m-tiles of 16 output channels (Cout <= 16*MC), NC n-tiles of 8 input channels (Cin == 8*NC)
template <int MC, int NC, int MINB>
__global__ void __launch_bounds__(128, MINB) pw_wgrad_stream_k(const PwWgP p) {
    constexpr int WARPS = 4;
    constexpr int PZ = 32 * MC + 16;                                  // dz row pitch (bytes): odd multiple of 16
    constexpr int PX = (NC % 2) ? 16 * NC + 32 : 16 * NC + 16;        // x row pitch (bytes):  odd multiple of 16, > 16*NC
    constexpr int CZ = 2 * MC;                                        // 16-byte chunks per staged dz row (padded)
    constexpr int CX = NC;                                            // 16-byte chunks per x row
    constexpr int CPR = CZ + CX;
    constexpr int NLD = (16 * CPR + 31) / 32;                         // staged chunks per lane per step
    __shared__ __align__(16) unsigned char s_z[WARPS][16 * PZ];
    __shared__ __align__(16) unsigned char s_x[WARPS][16 * PX];
    __shared__ float s_scale[8 * NC], s_shift[8 * NC];
    __shared__ float s_dw[16 * MC * 8 * NC];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int Cin = p.Cin, Cout = p.Cout;
    const bool xf = p.scale != nullptr;
    const int czv = Cout >> 3;                                        // valid dz chunks per row (Cout % 8 == 0)

    for (int i = threadIdx.x; i < 8 * NC; i += blockDim.x) {
        s_scale[i] = xf ? p.scale[i] : 1.f;
        s_shift[i] = xf ? p.shift[i] : 0.f;
    }
    for (int i = threadIdx.x; i < 16 * MC * 8 * NC; i += blockDim.x) s_dw[i] = 0.f;
    __syncthreads();
    float sc[NC], sh[NC];                                             // channel 8n+g of the B fragments
#pragma unroll
    for (int n = 0; n < NC; ++n) { sc[n] = s_scale[8 * n + g]; sh[n] = s_shift[8 * n + g]; }

    float acc[MC][NC][4];
#pragma unroll
    for (int m = 0; m < MC; ++m)
#pragma unroll
        for (int n = 0; n < NC; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;

    const long long nsteps = (p.M + 15) / 16;
    const long long wstride = (long long)gridDim.x * WARPS;
    long long step = (long long)blockIdx.x * WARPS + warp;

    uint4 nxt[NLD];
    auto load_step = [&](long long st) {
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            const int c = lane + 32 * i;
            const int r = c / CPR, cc = c % CPR;
            const long long row = st * 16 + r;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (c < 16 * CPR && row < p.M) {
                if (cc < CZ) {
                    if (cc < czv) v = __ldg(reinterpret_cast<const uint4*>(p.dz + row * Cout + cc * 8));
                } else {
                    v = __ldg(reinterpret_cast<const uint4*>(p.x + row * Cin + (cc - CZ) * 8));
                }
            }
            nxt[i] = v;
        }
    };
    if (step < nsteps) load_step(step);

    unsigned char* wz = s_z[warp];
    unsigned char* wx = s_x[warp];
    // ldmatrix lane addresses (matrix = lane>>3, row-in-matrix = lane&7)
    const int lm = lane >> 3, lr = lane & 7;
    const unsigned char* a_addr = wz + (lr + ((lm & 2) ? 8 : 0)) * PZ + ((lm & 1) ? 16 : 0);   // + 32*m per m-tile
    const unsigned char* b_addr = wx + (lr + ((lm & 1) ? 8 : 0)) * PX + ((lm & 2) ? 16 : 0);   // + 32*(n/2) per n pair

    for (; step < nsteps; step += wstride) {
        __syncwarp();                                                  // previous step's ldmatrix reads are done
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            const int c = lane + 32 * i;
            const int r = c / CPR, cc = c % CPR;
            if (c < 16 * CPR) {
                if (cc < CZ) *reinterpret_cast<uint4*>(wz + r * PZ + cc * 16) = nxt[i];
                else *reinterpret_cast<uint4*>(wx + r * PX + (cc - CZ) * 16) = nxt[i];
            }
        }
        __syncwarp();
        if (step + wstride < nsteps) load_step(step + wstride);       // lands while this step's MMAs run
        uint32_t b[NC + 1][2];
#pragma unroll
        for (int n = 0; n < NC; n += 2) {
            uint32_t r4[4];
            // matrices: (rows 0-7, ci 8n..), (rows 8-15, ci 8n..), (rows 0-7, ci 8n+8..), (rows 8-15, ci 8n+8..)
            // for an odd NC the last pair reads 16 bytes of row padding into the unused b[NC]
            ldmatrix_x4_trans(r4, b_addr + 16 * n);
            b[n][0] = r4[0]; b[n][1] = r4[1];
            b[n + 1][0] = r4[2]; b[n + 1][1] = r4[3];
        }
        if (xf) {
#pragma unroll
            for (int n = 0; n < NC; ++n) {
                b[n][0] = bn_relu_pair(b[n][0], sc[n], sc[n], sh[n], sh[n]);      // rows 2t, 2t+1   of channel 8n+g
                b[n][1] = bn_relu_pair(b[n][1], sc[n], sc[n], sh[n], sh[n]);      // rows 2t+8, 2t+9
            }
        }
#pragma unroll
        for (int m = 0; m < MC; ++m) {
            uint32_t a[4];
            // matrices: (rows 0-7, co 16m..), (rows 0-7, co 16m+8..), (rows 8-15, co 16m..), (rows 8-15, co 16m+8..)
            ldmatrix_x4_trans(a, a_addr + 32 * m);
#pragma unroll
            for (int n = 0; n < NC; ++n) mma16816(acc[m][n], a, b[n][0], b[n][1]);
        }
    }

    // CTA reduction, then one fp32 atomic per weight per CTA
#pragma unroll
    for (int m = 0; m < MC; ++m)
#pragma unroll
        for (int n = 0; n < NC; ++n) {
            const int co = 16 * m + g, ci = 8 * n + 2 * t;
            atomicAdd(&s_dw[co * 8 * NC + ci], acc[m][n][0]);
            atomicAdd(&s_dw[co * 8 * NC + ci + 1], acc[m][n][1]);
            atomicAdd(&s_dw[(co + 8) * 8 * NC + ci], acc[m][n][2]);
            atomicAdd(&s_dw[(co + 8) * 8 * NC + ci + 1], acc[m][n][3]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) atomicAdd(&p.dw[i], s_dw[i]);   // Cin == 8*NC
}

// ---------------------------------------------------------------------------------------------------------------
template <int KS, int NT, int MT, int MINB>
static int launch_pw(const PwP& p, int mode, cudaStream_t st) {
    const int grid = num_sms() * MINB;
    if (mode == PW_FWD) pw_stream_k<KS, NT, MT, PW_FWD, MINB><<<grid, 128, 0, st>>>(p);
    else pw_stream_k<KS, NT, MT, PW_DGRAD, MINB><<<grid, 128, 0, st>>>(p);
    MNB_LAUNCH_CHECK("pw_stream");
    return 0;
}

// K = channels of the streamed operand, NO = channels written
static int dispatch_pw(const PwP& p, int K, int NO, int mode, cudaStream_t st) {
    if (K == 16 && NO == 48) return launch_pw<1, 6, 2, 4>(p, mode, st);
    if (K == 48 && NO == 16) return launch_pw<3, 2, 2, 4>(p, mode, st);
    if (K == 32 && NO == 16) return launch_pw<2, 2, 2, 4>(p, mode, st);
    if (K == 16 && NO == 32) return launch_pw<1, 4, 2, 4>(p, mode, st);
    if (K == 24 && NO == 72) return launch_pw<2, 9, 1, 3>(p, mode, st);
    if (K == 72 && NO == 24) return launch_pw<5, 3, 1, 4>(p, mode, st);
    set_error("pw_stream: shape %d -> %d not instantiated", K, NO);
    return MNB_ERR_UNSUPPORTED;
}

static bool stream_shape_ok(const ConvP& p) {
    return p.k == 1 && p.stride == 1 && p.pad == 0 && !p.nchw_in && !p.out_f32 && !p.phase_mode && p.wpk != nullptr;
}

int conv_fwd_stream(const ConvP& c, cudaStream_t st) {
    if (!stream_shape_ok(c)) { set_error("pw_stream: needs a packed 1x1 NHWC bf16 problem"); return MNB_ERR_UNSUPPORTED; }
    PwP p = {};
    p.a = (const bf16*)c.x; p.scale = c.in_scale; p.shift = c.in_shift; p.wpk = (const bf16*)c.wpk; p.bias = c.bias;
    p.out = (bf16*)c.out; p.stats = c.stats; p.M = (long long)c.N * c.H * c.W; p.K = c.Cin;
    return dispatch_pw(p, c.Cin, c.Cout, PW_FWD, st);
}

int conv_dgrad_stream(const ConvP& c, cudaStream_t st) {
    if (!stream_shape_ok(c) || c.bn_z) { set_error("pw_stream: needs a packed 1x1 NHWC bf16 problem"); return MNB_ERR_UNSUPPORTED; }
    PwP p = {};
    p.a = (const bf16*)c.dz; p.wpk = (const bf16*)c.wpk; p.add = (const bf16*)c.add; p.out = (bf16*)c.out;
    p.M = (long long)c.N * c.H * c.W; p.K = c.Cout;
    return dispatch_pw(p, c.Cout, c.Cin, PW_DGRAD, st);
}

template <int MC, int NC, int MINB>
static int launch_wg(const PwWgP& p, cudaStream_t st) {
    pw_wgrad_stream_k<MC, NC, MINB><<<num_sms() * MINB, 128, 0, st>>>(p);
    MNB_LAUNCH_CHECK("pw_wgrad_stream");
    return 0;
}

int conv_wgrad_stream(const ConvP& c, cudaStream_t st) {
    if (c.k != 1 || c.stride != 1 || c.pad != 0 || c.nchw_in) { set_error("pw_stream: 1x1 NHWC only"); return MNB_ERR_UNSUPPORTED; }
    PwWgP p = {};
    p.x = (const bf16*)c.x; p.scale = c.in_scale; p.shift = c.in_shift; p.dz = (const bf16*)c.dz; p.dw = c.dw;
    p.M = (long long)c.N * c.H * c.W; p.Cin = c.Cin; p.Cout = c.Cout;
    const int Cin = c.Cin, Cout = c.Cout;
    if (Cin == 16 && Cout == 48) return launch_wg<3, 2, 4>(p, st);
    if (Cin == 48 && Cout == 16) return launch_wg<1, 6, 4>(p, st);
    if (Cin == 32 && Cout == 16) return launch_wg<1, 4, 4>(p, st);
    if (Cin == 24 && Cout == 72) return launch_wg<5, 3, 3>(p, st);
    // 72 -> 24 (MC = 2, NC = 9) measured slower than the tcgen05 kernel (109 vs 86 us at 256x56x56): not dispatched
    set_error("pw_wgrad_stream: shape %d -> %d not instantiated", Cin, Cout);
    return MNB_ERR_UNSUPPORTED;
}

}  // namespace mnb
