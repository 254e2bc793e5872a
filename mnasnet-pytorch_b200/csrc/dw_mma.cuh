// PTX helpers shared by the TMA + mma.sync depthwise kernels (dw_mma.cu, dw_mma_bwd.cu).  sm_100a only.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mnb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier (one per CTA: "the TMA boxes of this step have landed") ----------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a lost TMA completion traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
        // suspend-time hint (ns): the thread sleeps in hardware until the phase completes instead of spinning through issue slots
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
        if (spin > (1 << 20)) __trap();
    }
}

// ---- TMA: 4-D tiled boxes of an NHWC bf16 tensor, coordinates {c, w, h, n} --------------------------------
// Loads may start at negative coordinates / overhang the tensor: out-of-range elements are zero-filled, which IS the
// convolution's zero padding.  Stores clip on the high side only: negative start coordinates fault on B200
// (measured: scripts/ubench/ub_mma.cu), so every store in these kernels starts inside the tensor.
__device__ __forceinline__ void tma_load4(uint32_t dst, const CUtensorMap* tm, int c, int w, int h, int n, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(tm), "r"(c), "r"(w), "r"(h), "r"(n), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store4(const CUtensorMap* tm, int c, int w, int h, int n, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(tm), "r"(c), "r"(w), "r"(h), "r"(n), "r"(src) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (TMA store source, TMA load overwrite)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- 1-D bulk copies (contiguous tiles) ---------------------------------------------------------------------
__device__ __forceinline__ void bulk_store1(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_load1(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void ldsm1(uint32_t addr, uint32_t& r0) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(addr));
}

// ---- ldmatrix / mma.sync ------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm2t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// D(16x8) += A(16x8) * B(8x8)
__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// exact floor(n / d) for n < 65536 and d < 65536: q = (n * magic) >> 32, magic = floor(2^32 / d) + 1
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, uint32_t magic) { return __umulhi(n, magic); }

// ---- dense-GEMM helpers shared by c3_mma.cu / pw_proj_bwd.cu: weight rows [n][K] in shared memory, pitch 16 B x odd ----
__host__ __device__ constexpr int c3_odd16(int bytes) { return ((bytes / 16) | 1) * 16; }     // 16-byte multiple x odd: conflict-free ldmatrix rows
__host__ __device__ constexpr int c3_al128(int bytes) { return (bytes + 127) / 128 * 128; }

// B fragments of NT n-tiles for one k16 / k8 step at byte offset kb inside the weight rows
template <int NT, int WP>
__device__ __forceinline__ void c3_load_b16(uint32_t (&b)[NT][2], uint32_t b4, uint32_t b2, int kb) {
#pragma unroll
    for (int jp = 0; jp < NT / 2; ++jp) ldsm4(b4 + jp * 16 * WP + kb, b[2 * jp][0], b[2 * jp][1], b[2 * jp + 1][0], b[2 * jp + 1][1]);
    if constexpr (NT & 1) ldsm2(b2 + kb, b[NT - 1][0], b[NT - 1][1]);
}
template <int NT, int WP>
__device__ __forceinline__ void c3_load_b8(uint32_t (&b)[NT][2], uint32_t b8, int kb) {
#pragma unroll
    for (int q = 0; q < NT / 4; ++q) ldsm4(b8 + q * 32 * WP + kb, b[4 * q][0], b[4 * q + 1][0], b[4 * q + 2][0], b[4 * q + 3][0]);
    constexpr int R = NT & 3, Q = NT / 4 * 4;
    if constexpr (R >= 2) ldsm2(b8 + Q * 8 * WP + kb, b[Q][0], b[Q + 1][0]);
    if constexpr (R & 1) ldsm1(b8 + (NT - 1) * 8 * WP + kb, b[NT - 1][0]);
}

// ---- host: tensor maps ---------------------------------------------------------------------------------------
// 4-D map of an NHWC bf16 tensor with box {bc channels, bw columns, bh rows, 1 image}; cached per (pointer, shape,
// box) -- the engine's buffers are static, so every map is encoded once per plan.
int dwm_tensor_map(CUtensorMap* out, const void* ptr, int N, int H, int W, int C, int bc, int bw, int bh);

// channel-group / strip geometry shared by the forward and backward launchers: a CTA owns CG channels (NCH = CG/8
// chunks of 8, CG*2 bytes per pixel = 16 B x odd so that 8 consecutive pixels of one chunk fall into 8 different
// 16-byte bank groups: conflict-free ldmatrix without padding or swizzle) x TWS strips of 16 output columns; one warp
// per (chunk, strip).
struct DwmGeom {
    int CG, NCH, TWS, TW, HC, PITCH, tiles_w, cblocks, threads;
};
DwmGeom dwm_geometry(int C, int W, int K);

}  // namespace mnb
