// Micro-benchmarks that size the depthwise tensor-pipe kernels (csrc/dw_mma.cu) on B200:
//   1. mma.sync m16n8k16 / m16n8k8 bf16 issue rate per SM (the legacy HMMA path, the only MMA shape that can
//      carry a per-channel "diagonal" B operand without transposing NHWC data)
//   2. ldmatrix.x4 rate per SM
//   3. TMA 4-D tiled load with negative / out-of-range coordinates (zero fill = conv padding) and TMA store
//      with clipping, checked against the host: the exact tensor-map usage of the halo tiles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ub_mma ub_mma.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}

template <int MODE, int NACC>
__global__ void mma_rate_k(float* out, int iters, long long* clk) {
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = 0x3f803f80, b1 = 0x3f803f80;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) mma16816(c[i], a0, a1, a2, a3, b0, b1);
            else mma1688(c[i], a0, a1, b0);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

__global__ void ldsm_rate_k(uint32_t* out, int iters, long long* clk) {
    __shared__ __align__(128) unsigned char sm[16384];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) ((uint32_t*)sm)[i] = i;
    __syncthreads();
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x % 32) * 16 + (threadIdx.x / 32) * 512;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t r0, r1, r2, r3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr + ((u * 1024) & 8191)));
            acc += r0 ^ r1 ^ r2 ^ r3;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

// ---- TMA halo tile load + store ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void tma_tile_k(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, int c0, int w0,
                           int h0, int n, int box_bytes, unsigned short* dump, int ow0, int oh0, int mode) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (!(mode & 1)) {      // no load: fill the tile by hand
        for (int i = threadIdx.x; i < box_bytes / 2; i += blockDim.x) ((unsigned short*)sm)[i] = 0x3f80;
        __syncthreads();
    }
    if ((mode & 1) && threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(smem_u32(sm)), "l"(&tin), "r"(c0), "r"(w0), "r"(h0), "r"(n), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = (mode & 1) ? 0 : 1;
    int spins = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        if (++spins > 1000000) { if (threadIdx.x == 0) printf("TMA wait timed out\n"); return; }
    }
    for (int i = threadIdx.x; i < box_bytes / 2; i += blockDim.x) dump[i] = ((unsigned short*)sm)[i];
    // in-place edit through the generic proxy, then store the same box elsewhere through the async proxy
    for (int i = threadIdx.x; i < box_bytes / 2; i += blockDim.x) ((unsigned short*)sm)[i] ^= 0x8000;   // flip sign
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if ((mode & 2) && threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                     ::"l"(&tout), "r"(c0), "r"(ow0), "r"(oh0), "r"(n), "r"(smem_u32(sm)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeFn enc, CUtensorMap* m, void* ptr, int N, int H, int W, int C, int bc, int bw, int bh) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    return 0;
}

static int tma_test(int N, int H, int W, int C, int bc, int bw, int bh, int c0, int w0, int h0, int n, int mode = 3) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    EncodeFn enc = (EncodeFn)fn;
    size_t nel = (size_t)N * H * W * C;
    std::vector<unsigned short> hx(nel);
    for (size_t i = 0; i < nel; ++i) hx[i] = (unsigned short)(0x3f00 + (i * 2654435761u >> 20) % 0x7f);   // finite positive bf16 bits
    unsigned short *dx, *dy, *dd;
    CK(cudaMalloc(&dx, nel * 2)); CK(cudaMalloc(&dy, nel * 2));
    int box_el = bc * bw * bh;
    CK(cudaMalloc(&dd, box_el * 2));
    CK(cudaMemcpy(dx, hx.data(), nel * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dy, 0, nel * 2));
    CUtensorMap tin, tout;
    if (make_map(enc, &tin, dx, N, H, W, C, bc, bw, bh) || make_map(enc, &tout, dy, N, H, W, C, bc, bw, bh)) return 1;
    CK(cudaFuncSetAttribute(tma_tile_k, cudaFuncAttributeMaxDynamicSharedMemorySize, box_el * 2 + 128));
    tma_tile_k<<<1, 128, box_el * 2 + 128>>>(tin, tout, c0, w0, h0, n, box_el * 2, dd, w0, h0, mode);
    {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("TMA mode %d box(c%d,w%d,h%d) at (c%d,w%d,h%d): CUDA error %s\n", mode, bc, bw, bh, c0, w0, h0, cudaGetErrorString(e)); return 1; }
    }
    std::vector<unsigned short> hd(box_el), hy(nel);
    CK(cudaMemcpy(hd.data(), dd, box_el * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hy.data(), dy, nel * 2, cudaMemcpyDeviceToHost));
    long long bad = 0, bad_store = 0;
    for (int r = 0; r < bh; ++r)
        for (int cc = 0; cc < bw; ++cc)
            for (int ch = 0; ch < bc; ++ch) {
                int ih = h0 + r, iw = w0 + cc, ic = c0 + ch;
                bool in = ih >= 0 && ih < H && iw >= 0 && iw < W && ic >= 0 && ic < C;
                unsigned short exp = in ? hx[(((size_t)n * H + ih) * W + iw) * C + ic] : 0;
                unsigned short got = hd[((size_t)r * bw + cc) * bc + ch];
                if ((mode & 1) && exp != got) ++bad;
            }
    for (int ih = 0; ih < H; ++ih)
        for (int iw = 0; iw < W; ++iw)
            for (int ic = 0; ic < C; ++ic) {
                bool in = ih >= h0 && ih < h0 + bh && iw >= w0 && iw < w0 + bw && ic >= c0 && ic < c0 + bc;
                size_t idx = (((size_t)n * H + ih) * W + iw) * C + ic;
                unsigned short exp = in ? (unsigned short)(hx[idx] ^ 0x8000) : 0;
                if (!(mode & 1)) exp = in ? (unsigned short)(0x3f80 ^ 0x8000) : 0;
                if ((mode & 2) && hy[idx] != exp) ++bad_store;
            }
    printf("TMA N%d H%d W%d C%d box(c%d,w%d,h%d) at (c%d,w%d,h%d,n%d): load mismatches %lld, store mismatches %lld  [smem layout dense [h][w][c]]\n",
           N, H, W, C, bc, bw, bh, c0, w0, h0, n, bad, bad_store);
    cudaFree(dx); cudaFree(dy); cudaFree(dd);
    return (bad || bad_store) ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc > 1) {      // one TMA variant per process (a faulting kernel poisons the context)
        int v = atoi(argv[1]);
        switch (v) {
            case 0: return tma_test(2, 32, 32, 64, 64, 16, 8, 0, 0, 0, 0, 1);       // in-bounds load, 128-byte rows
            case 1: return tma_test(2, 32, 32, 64, 64, 16, 8, 0, 0, 0, 0, 2);       // in-bounds store
            case 2: return tma_test(2, 32, 32, 64, 24, 16, 8, 24, 4, 4, 1, 1);      // 48-byte rows, in-bounds load
            case 3: return tma_test(2, 14, 14, 48, 24, 20, 11, 24, -2, -2, 1, 1);   // negative coordinates, load
            case 4: return tma_test(2, 14, 14, 48, 24, 20, 11, 24, -2, -2, 1, 2);   // negative coordinates, store
            case 5: return tma_test(2, 14, 14, 48, 24, 20, 11, 0, 6, 9, 0, 3);
            case 6: return tma_test(1, 7, 7, 32, 40, 12, 9, 0, -1, -1, 0, 3);
            case 7: return tma_test(3, 30, 20, 72, 72, 36, 18, 0, -2, 14, 2, 3);
            case 8: return tma_test(2, 14, 14, 48, 24, 20, 11, 24, -2, -2, 1, 3);
        }
        return 0;
    }
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    printf("SMs %d, clock %d kHz\n", sms, khz);
    float* out; long long* clk;
    CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&clk, 8));
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) mma_rate_k<0, 8><<<sms, warps * 32>>>(out, iters, clk);
                else mma_rate_k<1, 8><<<sms, warps * 32>>>(out, iters, clk);
                CK(cudaDeviceSynchronize());
            }
            long long c;
            CK(cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost));
            double per_sm_per_clk = (double)iters * 8 * warps / (double)c;
            double flops = per_sm_per_clk * (mode == 0 ? 4096.0 : 2048.0);
            printf("mma.sync %s bf16: %2d warps/SM: %.3f MMA/clk/SM = %.0f FLOP/clk/SM = %.0f TFLOP/s at %.3f GHz\n",
                   mode == 0 ? "m16n8k16" : "m16n8k8 ", warps, per_sm_per_clk, flops, flops * sms * khz * 1e3 / 1e12, khz / 1e6);
        }
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            ldsm_rate_k<<<sms, warps * 32>>>((uint32_t*)out, iters, clk);
            CK(cudaDeviceSynchronize());
        }
        long long c;
        CK(cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost));
        double r = (double)iters * 8 * warps / (double)c;
        printf("ldmatrix.x4: %2d warps/SM: %.3f instr/clk/SM = %.0f B/clk/SM\n", warps, r, r * 512);
    }
    int fails = 0;
    printf(fails ? "UBENCH TMA FAIL\n" : "UBENCH TMA OK\n");
    return fails;
}
