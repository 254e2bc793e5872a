// How long does the per-CTA flush of BatchNorm partial sums take?  Every CTA adds 2*C values to the same 2*C addresses.
// Variants: fp64 atomics (what the kernels do), fp32 atomics, fp64 atomics on 16 replicated slots.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/ub_atomic ub_atomic.cu && ./build/ub_atomic
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void flush_k(double* out64, float* out32, int n2c, int slots) {
    // a little independent work so that CTAs do not all arrive in the same cycle
    float v = threadIdx.x * 1e-3f + blockIdx.x;
    for (int i = threadIdx.x; i < n2c; i += blockDim.x) {
        if (MODE == 0) atomicAdd(&out64[i], (double)v);
        if (MODE == 1) atomicAdd(&out32[i], v);
        if (MODE == 2) atomicAdd(&out64[(blockIdx.x % slots) * n2c + i], (double)v);
    }
}
__global__ void empty_k() {}

int main() {
    double* d64; float* d32;
    cudaMalloc(&d64, 64 * 4096 * sizeof(double)); cudaMalloc(&d32, 4096 * sizeof(float));
    cudaMemset(d64, 0, 64 * 4096 * sizeof(double)); cudaMemset(d32, 0, 4096 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](auto launch, const char* name, int ctas, int n2c) {
        for (int i = 0; i < 3; ++i) launch();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s ctas %4d  2C %4d : %7.2f us per launch\n", name, ctas, n2c, ms * 1000 / 20);
    };
    run([&] { empty_k<<<148, 256>>>(); }, "empty kernel", 148, 0);
    for (int n2c : {96, 480, 1152, 2304})
        for (int ctas : {148, 296, 444}) {
            run([&] { flush_k<0><<<ctas, 256>>>(d64, d32, n2c, 1); }, "fp64 atomics", ctas, n2c);
            run([&] { flush_k<1><<<ctas, 256>>>(d64, d32, n2c, 1); }, "fp32 atomics", ctas, n2c);
            run([&] { flush_k<2><<<ctas, 256>>>(d64, d32, n2c, 16); }, "fp64 atomics, 16 slots", ctas, n2c);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
