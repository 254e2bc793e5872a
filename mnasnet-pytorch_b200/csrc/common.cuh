// Shared helpers for the mnb200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mnb200.h"

namespace mnb {

typedef __nv_bfloat16 bf16;

void set_error(const char* fmt, ...);

// run-time kernel-selection switches (mnb_set_option / environment, see include/mnb200.h)
enum { OPT_PW_STREAM = 0, OPT_STEM_MMA, OPT_DW_MMA, OPT_DW_MMA_CG, OPT_DW_MMA_TWS, OPT_DW_MMA_SEG, OPT_BN_CTAS, OPT_C3_MMA,
       OPT_DW_SMALL, OPT_PWB_SLICE, OPT_PW_WIDE, OPT_COUNT };
int option_get(int id);

#define MNB_REQUIRE(cond, ...)               \
    do {                                     \
        if (!(cond)) {                       \
            mnb::set_error(__VA_ARGS__);     \
            return MNB_ERR_ARG;              \
        }                                    \
    } while (0)

#define MNB_LAUNCH_CHECK(name)                                            \
    do {                                                                  \
        cudaError_t e__ = cudaGetLastError();                             \
        if (e__ != cudaSuccess) {                                         \
            mnb::set_error("%s: %s", name, cudaGetErrorString(e__));      \
            return (int)e__;                                              \
        }                                                                 \
    } while (0)

static inline int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// ---- element access: everything is computed in fp32 ------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements (16 B for bf16, 32 B for fp32); pointers must be 16-B aligned.
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&v)[8]) {
    // bf16 -> fp32 is a 16-bit shift
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    unpack_bf16x8(u, v);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    return u;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = pack_bf16x8(v);
}

// 2 consecutive elements
__device__ __forceinline__ float2 load2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 load2(const bf16* p) {
    uint32_t u = *reinterpret_cast<const uint32_t*>(p);
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void store2(float* p, float a, float b) {
    *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
__device__ __forceinline__ void store2(bf16* p, float a, float b) {
    *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(a, b);
}

// largest divisor of cv that is <= cap (thread-x extent for [rows][cv] column-owner kernels)
static inline int largest_divisor_le(int cv, int cap) {
    for (int d = cap; d >= 1; --d)
        if (cv % d == 0) return d;
    return 1;
}

}  // namespace mnb
