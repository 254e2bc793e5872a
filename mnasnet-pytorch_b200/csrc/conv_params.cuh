// Parameter block shared by the SIMT (conv_simt.cu) and tcgen05 (gemm_tc.cu) dense-convolution paths.
#pragma once
#include "common.cuh"

namespace mnb {

enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

struct ConvP {
    const void* x;        // fwd/wgrad: input activations; dgrad: unused
    const float* in_scale;
    const float* in_shift;
    const float* w;       // [Cout,Cin,k,k]
    const void* wpk;      // optional bf16 K-major packing of w for this GEMM (mnb_pack_weights)
    const float* bias;
    const void* dz;       // dgrad/wgrad
    const void* add;      // dgrad residual
    void* out;            // fwd: z; dgrad: dx
    float* dw;            // wgrad
    double* stats;        // fwd: BN statistics of z; dgrad: fused BN-backward sums of the producer block
    const void* bn_z;     // dgrad: raw conv output of the producer block (NULL = no fused reduction)
    const float* bn_scale;
    const float* bn_shift;
    int N, H, W, Cin, Ho, Wo, Cout, k, stride, pad;
    int nchw_in;          // x is NCHW fp32
    // stride-2 3x3 backward-data, decomposed by input-position parity (ph_h, ph_w): only the taps with
    // (h+1-kh) and (w+1-kw) even contribute -> 4 dense sub-problems with 1/2/2/4 taps instead of 9 mostly-zero ones
    int phase_mode, ph_h, ph_w;
    int out_f32;          // fwd/dgrad: write fp32 rows straight from the accumulator (classifier GEMMs)
    int relu_out;         // with out_f32: apply ReLU
    long long kchunk;     // wgrad: positions per blockIdx.z
};

// tcgen05 path (gemm_tc.cu); returns MNB_ERR_UNSUPPORTED when the shape is not covered
int conv_fwd_tc(const ConvP& p, cudaStream_t st);
int conv_dgrad_tc(const ConvP& p, cudaStream_t st);
int conv_wgrad_tc(const ConvP& p, cudaStream_t st);

// warp-streaming mma.sync path for the small-channel 1x1 layers (pw_stream.cu); MNB_ERR_UNSUPPORTED otherwise
int conv_fwd_stream(const ConvP& p, cudaStream_t st);
int conv_dgrad_stream(const ConvP& p, cudaStream_t st);
int conv_wgrad_stream(const ConvP& p, cudaStream_t st);

// TMA + mma.sync path for the small-channel 3x3 stage transitions (c3_mma.cu); MNB_ERR_UNSUPPORTED otherwise
int conv_fwd_c3(const ConvP& p, cudaStream_t st);
int conv_dgrad_c3(const ConvP& p, cudaStream_t st);
int conv_wgrad_c3(const ConvP& p, cudaStream_t st);

// cp.async + mma.sync forward for the wide 1x1 layers of the 28x28 / 14x14 stages (pw_wide_fwd.cu); MNB_ERR_UNSUPPORTED otherwise
int conv_fwd_pwide(const ConvP& p, cudaStream_t st);

}  // namespace mnb
