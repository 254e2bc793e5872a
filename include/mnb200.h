/*
 * mnb200.h -- C ABI of libmnb200.so: hand-written sm_100a CUDA kernels for the MNASNet training step
 * of snakers4/mnasnet-pytorch (forward / backward / optimizer), NHWC activations.
 *
 * The reference has no FFI of its own; the boundary it exposes is the torch operator API reached from
 *   src/models/mnasnet.py:37-62   ConvBlock  = nn.Conv2d(bias) -> nn.BatchNorm2d -> nn.ReLU
 *   src/models/mnasnet.py:131-133 MBConv_block residual add
 *   src/models/classifiers.py:107-111 AdaptiveAvgPool2d(1) -> Dropout/Linear/ReLU head
 *   src/train.py:277,435-440      CrossEntropyLoss, backward, Adam.step
 * Every entry point below names the reference call it replaces.  See INTEGRATION.md for the ctypes
 * binding used by the drop-in `models/` package.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer owned by the caller; the library keeps
 *     no reference after the call returns and allocates nothing (except a per-process weight scratch, see
 *     mnb_gemm_workspace_bytes).
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*); no implicit sync.
 *   - return value: 0 ok; <0 argument error (MNB_ERR_*); >0 cudaError_t.  Message: mnb_last_error().
 *   - activations: NHWC, dtype MNB_F32 or MNB_BF16, C % 8 == 0 (except the NCHW fp32 network input);
 *     parameters, gradients of parameters, BN vectors: fp32; BN statistics accumulators: fp64.
 *   - "x-transform": when in_scale != NULL the kernel reads  a = max(in_scale[c]*x + in_shift[c], 0)
 *     (the previous ConvBlock's BN-apply + ReLU fused into the load); NULL -> x is used as is.
 *   - stats: double[2*C] = {sum z, sum z^2} per channel, ACCUMULATED (caller zeroes before the layer).
 */
#ifndef MNB200_H
#define MNB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MNB_F32 0
#define MNB_BF16 1

#define MNB_ERR_ARG (-1)        /* bad shape / alignment / dtype */
#define MNB_ERR_UNSUPPORTED (-2)

#define MNB_LAYOUT_NHWC 0
#define MNB_LAYOUT_NCHW_F32 1   /* network input: N x C x H x W fp32 (train.py:427) */
#define MNB_LAYOUT_NHWC_U8 2    /* network input: N x H x W x 3 uint8 as decoded (utils/datasets.py:456-462); the stem applies
                                   ToTensor + Normalize itself: in_scale = mean[3], in_shift = std[3] (classifiers.py:91-92) */

int mnb_version(void);
/* Kernel-selection switches (process-wide; an unset option takes MNB_<NAME> from the environment, then its default):
 *   "pw_stream" (default 1)  auto routes the low-channel bf16 1x1 layers to the warp-streaming kernels (pw_stream.cu)
 *   "stem_mma"  (default 1)  auto uses the tensor-pipe stem backward-weight kernel in bf16 mode
 *   "dw_mma" (default 1)     bf16 depthwise layers on the TMA + mma.sync kernels (dw_mma.cu): 1 = where they measured
 *                            faster than the tile kernels (forward / backward-data on maps of >= 12 rows, backward-weight
 *                            when additionally C % 24 == 0), 2 = every shape, 0 = never
 *   "dw_mma_cg" / "dw_mma_tws" / "dw_mma_seg" (default 0 = automatic)  geometry overrides of those kernels: channels per
 *                            CTA (24 | 40), 16-column strips per CTA (1 | 2), row blocks per work item
 *   "bn_ctas" (default 0 = sized to the tensor)  CTAs per SM of mnb_bn_bwd_reduce
 *   "dw_small"  (default 1)  whole-tile tensor-pipe depthwise kernels (dw_small.cu): forward, backward-data and backward-weight
 *                            on maps of 12..64 rows, the fused backward on every map of >= 12 rows (5x5: up to 64 rows);
 *                            2 = every shape, 0 = never (row-streaming / tile kernels)
 *   "pwb_slice" (default 0)  48 | 80: mnb_pw_bwd_fused also takes Cout = 40 layers whose Cin is a multiple of it, one CTA column
 *                            per slice of input channels (measured break-even, kept for experiments)
 *   "pw_wide"   (default 0)  1 = cp.async + mma.sync forward (pw_wide_fwd.cu) for the wide 1x1 layers 40<->240, 80<->480,
 *                            96<->576 under impl 0 / 3 (measured 15-35 % slower than the tcgen05 pipeline: kept for experiments)
 *   "c3_mma"    (default 1)  bulk-copy + mma.sync kernels for the stride-2 3x3 stage transitions 16->24, 24->40 (forward,
 *                            backward-data, backward-weight) and 40->80 (backward-data) under impl 0 / 3 (c3_mma.cu)
 * mnb_set_option returns 0 or MNB_ERR_ARG (unknown name); mnb_get_option the current value or MNB_ERR_ARG. */
int mnb_set_option(const char* name, int value);
int mnb_get_option(const char* name);
const char* mnb_last_error(void);
/* 1 if the running device is sm_100 (tcgen05 path usable) */
int mnb_device_is_sm100(void);

/* ---- dense convolution (groups=1; k in {1,3}; stride in {1,2}) : nn.Conv2d, mnasnet.py:48-54 -------------
 * z[n,ho,wo,co] = bias[co] + sum_{kh,kw,ci} a(n, ho*stride-pad+kh, wo*stride-pad+kw, ci) * w[co,ci,kh,kw]
 * w is the torch layout [Cout,Cin,k,k] fp32.  x_layout selects NHWC(dtype) or NCHW fp32 (stem only).
 * impl: 0 = auto: bf16 1x1 layers with Cin, Cout <= 72 and a packed weight take the warp-streaming mma.sync kernels
 *           (pw_stream.cu; backward-weight needs no packing), everything else bf16 takes tcgen05, fp32 takes SIMT;
 *       1 = force SIMT, 2 = force tcgen05 (bf16), 3 = like auto regardless of MNB_PW_STREAM / MNB_STEM_MMA.
 * mnb_set_option("pw_stream" / "stem_mma", 0) keeps auto on the tcgen05 / fp32-input stem kernels. */
int mnb_conv_fwd(const void* x, const float* in_scale, const float* in_shift, const float* w, const float* bias,
                 void* z, double* stats, int N, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                 int dtype, int x_layout, int impl, void* stream);
/* Optional bf16 re-layout of a dense conv weight for the tcgen05 path (call once per optimizer step):
 *   wpk_fwd  [Cout][k*k*Cin]  with kk = (kh*k+kw)*Cin + ci   (K-major B operand of the forward GEMM)
 *   wpk_dgrad[Cin ][k*k*Cout] with kk = (kh*k+kw)*Cout + co  (K-major B operand of the backward-data GEMM)
 * Either pointer may be NULL.  Register the result with mnb_conv_fwd_packed / mnb_conv_dgrad_packed. */
int mnb_pack_weights(const float* w, void* wpk_fwd, void* wpk_dgrad, int Cout, int Cin, int k, void* stream);
/* Same as mnb_conv_fwd / mnb_conv_dgrad (bf16, tcgen05) with the pre-packed weight: the GEMM producers then copy
 * the B operand with 16-byte cp.async instead of gathering fp32 scalars. */
int mnb_conv_fwd_packed(const void* x, const float* in_scale, const float* in_shift, const float* w,
                        const void* wpk_fwd, const float* bias, void* z, double* stats, int N, int H, int W, int Cin,
                        int Cout, int k, int stride, int pad, int dtype, int x_layout, int impl, void* stream);
int mnb_conv_dgrad_packed(const void* dz, const float* w, const void* wpk_dgrad, const void* add, void* dx,
                          const void* bn_z, const float* bn_scale, const float* bn_shift, double* bn_sums, int N, int H,
                          int W, int Cin, int Cout, int k, int stride, int pad, int dtype, int impl, void* stream);
/* dx[n,h,w,ci] = (add ? add[n,h,w,ci] : 0) + sum dz[n,ho,wo,co]*w[co,ci,kh,kw]   (conv backward-data)
 * Optional fused BatchNorm-backward reduction for the ConvBlock that PRODUCED this conv's input (dx is its dA):
 * when bn_z != NULL, bn_sums[0:Cin] += sum G and bn_sums[Cin:2Cin] += sum G*z with
 * G = dx * [bn_scale*z + bn_shift > 0], z = bn_z (same shape as dx) -- i.e. what mnb_bn_bwd_reduce computes. */
int mnb_conv_dgrad(const void* dz, const float* w, const void* add, void* dx, const void* bn_z,
                   const float* bn_scale, const float* bn_shift, double* bn_sums, int N, int H, int W, int Cin,
                   int Cout, int k, int stride, int pad, int dtype, int impl, void* stream);
/* dw[co,ci,kh,kw] += sum a(...)*dz[...]  (conv backward-weight; fp32 accumulate INTO dw) */
int mnb_conv_wgrad(const void* x, const float* in_scale, const float* in_shift, const void* dz, float* dw,
                   int N, int H, int W, int Cin, int Cout, int k, int stride, int pad, int dtype, int x_layout,
                   int impl, void* stream);

/* ---- depthwise convolution k in {3,5}, stride 1, pad k/2 : nn.Conv2d(groups=C), mnasnet.py:76-81,120-125 -- */
int mnb_dw_fwd(const void* x, const float* in_scale, const float* in_shift, const float* w /*[C,1,k,k]*/,
               const float* bias, void* z, double* stats, int N, int H, int W, int C, int k, int dtype,
               void* stream);
/* bn_z/bn_scale/bn_shift/bn_sums: optional fused BN-backward reduction, see mnb_conv_dgrad */
int mnb_dw_dgrad(const void* dz, const float* w, void* dx, const void* bn_z, const float* bn_scale,
                 const float* bn_shift, double* bn_sums, int N, int H, int W, int C, int k, int dtype,
                 void* stream);
int mnb_dw_wgrad(const void* x, const float* in_scale, const float* in_shift, const void* dz, float* dw,
                 int N, int H, int W, int C, int k, int dtype, void* stream);
/* Fused backward of one depthwise ConvBlock (bf16 only; TMA + mma.sync kernel, csrc/dw_mma.cu): the BatchNorm-backward
 * elementwise pass, conv backward-data, conv backward-weight and the BatchNorm-backward REDUCTIONS of the ConvBlock
 * that produced this block's input, in one pass over dA, z and x (4 tensor passes instead of 9).  Replaces what autograd
 * runs for ConvBlock.forward (src/models/mnasnet.py:58-62) of a depthwise block: threshold_backward +
 * native_batch_norm_backward + convolution_backward, and the reductions of the next native_batch_norm_backward.
 *   G  = dA * [scale*z + shift > 0]          sums = {sum G, sum G*z} of THIS block (mnb_bn_bwd_reduce or a fused producer)
 *   dZ = a*G + b*z + c  (a, b, c from sums / save_mean / save_invstd / scale / m, as mnb_bn_bwd_apply_fused)
 *   dgamma / dbeta / dbias += (each may be NULL)       dw[c,kh,kw] += sum dZ * relu(in_scale*x + in_shift) shifted
 *   dx = conv backward-data of dZ                       (dw may be NULL: frozen weight)
 *   in_sums[0:C] += sum dx', in_sums[C:2C] += sum dx'*x with dx' = dx * [in_scale*x + in_shift > 0]  (NULL = skip)
 * x is the RAW output of the producing block (its BN scale / shift given as in_scale / in_shift, NULL = x is the
 * activation itself and in_sums must be NULL).  Returns MNB_ERR_UNSUPPORTED for dtype != bf16. */
int mnb_dw_bwd_fused(const void* dA, const void* z, const float* scale, const float* shift, const double* sums,
                     const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, float* dbias,
                     const void* x, const float* in_scale, const float* in_shift, const float* w, void* dx, float* dw,
                     double* in_sums, int N, int H, int W, int C, int k, double m, int dtype, void* stream);
/* Fused backward of one pointwise (1x1) ConvBlock (bf16; csrc/pw_bwd_fused.cu), same contract as mnb_dw_bwd_fused with
 * the tensors seen as matrices: dA, z are M x Cout; x, dx (and the optional residual skip gradient `add`, summed into
 * dx) are M x Cin; w, dw are [Cout][Cin] (torch [Cout,Cin,1,1]).  Instantiated for the expand / project blocks of the
 * 112x112 and 56x56 stages (Cin -> Cout in {16->48, 48->16, 32->16, 24->72, 72->24}; src/models/mnasnet.py:116-129,
 * 82-85); other shapes return MNB_ERR_UNSUPPORTED and the caller keeps the unfused chain. */
int mnb_pw_bwd_fused(const void* dA, const void* z, const float* scale, const float* shift, const double* sums,
                     const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, float* dbias,
                     const void* x, const float* in_scale, const float* in_shift, const float* w, const void* add,
                     void* dx, float* dw, double* in_sums, long long M, int Cin, int Cout, double m, int dtype,
                     void* stream);
/* Backward of a "project" pointwise ConvBlock (wide input, narrow output) AFTER its BatchNorm-backward pass (bf16;
 * csrc/pw_proj_bwd.cu): from dz (M x Cout, the output of mnb_bn_bwd_apply_fused) one kernel computes
 *   dx = dz W (+ add)                      M x Cin, `add` = optional residual skip gradient
 *   dw[Cout][Cin] += dz^T relu(in_scale*x + in_shift)                       (dw may be NULL: frozen weight)
 *   in_sums[0:Cin] += sum dx', in_sums[Cin:2Cin] += sum dx'*x, dx' = dx * [in_scale*x + in_shift > 0]   (NULL = skip)
 * i.e. convolution_backward of a 1x1 conv plus the reductions of the next native_batch_norm_backward
 * (src/models/mnasnet.py:58-62,120-128 under autograd).  Instantiated for Cin -> Cout in {240->40, 480->80, 576->96}
 * (Cout == 96 | 80 with Cin % 96 == 0, Cout == 40 with Cin % 80 == 0); other shapes return MNB_ERR_UNSUPPORTED. */
int mnb_pw_proj_bwd(const void* dz, const void* x, const float* in_scale, const float* in_shift, const float* w,
                    const void* add, void* dx, float* dw, double* in_sums, long long M, int Cin, int Cout, int dtype,
                    void* stream);

/* ---- BatchNorm2d (train) : mnasnet.py:55,60 ; torch:nn/modules/batchnorm.py:163-178 ----------------------
 * finalize: mean/var from stats over m positions -> scale = gamma/sqrt(var+eps), shift = beta-mean*scale,
 * saves mean & invstd, updates running stats (unbiased var, momentum) and num_batches_tracked (+1) if given. */
int mnb_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, long long* num_batches_tracked, float* scale, float* shift,
                    float* save_mean, float* save_invstd, int C, double m, float eps, float momentum,
                    void* stream);
/* eval mode: scale/shift from running stats */
int mnb_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float* scale, float* shift, int C, float eps, void* stream);
/* Folded conv bias (conv_bias != NULL): the statistics were taken over z' = conv(x) WITHOUT its bias b (train-mode
 * BatchNorm cancels b exactly, so the conv kernels can skip the add).  scale/shift/save_mean refer to z';
 * running_mean tracks mean(z') + b as the reference's nn.Conv2d(bias=True) -> nn.BatchNorm2d would. */
int mnb_bn_finalize_fb(const double* stats, const float* gamma, const float* beta, const float* conv_bias,
                       float* running_mean, float* running_var, long long* num_batches_tracked, float* scale,
                       float* shift, float* save_mean, float* save_invstd, int C, double m, float eps,
                       float momentum, void* stream);
int mnb_bn_eval_coeffs_fb(const float* gamma, const float* beta, const float* conv_bias, const float* running_mean,
                          const float* running_var, float* scale, float* shift, int C, float eps, void* stream);
/* y = (residual ? residual : 0) + max(scale*z+shift, 0)     (BN-apply + ReLU + MBConv_block skip add) */
int mnb_bn_relu_apply(const void* z, const float* scale, const float* shift, const void* residual, void* y,
                      long long M, int C, int dtype, void* stream);
/* backward reductions: sums[0:C] += sum G, sums[C:2C] += sum G*z with G = dA*[scale*z+shift>0] */
int mnb_bn_bwd_reduce(const void* dA, const void* z, const float* scale, const float* shift, double* sums,
                      long long M, int C, int dtype, void* stream);
/* dgamma += , dbeta += , dbias += (analytically 0), coef[3C] = {a,b,c} with dZ = a*G + b*z + c */
int mnb_bn_bwd_finalize(const double* sums, const float* scale, const float* save_mean,
                        const float* save_invstd, float* dgamma, float* dbeta, float* dbias, float* coef, int C,
                        double m, void* stream);
/* bn_bwd_finalize + bn_bwd_apply in one launch: every CTA derives a,b,c for its channels from `sums`; dgamma /
 * dbeta / dbias (nullable) are accumulated once. */
int mnb_bn_bwd_apply_fused(const void* dA, const void* z, const float* scale, const float* shift, const double* sums,
                           const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta,
                           float* dbias, void* dZ, long long M, int C, double m, int dtype, void* stream);
/* dZ = a*G + b*z + c  (materialised) */
int mnb_bn_bwd_apply(const void* dA, const void* z, const float* scale, const float* shift, const float* coef,
                     void* dZ, long long M, int C, int dtype, void* stream);

/* ---- head : classifiers.py:107-111 -------------------------------------------------------------------- */
/* f[n,c] = mean_hw max(scale*z+shift,0)  (fp32 out)  : AdaptiveAvgPool2d(1) fused with the last BN+ReLU */
int mnb_gap_fwd(const void* z, const float* scale, const float* shift, float* f, int N, int HW, int C,
                int dtype, void* stream);
/* Global average pooling fused with the first Linear of the head (classifiers.py:107-111: pooling -> view -> Dropout ->
 * Linear): xb[n,c] = bf16( dropout( mean_hw max(scale*z+shift,0) ) ) is the A operand of the tcgen05 FC GEMM
 * (mnb_fc_fwd_tc), written straight from the pooling reduction; mask (N x C keep flags, NULL = no dropout) and
 * mask_scale = 1/(1-p) as in mnb_fc_prep_bf16.  f (fp32 pooled features) is optional (NULL = not wanted). */
int mnb_gap_fc_prep(const void* z, const float* scale, const float* shift, float* f, const unsigned char* mask,
                    float mask_scale, void* xb, int N, int HW, int C, int dtype, void* stream);
/* dA[n,hw,c] = df[n,c]/HW */
int mnb_gap_bwd(const float* df, void* dA, int N, int HW, int C, int dtype, void* stream);
/* Bernoulli keep-mask (1 = keep) with prob 1-p from a counter-based RNG (seed, offset) : nn.Dropout.
 * dev_step (nullable, device): element i uses counter offset + (*dev_step)*n + i, so a captured CUDA graph
 * draws fresh masks on every replay. */
int mnb_dropout_mask(unsigned char* mask, long long n, float p, unsigned long long seed,
                     unsigned long long offset, const long long* dev_step, void* stream);
/* y[n,o] = b[o] + sum_k (x[n,k]*mask[n,k]*mask_scale) * w[o,k] ; relu_out -> y = max(y,0)  : nn.Linear */
int mnb_fc_fwd(const float* x, const unsigned char* mask, float mask_scale, const float* w, const float* b,
               float* y, int relu_out, int N, int K, int O, void* stream);
/* dx[n,k] = (sum_o dy[n,o]*w[o,k]) * mask[n,k]*mask_scale * (relu_ref ? relu_ref[n,k]>0 : 1) */
int mnb_fc_dgrad(const float* dy, const float* w, const unsigned char* mask, float mask_scale,
                 const float* relu_ref, float* dx, int N, int K, int O, void* stream);
/* dw[o,k] += sum_n dy[n,o]*x[n,k]*mask*mask_scale ; db[o] += sum_n dy[n,o] */
int mnb_fc_wgrad(const float* x, const unsigned char* mask, float mask_scale, const float* dy, float* dw,
                 float* db, int N, int K, int O, void* stream);

/* ---- classifier on the tensor pipe (bf16 mode; nn.Linear, classifiers.py:56-89) -------------------------------
 * xb[i] = bf16( x[i] * (mask ? mask[i]*mask_scale : 1) )      operand prep (dropout folded in) */
int mnb_fc_prep_bf16(const float* x, const unsigned char* mask, float mask_scale, void* xb, long long n, void* stream);
/* y[n,o] = b[o] + sum_k xb[n,k]*w[o,k] (fp32 out, optional ReLU) as a tcgen05 GEMM; wpk = mnb_pack_weights(w, O, K, 1)
 * forward packing (bf16 [O][K]).  Requires K % 8 == 0. */
int mnb_fc_fwd_tc(const void* xb, const float* w, const void* wpk_fwd, const float* b, float* y, int relu_out, int N,
                  int K, int O, void* stream);
/* dx[n,k] = sum_o dyb[n,o]*w[o,k] (fp32 out, ungated) ; wpk_dgrad = bf16 [K][O].  Requires K % 8 == 0, O % 8 == 0. */
int mnb_fc_dgrad_tc(const void* dyb, const float* w, const void* wpk_dgrad, float* dx, int N, int K, int O,
                    void* stream);
/* dx[i] *= mask[i]*mask_scale * (relu_ref ? relu_ref[i] > 0 : 1)   (dropout / ReLU backward of the FC input) */
int mnb_fc_gate(float* dx, const unsigned char* mask, float mask_scale, const float* relu_ref, long long n,
                void* stream);
/* db[o] += sum_n dy[n,o] */
int mnb_fc_bias_grad(const float* dy, float* db, int N, int O, void* stream);

/* ---- loss : nn.CrossEntropyLoss(mean), train.py:277,435 --------------------------------------------------
 * loss[0] += mean_n( -log softmax(logits_n)[target_n] ) (caller zeroes loss);
 * dlogits (nullable) = (softmax - onehot) * grad_scale / N */
int mnb_xent_fwd_bwd(const float* logits, const long long* target, float* loss, float* dlogits, int N, int O,
                     float grad_scale, void* stream);

/* ---- optimizer : torch.optim.Adam defaults, train.py:219-221,440 (flat buffers) ------------------------ */
/* dev_lr / dev_step (nullable, device scalars) override lr / step so that a captured CUDA graph follows the
 * host-side LR schedule (train.py:282-302,334) and step count without re-capture. */
int mnb_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, int step, float grad_scale, const float* dev_lr, const long long* dev_step,
                  void* stream);
/* torch.optim.SGD(lr) (momentum 0) and torch.optim.RMSprop(lr) (alpha 0.99, eps 1e-8, not centered, momentum 0):
 * the other optimizers train.py:222-229 can select.  p -= lr * g  /  sq = alpha*sq + (1-alpha)*g*g; p -= lr*g/(sqrt(sq)+eps) */
int mnb_sgd_step(float* p, const float* g, long long n, float lr, float grad_scale, const float* dev_lr, void* stream);
int mnb_rmsprop_step(float* p, const float* g, float* square_avg, long long n, float lr, float alpha, float eps,
                     float grad_scale, const float* dev_lr, void* stream);
/* *counter += 1 on the device (step counter feeding mnb_adam_step / mnb_dropout_mask inside a graph) */
int mnb_counter_inc(long long* counter, void* stream);

/* ---- layout helpers --------------------------------------------------------------------------------- */
int mnb_nhwc_to_nchw_f32(const void* x, float* y, int N, int H, int W, int C, int dtype, void* stream);
/* input pipeline: y[n,c,h,w] = (x[n,h,w,c] / 255 - mean[c]) / std[c], uint8 HWC -> fp32 NCHW, i.e. transforms.ToTensor +
 * transforms.Normalize (utils/datasets.py:460-462) with FineTuneModelPool.mean / .std (classifiers.py:91-92) */
int mnb_u8hwc_to_nchw_f32(const unsigned char* x, const float* mean, const float* std, float* y, int N, int H, int W,
                          int C, void* stream);
int mnb_nchw_f32_to_nhwc(const float* x, void* y, int N, int H, int W, int C, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MNB200_H */
