// tcgen05 / TMEM dense-convolution GEMMs (bf16).  Round-1 placeholder: filled in below.
#include "conv_params.cuh"

namespace mnb {
int conv_fwd_tc(const ConvP&, cudaStream_t) { set_error("tcgen05 fwd: not built"); return MNB_ERR_UNSUPPORTED; }
int conv_dgrad_tc(const ConvP&, cudaStream_t) { set_error("tcgen05 dgrad: not built"); return MNB_ERR_UNSUPPORTED; }
int conv_wgrad_tc(const ConvP&, cudaStream_t) { set_error("tcgen05 wgrad: not built"); return MNB_ERR_UNSUPPORTED; }
}  // namespace mnb
