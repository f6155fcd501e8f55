// tcgen05 convolution path -- placeholder until the tensor-core kernels land: reports "unsupported"
// so the dispatcher routes every shape to the fp32 SIMT kernels.
#include "common.cuh"
#include "conv.h"
namespace sg2 {
bool conv_tc_supported(int, int, int, int, int, int) { return false; }
bool wgrad_tc_supported(int, int, int, int, int, int) { return false; }
int conv_fwd_tc(const ConvParams&, cudaStream_t) { return fail(SG2_ENOTSUP, "conv_fwd_tc: not built"); }
int conv_wgrad_tc(WgradParams, int, cudaStream_t) { return fail(SG2_ENOTSUP, "conv_wgrad_tc: not built"); }
int conv_pack_tc(const float*, void*, int, int, int, float, int, cudaStream_t) { return fail(SG2_ENOTSUP, "conv_pack_tc: not built"); }
long long conv_packed_bytes_tc(int co, int ci, int k) { return (long long)co * ci * k * k * 4; }
}  // namespace sg2
