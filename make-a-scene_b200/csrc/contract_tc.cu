// tcgen05 (TF32) contraction kernels — placeholder until the tensor path lands: every entry reports
// "unsupported" so that MAS_IMPL_AUTO falls through to the SIMT kernels.
#include "mas_common.cuh"
namespace mas {
int conv3x3_fprop_tc_launch(const float*, mas_tensor4, const float*, const float*, const float*, float*, mas_tensor4, int, cudaStream_t) {
  return fail(MAS_ERR_UNSUPPORTED, "tcgen05 conv path not built");
}
int gemm_tc_launch(const float*, const float*, float*, int, int, int, int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, int,
                   float, const float*, const float*, cudaStream_t) {
  return fail(MAS_ERR_UNSUPPORTED, "tcgen05 gemm path not built");
}
size_t conv_wgrad_tc_ws(mas_tensor4, mas_tensor4, int) { return 0; }
int conv_wgrad_tc_launch(const float*, mas_tensor4, const float*, mas_tensor4, float*, int, void*, size_t, cudaStream_t) {
  return fail(MAS_ERR_UNSUPPORTED, "tcgen05 wgrad path not built");
}
}  // namespace mas
