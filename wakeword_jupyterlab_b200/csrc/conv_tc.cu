// K3 (tensor-core path): placeholder until the tcgen05 kernels land.
#include "ctx.cuh"
size_t ww_conv_tc_act2_bytes_per_clip(const ww_ctx* c) { return 16; }
int ww_conv_tc_prepare(ww_ctx* c, cudaStream_t) { return WW_OK; }
int ww_launch_conv_tc(ww_ctx* c, const float*, int, cudaStream_t) {
  c->set_error("conv_tc: tcgen05 path not built in this revision");
  return WW_ERR_INVALID;
}
