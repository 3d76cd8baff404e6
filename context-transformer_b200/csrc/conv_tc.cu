// conv_tc.cu — tcgen05 implicit-GEMM convolution (placeholder until the kernel lands)
#include "common.cuh"
extern "C" int ctx_conv2d_tc_supported(const CtxConvParams*) { return 0; }
extern "C" int ctx_conv2d_tc_plan_create(const CtxConvParams*, void** plan_out) {
  if (plan_out) *plan_out = nullptr;
  ctx::set_error("ctx_conv2d_tc: not built");
  return CTX_ERR_UNSUPPORTED;
}
extern "C" int ctx_conv2d_tc_plan_run(void*, void*) { ctx::set_error("ctx_conv2d_tc: not built"); return CTX_ERR_UNSUPPORTED; }
extern "C" void ctx_conv2d_tc_plan_destroy(void*) {}
