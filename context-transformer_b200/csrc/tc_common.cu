// tc_common.cu — host-side TMA descriptor encoding shared by the tensor-core kernels.
// cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint so the library does not link libcuda.
#include "tc_common.cuh"

#include <mutex>

namespace ctx {

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  });
  return fn;
}

int encode_2d_sw128(CUtensorMap* out, const void* base, bool bf16, unsigned long long rows, unsigned long long cols,
                    unsigned box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return CTX_ERR_CUDA; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {64u, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base),
                   gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return CTX_ERR_CUDA; }
  return CTX_OK;
}

// NHWC activation [N][H][W][C] 16-bit as a 4-D tensor (C innermost, pixels C elements apart, only channels < c_limit
// addressable); box = 64 channels x bw x bh pixels of one image, SWIZZLE_128B.  Out-of-range pixels (conv padding) and
// channels >= c_limit (the tail of a channel slice that is not a multiple of 64) read as zero.
int encode_nhwc_sw128(CUtensorMap* out, const void* base, bool bf16, int N, int H, int W, int C, int c_limit, unsigned bw, unsigned bh,
                      unsigned step) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return CTX_ERR_CUDA; }
  cuuint64_t gdim[4] = {(cuuint64_t)c_limit, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  // step > 1: every step-th pixel in x and y (a box dimension counts the elements walked over, ceil(box / step) of them are loaded)
  cuuint32_t box[4] = {64u, bw * step, bh * step, 1u};
  cuuint32_t estr[4] = {1, step, step, 1};
  if (box[1] > 256u || box[2] > 256u) { set_error("encode_nhwc_sw128: strided box %u x %u exceeds 256", box[1], box[2]); return CTX_ERR_UNSUPPORTED; }
  CUresult r = enc(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base),
                   gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (NHWC) failed (CUresult %d)", (int)r); return CTX_ERR_CUDA; }
  return CTX_OK;
}

}  // namespace ctx
