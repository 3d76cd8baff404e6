// common.cuh — shared helpers for libctx_b200 (error plumbing, launch counting, small device utils)
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ctx_b200.h"

namespace ctx {

// thread-local last error message (never printed by the library)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define CTX_CUDA_TRY(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ctx::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CTX_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

#define CTX_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ctx::set_error(__VA_ARGS__);             \
      return CTX_ERR_INVALID;                  \
    }                                          \
  } while (0)

// check the launch that was just issued
#define CTX_LAUNCH_CHECK()                                                                   \
  do {                                                                                       \
    ctx::count_launch();                                                                     \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      ctx::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return CTX_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// Store one fp32 value into a tensor of run-time dtype.
__device__ __forceinline__ void store_as(void* base, long long idx, int dtype, float v) {
  if (dtype == CTX_F32) reinterpret_cast<float*>(base)[idx] = v;
  else if (dtype == CTX_BF16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
}
__device__ __forceinline__ float load_as(const void* base, long long idx, int dtype) {
  if (dtype == CTX_F32) return reinterpret_cast<const float*>(base)[idx];
  if (dtype == CTX_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
  return __half2float(reinterpret_cast<const __half*>(base)[idx]);
}
static inline int dtype_size(int dtype) { return dtype == CTX_F32 ? 4 : 2; }

// Device copy of the output-segment table of a conv (CtxOutSeg is plain data).
struct SegTable {
  int nseg;
  CtxOutSeg seg[3];
};

}  // namespace ctx
