// MUFU.EX2 throughput per SM (and with the FFMA + F2FP packing that the softmax loop adds).  nvcc -arch=sm_100a -O3 ex2_probe.cu
#include <cstdio>
#include <cuda_fp16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters, float a) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = a * (threadIdx.x + i);
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0) { v[i] = ex2(v[i]); v[i + 1] = ex2(v[i + 1]); }
      else {                                   // softmax inner loop: fma, ex2, pack to half2
        float p0 = ex2(fmaf(v[i], 1.4427f, -a)), p1 = ex2(fmaf(v[i + 1], 1.4427f, -a));
        __half2 h = __floats2half2_rn(p0, p1);
        acc ^= *reinterpret_cast<unsigned*>(&h);
        v[i] += 1e-3f; v[i + 1] -= 1e-3f;
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16, 32}) {
      const int iters = 4096;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 0.001f); else k<1><<<148, warps * 32>>>(out, iters, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double n = 148.0 * warps * 32 * iters * 16;
      printf("mode %d (%s) warps/SM %2d: %.2f ex2/clk/SM (at 1.965 GHz), %.3f ms\n", mode, mode ? "fma+ex2+pack" : "ex2 only", warps, n / (ms * 1e-3) / 148 / 1.965e9, ms);
    }
  return 0;
}
