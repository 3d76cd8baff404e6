// Probe: 4-D fp32 TMA box loads from an NCHW image (no swizzle) — which box widths / coordinates / shared-memory alignments work.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_raw_probe tma_raw_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap m, int x0, int y0, int n, int bytes, uint32_t dst_off, float* out, int count) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bar = base, dst = base + 1024 + dst_off;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&m)), "r"(bar), "r"(x0), "r"(y0), "r"(0), "r"(n) : "memory");
  }
  uint32_t done = 0;
  for (int spin = 0; !done && spin < (1 << 20); ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(0) : "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < count; i += blockDim.x) out[i] = done ? reinterpret_cast<float*>(smem + 1024 + dst_off)[i] : -777.f;
}

int main() {
  const int N = 2, H = 40, W = 44;
  std::vector<float> h(N * 3 * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, 4096 * 4);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  Enc enc = (Enc)sym;
  const int bws[3] = {12, 16, 32};
  for (int bi = 0; bi < 3; ++bi) {
    const int bw = bws[bi], bh = 20;
    CUtensorMap m;
    cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)N};
    cuuint64_t gs[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, 3, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gdim, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box width %d: encode -> %d\n", bw, (int)r);
    if (r) continue;
    const int coords[3][2] = {{0, 0}, {-2, -2}, {6, 14}};
    const uint32_t offs[3] = {0, 3072, 128};
    for (int ci = 0; ci < 3; ++ci)
      for (int oi = 0; oi < 3; ++oi) {
        const int bytes = bw * bh * 3 * 4;
        cudaMemset(out, 0, 4096 * 4);
        probe<<<1, 128, 16384>>>(m, coords[ci][0], coords[ci][1], 1, bytes, offs[oi], out, bw * bh * 3);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> o(bw * bh * 3);
        if (e == cudaSuccess) cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        if (e == cudaSuccess)
          for (int c = 0; c < 3; ++c) for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
            const int gy = coords[ci][1] + y, gx = coords[ci][0] + x;
            const float want = (gy < 0 || gy >= H || gx < 0 || gx >= W) ? 0.f : h[((1 * 3 + c) * H + gy) * W + gx];
            bad += o[(c * bh + y) * bw + x] != want;
          }
        printf("  coords (%d,%d) smem offset %u: %s, %d mismatches\n", coords[ci][0], coords[ci][1], offs[oi], cudaGetErrorString(e), bad);
        if (e != cudaSuccess) { printf("  (context lost)\n"); return 0; }
      }
  }
  return 0;
}
