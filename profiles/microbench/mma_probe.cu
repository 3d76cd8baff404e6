// mma_probe.cu — standalone sm_100a probe (development aid, not part of the library):
//   (1) tcgen05.mma issue cost and issue->commit-arrival latency for M128 x N x K16 bf16 MMAs, N = 64 / 128 / 256;
//   (2) steady-state throughput of {4 MMAs + commit} groups (the conv kernel's K-step);
//   (3) whether a SWIZZLE_128B K-major operand descriptor may START at any 128-byte row (not only at a 1024-byte swizzle
//       atom) and use a stride between 8-row groups that is not 1024 B: the "halo window" addressing of a 3x3 conv whose
//       activation patch is staged once and read under nine shifted descriptors.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_probe mma_probe.cu ; run: ./mma_probe
#include "../../context-transformer_b200/csrc/tc_common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <vector>

using namespace ctx;

namespace ctx { void set_error(const char*, ...) {} void count_launch(int) {} }

// busy poll (mbarrier.test_wait never suspends the thread): the probe measures arrival latency itself
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// ---- (1) + (2): timing -----------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) timing_kernel(long long* out, int N, int shift_rows, int group_rows) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 49152, bars = base + 49152 + 32768, slot = bars + 256;
  for (int i = threadIdx.x; i < (49152 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0u;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) { for (int b = 0; b < 32; ++b) mbar_init(bars + 8 * b, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(slot, 512); tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(true, 128, N);
    const uint64_t ad = desc_sw128(sA + shift_rows * 128, group_rows * 128), bd = desc_sw128(sB, 1024);
    int o = 0, bar = 0;
    uint32_t ph[32];
    for (int b = 0; b < 32; ++b) ph[b] = 0;
    // (1) n MMAs + one commit: issue clocks, clocks until the commit's mbarrier arrival is visible
    const int counts[8] = {1, 2, 4, 8, 16, 36, 72, 144};
    for (int rep = 0; rep < 2; ++rep)
      for (int c = 0; c < 8; ++c) {
        const long long t0 = clock64();
        for (int i = 0; i < counts[c]; ++i) umma_f16(tmem, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, i ? 1u : 0u);
        umma_commit(bars + 8 * bar);
        const long long t1 = clock64();
        mbar_spin(bars + 8 * bar, ph[bar]); ph[bar] ^= 1u;
        const long long t2 = clock64();
        if (rep == 1) { out[o++] = counts[c]; out[o++] = t1 - t0; out[o++] = t2 - t0; }
      }
    // (2) G groups of {n MMAs + commit} issued back to back on a ring of `ring` barriers (waiting for group g - ring before
    //     issuing g, like the conv kernel's smem ring): clocks per group in steady state
    for (int ring = 2; ring <= 8; ring *= 2) {
      const int G = 64;
      const long long t0 = clock64();
      for (int g = 0; g < G; ++g) {
        const int b = g % ring;
        if (g >= ring) { mbar_spin(bars + 8 * b, ph[b]); ph[b] ^= 1u; }
        for (int i = 0; i < 4; ++i) umma_f16(tmem, ad + 2 * i, bd + 2 * i, idesc, (g | i) ? 1u : 0u);
        umma_commit(bars + 8 * b);
      }
      for (int b = 0; b < ring; ++b) { mbar_spin(bars + 8 * b, ph[b]); ph[b] ^= 1u; }
      const long long t1 = clock64();
      out[o++] = -ring; out[o++] = G; out[o++] = t1 - t0;
    }
    // (2b) the same with n = 8, 12, 16, 36 MMAs per commit (ring of 8), and 36 MMAs followed by two commits
    const int per[5] = {8, 12, 16, 36, 36};
    for (int c = 0; c < 5; ++c) {
      const int G = 32, ring = 8;
      const long long t0 = clock64();
      for (int g = 0; g < G; ++g) {
        const int b = g % ring;
        if (g >= ring) { mbar_spin(bars + 8 * b, ph[b]); ph[b] ^= 1u; if (c == 4) { mbar_spin(bars + 8 * (b + 8), ph[b + 8]); ph[b + 8] ^= 1u; } }
        for (int i = 0; i < per[c]; ++i) umma_f16(tmem, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, (g | i) ? 1u : 0u);
        umma_commit(bars + 8 * b);
        if (c == 4) umma_commit(bars + 8 * (b + 8));
      }
      for (int b = 0; b < ring; ++b) { mbar_spin(bars + 8 * b, ph[b]); ph[b] ^= 1u; if (c == 4) { mbar_spin(bars + 8 * (b + 8), ph[b + 8]); ph[b + 8] ^= 1u; } }
      const long long t1 = clock64();
      out[o++] = -(1000 + per[c] + (c == 4 ? 100 : 0)); out[o++] = G; out[o++] = t1 - t0;
    }
    // (2c) variations of 32 x 36 MMAs: V1 one commit at the very end; V2 commit per group, no barrier polling inside the
    //      loop; V3 commit per group, groups alternate between two accumulators; V4 commit per group to an unused barrier
    for (int v = 1; v <= 4; ++v) {
      const int G = 32;
      const long long t0 = clock64();
      for (int g = 0; g < G; ++g) {
        const uint32_t d = tmem + ((v == 3 && (g & 1)) ? 256u : 0u);
        for (int i = 0; i < 36; ++i) umma_f16(d, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, (g | i) ? 1u : 0u);
        if (v == 2 || v == 3) umma_commit(bars + 8 * (16 + (g & 7)));
        if (v == 4) umma_commit(bars + 8 * 31);
      }
      umma_commit(bars + 8 * 30);
      mbar_spin(bars + 8 * 30, ph[30]); ph[30] ^= 1u;
      const long long t1 = clock64();
      out[o++] = -(2000 + v); out[o++] = G; out[o++] = t1 - t0;
    }
    // (2d) 288 groups of 4 MMAs: V5 + commit; V6 + commit + one mbarrier.test_wait on a long-completed barrier; V7 + commit +
    //      a volatile shared-memory load; V8 test_wait only (no commit)
    mbar_arrive(bars + 8 * 29);                                  // barrier 29: phase 0 complete from here on
    for (int v = 5; v <= 8; ++v) {
      const int G = 288;
      uint32_t sink = 0;
      const long long t0 = clock64();
      for (int g = 0; g < G; ++g) {
        for (int i = 0; i < 4; ++i) umma_f16(tmem, ad + 2 * i, bd + 2 * i, idesc, (g | i) ? 1u : 0u);
        if (v != 8) umma_commit(bars + 8 * (16 + (g & 7)));
        if (v == 6 || v == 8) {
          uint32_t done;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done) : "r"(bars + 8 * 29), "r"(0) : "memory");
          sink += done;
        }
        if (v == 7) { uint32_t x; asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(x) : "r"(bars + 8 * 29) : "memory"); sink += x; }
      }
      umma_commit(bars + 8 * 30);
      mbar_spin(bars + 8 * 30, ph[30]); ph[30] ^= 1u;
      const long long t1 = clock64();
      out[o++] = -(2000 + v); out[o++] = G; out[o++] = (t1 - t0) * 9 + (sink == 12345u);      // x9: printed per 36 MMAs
    }
    out[o++] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---- (3) shifted / strided descriptor windows ---------------------------------------------------------
// smem A: R rows of 64 bf16 (128 B) in the SWIZZLE_128B pattern of a 1024-B aligned buffer: A[r][k] = value(r, k).
// B = 64 x 64 identity.  D = A_window * B^T  ->  D[m][n] = A[row(m)][n], row(m) = shift + (m / 8) * group_rows + m % 8.
__global__ void __launch_bounds__(128, 1) window_kernel(float* out, int shift, int group_rows) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const int R = 320;
  const uint32_t sA = base, sB = base + R * 128, bars = sB + 64 * 128, slot = bars + 64;
  for (int i = threadIdx.x; i < R * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    const float v = (float)((r * 3 + k * 5) % 251) - 125.f;          // exactly representable in bf16 (|v| < 256, integer)
    const uint32_t off = r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1));
    *reinterpret_cast<__nv_bfloat16*>(gen + off) = __float2bfloat16_rn(v);
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    const uint32_t off = R * 128 + r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1));
    *reinterpret_cast<__nv_bfloat16*>(gen + off) = __float2bfloat16_rn(r == k ? 1.f : 0.f);
  }
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) { mbar_init(bars, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(slot, 64); tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(true, 128, 64);
    const uint64_t ad = desc_sw128(sA + shift * 128, group_rows * 128), bd = desc_sw128(sB, 1024);
    for (int k = 0; k < 4; ++k) umma_f16(tmem, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : 0u);
    umma_commit(bars);
    mbar_spin(bars, 0);
  }
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cb = 0; cb < 2; ++cb) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + cb * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  long long* d_out; float* d_f;
  cudaMalloc(&d_out, 4096 * sizeof(long long));
  cudaMalloc(&d_f, 128 * 64 * sizeof(float));
  cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  cudaFuncSetAttribute(window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int Ns[3] = {64, 128, 256};
  const int variants[6][2] = {{0, 8}, {1, 8}, {8, 8}, {0, 10}, {11, 10}, {0, 16}};      // {start row, rows between 8-row groups}
  for (int v = 0; v < 6; ++v)
  for (int n = 0; n < 3; ++n) {
    cudaMemset(d_out, 0, 4096 * sizeof(long long));
    timing_kernel<<<1, 128, 96 * 1024>>>(d_out, Ns[n], variants[v][0], variants[v][1]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("timing N=%d failed: %s\n", Ns[n], cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(4096);
    cudaMemcpy(h.data(), d_out, 4096 * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("== N = %d (floor %d clk / MMA), A window: start row %d, 8-row groups %d rows apart\n", Ns[n], Ns[n] / 2, variants[v][0], variants[v][1]);
    for (int o = 0; h[o] != 0; o += 3) {
      if (v > 0 && !(h[o] == 144 || h[o] == -8)) continue;
      if (v > 0 && h[o] <= -1000) continue;
      if (h[o] > 0) printf("  %3lld MMAs + commit: issue %5lld clk, issue->arrival %6lld clk (%.1f clk/MMA)\n", h[o], h[o + 1], h[o + 2], (double)h[o + 2] / h[o]);
      else if (h[o] <= -2000) printf("  variation V%lld: %lld groups: %.1f clk/MMA\n", -h[o] - 2000, h[o + 1], (double)h[o + 2] / (36.0 * h[o + 1]));
      else if (h[o] > -1000) printf("  ring of %lld: %lld groups of {4 MMA + commit}: %lld clk = %.1f clk/group\n", -h[o], h[o + 1], h[o + 2], (double)h[o + 2] / h[o + 1]);
      else printf("  ring of 8: %lld groups of {%lld MMA + %d commit}: %lld clk = %.1f clk/group\n", h[o + 1], (-h[o] - 1000) % 100, -h[o] - 1000 >= 100 ? 2 : 1, h[o + 2], (double)h[o + 2] / h[o + 1]);
    }
  }
  // (3)
  const int shifts[8] = {0, 1, 2, 3, 7, 8, 9, 11};
  const int groups[3] = {8, 10, 18};
  std::vector<float> hf(128 * 64);
  for (int gi = 0; gi < 3; ++gi)
    for (int si = 0; si < 8; ++si) {
      cudaMemset(d_f, 0, 128 * 64 * sizeof(float));
      window_kernel<<<1, 128, 64 * 1024>>>(d_f, shifts[si], groups[gi]);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("window shift %d group %d failed: %s\n", shifts[si], groups[gi], cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hf.data(), d_f, hf.size() * sizeof(float), cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m)
        for (int nn = 0; nn < 64; ++nn) {
          const int r = shifts[si] + (m / 8) * groups[gi] + m % 8;
          const float want = (float)((r * 3 + nn * 5) % 251) - 125.f;
          if (hf[m * 64 + nn] != want) { if (first < 0) first = m * 64 + nn; ++bad; }
        }
      printf("window: start row %2d, 8-row groups %2d rows apart: %s (%d mismatches%s)\n", shifts[si], groups[gi], bad ? "WRONG" : "exact", bad,
             bad ? "" : "");
      if (bad && first >= 0) printf("    first mismatch at m=%d n=%d: got %g\n", first / 64, first % 64, hf[first]);
    }
  return 0;
}
