// sort.cuh — device helpers shared by post.cu and train.cu: monotone float keys, a block-wide
// exclusive scan and a block bitonic sort of 64-bit keys (descending).
#pragma once
#include "common.cuh"

namespace ctx {

__device__ __forceinline__ uint32_t float_order_bits(float f) {   // monotone float -> uint32
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t idx) {
  // descending sort on this key == score descending, index ascending
  return ((uint64_t)float_order_bits(score) << 32) | (uint64_t)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ uint32_t key_index(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)k; }

// block-wide exclusive scan of one int per thread; returns exclusive prefix, *total = block sum.
// scratch: >= 33 ints of shared memory.  All threads of the block must call.
__device__ __forceinline__ int block_exclusive_scan_n(int v, int* scratch, int* total, int nwarp);
__device__ __forceinline__ int block_exclusive_scan(int v, int* scratch, int* total) {
  return block_exclusive_scan_n(v, scratch, total, (blockDim.x + 31) >> 5);
}
// same, for the first `nwarp` warps of a CTA whose other warps have exited
__device__ __forceinline__ int block_exclusive_scan_n(int v, int* scratch, int* total, int nwarp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();                       // scratch may still be read from a previous call
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarp ? scratch[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    scratch[lane] = wi - w;              // exclusive warp offsets
    if (lane == 31) scratch[32] = wi;
  }
  __syncthreads();
  *total = scratch[32];
  return scratch[warp] + incl - v;
}

// ------------------------------------------------------------------------------------------------
// Block bitonic sort (descending) of n 64-bit keys.  Rows of up to kSortSmem keys are sorted
// entirely in shared memory; longer rows run the wide compare-exchange passes in global memory
// (row capacity must be >= next_pow2(n)) and the narrow ones per shared-memory block.
// Returns the pointer (shared or global) that holds the sorted keys.
// ------------------------------------------------------------------------------------------------
constexpr int kSortSmem = 4096;      // keys kept in shared memory (32 KB)

__device__ __forceinline__ void cmpxchg_desc(uint64_t* a, int i, int j, bool desc) {
  uint64_t x = a[i], y = a[j];
  if ((x < y) == desc) { a[i] = y; a[j] = x; }
}

static __device__ const uint64_t* block_sort_desc(uint64_t* gkeys, int n, uint64_t* skeys) {
  const int tid = threadIdx.x, T = blockDim.x;
  int npad = 2;
  while (npad < n) npad <<= 1;
  if (npad <= kSortSmem) {
    for (int i = tid; i < npad; i += T) skeys[i] = i < n ? gkeys[i] : 0ull;
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < (npad >> 1); t += T) {
          int i = 2 * t - (t & (j - 1));
          cmpxchg_desc(skeys, i, i + j, (i & k) == 0);
        }
        __syncthreads();
      }
    return skeys;
  }
  for (int i = n + tid; i < npad; i += T) gkeys[i] = 0ull;
  __syncthreads();
  // phase 1: every kSortSmem-aligned block fully sorted (all k <= kSortSmem) in shared memory
  // phase 2: for k > kSortSmem: wide passes (j >= kSortSmem) in global, then j < kSortSmem per block
  for (int k = kSortSmem; k <= npad; k <<= 1) {
    if (k > kSortSmem) {
      for (int j = k >> 1; j >= kSortSmem; j >>= 1) {
        for (int t = tid; t < (npad >> 1); t += T) {
          int i = 2 * t - (t & (j - 1));
          cmpxchg_desc(gkeys, i, i + j, (i & k) == 0);
        }
        __syncthreads();
      }
    }
    for (int blk = 0; blk < npad; blk += kSortSmem) {
      for (int i = tid; i < kSortSmem; i += T) skeys[i] = gkeys[blk + i];
      __syncthreads();
      const int k_lo = (k == kSortSmem) ? 2 : k;
      for (int kk = k_lo; kk <= k; kk <<= 1) {
        const int j_hi = (kk == k && k > kSortSmem) ? (kSortSmem >> 1) : (kk >> 1);
        for (int j = j_hi; j > 0; j >>= 1) {
          for (int t = tid; t < (kSortSmem >> 1); t += T) {
            int i = 2 * t - (t & (j - 1));
            cmpxchg_desc(skeys, i, i + j, ((blk + i) & kk) == 0);
          }
          __syncthreads();
        }
      }
      for (int i = tid; i < kSortSmem; i += T) gkeys[blk + i] = skeys[i];
      __syncthreads();
    }
  }
  return gkeys;
}


static inline int next_pow2(int n) { int p = 2; while (p < n) p <<= 1; return p; }

}  // namespace ctx
