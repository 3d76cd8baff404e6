// attention_tc.cu — Context-Transformer block (models/RFB_Net_vgg.py:253-271) on the sm_100a tensor
// cores: theta projection, softmax(Q K^T) V over the spatially pooled keys, residual, L2-norm, cosine
// classifier and the class softmax as ONE warp-specialised kernel; neither Q nor the [B, P, Pk]
// affinity matrix (86 MB / image upstream) ever exists in HBM.
//
// Numerics.  The reference's logits are un-scaled dot products of raw class scores (|s| up to ~1.5e3 on
// the seeded weights), so a plain 16-bit Q K^T is not accurate enough.  Every fp32 operand X is split
// into fp16 hi + lo parts and X Y^T = Xh Yh^T + Xl Yh^T + Xh Yl^T is accumulated in fp32 by three
// tcgen05 MMAs (error ~2^-22 relative; measured final |dconf| 4e-5 against the fp32 reference).
//
//   proj_kv kernel (CUDA cores, fp32, tiny): K = phi(pool)+pool as fp16 hi/lo, V = g(pool)+pool as fp16 V^T.
//   attention kernel: CTA = 2 x 128 queries of one image (two "q-tiles" that share the key stream and
//   interleave on the tensor core), keys streamed in tiles of 128 through a 2/3-stage TMA ring.  320 threads:
//     warps 0-7  softmax (4 per q-tile, one query row per thread)
//                prologue: load the conf row, split hi/lo, stage it as an A operand; after the projection
//                MMA (Q = conf (theta+I)^T, 12 UMMAs) read Q from TMEM, add the bias, split hi/lo and stage
//                it as the A operand of the logits MMAs.
//                key loop: logits (4 UMMAs / tile, 12 in split mode), p = ex2(s*log2e - ref) against a reference that
//                only moves when a tile's maximum exceeds it by more than 2^tau (online softmax with LAZY rescale:
//                the first half tile sets it, ref_up above its maximum; a later jump rescales the row's O / l accumulator in TMEM and the half
//                tile of P already staged — a handful of times per row, so PV of tile j still overlaps the
//                exponentials of tile j+1 and no separate maximum pass over the keys exists); P is written as the
//                fp16 A operand of the PV MMA.
//                epilogue: O/l, z = conf + O*Wz, z/||z||, OBJ_Target*scale [, fc_base(conf)+conf], class softmax.
//     warp 8     TMA producer: theta weights once, then {Kh, V^T[, Kl]} tiles.
//     warp 9     TMEM allocator + MMA issuer (one elected lane); tcgen05.commit drives every hand-off mbarrier.
#include "tc_common.cuh"

#include <stdlib.h>

namespace ctx {

constexpr int AT_BQ = 128;        // queries per q-tile
constexpr int AT_QT = 2;          // q-tiles per CTA
constexpr int AT_BK = 128;        // keys per tile (UMMA N = 128: half the shared-memory operand traffic per MAC of N = 64)
constexpr int AT_DP = 64;         // padded feature dim
constexpr int AT_RING = 96 * 1024; // key/value ring bytes: 2 stages {Kh, Vt, Kl} (split) or 3 stages {Kh, Vt}
constexpr int AT_THREADS = 320;
constexpr int AT_TILE_Q = AT_BQ * AT_DP * 2;      // 16 KB
constexpr int AT_TILE_K = AT_BK * AT_DP * 2;      // 16 KB: 128 keys x 64 features
constexpr int AT_TILE_W = AT_DP * AT_DP * 2;      // 8 KB: 64 x 64 (theta' half, V^T key block)
constexpr int AT_STAGE_SPLIT = 3 * AT_TILE_K, AT_STAGE_FAST = 2 * AT_TILE_K;   // stage layout: Kh | Vt[2] | Kl
constexpr float AT_LOG2E = 1.4426950408889634f;
constexpr float AT_TAU = 12.0f;   // precise mode: a row's softmax reference follows the running maximum only in jumps of > 2^12 (the other modes: AttnTcParams::tau)

// ---- K / V projection (fp32 CUDA cores) -> fp16 operands --------------------------------------------
// rows: B*Pk_pad keys (rows >= Pk of an image are zero).  Khl: [2][B*Pk_pad][64]; Vt: [B][64][Pk_pad].
template <int D>
__global__ void __launch_bounds__(128)
proj_kv_kernel(const float* __restrict__ pooled, const float* __restrict__ phi_w, const float* __restrict__ phi_b,
               const float* __restrict__ g_w, const float* __restrict__ g_b, int B, int Pk, int Pk_pad,
               __half* __restrict__ khl, __half* __restrict__ vt, __half* __restrict__ vt_lo) {
  __shared__ __align__(16) float s_phi[D * D], s_g[D * D], s_pb[D], s_gb[D];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) { s_phi[i] = phi_w[i]; s_g[i] = g_w[i]; }
  for (int i = threadIdx.x; i < D; i += blockDim.x) { s_pb[i] = phi_b[i]; s_gb[i] = g_b[i]; }
  __syncthreads();
  const long long rows = (long long)B * Pk_pad;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int b = (int)(r / Pk_pad), j = (int)(r - (long long)b * Pk_pad);
  const bool valid = j < Pk;
  float xv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) xv[d] = valid ? pooled[((long long)b * Pk + j) * D + d] : 0.f;
  __half* hi = khl + r * AT_DP;
  __half* lo = khl + (rows + r) * AT_DP;
  for (int o0 = 0; o0 < AT_DP; o0 += 8) {
    __align__(16) __half h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int o = o0 + e;
      float k = 0.f, v = 0.f;
      if (o < D && valid) {
        float ak = 0.f, av = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) { ak = fmaf(s_phi[o * D + d], xv[d], ak); av = fmaf(s_g[o * D + d], xv[d], av); }
        k = (ak + s_pb[o]) + xv[o];
        v = (av + s_gb[o]) + xv[o];
      }
      h[e] = __float2half_rn(k);
      l[e] = __float2half_rn(k - __half2float(h[e]));
      if (o == D) v = valid ? 1.f : 0.f;                     // "ones" feature: the PV MMA then accumulates sum_j p_j in O[:, D]
      const __half vh = __float2half_rn(v);
      vt[((long long)b * AT_DP + o) * Pk_pad + j] = vh;
      if (vt_lo) vt_lo[((long long)b * AT_DP + o) * Pk_pad + j] = __float2half_rn(v - __half2float(vh));
    }
    *reinterpret_cast<uint4*>(hi + o0) = *reinterpret_cast<uint4*>(h);
    *reinterpret_cast<uint4*>(lo + o0) = *reinterpret_cast<uint4*>(l);
  }
}

// W' = W + I (the "+x" residual of theta(x) + x, phi(x) + x, g(x) + x), padded to 64 x 64, as fp16 hi/lo: wq[3][2][64][64] for
// theta, phi, g
__global__ void prep_wq_kernel(const float* __restrict__ theta_w, const float* __restrict__ phi_w, const float* __restrict__ g_w, int D,
                               __half* __restrict__ wq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * AT_DP * AT_DP) return;
  const int m = i / (AT_DP * AT_DP), e = i - m * (AT_DP * AT_DP);
  const int o = e / AT_DP, d = e - o * AT_DP;
  const float* src = m == 0 ? theta_w : (m == 1 ? phi_w : g_w);
  float w = 0.f;
  if (o < D && d < D) w = src[o * D + d] + (o == d ? 1.f : 0.f);
  const __half h = __float2half_rn(w);
  wq[(2 * m) * AT_DP * AT_DP + e] = h;
  wq[(2 * m + 1) * AT_DP * AT_DP + e] = __float2half_rn(w - __half2float(h));
}

// ---- fused attention --------------------------------------------------------------------------------
struct AttnTcParams {
  int B, P, Pk, Pk_pad, ntiles;
  int num_novel, incre, apply_softmax;
  int bulk_x;                       // conf blocks are 16-byte aligned / sized: staged with 1-D bulk copies
  int split;                        // 1: logits = Qh Kh^T + Ql Kh^T + Qh Kl^T (fp32-grade); 0: Qh Kh^T only (fp16-grade, 3x fewer MMAs)
  int precise;                      // 1 (implies split): 64-key tiles, P and V as fp16 hi/lo pairs too: O = Pl Vh + Ph Vl + Ph Vh (fp32-grade output)
  long long k_rows;                 // B*Pk_pad (row offset of the "lo" half of K)
  const float* conf;                // [B,P,D] fp32: projection input and residual x
  const float *theta_b, *Wz, *obj_w, *fc_w, *fc_b;
  float scale;
  float tau, ref_up;                // softmax reference window (log2 units), see the key loop
  float* out;
  long long* dbg;                   // optional timeline buffer (ctx_debug_set_buffer), CTA (0,0) only
};

#define AT_DBG(slot) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0) p.dbg[(slot)] = clock64(); } while (0)

// one row (64 fp32 values) -> fp16 hi and lo rows of two 128B-swizzled K-major operand tiles
template <bool WITH_LO = true>
__device__ __forceinline__ void store_split_row(uint32_t tile_hi, uint32_t tile_lo, int r, const float (&x)[64]) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = x[ch * 8 + 2 * e], b = x[ch * 8 + 2 * e + 1];
      const __half2 hh = __floats2half2_rn(a, b);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
      h[e] = *reinterpret_cast<const uint32_t*>(&hh);
      l[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const uint32_t off = r * 128 + ((ch ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_hi + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
    if (WITH_LO) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_lo + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
  }
}

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
  tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
  tmem_ld_wait();
}

// ---- K / V projection on the tensor cores -----------------------------------------------------------
// The same arithmetic as the Q projection inside the fused kernel: a 128-row tile of pooled conf rows, split into fp16 hi / lo
// A tiles, times phi' = phi + I and g' = g + I (hi / lo, prepared by prep_wq_kernel): x_hi W_hi + x_lo W_hi + x_hi W_lo in fp32
// (12 UMMAs per projection), bias added on the way out of TMEM.  K goes out as fp16 hi / lo rows, V transposed (+ its lo plane
// in the precise mode), rows past an image's Pk as zeros, feature D of V as the "ones" column.  One CTA per tile, 128 threads,
// thread = row = TMEM lane.  The CUDA-core kernel above needs 7 200 FMAs per row at 0.4 IPC: 61 us for 61 k rows, 11 % of the
// whole Context-Transformer op.
constexpr int PJ_TILES = 2;       // 128-row tiles per CTA (they share the four weight tiles): 240 CTAs of 256 threads, two per SM = one wave
template <int D>
__global__ void __launch_bounds__(128 * PJ_TILES)
proj_kv_tc_kernel(const __grid_constant__ CUtensorMap tm_w, const float* __restrict__ pooled, const float* __restrict__ phi_b,
                  const float* __restrict__ g_b, int B, int Pk, int Pk_pad, __half* __restrict__ khl, __half* __restrict__ vt,
                  __half* __restrict__ vt_lo, int vec_x) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // A hi | A lo per tile, then phi' hi | phi' lo | g' hi | g' lo
  const uint32_t sA = base, sW = sA + PJ_TILES * 2 * AT_TILE_Q, bars = sW + 4 * AT_TILE_W;
  const uint32_t w_full = bars, mma_done = bars + 8, tmem_slot = bars + 16;
  float* s_b = reinterpret_cast<float*>(smem_raw + (bars + 32 - smem_u32(smem_raw)));   // [2][64] biases, zero padded
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = warp >> 2, r = threadIdx.x & 127;                 // tile of the CTA, row of the tile = TMEM lane
  const long long rows = (long long)B * Pk_pad;
  const long long row = ((long long)blockIdx.x * PJ_TILES + t) * 128 + r;
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_w);
      mbar_init(w_full, 1); mbar_init(mma_done, 1);
      fence_barrier_init();
      mbar_arrive_expect_tx(w_full, 4 * AT_TILE_W);
      for (int k = 0; k < 4; ++k) tma_load_2d(sW + k * AT_TILE_W, &tm_w, 0, (2 + k) * AT_DP, w_full);   // rows 128 .. 383 of wq
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 128 * PJ_TILES);
    tmem_relinquish();
  }
  if (threadIdx.x < 64) { s_b[threadIdx.x] = threadIdx.x < D ? phi_b[threadIdx.x] : 0.f; s_b[64 + threadIdx.x] = threadIdx.x < D ? g_b[threadIdx.x] : 0.f; }
  const int b = row < rows ? (int)(row / Pk_pad) : 0, j = row < rows ? (int)(row - (long long)b * Pk_pad) : Pk;
  const bool valid = j < Pk;
  {
    float x[64];
    const float* src = pooled + ((long long)b * Pk + (valid ? j : 0)) * D;
    if (D % 4 == 0 && vec_x) {                                    // 16-byte row pieces: a quarter of the requests
#pragma unroll
      for (int d = 0; d < 64; d += 4) {
        const float4 q = (d < D && valid) ? __ldg(reinterpret_cast<const float4*>(src + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
        x[d] = q.x; x[d + 1] = q.y; x[d + 2] = q.z; x[d + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int d = 0; d < 64; ++d) x[d] = (d < D && valid) ? src[d] : 0.f;
    }
    store_split_row(sA + t * 2 * AT_TILE_Q, sA + (t * 2 + 1) * AT_TILE_Q, r, x);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  if (warp == 0) {
    mbar_wait(w_full, 0);
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(false, AT_BQ, AT_DP);
      const uint64_t w0 = make_sw128_desc(sW);
#pragma unroll
      for (int tt = 0; tt < PJ_TILES; ++tt) {
        const uint64_t xh = make_sw128_desc(sA + tt * 2 * AT_TILE_Q), xl = xh + (uint64_t)(AT_TILE_Q >> 4);
#pragma unroll
        for (int m = 0; m < 2; ++m) {                            // K = x phi'^T -> columns 0..63 of the tile's 128, V = x g'^T -> 64..127
          const uint64_t wh = w0 + (uint64_t)((2 * m * AT_TILE_W) >> 4), wl = wh + (uint64_t)(AT_TILE_W >> 4);
          const uint32_t d = tmem + tt * 128 + m * AT_DP;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, xh + 2 * k, wh + 2 * k, idesc, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, xl + 2 * k, wh + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, xh + 2 * k, wl + 2 * k, idesc, 1u);
        }
      }
      umma_commit(mma_done);
    }
    __syncwarp();
  }
  mbar_wait(mma_done, 0);
  tc_fence_after();
  const uint32_t trow = tmem + t * 128 + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[64];
  tmem_ld64(trow, v);
  if (row < rows) {
    __half* hi = khl + row * AT_DP;
    __half* lo = khl + (rows + row) * AT_DP;
#pragma unroll
    for (int o0 = 0; o0 < AT_DP; o0 += 8) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float k0 = valid ? __uint_as_float(v[o0 + 2 * e]) + s_b[o0 + 2 * e] : 0.f;
        const float k1 = valid ? __uint_as_float(v[o0 + 2 * e + 1]) + s_b[o0 + 2 * e + 1] : 0.f;
        const __half2 hh = __floats2half2_rn(k0, k1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(k0 - hf.x, k1 - hf.y);
        h[e] = *reinterpret_cast<const uint32_t*>(&hh);
        l[e] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      *reinterpret_cast<uint4*>(hi + o0) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(lo + o0) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
  tmem_ld64(trow + AT_DP, v);
  if (row < rows) {
#pragma unroll
    for (int o = 0; o < AT_DP; ++o) {
      float val = valid ? __uint_as_float(v[o]) + s_b[64 + o] : 0.f;
      if (o == D) val = valid ? 1.f : 0.f;                       // "ones" feature: the PV MMA then accumulates sum_j p_j in O[:, D]
      if (o > D) val = 0.f;
      const __half vh = __float2half_rn(val);
      vt[((long long)b * AT_DP + o) * Pk_pad + j] = vh;
      if (vt_lo) vt_lo[((long long)b * AT_DP + o) * Pk_pad + j] = __float2half_rn(val - __half2float(vh));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128 * PJ_TILES);
  }
}

// QT = 2: CTA = two q-tiles sharing the key stream, one CTA per SM (all modes).  QT = 1 (fp16-logit mode only): CTA = one q-tile with
// half of TMEM and < half of the shared memory, TWO CTAs per SM — the prologue (projection) and the per-row epilogue of one CTA, 27 %
// of its life with the tensor cores and the SFU idle, then run beside the key loop of the other.
template <int D, int NN, int QT>       // feature dim, novel classes (rows of OBJ_Target), q-tiles per CTA
__global__ void __launch_bounds__((4 * QT + 2) * 32, QT == 2 ? 1 : 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_vl, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // (QT = 1 has no room for alignment slack: the dynamic shared-memory window of a kernel without static shared memory starts 1 KB
  // into the CTA's allocation, which is 1024-aligned; checked below)
  const uint32_t base = QT == 2 ? (smem_u32(smem_raw) + 1023u) & ~1023u : smem_u32(smem_raw);
  if (QT == 1 && (base & 1023u)) __trap();
  constexpr int NQ = QT == 2 ? 4 : 1;                  // Q tiles: hi / lo per q-tile, or hi only
  constexpr int RING = QT == 2 ? AT_RING : 64 * 1024;
  constexpr int NSB = QT == 2 ? 3 : 1;                 // S buffers (128 TMEM columns each)
  constexpr int W_TMA = 4 * QT, W_MMA = 4 * QT + 1;
  // [QhA QlA QhB QlB] 64 KB (epilogue: classifier weights per q-tile)
  // [P_A P_B] 2 x 32 KB, each two 128x64 K-blocks (prologue: conf hi / lo tiles of the projection)
  // ring 96 KB: 2 x {Kh 16, Vt 2 x 8, Kl 16} or 3 x {Kh, Vt} (prologue: the last stage holds theta' hi/lo)
  const uint32_t sQ = base, sP = sQ + NQ * AT_TILE_Q, sKV = sP + 2 * QT * AT_TILE_Q;
  const uint32_t bars = sKV + RING;
  // precise mode: 64-key tiles, stage = {Kh 8, Kl 8, Vh^T 8, Vl^T 8} KB, three stages
  // (the one-q-tile kernel only exists for the fp16-logit mode: the other modes' code is compiled out of it)
  const bool precise = QT == 2 && p.precise != 0, split = QT == 2 && p.split != 0;
  const int KT = precise ? 64 : AT_BK;                 // keys per tile
  const int NST = QT == 1 ? 2 : (precise ? 3 : (split ? 2 : 3));     // ring depth
  const uint32_t STAGE = precise ? 4u * AT_TILE_W : (split ? AT_STAGE_SPLIT : AT_STAGE_FAST);
  const uint32_t w_full = bars, x_full = bars + 8, xq_done = bars + 24, q_ready = bars + 40, kv_full = bars + 56,
                 kv_empty = kv_full + 8 * 3, s_full = kv_empty + 8 * 3, s_empty = s_full + 32,
                 p_full = s_empty + 32, pv_done = p_full + 16, tmem_slot = pv_done + 16, xin_full = tmem_slot + 8;
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 256 - smem_u32(smem_raw)));   // [64] theta bias

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, q0 = blockIdx.x * (QT * AT_BQ);
  // staging area of a q-tile's fp32 conf rows (30 KB): its Q hi / lo tiles, or (QT = 1: Q is one tile) the first ring stage
  auto xstage = [&](int t) { return QT == 2 ? sQ + 2 * t * AT_TILE_Q : sKV; };
  const int T = p.ntiles;

  if (warp == W_TMA && lane == 0) { tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); if (precise) tma_prefetch_desc(&tm_vl); }
  if (warp == W_MMA) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int t = 0; t < QT; ++t) {
        mbar_init(x_full + 8 * t, 4); mbar_init(xq_done + 8 * t, 1); mbar_init(q_ready + 8 * t, 4);
        mbar_init(p_full + 8 * t, 4); mbar_init(pv_done + 8 * t, 1);
        for (int sb = 0; sb < 2; ++sb) { mbar_init(s_full + 8 * (2 * t + sb), 1); mbar_init(s_empty + 8 * (2 * t + sb), 4); }   // (three of the four are used)
        mbar_init(xin_full + 8 * t, 1);
      }
      for (int s = 0; s < 3; ++s) { mbar_init(kv_full + 8 * s, 1); mbar_init(kv_empty + 8 * s, 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, QT == 2 ? 512 : 256);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_bias[i] = i < D ? p.theta_b[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  // TMEM columns: three rotating S buffers (128 keys each) at 0 / 128 / 256 (the projection of q-tile t lands in buffer t);
  // O of q-tile t at 384 + 64 t
  const uint32_t tO = tmem + NSB * AT_BK;

  if (warp == W_TMA) {
    // ================= TMA producer =================
    const uint32_t wstage = sKV + (NST - 1) * STAGE;        // theta' lives in the last ring stage during the projection
    if (elect_one()) {
      for (int t = 0; t < QT; ++t) {                        // conf rows of q-tile t: one contiguous block -> its staging area
        const int rows = min(AT_BQ, p.P - (q0 + t * AT_BQ));
        const uint32_t bytes = rows > 0 ? (uint32_t)rows * D * 4u : 0u;
        mbar_arrive_expect_tx(xin_full + 8 * t, p.bulk_x ? bytes : 0u);
        if (p.bulk_x && bytes) bulk_load(xstage(t), p.conf + ((size_t)b * p.P + q0 + t * AT_BQ) * D, bytes, xin_full + 8 * t);
      }
      mbar_arrive_expect_tx(w_full, 2 * AT_TILE_W);
      tma_load_2d(wstage, &tm_w, 0, 0, w_full);                   // theta' hi
      tma_load_2d(wstage + AT_TILE_W, &tm_w, 0, AT_DP, w_full);   // theta' lo
    }
    __syncwarp();
    for (int j = 0, s = 0, ph = 1; j < T; ++j) {                // ring slot / EMPTY parity tracked incrementally (no division)
      if (j == NST - 1) { for (int t = 0; t < QT; ++t) mbar_wait(xq_done + 8 * t, 0); }     // theta' (last ring stage) has been consumed
      if (QT == 1 && j == 0) mbar_wait(x_full, 0);               // the conf rows staged in ring stage 0 have been read
      mbar_wait(kv_empty + 8 * s, ph);
      if (elect_one()) {
        const uint32_t dst = sKV + s * STAGE;
        mbar_arrive_expect_tx(kv_full + 8 * s, STAGE);
        if (precise) {                                        // (tm_k is encoded with 64-row boxes in this mode)
          tma_load_2d(dst, &tm_k, 0, b * p.Pk_pad + j * 64, kv_full + 8 * s);
          tma_load_2d(dst + AT_TILE_W, &tm_k, 0, (int)p.k_rows + b * p.Pk_pad + j * 64, kv_full + 8 * s);
          tma_load_2d(dst + 2 * AT_TILE_W, &tm_v, j * 64, b * AT_DP, kv_full + 8 * s);
          tma_load_2d(dst + 3 * AT_TILE_W, &tm_vl, j * 64, b * AT_DP, kv_full + 8 * s);
        } else {
          tma_load_2d(dst, &tm_k, 0, b * p.Pk_pad + j * AT_BK, kv_full + 8 * s);
          tma_load_2d(dst + AT_TILE_K, &tm_v, j * AT_BK, b * AT_DP, kv_full + 8 * s);
          tma_load_2d(dst + AT_TILE_K + AT_TILE_W, &tm_v, j * AT_BK + 64, b * AT_DP, kv_full + 8 * s);
          if (split) tma_load_2d(dst + 2 * AT_TILE_K, &tm_k, 0, (int)p.k_rows + b * p.Pk_pad + j * AT_BK, kv_full + 8 * s);
        }
      }
      __syncwarp();
      if (++s == NST) { s = 0; ph ^= 1; }
    }
  } else if (warp == W_MMA) {
    // ================= MMA issuer ================= (whole warp converged, one elected lane issues)
    const uint32_t idesc_s = make_idesc_f16(false, AT_BQ, KT);       // logits: M128 x N = keys per tile
    const uint32_t idesc_d = make_idesc_f16(false, AT_BQ, AT_DP);    // projection / PV: M128 x N64 features
    const uint64_t dQ = make_sw128_desc(sQ), dP = make_sw128_desc(sP), dKV = make_sw128_desc(sKV);
    // ---- projection: Q_t = conf_t (theta + I)^T, hi/lo split, into the first 64 columns of S_t
    mbar_wait(w_full, 0);
    for (int t = 0; t < QT; ++t) {
      mbar_wait(x_full + 8 * t, 0);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t xh = dP + (uint64_t)((t * 2 * AT_TILE_Q) >> 4), xl = xh + (uint64_t)(AT_TILE_Q >> 4);
        const uint64_t wh = dKV + (uint64_t)(((NST - 1) * STAGE) >> 4), wl = wh + (uint64_t)(AT_TILE_W >> 4);
        const uint32_t d = tmem + t * AT_BK;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(d, xh + 2 * k, wh + 2 * k, idesc_d, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(d, xl + 2 * k, wh + 2 * k, idesc_d, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(d, xh + 2 * k, wl + 2 * k, idesc_d, 1u);
        umma_commit(xq_done + 8 * t);
      }
      __syncwarp();
    }
    // ---- key loop over "uses" u = 2 j + t (key tile j, q-tile t): logits S_u (4 / 12 UMMAs) into S buffer u % 3, and — two uses
    // later, once the softmax warps have turned S_u into P — O_t += P V (8 UMMAs).  Three S buffers rotate between the two
    // q-tiles, so the logits of a q-tile's NEXT key tile are already in TMEM when its softmax warps finish the current one
    // (with one buffer per q-tile they idled for an MMA round trip per tile).
    int sq = 0, phq = 0, sp = 0;                                 // ring slot / FULL parity on the QK side; ring slot on the PV side
    int buf = 0, bpar = 0;                                       // S buffer u % 3 and parity of its (u / 3)-th use
    for (int u = 0; u < QT * T + QT; ++u) {
      if (u < QT * T) {
        const int t = u % QT, j = u / QT;
        if (t == 0) { mbar_wait(kv_full + 8 * sq, phq); if (lane == 0) AT_DBG(j * 16 + 0); }
        if (j == 0) mbar_wait(q_ready + 8 * t, 0);
        mbar_wait(s_empty + 8 * buf, bpar ^ 1);
        tc_fence_after();
        if (lane == 0) AT_DBG(j * 16 + 1 + t);
        if (elect_one()) {
          const uint64_t kh = dKV + (uint64_t)((sq * STAGE) >> 4), kl = kh + (uint64_t)((precise ? AT_TILE_W : 2 * AT_TILE_K) >> 4);
          const uint64_t qh = dQ + (uint64_t)(((QT == 2 ? 2 * t : 0) * AT_TILE_Q) >> 4), ql = qh + (uint64_t)(AT_TILE_Q >> 4);
          const uint32_t d = tmem + buf * AT_BK;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, qh + 2 * k, kh + 2 * k, idesc_s, k ? 1u : 0u);
          if (split) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d, ql + 2 * k, kh + 2 * k, idesc_s, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d, qh + 2 * k, kl + 2 * k, idesc_s, 1u);
          }
          umma_commit(s_full + 8 * buf);
        }
        __syncwarp();
        if (t == QT - 1 && ++sq == NST) { sq = 0; phq ^= 1; }
        if (++buf == NSB) { buf = 0; bpar ^= 1; }
      }
      if (u >= QT) {
        const int uu = u - QT, t = uu % QT, jj = uu / QT;
        if (lane == 0 && t == 0) AT_DBG(jj * 16 + 3);
        mbar_wait(p_full + 8 * t, jj & 1);
        tc_fence_after();
        if (lane == 0) AT_DBG(jj * 16 + 4 + t);
        if (elect_one()) {
          const uint64_t vt = dKV + (uint64_t)((sp * STAGE + (precise ? 2 * AT_TILE_W : AT_TILE_K)) >> 4), pp = dP + (uint64_t)((t * 2 * AT_TILE_Q) >> 4);
          if (precise) {                 // 64 keys: O += Pl Vh + Ph Vl + Ph Vh (small terms first)
            const uint64_t pl = pp + (uint64_t)(AT_TILE_Q >> 4), vl = vt + (uint64_t)(AT_TILE_W >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tO + t * AT_DP, pl + 2 * k, vt + 2 * k, idesc_d, (jj | k) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tO + t * AT_DP, pp + 2 * k, vl + 2 * k, idesc_d, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tO + t * AT_DP, pp + 2 * k, vt + 2 * k, idesc_d, 1u);
          } else
#pragma unroll
          for (int k = 0; k < 8; ++k)      // keys 0-63: P / Vt block 0, keys 64-127: block 1
            umma_f16(tO + t * AT_DP, pp + (uint64_t)((k >> 2) * (AT_TILE_Q >> 4)) + 2 * (k & 3),
                     vt + (uint64_t)((k >> 2) * (AT_TILE_W >> 4)) + 2 * (k & 3), idesc_d, (jj | k) ? 1u : 0u);
          umma_commit(pv_done + 8 * t);
          if (t == QT - 1) umma_commit(kv_empty + 8 * sp);
        }
        __syncwarp();
        if (lane == 0 && t == 1) AT_DBG(jj * 16 + 6);
        if (t == QT - 1 && ++sp == NST) sp = 0;
      }
    }
  } else {
    // ================= softmax warps (4 per q-tile) + epilogue =================
    const int t = warp >> 2, qd = warp & 3;
    const int r = qd * 32 + lane;                        // row inside the q-tile == TMEM lane
    const int q = q0 + t * AT_BQ + r;                    // query (prior) index inside the image
    const bool q_ok = q < p.P;
    const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
    const uint32_t tS = tmem + lane_sel + t * AT_BK;
    const uint32_t sPt = sP + t * 2 * AT_TILE_Q;
    const float* xrow = p.conf + ((size_t)b * p.P + (q_ok ? q : 0)) * D;
    uint32_t v[64];
    if (warp == 0 && lane == 0) AT_DBG(500);

    // ---- prologue: stage the conf row (hi / lo -> the two K-blocks of P_t), then Q hi/lo from the projection
    {
      float x[64];
      mbar_wait(xin_full + 8 * t, 0);
      const float* xs = p.bulk_x ? reinterpret_cast<const float*>(smem_raw + (xstage(t) - smem_u32(smem_raw))) + r * D : xrow;
#pragma unroll
      for (int d = 0; d < 64; ++d) x[d] = (d < D && q_ok) ? xs[d] : 0.f;
      if (warp == 0 && lane == 0) AT_DBG(501);
      store_split_row(sPt, sPt + AT_TILE_Q, r, x);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_full + 8 * t);
      if (warp == 0 && lane == 0) AT_DBG(502);
      mbar_wait(xq_done + 8 * t, 0);
      tc_fence_after();
      if (warp == 0 && lane == 0) AT_DBG(503);
      tmem_ld64(tS, v);
#pragma unroll
      for (int d = 0; d < 64; ++d) x[d] = __uint_as_float(v[d]) + s_bias[d];
      if (QT == 2) store_split_row(sQ + (2 * t) * AT_TILE_Q, sQ + (2 * t + 1) * AT_TILE_Q, r, x);
      else store_split_row<false>(sQ, 0u, r, x);                 // fp16-logit mode: the lo half of Q is never read
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready + 8 * t);
    }

    if (warp == 0 && lane == 0) AT_DBG(504);
    // ---- key loop: p = 2^(s log2e - ref2); ref2 (per row, log2 units) trails the running maximum in jumps of > AT_TAU
    float ref2 = 0.f;
    const uint32_t tOrow = tO + lane_sel + t * AT_DP;
    for (int j = 0; j < T; ++j) {
      if (warp == 0 && lane == 0) AT_DBG(j * 16 + 7);
      const int u = QT * j + t, ub = u % NSB;                  // this q-tile's use of S buffer u % NSB (see the MMA issuer)
      const uint32_t tSu = tmem + lane_sel + ub * AT_BK;
      mbar_wait(s_full + 8 * ub, (u / NSB) & 1);
      tc_fence_after();
      if (warp == 0 && lane == 0) AT_DBG(j * 16 + 8);
      const int NH = precise ? 1 : 2;                          // 64-key halves per tile
      for (int hb = 0; hb < NH; ++hb) {
        tmem_ld64(tSu + hb * 64, v);
        if (warp == 0 && lane == 0) AT_DBG(j * 16 + 9 + 3 * hb);
        if (hb == NH - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty + 8 * ub);
        }
        const int nvalid = p.Pk - j * KT - hb * 64;
        if (nvalid < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= nvalid) v[c] = 0xff800000u;               // -inf -> ex2 gives exactly 0
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // independent chains for ILP
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
          m4[0] = fmaxf(m4[0], fmaxf(__uint_as_float(v[c]), __uint_as_float(v[c + 1])));
          m4[1] = fmaxf(m4[1], fmaxf(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])));
          m4[2] = fmaxf(m4[2], fmaxf(__uint_as_float(v[c + 4]), __uint_as_float(v[c + 5])));
          m4[3] = fmaxf(m4[3], fmaxf(__uint_as_float(v[c + 6]), __uint_as_float(v[c + 7])));
        }
        const float m2 = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * AT_LOG2E;
        // PV_{j-1} must have finished reading P_t before this tile's P is stored, and writing O_t before O_t is rescaled: the wait
        // sits in front of those two places, not here — the 64 exponentials of the first half run while PV_{j-1} completes
        bool pv_ok = j == 0 || hb > 0;
        auto need_pv = [&] { if (!pv_ok) { mbar_wait(pv_done + 8 * t, (j - 1) & 1); tc_fence_after(); pv_ok = true; } };
        if (j == 0 && hb == 0) {
          ref2 = m2 + p.ref_up;                                // first half tile: the reference starts ref_up above its maximum
        } else {
          const bool jump = m2 > ref2 + p.tau;
          if (__any_sync(0xffffffffu, jump)) {
            // rare: move the reference of the rows that jumped and rescale what they have accumulated under the old one
            const float nref = jump ? m2 + p.ref_up : ref2;
            const float f = fast_exp2(ref2 - nref);            // 1 for rows that stay
            ref2 = nref;
            need_pv();
            if (j > 0) {                                       // O_t / l (TMEM): all MMAs that wrote it are complete
              uint32_t o[32];
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                tmem_ld32(tOrow + hh * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
                tmem_st32(tOrow + hh * 32, o);
              }
              tmem_st_wait();
            }
            if (hb == 1) {                                     // the first half of this tile's P is already staged: scale it in place
              // (in fp32: f < 2^-tau is a subnormal with a few significant bits — or zero — as an fp16 factor, and the staged
              // values it scales can weigh as much as the new maximum)
              const uint32_t prow0 = sPt + r * 128;
#pragma unroll
              for (int ch = 0; ch < 8; ++ch) {
                uint32_t w[4];
                const uint32_t a = prow0 + ((ch ^ (r & 7)) << 4);
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(a) : "memory");
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 x2 = __half22float2(*reinterpret_cast<__half2*>(&w[k]));
                  __half2 h = __floats2half2_rn(x2.x * f, x2.y * f);
                  w[k] = *reinterpret_cast<uint32_t*>(&h);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
              }
            }
          }
        }
        if (precise) {                     // P as fp16 hi / lo pairs: the two K-blocks of P_t hold Ph and Pl of the same 64 keys
#pragma unroll
          for (int c = 0; c < 64; ++c) v[c] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(v[c]), AT_LOG2E, -ref2)));
          need_pv();
          store_split_row(sPt, sPt + AT_TILE_Q, r, *reinterpret_cast<const float(*)[64]>(&v[0]));
          continue;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) {       // (ex2.approx.f16x2 is no faster: it issues one MUFU.EX2.F16 per half — checked in SASS)
          const float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * e]), AT_LOG2E, -ref2));
          const float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * e + 1]), AT_LOG2E, -ref2));
          __half2 hh = __floats2half2_rn(p0, p1);
          v[e] = *reinterpret_cast<uint32_t*>(&hh);
        }
        if (warp == 0 && lane == 0) AT_DBG(j * 16 + 10 + 3 * hb);
        if (warp == 0 && lane == 0) AT_DBG(j * 16 + 11 + 3 * hb);
        need_pv();
        const uint32_t prow = sPt + hb * AT_TILE_Q + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((ch ^ (r & 7)) << 4)), "r"(v[4 * ch]),
                       "r"(v[4 * ch + 1]), "r"(v[4 * ch + 2]), "r"(v[4 * ch + 3]) : "memory");
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + 8 * t);
    }
    if (warp == 0 && lane == 0) AT_DBG(506);

    // ---- epilogue: z = x + (O/l)*Wz ; z/||z|| ; OBJ_Target*scale ; [fc_base] ; [softmax] ----
    // classifier weights: fetched into registers while the last PV MMAs still run, stored once P_t is free
    constexpr int W_PER_THREAD = (NN * D + D + AT_BQ - 1) / AT_BQ;
    float wreg[W_PER_THREAD];
#pragma unroll
    for (int k = 0; k < W_PER_THREAD; ++k) {
      const int i = r + k * AT_BQ;
      wreg[k] = i < NN * D ? p.obj_w[i] : (i < NN * D + D ? p.Wz[i - NN * D] : 0.f);
    }
    mbar_wait(pv_done + 8 * t, (T - 1) & 1);                 // all MMAs of this q-tile are complete: Q_t is dead
    tc_fence_after();
    // Q_t and P_t are free now.  Q_t <- the conf rows again (bulk copy), P_t <- classifier weights + output staging.
    if (p.bulk_x && r == 0) {
      const int rows = min(AT_BQ, p.P - (q0 + t * AT_BQ));
      const uint32_t bytes = rows > 0 ? (uint32_t)rows * D * 4u : 0u;
      mbar_arrive_expect_tx(xin_full + 8 * t, bytes);
      if (bytes) bulk_load(xstage(t), p.conf + ((size_t)b * p.P + q0 + t * AT_BQ) * D, bytes, xin_full + 8 * t);
    }
    float* s_obj = reinterpret_cast<float*>(smem_raw + (sPt - smem_u32(smem_raw)));      // [num_novel][D]
    float* s_wz = s_obj + NN * D;                                                // [D]
    float* s_fc = s_wz + D;                                                               // incre: [D][D] + [D]
    const int n_out = NN + (p.incre ? D : 0);
    float* s_out = s_fc + (p.incre ? D * D + D : 0);                                      // [128][n_out]
#pragma unroll
    for (int k = 0; k < W_PER_THREAD; ++k) {                 // s_wz directly follows s_obj
      const int i = r + k * AT_BQ;
      if (i < NN * D + D) s_obj[i] = wreg[k];
    }
    if (p.incre) {
      for (int i = r; i < D * D; i += AT_BQ) s_fc[i] = p.fc_w[i];
      for (int i = r; i < D; i += AT_BQ) s_fc[D * D + i] = p.fc_b[i];
    }
    if (warp == 0 && lane == 0) AT_DBG(507);
    asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");
    tmem_ld64(tO + lane_sel + t * AT_DP, v);
    if (p.bulk_x) mbar_wait(xin_full + 8 * t, 1);
    if (warp == 0 && lane == 0) AT_DBG(508);
    if (q_ok) {
      const float* xs = p.bulk_x ? reinterpret_cast<const float*>(smem_raw + (xstage(t) - smem_u32(smem_raw))) + r * D : xrow;
      static_assert(D < AT_DP, "the row sum lives in the padding feature column D");
      const float inv_l = 1.0f / __uint_as_float(v[D]);        // sum_j p_j, accumulated by the PV MMA (ones feature of V)
      float x[D], z[D];
      float nrm = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        x[d] = xs[d];
        z[d] = x[d] + (__uint_as_float(v[d]) * inv_l) * s_wz[d];
        nrm = fmaf(z[d], z[d], nrm);
      }
      if (warp == 0 && lane == 0) AT_DBG(510);
      const float inv_n = 1.0f / sqrtf(nrm);
#pragma unroll
      for (int d = 0; d < D; ++d) z[d] *= inv_n;
      float* orow_s = s_out + r * n_out;
      if (p.incre) {
        for (int c = 0; c < D; ++c) {
          float a = 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) a = fmaf(s_fc[c * D + d], x[d], a);
          orow_s[c] = (a + s_fc[D * D + c]) + x[c];
        }
      }
      // cosine classifier (OBJ_Target rows are unit vectors).  Feature chunks outermost, classes innermost: NN independent
      // accumulation chains are in flight and each 128-bit weight load feeds four FMAs (per class the sum still runs over
      // d in ascending order).  With classes outermost this loop was one D-deep dependent chain after another and the
      // whole CTA waited on it at its end.
      float nov[NN];
#pragma unroll
      for (int c = 0; c < NN; ++c) nov[c] = 0.f;
      static_assert(D % 4 == 0 || D == 15, "classifier loop assumes 16-byte rows (D % 4 == 0) or the scalar path");
      if (D % 4 == 0 && NN % 4 == 0) {
        // four classes per trip of a ROLLED loop (the raw sums pass through the row's output staging): 0.3 k instructions of
        // code instead of 1.5 k of straight line that every warp fetches once per CTA — the unrolled form spent a quarter of
        // the epilogue's samples in stall_no_inst (0.535 -> 0.523 ms for the kernel, same box).  Per class the sum still runs
        // over d in ascending order: same bits.
        float* stage = orow_s + (p.incre ? D : 0);
#pragma unroll 1
        for (int g = 0; g < NN; g += 4) {
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          const float* wg = s_obj + g * D;
#pragma unroll
          for (int d = 0; d < D; d += 4) {
            const float4 w0 = *reinterpret_cast<const float4*>(wg + d), w1 = *reinterpret_cast<const float4*>(wg + D + d);
            const float4 w2 = *reinterpret_cast<const float4*>(wg + 2 * D + d), w3 = *reinterpret_cast<const float4*>(wg + 3 * D + d);
            a0 = fmaf(w0.x, z[d], a0); a0 = fmaf(w0.y, z[d + 1], a0); a0 = fmaf(w0.z, z[d + 2], a0); a0 = fmaf(w0.w, z[d + 3], a0);
            a1 = fmaf(w1.x, z[d], a1); a1 = fmaf(w1.y, z[d + 1], a1); a1 = fmaf(w1.z, z[d + 2], a1); a1 = fmaf(w1.w, z[d + 3], a1);
            a2 = fmaf(w2.x, z[d], a2); a2 = fmaf(w2.y, z[d + 1], a2); a2 = fmaf(w2.z, z[d + 2], a2); a2 = fmaf(w2.w, z[d + 3], a2);
            a3 = fmaf(w3.x, z[d], a3); a3 = fmaf(w3.y, z[d + 1], a3); a3 = fmaf(w3.z, z[d + 2], a3); a3 = fmaf(w3.w, z[d + 3], a3);
          }
          stage[g] = a0; stage[g + 1] = a1; stage[g + 2] = a2; stage[g + 3] = a3;
        }
#pragma unroll
        for (int c = 0; c < NN; ++c) nov[c] = stage[c];
      } else if (D % 4 == 0) {
#pragma unroll
        for (int d = 0; d < D; d += 4) {
#pragma unroll
          for (int c = 0; c < NN; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(s_obj + c * D + d);
            nov[c] = fmaf(w.x, z[d], nov[c]); nov[c] = fmaf(w.y, z[d + 1], nov[c]);
            nov[c] = fmaf(w.z, z[d + 2], nov[c]); nov[c] = fmaf(w.w, z[d + 3], nov[c]);
          }
        }
      } else {
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
          for (int c = 0; c < NN; ++c) nov[c] = fmaf(s_obj[c * D + d], z[d], nov[c]);
      }
      if (warp == 0 && lane == 0) AT_DBG(511);
#pragma unroll
      for (int c = 0; c < NN; ++c) nov[c] *= p.scale;
      if (p.apply_softmax) {
        float mxo = -INFINITY;
#pragma unroll
        for (int c = 0; c < NN; ++c) mxo = fmaxf(mxo, nov[c]);
        if (p.incre) for (int c = 0; c < D; ++c) mxo = fmaxf(mxo, orow_s[c]);
        float sm = 0.f;
#pragma unroll
        for (int c = 0; c < NN; ++c) { nov[c] = __expf(nov[c] - mxo); sm += nov[c]; }
        if (p.incre) for (int c = 0; c < D; ++c) { const float e = __expf(orow_s[c] - mxo); orow_s[c] = e; sm += e; }
        const float inv = 1.0f / sm;
#pragma unroll
        for (int c = 0; c < NN; ++c) nov[c] *= inv;
        if (p.incre) for (int c = 0; c < D; ++c) orow_s[c] *= inv;
      }
      const int off = p.incre ? D : 0;
#pragma unroll
      for (int c = 0; c < NN; ++c) orow_s[off + c] = nov[c];
    }
    if (warp == 0 && lane == 0) AT_DBG(512);
    asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");
    if (warp == 0 && lane == 0) AT_DBG(513);
    {   // the q-tile's output rows are one contiguous block: coalesced copy
      const int rows = min(AT_BQ, p.P - (q0 + t * AT_BQ));
      float* oblk = p.out + ((size_t)b * p.P + q0 + t * AT_BQ) * n_out;
      for (int i = r; i < rows * n_out; i += AT_BQ) oblk[i] = s_out[i];
    }
    if (warp == 0 && lane == 0) AT_DBG(509);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem, QT == 2 ? 512 : 256);
  }
}

// The one-q-tile kernel has no room for alignment slack in its 112.6 KB (two CTAs must fit an SM): it relies on the dynamic
// shared-memory window of a kernel without static shared memory starting on a 1 KB boundary (it does: the first KB of a CTA's
// allocation is reserved on sm_90+).  Checked once per process with a probe kernel instead of trusted; if it ever did not hold,
// or the probe cannot run (no device, or the first call arrives inside a stream capture), the two-q-tile kernel is used.
__global__ void smem_base_probe_kernel(unsigned* out) {
  extern __shared__ uint8_t probe_smem[];
  *out = smem_u32(probe_smem);
}
static int qt1_smem_base_ok(cudaStream_t st) {
  static int ok = -1;
  if (ok >= 0) return ok;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return 0; }   // undecided: ask again later
  unsigned* d = nullptr;
  unsigned h = 1u;
  if (cudaMalloc(&d, sizeof(unsigned)) == cudaSuccess) {
    smem_base_probe_kernel<<<1, 32, 4096, st>>>(d);
    if (cudaMemcpyAsync(&h, d, sizeof(unsigned), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) h = 1u;
    cudaFree(d);
  }
  cudaGetLastError();
  ok = (h & 1023u) == 0u ? 1 : 0;
  return ok;
}

static size_t attn_tc_smem(int D, int num_novel, int incre) {
  (void)D; (void)num_novel; (void)incre;      // classifier weights alias the Q region in the epilogue (<= 32 KB per q-tile)
  return 1024 + 4 * AT_TILE_Q + 4 * AT_TILE_Q + AT_RING + 256 + 256 + 64;
}

size_t attention_tc_workspace_bytes(int B, int P, int Pk) {
  const size_t Pk_pad = (size_t)(Pk + AT_BK - 1) / AT_BK * AT_BK;
  (void)P;                                  // K hi/lo, V^T hi, V^T lo (precise mode), theta' / phi' / g' hi/lo
  return align_up((size_t)2 * B * Pk_pad * AT_DP * 2, 1024) + 2 * align_up((size_t)B * AT_DP * Pk_pad * 2, 1024) +
         align_up((size_t)6 * AT_DP * AT_DP * 2, 1024);
}

static long long* g_attn_dbg = nullptr;
void attention_set_debug_buffer(void* p) { g_attn_dbg = (long long*)p; }

template <int D, int NN>
static int attention_tc_launch_t(const CtxAttnParams* a, cudaStream_t st) {
  const int B = a->batch, P = a->num_priors, Pk = a->num_pooled;
  const int Pk_pad = (Pk + AT_BK - 1) / AT_BK * AT_BK;
  const size_t need = attention_tc_workspace_bytes(B, P, Pk);
  if (!a->workspace || a->workspace_bytes < need) {
    set_error("attention (tensor-core path): workspace %zu < required %zu", a->workspace_bytes, need);
    return CTX_ERR_WORKSPACE;
  }
  CTX_REQUIRE(((uintptr_t)a->workspace) % 1024 == 0, "attention: workspace must be 1024-byte aligned");
  CTX_REQUIRE((long long)2 * B * Pk_pad < (1ll << 31), "attention: too many rows");
  char* ws = (char*)a->workspace;
  __half* khl = (__half*)ws;
  __half* vt = (__half*)(ws + align_up((size_t)2 * B * Pk_pad * AT_DP * 2, 1024));
  __half* vt_lo = (__half*)((char*)vt + align_up((size_t)B * AT_DP * Pk_pad * 2, 1024));
  __half* wq = (__half*)((char*)vt_lo + align_up((size_t)B * AT_DP * Pk_pad * 2, 1024));
  const bool precise = a->use_tensor_cores == 3;
  const int KT = precise ? 64 : AT_BK;
  const long long k_rows = (long long)B * Pk_pad;
  prep_wq_kernel<<<cdiv(3 * AT_DP * AT_DP, 256), 256, 0, st>>>(a->theta_w, a->phi_w, a->g_w, D, wq);
  CTX_LAUNCH_CHECK();
  CUtensorMap tw, tk, tv, tvl;
  int rc = encode_2d_sw128(&tw, wq, false, 6ull * AT_DP, AT_DP, AT_DP);
  if (rc) return rc;
  static const int proj_tc = [] { const char* e = getenv("CTX_ATTN_PROJ_TC"); return (e && e[0] == '0') ? 0 : 1; }();
  if (proj_tc) {
    const size_t psmem = 1024 + PJ_TILES * 2 * AT_TILE_Q + 4 * AT_TILE_W + 32 + 2 * 64 * sizeof(float);
    static const int vec_env = [] { const char* e = getenv("CTX_ATTN_PROJ_VECX"); return (e && e[0] == '0') ? 0 : 1; }();
    const int vec_x = vec_env && ((uintptr_t)a->pooled % 16 == 0) && ((size_t)D * 4 % 16 == 0);
    CTX_CUDA_TRY(cudaFuncSetAttribute(proj_kv_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    proj_kv_tc_kernel<D><<<cdiv(k_rows, 128 * PJ_TILES), 128 * PJ_TILES, psmem, st>>>(tw, a->pooled, a->phi_b, a->g_b, B, Pk, Pk_pad, khl, vt,
                                                                                   precise ? vt_lo : nullptr, vec_x);
  } else {
    proj_kv_kernel<D><<<cdiv(k_rows, 128), 128, 0, st>>>(a->pooled, a->phi_w, a->phi_b, a->g_w, a->g_b, B, Pk, Pk_pad, khl, vt, precise ? vt_lo : nullptr);
  }
  CTX_LAUNCH_CHECK();
  if (!rc) rc = encode_2d_sw128(&tk, khl, false, 2ull * k_rows, AT_DP, (unsigned)KT);
  if (!rc) rc = encode_2d_sw128(&tv, vt, false, (unsigned long long)B * AT_DP, (unsigned long long)Pk_pad, AT_DP);   // box: 64 keys x 64 features
  if (!rc) rc = encode_2d_sw128(&tvl, precise ? vt_lo : vt, false, (unsigned long long)B * AT_DP, (unsigned long long)Pk_pad, AT_DP);
  if (rc) return rc;
  AttnTcParams p;
  p.B = B; p.P = P; p.Pk = Pk; p.Pk_pad = Pk_pad; p.ntiles = Pk_pad / KT;
  p.precise = precise;
  p.num_novel = a->num_novel; p.incre = a->incre; p.apply_softmax = a->apply_softmax;
  p.k_rows = k_rows;
  p.split = a->use_tensor_cores >= 2;
  // P is fp16: normal from 2^-14, subnormal (absolute precision 2^-24) below, finite below 2^16.  A row's reference starts
  // ref_up ABOVE the first maximum it sees and moves (rescaling O, 64 TMEM columns per row) only when a later maximum exceeds it
  // by more than tau: the row maximum may grow by ref_up + tau before the first rescale, and the largest P of a row lies in
  // [2^-ref_up, 2^tau].  With ref_up = 0, tau = 12 (round 1) 46 % of the 64-key half tiles of the benchmark input had a row
  // of their warp jump, and every jump rescales all 32 rows of the warp.  The precise mode splits P into fp16 hi + lo and needs
  // lo = 2^-11 P to stay normal: it keeps ref_up = 0.
  static const float tau_env = [] { const char* e = getenv("CTX_ATTN_TAU"); return e ? (float)atof(e) : -1.f; }();
  static const float up_env = [] { const char* e = getenv("CTX_ATTN_REFUP"); return e ? (float)atof(e) : -1.f; }();
  p.tau = tau_env >= 0.f ? tau_env : (precise ? AT_TAU : 15.0f);
  p.ref_up = up_env >= 0.f ? up_env : (precise ? 0.0f : 6.0f);
  p.bulk_x = ((uintptr_t)a->conf % 16 == 0) && ((size_t)P * D * 4 % 16 == 0) && ((size_t)AT_BQ * D * 4 % 16 == 0) &&
             ((size_t)(P % AT_BQ) * D * 4 % 16 == 0);
  p.conf = a->conf; p.theta_b = a->theta_b; p.Wz = a->Wz; p.obj_w = a->obj_target_w; p.fc_w = a->fc_base_w; p.fc_b = a->fc_base_b;
  p.scale = a->scale; p.out = a->out;
  p.dbg = g_attn_dbg;
  static const int qt1_env = [] { const char* e = getenv("CTX_ATTN_QT1"); return (e && e[0] == '0') ? 0 : 1; }();
  if (!p.split && qt1_env && qt1_smem_base_ok(st)) {
    // fp16-logit mode: one q-tile per CTA, two CTAs per SM
    const size_t smem1 = (size_t)AT_TILE_Q + 2 * AT_TILE_Q + 64 * 1024 + 256 + 256 + 64;
    CTX_CUDA_TRY(cudaFuncSetAttribute(attention_tc_kernel<D, NN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    CTX_CUDA_TRY(cudaFuncSetAttribute(attention_tc_kernel<D, NN, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    attention_tc_kernel<D, NN, 1><<<dim3(cdiv(P, AT_BQ), B), 6 * 32, smem1, st>>>(tw, tk, tv, tvl, p);
    CTX_LAUNCH_CHECK();
    return CTX_OK;
  }
  const size_t smem = attn_tc_smem(D, a->num_novel, a->incre);
  CTX_CUDA_TRY(cudaFuncSetAttribute(attention_tc_kernel<D, NN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_tc_kernel<D, NN, 2><<<dim3(cdiv(P, AT_QT * AT_BQ), B), AT_THREADS, smem, st>>>(tw, tk, tv, tvl, p);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

int attention_tc_launch(const CtxAttnParams* a, cudaStream_t st) {
  if (a->dim == 60 && a->num_novel == 20) return attention_tc_launch_t<60, 20>(a, st);     // transfer (RFB_Net_vgg.py:157-163)
  if (a->dim == 15 && a->num_novel == 5) return attention_tc_launch_t<15, 5>(a, st);       // incre (:172-179)
  if (a->dim == 20 && a->num_novel == 20) return attention_tc_launch_t<20, 20>(a, st);
  set_error("attention: (dim %d, novel %d) not instantiated (60/20 transfer, 15/5 incre, 20/20)", a->dim, a->num_novel);
  return CTX_ERR_UNSUPPORTED;
}

}  // namespace ctx
