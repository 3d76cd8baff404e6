// attention_tc.cu — Context-Transformer block (models/RFB_Net_vgg.py:253-271) on the sm_100a tensor
// cores: spatially pooled keys/values, softmax(Q K^T) V, residual, L2-norm, cosine classifier and the
// class softmax as ONE warp-specialised kernel; the [B, P, Pk] affinity matrix never leaves the SM.
//
//   proj kernels (CUDA cores, fp32): Q = theta(conf)+conf, K = phi(pool)+pool, V = g(pool)+pool, each
//        written as fp16 operands.  The reference's logits are un-scaled dot products of raw class
//        scores (|s| up to ~1.5e3), so a plain 16-bit Q K^T is not accurate enough; Q and K are
//        therefore split into fp16 hi + lo parts and S = Qh Kh^T + Ql Kh^T + Qh Kl^T is accumulated in
//        fp32 by three tcgen05 MMAs (error ~2^-22 relative; measured final |dconf| 4e-5 vs fp32).
//   attention kernel: CTA = 128 queries of one image, keys streamed in tiles of 64.
//        warp 4  TMA producer   : Qh/Ql tile once, then {Kh, Kl, V^T} tiles through a 2-stage ring
//        warp 5  MMA issuer     : S[j%2] = Q K_j^T (12 UMMAs, M128 N64 K16) into a double-buffered TMEM
//                                 accumulator, then O += P_j V_j (4 UMMAs) once the softmax warps have
//                                 published P_j; tcgen05.commit drives every hand-off mbarrier
//        warps 0-3 softmax      : one query row per thread: tcgen05.ld the S row, streaming softmax with
//                                 a lazily updated reference maximum (O in TMEM is rescaled only when a
//                                 row's maximum grows by more than 8), p = ex2(s*log2e - ref), P written
//                                 as the fp16 A operand of the PV MMA (128B-swizzled K-major tile);
//                                 finally the epilogue: O/l, z = conf + O*Wz, z/||z||, OBJ_Target * scale
//                                 [, fc_base(conf)+conf for 'incre'], class softmax, store.
#include "tc_common.cuh"

namespace ctx {

constexpr int AT_BQ = 128;        // queries per CTA
constexpr int AT_BK = 64;         // keys per tile
constexpr int AT_DP = 64;         // padded feature dim
constexpr int AT_THREADS = 192;
constexpr int AT_TILE_Q = AT_BQ * AT_DP * 2;      // 16 KB
constexpr int AT_TILE_K = AT_BK * AT_DP * 2;      // 8 KB
constexpr float AT_LOG2E = 1.4426950408889634f;
constexpr float AT_RESCALE = 8.0f;                // lazy-rescale threshold (natural-log units)

// ---- projections (fp32 CUDA cores) -> fp16 hi/lo operands ------------------------------------------
// rows: B*P queries.  Qhl: [2][B*P][64] fp16 (hi, lo), feature columns D..63 zero.
template <int D>
__global__ void __launch_bounds__(128)
proj_q_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, long long rows,
              __half* __restrict__ qhl) {
  __shared__ float s_w[D * D], s_b[D];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) s_w[i] = w[i];
  for (int i = threadIdx.x; i < D; i += blockDim.x) s_b[i] = bias[i];
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float xv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) xv[d] = x[r * D + d];
  __half* hi = qhl + r * AT_DP;
  __half* lo = qhl + (rows + r) * AT_DP;
  for (int o0 = 0; o0 < AT_DP; o0 += 8) {
    __align__(16) __half h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int o = o0 + e;
      float q = 0.f;
      if (o < D) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) a = fmaf(s_w[o * D + d], xv[d], a);
        q = (a + s_b[o]) + xv[o];
      }
      h[e] = __float2half_rn(q);
      l[e] = __float2half_rn(q - __half2float(h[e]));
    }
    *reinterpret_cast<uint4*>(hi + o0) = *reinterpret_cast<uint4*>(h);
    *reinterpret_cast<uint4*>(lo + o0) = *reinterpret_cast<uint4*>(l);
  }
}

// rows: B*Pk_pad keys (rows >= Pk of an image are zero).  Khl: [2][B*Pk_pad][64]; Vt: [B][64][Pk_pad].
template <int D>
__global__ void __launch_bounds__(128)
proj_kv_kernel(const float* __restrict__ pooled, const float* __restrict__ phi_w, const float* __restrict__ phi_b,
               const float* __restrict__ g_w, const float* __restrict__ g_b, int B, int Pk, int Pk_pad,
               __half* __restrict__ khl, __half* __restrict__ vt) {
  __shared__ float s_phi[D * D], s_g[D * D], s_pb[D], s_gb[D];
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
      vt[((long long)b * AT_DP + o) * Pk_pad + j] = __float2half_rn(v);
    }
    *reinterpret_cast<uint4*>(hi + o0) = *reinterpret_cast<uint4*>(h);
    *reinterpret_cast<uint4*>(lo + o0) = *reinterpret_cast<uint4*>(l);
  }
}

// ---- fused attention --------------------------------------------------------------------------------
struct AttnTcParams {
  int B, P, Pk, Pk_pad, ntiles;
  int num_novel, incre, apply_softmax;
  long long q_rows, k_rows;         // B*P, B*Pk_pad (row offset of the "lo" halves)
  const float* conf;                // [B,P,D] fp32 (residual input x)
  const float *Wz, *obj_w, *fc_w, *fc_b;
  float scale;
  float* out;
};

template <int D>
__global__ void __launch_bounds__(AT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQh = base, sQl = base + AT_TILE_Q;
  const uint32_t sKV = base + 2 * AT_TILE_Q;                 // 2 stages x {Kh, Kl, Vt} = 2 x 24 KB
  const uint32_t sP = sKV + 2 * 3 * AT_TILE_K;               // 128 x 64 fp16 = 16 KB
  const uint32_t bars = sP + AT_TILE_Q;
  const uint32_t q_full = bars, kv_full = bars + 8, kv_empty = bars + 24, s_full = bars + 40, s_empty = bars + 56,
                 p_full = bars + 72, pv_done = bars + 80, tmem_slot = bars + 88;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  float* s_obj = reinterpret_cast<float*>(gen_base + (bars - base) + 128);     // epilogue weights: obj [n_novel][D], fc [D][D]+[D]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, q0 = blockIdx.x * AT_BQ;
  const int T = p.ntiles;

  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); }
  if (warp == 5) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(kv_full + 8 * s, 1); mbar_init(kv_empty + 8 * s, 1);
        mbar_init(s_full + 8 * s, 1); mbar_init(s_empty + 8 * s, 4);
      }
      mbar_init(p_full, 4);
      mbar_init(pv_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  // epilogue weights (read long after this barrier)
  for (int i = threadIdx.x; i < p.num_novel * D; i += blockDim.x) s_obj[i] = p.obj_w[i];
  if (p.incre) {
    float* s_fc = s_obj + p.num_novel * D;
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) s_fc[i] = p.fc_w[i];
    for (int i = threadIdx.x; i < D; i += blockDim.x) s_fc[D * D + i] = p.fc_b[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tS = tmem, tO = tmem + 128;              // S0: cols 0-63, S1: 64-127, O: 128-191

  if (warp == 4) {
    // ================= TMA producer ================= (whole warp converged, one elected lane issues)
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * AT_TILE_Q);
      tma_load_2d(sQh, &tm_q, 0, b * p.P + q0, q_full);
      tma_load_2d(sQl, &tm_q, 0, (int)p.q_rows + b * p.P + q0, q_full);
    }
    __syncwarp();
    for (int j = 0; j < T; ++j) {
      const int s = j & 1;
      mbar_wait(kv_empty + 8 * s, ((j >> 1) & 1) ^ 1);
      if (elect_one()) {
        const uint32_t dst = sKV + s * 3 * AT_TILE_K;
        mbar_arrive_expect_tx(kv_full + 8 * s, 3 * AT_TILE_K);
        tma_load_2d(dst, &tm_k, 0, b * p.Pk_pad + j * AT_BK, kv_full + 8 * s);
        tma_load_2d(dst + AT_TILE_K, &tm_k, 0, (int)p.k_rows + b * p.Pk_pad + j * AT_BK, kv_full + 8 * s);
        tma_load_2d(dst + 2 * AT_TILE_K, &tm_v, j * AT_BK, b * AT_DP, kv_full + 8 * s);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ================= MMA issuer ================= (whole warp converged, one elected lane issues)
    const uint32_t idesc = make_idesc_f16(false, AT_BQ, AT_BK);     // fp16 operands, M128 x N64 (keys or features)
    const uint64_t dQh = make_sw128_desc(sQh), dQl = make_sw128_desc(sQl), dP = make_sw128_desc(sP), dKV = make_sw128_desc(sKV);
    mbar_wait(q_full, 0);
    for (int j = 0; j <= T; ++j) {
      if (j < T) {
        const int s = j & 1;
        mbar_wait(kv_full + 8 * s, (j >> 1) & 1);
        mbar_wait(s_empty + 8 * s, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t kh = dKV + (uint64_t)((s * 3 * AT_TILE_K) >> 4), kl = kh + (uint64_t)(AT_TILE_K >> 4);
          const uint32_t d = tS + s * AT_BK;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, dQh + 2 * k, kh + 2 * k, idesc, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, dQl + 2 * k, kh + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d, dQh + 2 * k, kl + 2 * k, idesc, 1u);
          umma_commit(s_full + 8 * s);
        }
        __syncwarp();
      }
      if (j > 0) {
        const int jj = j - 1, s = jj & 1;
        mbar_wait(p_full, jj & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t vt = dKV + (uint64_t)((s * 3 * AT_TILE_K + 2 * AT_TILE_K) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tO, dP + 2 * k, vt + 2 * k, idesc, (jj | k) ? 1u : 0u);
          umma_commit(kv_empty + 8 * s);
          umma_commit(pv_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= softmax warps + epilogue =================
    const int r = warp * 32 + lane;
    const int q = q0 + r;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    float ref = -INFINITY;          // reference maximum (natural-log units), lazily updated
    float l = 0.f;
    for (int j = 0; j < T; ++j) {
      const int s = j & 1;
      mbar_wait(s_full + 8 * s, (j >> 1) & 1);
      tc_fence_after();
      uint32_t v[64];
      tmem_ld32(tS + lane_sel + s * AT_BK, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(tS + lane_sel + s * AT_BK + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty + 8 * s);
      const int nvalid = p.Pk - j * AT_BK;                   // keys of this tile that exist
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float sv = __uint_as_float(v[c]);
        if (c >= nvalid) sv = -INFINITY;
        v[c] = __float_as_uint(sv);
        mx = fmaxf(mx, sv);
      }
      float factor = 1.f;
      const bool grow = mx > ref + AT_RESCALE;
      if (grow) {
        factor = (ref == -INFINITY) ? 0.f : __expf(ref - mx);
        ref = mx;
        l *= factor;
      }
      // the P buffer and O are free once PV_{j-1} has completed
      if (j > 0) mbar_wait(pv_done, (j - 1) & 1);
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        tc_fence_after();
        uint32_t o[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tmem_ld32(tO + lane_sel + h * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * factor);
          tmem_st32(tO + lane_sel + h * 32, o);
        }
        tmem_st_wait();
      }
      const float ref2 = ref * AT_LOG2E;
      float sum = 0.f;
      const uint32_t prow = sP + r * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(v[ch * 8 + 2 * e]), AT_LOG2E, -ref2));
          const float p1 = fast_exp2(fmaf(__uint_as_float(v[ch * 8 + 2 * e + 1]), AT_LOG2E, -ref2));
          sum += p0 + p1;
          __half2 hh = __floats2half2_rn(p0, p1);
          pk[e] = *reinterpret_cast<uint32_t*>(&hh);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((ch ^ (r & 7)) << 4)), "r"(pk[0]), "r"(pk[1]),
                     "r"(pk[2]), "r"(pk[3]) : "memory");
      }
      l += sum;
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }

    // ---- epilogue: z = x + (O/l)*Wz ; z/||z|| ; OBJ_Target*scale ; [fc_base] ; [softmax] ----
    mbar_wait(pv_done, (T - 1) & 1);
    tc_fence_after();
    uint32_t o[64];
    tmem_ld32(tO + lane_sel, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
    tmem_ld32(tO + lane_sel + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
    tmem_ld_wait();
    if (q < p.P) {
      const float* xrow = p.conf + ((size_t)b * p.P + q) * D;
      const float inv_l = 1.0f / l;
      float x[D], z[D];
      float nrm = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        x[d] = xrow[d];
        z[d] = x[d] + (__uint_as_float(o[d]) * inv_l) * p.Wz[d];
        nrm = fmaf(z[d], z[d], nrm);
      }
      const float inv_n = 1.0f / sqrtf(nrm);
      const int n_out = p.num_novel + (p.incre ? D : 0);
      float* orow = p.out + ((size_t)b * p.P + q) * n_out;
      float outv[64];
      int no = 0;
      if (p.incre) {
        const float* s_fc = s_obj + p.num_novel * D;
        for (int c = 0; c < D; ++c) {
          float a = 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) a = fmaf(s_fc[c * D + d], x[d], a);
          outv[no++] = (a + s_fc[D * D + c]) + x[c];
        }
      }
      for (int c = 0; c < p.num_novel; ++c) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) a = fmaf(s_obj[c * D + d], z[d] * inv_n, a);
        outv[no++] = a * p.scale;
      }
      if (p.apply_softmax) {
        float mxo = -INFINITY;
        for (int c = 0; c < no; ++c) mxo = fmaxf(mxo, outv[c]);
        float sm = 0.f;
        for (int c = 0; c < no; ++c) { outv[c] = expf(outv[c] - mxo); sm += outv[c]; }
        const float inv = 1.0f / sm;
        for (int c = 0; c < no; ++c) outv[c] *= inv;
      }
      for (int c = 0; c < no; ++c) orow[c] = outv[c];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

static size_t attn_tc_smem(int D, int num_novel, int incre) {
  return 1024 + 2 * AT_TILE_Q + 2 * 3 * AT_TILE_K + AT_TILE_Q + 128 + sizeof(float) * (num_novel * D + (incre ? D * D + D : 0)) + 64;
}

size_t attention_tc_workspace_bytes(int B, int P, int Pk) {
  const size_t Pk_pad = (size_t)(Pk + AT_BK - 1) / AT_BK * AT_BK;
  return align_up((size_t)2 * B * P * AT_DP * 2, 1024) + align_up((size_t)2 * B * Pk_pad * AT_DP * 2, 1024) +
         align_up((size_t)B * AT_DP * Pk_pad * 2, 1024);
}

template <int D>
static int attention_tc_launch_t(const CtxAttnParams* a, cudaStream_t st) {
  const int B = a->batch, P = a->num_priors, Pk = a->num_pooled;
  const int Pk_pad = (Pk + AT_BK - 1) / AT_BK * AT_BK;
  const size_t need = attention_tc_workspace_bytes(B, P, Pk);
  if (!a->workspace || a->workspace_bytes < need) {
    set_error("attention (tensor-core path): workspace %zu < required %zu", a->workspace_bytes, need);
    return CTX_ERR_WORKSPACE;
  }
  CTX_REQUIRE(((uintptr_t)a->workspace) % 1024 == 0, "attention: workspace must be 1024-byte aligned");
  CTX_REQUIRE((long long)2 * B * P < (1ll << 31) && (long long)2 * B * Pk_pad < (1ll << 31), "attention: too many rows");
  char* ws = (char*)a->workspace;
  __half* qhl = (__half*)ws;
  __half* khl = (__half*)(ws + align_up((size_t)2 * B * P * AT_DP * 2, 1024));
  __half* vt = (__half*)((char*)khl + align_up((size_t)2 * B * Pk_pad * AT_DP * 2, 1024));
  const long long q_rows = (long long)B * P, k_rows = (long long)B * Pk_pad;
  proj_q_kernel<D><<<cdiv(q_rows, 128), 128, 0, st>>>(a->conf, a->theta_w, a->theta_b, q_rows, qhl);
  CTX_LAUNCH_CHECK();
  proj_kv_kernel<D><<<cdiv(k_rows, 128), 128, 0, st>>>(a->pooled, a->phi_w, a->phi_b, a->g_w, a->g_b, B, Pk, Pk_pad, khl, vt);
  CTX_LAUNCH_CHECK();
  CUtensorMap tq, tk, tv;
  int rc = encode_2d_sw128(&tq, qhl, false, 2ull * q_rows, AT_DP, AT_BQ);
  if (!rc) rc = encode_2d_sw128(&tk, khl, false, 2ull * k_rows, AT_DP, AT_BK);
  if (!rc) rc = encode_2d_sw128(&tv, vt, false, (unsigned long long)B * AT_DP, (unsigned long long)Pk_pad, AT_DP);
  if (rc) return rc;
  AttnTcParams p;
  p.B = B; p.P = P; p.Pk = Pk; p.Pk_pad = Pk_pad; p.ntiles = Pk_pad / AT_BK;
  p.num_novel = a->num_novel; p.incre = a->incre; p.apply_softmax = a->apply_softmax;
  p.q_rows = q_rows; p.k_rows = k_rows;
  p.conf = a->conf; p.Wz = a->Wz; p.obj_w = a->obj_target_w; p.fc_w = a->fc_base_w; p.fc_b = a->fc_base_b;
  p.scale = a->scale; p.out = a->out;
  const size_t smem = attn_tc_smem(D, a->num_novel, a->incre);
  CTX_CUDA_TRY(cudaFuncSetAttribute(attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_tc_kernel<D><<<dim3(cdiv(P, AT_BQ), B), AT_THREADS, smem, st>>>(tq, tk, tv, p);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

int attention_tc_launch(const CtxAttnParams* a, cudaStream_t st) {
  if (a->dim == 60) return attention_tc_launch_t<60>(a, st);
  if (a->dim == 15) return attention_tc_launch_t<15>(a, st);
  if (a->dim == 20) return attention_tc_launch_t<20>(a, st);
  set_error("attention: dim %d not instantiated (60 transfer / 15 incre / 20)", a->dim);
  return CTX_ERR_UNSUPPORTED;
}

}  // namespace ctx
