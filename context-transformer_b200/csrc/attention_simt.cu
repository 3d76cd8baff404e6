// attention_simt.cu — Context-Transformer block (models/RFB_Net_vgg.py:253-271) in fp32 on CUDA
// cores, fused so the [B, P, Pk] affinity matrix (86 MB / image at 300x300 upstream) never exists.
//
//   K = phi(pool)+pool, V = g(pool)+pool                         kv_project_kernel   (tiny)
//   Q = theta(conf)+conf ; W = softmax(Q K^T) (no 1/sqrt(d)) ; delta = (W V) * Wz ;
//   z = conf + delta ; z /= ||z|| ; novel = OBJ_Target(z) * scale ;
//   conf = novel (transfer) | cat(fc_base(conf)+conf, novel) (incre) ; [softmax in eval]
//                                                                 attention_kernel
// One thread owns one query row (q[D] and the output accumulator in registers); K/V tiles are
// staged in shared memory and read with warp-broadcast vector loads; streaming (online) softmax
// with the running maximum updated once per group of 8 keys.  This is the fp32 precision mode;
// the tensor-core mode lives in attention_tc.cu.
#include "common.cuh"

namespace ctx {

constexpr int kAttnThreads = 128;
constexpr int kKeyTile = 64;

template <int D>
__global__ void __launch_bounds__(128)
kv_project_kernel(const float* __restrict__ pooled, const float* __restrict__ phi_w, const float* __restrict__ phi_b,
                  const float* __restrict__ g_w, const float* __restrict__ g_b, long long rows,
                  float* __restrict__ k_out, float* __restrict__ v_out) {
  __shared__ float s_phi[D * D], s_g[D * D], s_pb[D], s_gb[D];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) { s_phi[i] = phi_w[i]; s_g[i] = g_w[i]; }
  for (int i = threadIdx.x; i < D; i += blockDim.x) { s_pb[i] = phi_b[i]; s_gb[i] = g_b[i]; }
  __syncthreads();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float x[D];
#pragma unroll
  for (int d = 0; d < D; ++d) x[d] = pooled[r * D + d];
  for (int o = 0; o < D; ++o) {
    float ak = 0.f, av = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) { ak = fmaf(s_phi[o * D + d], x[d], ak); av = fmaf(s_g[o * D + d], x[d], av); }
    k_out[r * D + o] = (ak + s_pb[o]) + x[o];
    v_out[r * D + o] = (av + s_gb[o]) + x[o];
  }
}

template <int D>
__global__ void __launch_bounds__(kAttnThreads)
attention_kernel(CtxAttnParams p) {
  constexpr int DP = (D + 3) & ~3;                       // padded row length in shared memory
  __shared__ __align__(16) float s_k[kKeyTile * DP];
  __shared__ __align__(16) float s_v[kKeyTile * DP];
  const int b = blockIdx.y;
  const int row = blockIdx.x * kAttnThreads + threadIdx.x;
  const bool valid = row < p.num_priors;
  const float* xrow = p.conf + ((size_t)b * p.num_priors + (valid ? row : 0)) * D;

  // ---- Q = theta(x) + x  (theta staged through s_k/s_v: D*D floats <= 2*kKeyTile*DP)
  static_assert(D * D <= 2 * kKeyTile * DP, "theta does not fit the staging buffer");
  // theta rows that fit go to s_k, the rest to s_v
  constexpr int ROWS_A = (kKeyTile * DP) / D;             // theta rows that fit in s_k
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    int o = i / D;
    if (o < ROWS_A) s_k[i] = p.theta_w[i]; else s_v[i - ROWS_A * D] = p.theta_w[i];
  }
  __syncthreads();
  float q[DP];
#pragma unroll
  for (int d = D; d < DP; ++d) q[d] = 0.f;
  {
    float x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = xrow[d];
#pragma unroll
    for (int o = 0; o < D; ++o) {
      const float* w = (o < ROWS_A) ? (s_k + o * D) : (s_v + (o - ROWS_A) * D);
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) a = fmaf(w[d], x[d], a);
      q[o] = (a + p.theta_b[o]) + x[o];
    }
  }
  __syncthreads();

  // ---- streaming softmax(Q K^T) V
  float acc[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) acc[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  const float* kv_scratch = reinterpret_cast<const float*>(p.workspace);
  const float* kb = kv_scratch + (size_t)b * p.num_pooled * D;
  const float* vb = kv_scratch + (size_t)p.batch * p.num_pooled * D + (size_t)b * p.num_pooled * D;
  for (int t0 = 0; t0 < p.num_pooled; t0 += kKeyTile) {
    const int nt = min(kKeyTile, p.num_pooled - t0);
    for (int i = threadIdx.x; i < kKeyTile * DP; i += blockDim.x) {
      int j = i / DP, d = i - j * DP;
      float kvv = 0.f, vvv = 0.f;
      if (j < nt && d < D) { kvv = kb[(size_t)(t0 + j) * D + d]; vvv = vb[(size_t)(t0 + j) * D + d]; }
      s_k[i] = kvv; s_v[i] = vvv;
    }
    __syncthreads();
    for (int j0 = 0; j0 < nt; j0 += 8) {
      float s[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float4* kr = reinterpret_cast<const float4*>(s_k + (j0 + jj) * DP);
        float a = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < DP / 4; ++d4) {
          const float4 kv = kr[d4];
          a = fmaf(q[4 * d4], kv.x, a); a = fmaf(q[4 * d4 + 1], kv.y, a);
          a = fmaf(q[4 * d4 + 2], kv.z, a); a = fmaf(q[4 * d4 + 3], kv.w, a);
        }
        s[jj] = (j0 + jj < nt) ? a : -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int jj = 1; jj < 8; ++jj) mx = fmaxf(mx, s[jj]);
      const float mn = fmaxf(m, mx);
      const float corr = expf(m - mn);                   // m = -inf on the first group -> 0
      l *= corr;
#pragma unroll
      for (int d = 0; d < DP; ++d) acc[d] *= corr;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float pj = expf(s[jj] - mn);               // -inf -> 0 for the ragged tail
        l += pj;
        const float4* vr = reinterpret_cast<const float4*>(s_v + (j0 + jj) * DP);
#pragma unroll
        for (int d4 = 0; d4 < DP / 4; ++d4) {
          const float4 vv = vr[d4];
          acc[4 * d4] = fmaf(pj, vv.x, acc[4 * d4]); acc[4 * d4 + 1] = fmaf(pj, vv.y, acc[4 * d4 + 1]);
          acc[4 * d4 + 2] = fmaf(pj, vv.z, acc[4 * d4 + 2]); acc[4 * d4 + 3] = fmaf(pj, vv.w, acc[4 * d4 + 3]);
        }
      }
      m = mn;
    }
    __syncthreads();
  }

  // ---- epilogue: z = x + (acc/l)*Wz ; z/||z|| ; OBJ_Target ; [fc_base] ; [softmax]
  const int n_novel = p.num_novel;
  float* s_obj = s_k;                                    // [n_novel][D]
  for (int i = threadIdx.x; i < n_novel * D; i += blockDim.x) s_obj[i] = p.obj_target_w[i];
  float* s_fc = s_v;                                     // [D][D] (+ bias) for incre
  if (p.incre) {
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) s_fc[i] = p.fc_base_w[i];
    for (int i = threadIdx.x; i < D; i += blockDim.x) s_fc[D * D + i] = p.fc_base_b[i];
  }
  __syncthreads();
  if (!valid) return;
  const float inv_l = 1.0f / l;
  float nrm = 0.f;
  float x[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    x[d] = xrow[d];
    const float z = x[d] + (acc[d] * inv_l) * p.Wz[d];
    acc[d] = z;
    nrm = fmaf(z, z, nrm);
  }
  const float inv_n = 1.0f / sqrtf(nrm);
  constexpr int MAXOUT = 64;
  float o[MAXOUT];
  int n_out = 0;
  if (p.incre) {
    for (int c = 0; c < D; ++c) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) a = fmaf(s_fc[c * D + d], x[d], a);
      o[n_out++] = (a + s_fc[D * D + c]) + x[c];
    }
  }
  for (int c = 0; c < n_novel; ++c) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) a = fmaf(s_obj[c * D + d], acc[d] * inv_n, a);
    o[n_out++] = a * p.scale;
  }
  if (p.apply_softmax) {
    float mx = -INFINITY;
    for (int c = 0; c < n_out; ++c) mx = fmaxf(mx, o[c]);
    float sum = 0.f;
    for (int c = 0; c < n_out; ++c) { o[c] = expf(o[c] - mx); sum += o[c]; }
    const float inv = 1.0f / sum;
    for (int c = 0; c < n_out; ++c) o[c] *= inv;
  }
  float* orow = p.out + ((size_t)b * p.num_priors + row) * n_out;
  for (int c = 0; c < n_out; ++c) orow[c] = o[c];
}

int attention_tc_launch(const CtxAttnParams* a, cudaStream_t st);
size_t attention_tc_workspace_bytes(int B, int P, int Pk);

template <int D>
static int attention_launch_t(const CtxAttnParams* p, cudaStream_t st) {
  const long long rows = (long long)p->batch * p->num_pooled;
  float* k_out = reinterpret_cast<float*>(p->workspace);
  float* v_out = k_out + rows * D;
  kv_project_kernel<D><<<cdiv(rows, 128), 128, 0, st>>>(p->pooled, p->phi_w, p->phi_b, p->g_w, p->g_b, rows, k_out, v_out);
  CTX_LAUNCH_CHECK();
  dim3 grid(cdiv(p->num_priors, kAttnThreads), p->batch);
  attention_kernel<D><<<grid, kAttnThreads, 0, st>>>(*p);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

int attention_simt_launch(const CtxAttnParams* p, cudaStream_t st) {
  CTX_REQUIRE(p, "attention: null params");
  CTX_REQUIRE(p->conf && p->pooled && p->theta_w && p->theta_b && p->phi_w && p->phi_b && p->g_w && p->g_b && p->Wz &&
              p->obj_target_w && p->workspace && p->out, "attention: null pointer");
  CTX_REQUIRE(!p->incre || (p->fc_base_w && p->fc_base_b), "attention: incre needs fc_base");
  CTX_REQUIRE(p->batch > 0 && p->num_priors > 0 && p->num_pooled > 0, "attention: bad sizes");
  CTX_REQUIRE(p->num_novel > 0 && p->num_novel + (p->incre ? p->dim : 0) <= 64, "attention: too many output classes");
  if (p->use_tensor_cores) {
    CTX_REQUIRE(p->num_novel <= 32, "attention (tensor-core path): num_novel must be <= 32");
    return attention_tc_launch(p, st);
  }
  CTX_REQUIRE(p->workspace_bytes >= sizeof(float) * 2 * (size_t)p->batch * p->num_pooled * p->dim, "attention: workspace too small");
  if (p->dim == 60) return attention_launch_t<60>(p, st);
  if (p->dim == 15) return attention_launch_t<15>(p, st);
  if (p->dim == 20) return attention_launch_t<20>(p, st);
  set_error("attention: dim %d not instantiated (60 transfer / 15 incre / 20)", p->dim);
  return CTX_ERR_UNSUPPORTED;
}

}  // namespace ctx

namespace ctx { void attention_set_debug_buffer(void* p); }
/* development aid: device buffer of >= 512 int64 that receives clock64() stamps of CTA (0,0) of the tensor-core kernel */
extern "C" void ctx_debug_set_attention_timeline(void* device_buffer) { ctx::attention_set_debug_buffer(device_buffer); }

extern "C" size_t ctx_attention_workspace_bytes(const CtxAttnParams* p) {
  if (!p || p->batch <= 0 || p->num_priors <= 0 || p->num_pooled <= 0 || p->dim <= 0) return 1024;
  if (p->use_tensor_cores) return ctx::attention_tc_workspace_bytes(p->batch, p->num_priors, p->num_pooled);
  return ctx::align_up(sizeof(float) * 2 * (size_t)p->batch * p->num_pooled * p->dim, 1024);
}

extern "C" int ctx_attention_forward(const CtxAttnParams* p, void* stream) {
  return ctx::attention_simt_launch(p, (cudaStream_t)stream);
}
