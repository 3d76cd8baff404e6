// train.cu — target assignment and hard-negative ranking of the few-shot fine-tune loop for sm_100a.
//
// Replaces, batched and on device:
//   utils/box_utils.py:83-132 (match), :5-14 (point_form), :29-68 (intersect / jaccard),
//   :135-156 (encode)  — the per-image python loop of multibox_loss_combined.py:70-74;
//   layers/modules/multibox_loss_combined.py:91-93 (the two full sorts that turn the mining loss
//   into a rank).
//
// Integer results (matched truth index, labels, obj flags, ranks) are bit-exact against the CPU
// oracle: every IoU is evaluated in fp32 with explicit round-to-nearest intrinsics in the
// reference's operation order, arg-max ties resolve to the first index (torch.max on CPU), and the
// forced matches are applied in ground-truth order so the last ground truth wins a collision
// (box_utils.py:122-123).
#include "common.cuh"
#include "sort.cuh"

namespace ctx {

constexpr int kMatchThreads = 1024;
constexpr int kMaxObj = 128;

// jaccard of a ground-truth box (corner form) and a prior (corner form) — box_utils.py:29-68
__device__ __forceinline__ float jaccard_one(float4 t, float area_t, float4 p) {
  float w = fmaxf(__fsub_rn(fminf(t.z, p.z), fmaxf(t.x, p.x)), 0.0f);
  float h = fmaxf(__fsub_rn(fminf(t.w, p.w), fmaxf(t.y, p.y)), 0.0f);
  float inter = __fmul_rn(w, h);
  float area_p = __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y));
  float uni = __fsub_rn(__fadd_rn(area_t, area_p), inter);
  return __fdiv_rn(inter, uni);
}

// point_form — box_utils.py:5-14
__device__ __forceinline__ float4 point_form(float4 p) {
  float hw = __fmul_rn(p.z, 0.5f), hh = __fmul_rn(p.w, 0.5f);
  return make_float4(__fsub_rn(p.x, hw), __fsub_rn(p.y, hh), __fadd_rn(p.x, hw), __fadd_rn(p.y, hh));
}

// One CTA per image.
__global__ void __launch_bounds__(kMatchThreads)
match_encode_kernel(const float* __restrict__ truths, const int* __restrict__ num_obj, int max_obj,
                    const float* __restrict__ priors, int P, float threshold, float v0, float v1,
                    float* __restrict__ loc_t, float* __restrict__ conf_t, unsigned char* __restrict__ obj_t,
                    int* __restrict__ best_truth_idx_out, float* __restrict__ best_truth_overlap_out) {
  __shared__ float s_t[kMaxObj][6];
  __shared__ float s_area[kMaxObj];
  __shared__ int s_best_prior[kMaxObj];
  __shared__ unsigned long long s_red[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int n = num_obj ? num_obj[b] : max_obj;
  n = min(max(n, 0), min(max_obj, kMaxObj));
  for (int i = tid; i < n * 6; i += blockDim.x) s_t[i / 6][i % 6] = truths[((size_t)b * max_obj) * 6 + i];
  __syncthreads();
  if (tid < n) s_area[tid] = __fmul_rn(__fsub_rn(s_t[tid][2], s_t[tid][0]), __fsub_rn(s_t[tid][3], s_t[tid][1]));
  __syncthreads();

  // best prior per ground truth: arg-max over priors, first index on ties (box_utils.py:108)
  for (int j = 0; j < n; ++j) {
    const float4 t = make_float4(s_t[j][0], s_t[j][1], s_t[j][2], s_t[j][3]);
    const float at = s_area[j];
    unsigned long long best = 0ull;                 // (iou bits << 32) | (~p): max == larger iou, then smaller p
    for (int p = tid; p < P; p += blockDim.x) {
      float4 pr = point_form(reinterpret_cast<const float4*>(priors)[p]);
      float ov = jaccard_one(t, at, pr);
      unsigned long long key = ((unsigned long long)float_order_bits(ov) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)p);
      best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    if (warp == 0) {
      best = lane < (int)(blockDim.x >> 5) ? s_red[lane] : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
      }
      if (lane == 0) s_best_prior[j] = (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull));
    }
    __syncthreads();
  }

  // best truth per prior (first index on ties, :110), forced matches (:119-123), labels, encode
  for (int p = tid; p < P; p += blockDim.x) {
    const float4 pc = reinterpret_cast<const float4*>(priors)[p];
    const float4 pr = point_form(pc);
    float best_ov = -INFINITY;
    int best_j = 0;
    for (int j = 0; j < n; ++j) {
      float ov = jaccard_one(make_float4(s_t[j][0], s_t[j][1], s_t[j][2], s_t[j][3]), s_area[j], pr);
      if (ov > best_ov) { best_ov = ov; best_j = j; }
    }
    const size_t at = (size_t)b * P + p;
    if (best_truth_overlap_out) best_truth_overlap_out[at] = n > 0 ? best_ov : 0.f;   // `overlap[idx]`, :115-116
    for (int j = 0; j < n; ++j)
      if (s_best_prior[j] == p) { best_ov = 2.0f; best_j = j; }
    float label = 0.f, weight = 1.f;
    float4 enc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n > 0) {
      label = s_t[best_j][4]; weight = s_t[best_j][5];
      if (best_ov < threshold) { label = 0.f; weight = 1.f; }
      // encode — box_utils.py:135-156
      const float mx1 = s_t[best_j][0], my1 = s_t[best_j][1], mx2 = s_t[best_j][2], my2 = s_t[best_j][3];
      float gx = __fsub_rn(__fmul_rn(__fadd_rn(mx1, mx2), 0.5f), pc.x);
      float gy = __fsub_rn(__fmul_rn(__fadd_rn(my1, my2), 0.5f), pc.y);
      gx = __fdiv_rn(gx, __fmul_rn(v0, pc.z));
      gy = __fdiv_rn(gy, __fmul_rn(v0, pc.w));
      float gw = __fdiv_rn(logf(__fdiv_rn(__fsub_rn(mx2, mx1), pc.z)), v1);
      float gh = __fdiv_rn(logf(__fdiv_rn(__fsub_rn(my2, my1), pc.w)), v1);
      enc = make_float4(gx, gy, gw, gh);
    }
    reinterpret_cast<float4*>(loc_t)[at] = enc;
    reinterpret_cast<float2*>(conf_t)[at] = make_float2(label, weight);
    obj_t[at] = label != 0.f ? 1 : 0;
    if (best_truth_idx_out) best_truth_idx_out[at] = best_j;
  }
}

// One CTA per image: rank[p] = position of p in the stable descending sort of loss[b, :].
__global__ void __launch_bounds__(1024)
hard_negative_rank_kernel(const float* __restrict__ loss, int P, int key_stride, uint64_t* __restrict__ keys_ws,
                          int* __restrict__ rank) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw);
  const int b = blockIdx.x;
  uint64_t* gkeys = keys_ws + (size_t)b * key_stride;
  for (int p = threadIdx.x; p < P; p += blockDim.x) gkeys[p] = make_key(loss[(size_t)b * P + p], (uint32_t)p);
  __syncthreads();
  const uint64_t* sorted = block_sort_desc(gkeys, P, skeys);
  for (int i = threadIdx.x; i < P; i += blockDim.x) rank[(size_t)b * P + key_index(sorted[i])] = i;
}

}  // namespace ctx

using namespace ctx;

extern "C" int ctx_match_encode(const float* truths, const int* num_obj, int max_obj, const float* priors, int batch,
                                int num_priors, float threshold, float var0, float var1, float* loc_t, float* conf_t,
                                unsigned char* obj_t, int* best_truth_idx, float* best_truth_overlap, void* stream) {
  CTX_REQUIRE(batch >= 0 && num_priors >= 1 && max_obj >= 0, "ctx_match_encode: bad sizes");
  CTX_REQUIRE(max_obj <= kMaxObj, "ctx_match_encode: max_obj %d exceeds the supported %d", max_obj, kMaxObj);
  CTX_REQUIRE(priors && loc_t && conf_t && obj_t && (truths || max_obj == 0), "ctx_match_encode: null pointer");
  if (batch == 0) return CTX_OK;
  match_encode_kernel<<<batch, kMatchThreads, 0, (cudaStream_t)stream>>>(truths, num_obj, max_obj, priors, num_priors,
                                                                        threshold, var0, var1, loc_t, conf_t, obj_t,
                                                                        best_truth_idx, best_truth_overlap);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

extern "C" size_t ctx_rank_workspace_bytes(int batch, int num_priors) {
  if (batch <= 0 || num_priors <= 0) return 256;
  return align_up(sizeof(uint64_t) * (size_t)batch * next_pow2(num_priors), 256);
}

extern "C" int ctx_hard_negative_rank(const float* loss, int batch, int num_priors, int* rank, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  CTX_REQUIRE(batch >= 0 && num_priors >= 1, "ctx_hard_negative_rank: bad sizes");
  CTX_REQUIRE(loss && rank, "ctx_hard_negative_rank: null pointer");
  if (batch == 0) return CTX_OK;
  const size_t need = ctx_rank_workspace_bytes(batch, num_priors);
  if (!workspace || workspace_bytes < need) {
    set_error("ctx_hard_negative_rank: workspace %zu < required %zu", workspace_bytes, need);
    return CTX_ERR_WORKSPACE;
  }
  const int smem = kSortSmem * (int)sizeof(uint64_t);
  hard_negative_rank_kernel<<<batch, 1024, smem, (cudaStream_t)stream>>>(loss, num_priors, next_pow2(num_priors),
                                                                         (uint64_t*)workspace, rank);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}


// ---- OBJ(Target) prototype initialisation : train.py:252-286 (init_reweight) -------------------------------------------
// Upstream gathers, per foreground class, the raw conf features of every prior matched to that class over <= init_iter
// batches (boolean-mask gathers + torch.cat growing lists), L2-normalises each row, averages per class and normalises the
// mean.  Here one pass over the match labels adds each positive prior's normalised feature row to a per-class fp64 sum
// (positives are a few dozen per image: the pass is bound by reading the [B, P] label plane), a tiny second kernel
// produces the normalised class means.
namespace ctx {
__global__ void __launch_bounds__(256)
prototype_accumulate_kernel(const float* __restrict__ feat, const float* __restrict__ conf_t, long long rows, int dim, int num_fg,
                            double* __restrict__ sums, int* __restrict__ counts) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float label = conf_t[2 * r];                           // conf_t[..., 0]: 0 background, -1 ignore, 1..num_fg class
  const int c = (int)label;
  if (!(label >= 1.f) || (float)c != label || c > num_fg) return;
  const float* x = feat + r * dim;
  float ss = 0.f;
  for (int d = 0; d < dim; ++d) ss = fmaf(x[d], x[d], ss);
  const float nrm = sqrtf(ss);                                 // torch: item / item.norm(dim=1, keepdim=True), no eps
  double* dst = sums + (long long)(c - 1) * dim;
  for (int d = 0; d < dim; ++d) atomicAdd(dst + d, (double)(x[d] / nrm));
  atomicAdd(counts + (c - 1), 1);
}

__global__ void prototype_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ counts, int dim, int first_class,
                                          float* __restrict__ out) {
  // one warp per output class: mean over the class's rows (empty class -> NaN, as torch's mean of an empty tensor), / its norm
  const int c = first_class + blockIdx.x, lane = threadIdx.x;
  const double n = (double)counts[c];
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) { const float m = (float)(sums[(long long)c * dim + d] / n); ss = fmaf(m, m, ss); }
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float nrm = sqrtf(ss);
  for (int d = lane; d < dim; d += 32) out[(long long)blockIdx.x * dim + d] = (float)(sums[(long long)c * dim + d] / n) / nrm;
}
}  // namespace ctx

extern "C" int ctx_prototype_accumulate(const float* feat, const float* conf_t, int batch, int num_priors, int dim, int num_fg,
                                        double* sums, int* counts, void* stream) {
  CTX_REQUIRE(batch >= 0 && num_priors >= 1 && dim >= 1 && num_fg >= 1, "ctx_prototype_accumulate: bad sizes");
  CTX_REQUIRE(feat && conf_t && sums && counts, "ctx_prototype_accumulate: null pointer");
  const long long rows = (long long)batch * num_priors;
  if (rows == 0) return CTX_OK;
  prototype_accumulate_kernel<<<cdiv(rows, 256), 256, 0, (cudaStream_t)stream>>>(feat, conf_t, rows, dim, num_fg, sums, counts);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

extern "C" int ctx_prototype_finalize(const double* sums, const int* counts, int num_fg, int dim, int first_class, float* out,
                                      void* stream) {
  CTX_REQUIRE(sums && counts && out, "ctx_prototype_finalize: null pointer");
  CTX_REQUIRE(num_fg >= 1 && dim >= 1 && first_class >= 0 && first_class < num_fg, "ctx_prototype_finalize: bad sizes");
  prototype_finalize_kernel<<<num_fg - first_class, 32, 0, (cudaStream_t)stream>>>(sums, counts, dim, first_class, out);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}


// ---- MultiBoxLoss_combined, forward + backward : layers/modules/multibox_loss_combined.py:76-122 ------------------------
// Upstream is ~40 framework ops with boolean-mask gathers (dynamic shapes -> host syncs) and an autograd tape.  Here:
//   ctx_loss_mining            the no-grad objectness CE used for mining (:88-90) + the per-image weighted positive count (:77)
//   (ctx_hard_negative_rank    the two sorts, :91-93)
//   ctx_loss_forward_backward  one pass over the priors: smooth-L1 on positives (:81-85), objectness CE on pos | neg (:98-100),
//                              class CE on the logit-combined scores [obj0 + log sum exp(conf), obj1 + conf_k] (:106-117), the
//                              three sums in fp64, and — since every term is a closed-form function of one prior's row — the
//                              gradients w.r.t. loc / conf / obj in the same pass (no tape, no second read).
namespace ctx {
constexpr int kLossThreads = 256;
constexpr int kMaxLossClasses = 64;

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  v = lane < (int)(blockDim.x >> 5) ? s_red[lane] : 0.0;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;                                   // valid in thread 0
}

__global__ void __launch_bounds__(kLossThreads)
loss_mining_kernel(const float* __restrict__ obj_p, const float* __restrict__ conf_t, const unsigned char* __restrict__ obj_t, int P,
                   float* __restrict__ mining, double* __restrict__ num_pos_w) {
  __shared__ double s_red[32];
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double w = 0.0;
  if (p < P) {
    const long long i = (long long)b * P + p;
    const float2 o = reinterpret_cast<const float2*>(obj_p)[i];
    const float2 t = reinterpret_cast<const float2*>(conf_t)[i];
    float l = 0.f;
    if (!obj_t[i]) {                          // target 0: lse(o) - o0; positives and ignored boxes do not compete (:90)
      const float m = fmaxf(o.x, o.y);
      l = (m + logf(expf(o.x - m) + expf(o.y - m))) - o.x;
    }
    mining[i] = l;
    if (t.x > 0.f) w = (double)t.y;
  }
  const double s = block_sum(w, s_red);
  if (threadIdx.x == 0 && s != 0.0) atomicAdd(num_pos_w + b, s);
}

template <int C>
__global__ void __launch_bounds__(kLossThreads)
loss_fwd_bwd_kernel(const float* __restrict__ loc_p, const float* __restrict__ conf_p, const float* __restrict__ obj_p,
                    const float* __restrict__ loc_t, const float* __restrict__ conf_t, const unsigned char* __restrict__ obj_t,
                    const int* __restrict__ rank, const long long* __restrict__ num_neg, int P,
                    double* __restrict__ sums, float* __restrict__ dloc, float* __restrict__ dconf, float* __restrict__ dobj_c,
                    float* __restrict__ dobj_o) {
  __shared__ double s_red[32];
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double sl = 0.0, sc = 0.0, so = 0.0;
  if (p < P) {
    const long long i = (long long)b * P + p;
    const float2 t = reinterpret_cast<const float2*>(conf_t)[i];
    const float w = t.y;
    const bool pos = t.x > 0.f;
    const bool neg = (long long)rank[i] < num_neg[b];
    float4 gl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pos) {                                 // smooth-L1 (beta = 1), weighted (:81-85)
      const float4 a = reinterpret_cast<const float4*>(loc_p)[i], g = reinterpret_cast<const float4*>(loc_t)[i];
      const float d[4] = {a.x - g.x, a.y - g.y, a.z - g.z, a.w - g.w};
      float s = 0.f, gd[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float ad = fabsf(d[k]);
        s += ad < 1.f ? 0.5f * d[k] * d[k] : ad - 0.5f;
        gd[k] = (ad < 1.f ? d[k] : (d[k] > 0.f ? 1.f : -1.f)) * w;
      }
      sl = (double)(s * w);
      gl = make_float4(gd[0], gd[1], gd[2], gd[3]);
    }
    reinterpret_cast<float4*>(dloc)[i] = gl;
    float2 go = make_float2(0.f, 0.f), gc = make_float2(0.f, 0.f);
    float* dcr = dconf + i * C;
    const long long label = (long long)t.x;     // conf_t[..., 0].long()
    if ((pos || neg) && label >= 0 && label <= C) {
      const float2 o = reinterpret_cast<const float2*>(obj_p)[i];
      // objectness CE against obj_t (:98-100)
      {
        const float m = fmaxf(o.x, o.y), e0 = expf(o.x - m), e1 = expf(o.y - m), z = e0 + e1;
        const bool tt = obj_t[i] != 0;
        so = (double)(((m + logf(z)) - (tt ? o.y : o.x)) * w);
        go = make_float2((e0 / z - (tt ? 0.f : 1.f)) * w, (e1 / z - (tt ? 1.f : 0.f)) * w);
      }
      // class CE on the combined logits (:106-117); log(sum(exp(conf))) un-stabilised, as upstream
      float c[C], S = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) { c[k] = conf_p[i * C + k]; S += expf(c[k]); }
      const float l0 = o.x + logf(S);
      float M = l0;
#pragma unroll
      for (int k = 0; k < C; ++k) M = fmaxf(M, o.y + c[k]);
      float Z = expf(l0 - M);
#pragma unroll
      for (int k = 0; k < C; ++k) Z += expf(o.y + c[k] - M);
      const float l_lab = label == 0 ? l0 : o.y + c[label - 1];
      sc = (double)(((M + logf(Z)) - l_lab) * w);
      const float g0 = (expf(l0 - M) / Z - (label == 0 ? 1.f : 0.f)) * w;
      float gsum = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const float gk = (expf(o.y + c[k] - M) / Z - (label == k + 1 ? 1.f : 0.f)) * w;
        gsum += gk;
        dcr[k] = gk + g0 * (expf(c[k]) / S);    // d l0 / d conf_k = softmax(conf)_k
      }
      gc = make_float2(g0, gsum);
    }                                           // rows outside pos | neg stay zero (dconf is cleared by the launcher)
    reinterpret_cast<float2*>(dobj_o)[i] = go;
    reinterpret_cast<float2*>(dobj_c)[i] = gc;
  }
  double v = block_sum(sl, s_red);
  if (threadIdx.x == 0 && v != 0.0) atomicAdd(sums + 0, v);
  v = block_sum(sc, s_red);
  if (threadIdx.x == 0 && v != 0.0) atomicAdd(sums + 1, v);
  v = block_sum(so, s_red);
  if (threadIdx.x == 0 && v != 0.0) atomicAdd(sums + 2, v);
}
}  // namespace ctx

extern "C" int ctx_loss_mining(const float* obj_p, const float* conf_t, const unsigned char* obj_t, int batch, int num_priors,
                               float* mining, double* num_pos_w, void* stream) {
  CTX_REQUIRE(obj_p && conf_t && obj_t && mining && num_pos_w, "ctx_loss_mining: null pointer");
  CTX_REQUIRE(batch >= 0 && num_priors >= 1, "ctx_loss_mining: bad sizes");
  if (batch == 0) return CTX_OK;
  CTX_CUDA_TRY(cudaMemsetAsync(num_pos_w, 0, sizeof(double) * batch, (cudaStream_t)stream));
  loss_mining_kernel<<<dim3(cdiv(num_priors, kLossThreads), batch), kLossThreads, 0, (cudaStream_t)stream>>>(obj_p, conf_t, obj_t, num_priors, mining, num_pos_w);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

extern "C" int ctx_loss_forward_backward(const float* loc_p, const float* conf_p, const float* obj_p, const float* loc_t, const float* conf_t,
                                         const unsigned char* obj_t, const int* rank, const long long* num_neg, int batch, int num_priors,
                                         int num_fg_classes, double* sums3, float* dloc, float* dconf, float* dobj_cls, float* dobj_obj,
                                         void* stream) {
  CTX_REQUIRE(loc_p && conf_p && obj_p && loc_t && conf_t && obj_t && rank && num_neg && sums3 && dloc && dconf && dobj_cls && dobj_obj,
              "ctx_loss_forward_backward: null pointer");
  CTX_REQUIRE(batch >= 0 && num_priors >= 1, "ctx_loss_forward_backward: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  CTX_CUDA_TRY(cudaMemsetAsync(sums3, 0, 3 * sizeof(double), st));
  if (batch == 0) return CTX_OK;
  CTX_CUDA_TRY(cudaMemsetAsync(dconf, 0, sizeof(float) * (size_t)batch * num_priors * num_fg_classes, st));
  const dim3 grid(cdiv(num_priors, kLossThreads), batch);
#define CTX_LOSS(CN) loss_fwd_bwd_kernel<CN><<<grid, kLossThreads, 0, st>>>(loc_p, conf_p, obj_p, loc_t, conf_t, obj_t, rank, num_neg, num_priors, sums3, dloc, dconf, dobj_cls, dobj_obj)
  switch (num_fg_classes) {
    case 20: CTX_LOSS(20); break;              // VOC (num_classes 21)
    case 60: CTX_LOSS(60); break;              // COCO source classes
    case 15: CTX_LOSS(15); break;
    case 5: CTX_LOSS(5); break;
    default: set_error("ctx_loss_forward_backward: %d foreground classes not instantiated (5, 15, 20, 60)", num_fg_classes); return CTX_ERR_UNSUPPORTED;
  }
#undef CTX_LOSS
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}


// ---- stand-alone box algebra : utils/box_utils.py:5-14 (point_form), :50-68 (jaccard), :135-156 (encode) ----------------
// The same device functions the match kernel uses, exposed for callers of the reference's helper API.
namespace ctx {
__global__ void __launch_bounds__(256) point_form_kernel(const float4* __restrict__ boxes, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = point_form(boxes[i]);
}
__global__ void __launch_bounds__(256) jaccard_kernel(const float4* __restrict__ a, int A, const float4* __restrict__ b, int B, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)A * B) return;
  const float4 t = a[i / B], p = b[i % B];
  out[i] = jaccard_one(t, __fmul_rn(__fsub_rn(t.z, t.x), __fsub_rn(t.w, t.y)), p);
}
__global__ void __launch_bounds__(256) encode_kernel(const float4* __restrict__ matched, const float4* __restrict__ priors, int n, float v0, float v1,
                                                     float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 m = matched[i], pc = priors[i];
  float gx = __fsub_rn(__fmul_rn(__fadd_rn(m.x, m.z), 0.5f), pc.x);
  float gy = __fsub_rn(__fmul_rn(__fadd_rn(m.y, m.w), 0.5f), pc.y);
  gx = __fdiv_rn(gx, __fmul_rn(v0, pc.z));
  gy = __fdiv_rn(gy, __fmul_rn(v0, pc.w));
  out[i] = make_float4(gx, gy, __fdiv_rn(logf(__fdiv_rn(__fsub_rn(m.z, m.x), pc.z)), v1), __fdiv_rn(logf(__fdiv_rn(__fsub_rn(m.w, m.y), pc.w)), v1));
}
}  // namespace ctx

extern "C" int ctx_point_form(const float* boxes, int n, float* out, void* stream) {
  CTX_REQUIRE(n >= 0 && (n == 0 || (boxes && out)), "ctx_point_form: bad arguments");
  if (n == 0) return CTX_OK;
  point_form_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)boxes, n, (float4*)out);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}
extern "C" int ctx_jaccard(const float* box_a, int na, const float* box_b, int nb, float* out, void* stream) {
  CTX_REQUIRE(na >= 0 && nb >= 0 && ((long long)na * nb == 0 || (box_a && box_b && out)), "ctx_jaccard: bad arguments");
  if ((long long)na * nb == 0) return CTX_OK;
  jaccard_kernel<<<cdiv((long long)na * nb, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)box_a, na, (const float4*)box_b, nb, out);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}
extern "C" int ctx_encode(const float* matched, const float* priors, int n, float var0, float var1, float* out, void* stream) {
  CTX_REQUIRE(n >= 0 && (n == 0 || (matched && priors && out)), "ctx_encode: bad arguments");
  if (n == 0) return CTX_OK;
  encode_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)matched, (const float4*)priors, n, var0, var1, (float4*)out);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}
