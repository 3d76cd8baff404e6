// post.cu — prior-box decode, score combine, per-class (soft-)NMS and per-image top-k for sm_100a.
//
// Replaces, on device and without host round trips:
//   layers/functions/detection.py:18-55 (Detect.forward), utils/box_utils.py:184-202 (decode),
//   test.py:133-161 (scale, per-class threshold, nms(), max_per_image cut),
//   utils/nms/nms_kernel.cu:34-144 + gpu_nms.pyx:16-31 (gpu_nms), cpu_nms.pyx:17-68 (cpu_nms),
//   cpu_nms.pyx:70-163 (cpu_soft_nms).
//
// Arithmetic that decides an integer result (IoU vs threshold, score vs threshold) is written with
// explicit round-to-nearest intrinsics in the reference's operation order, so kept indices are
// bit-exact against the CPU oracle; no FMA contraction can change a comparison.
//
// HBM-bound pieces (decode/score, candidate selection) are coalesced one-pass kernels; the NMS
// itself is latency/on-chip bound: one CTA per (image, class) that never materialises the n x n
// IoU bit-matrix of the reference (nms_kernel.cu:110-122) — candidates are streamed in chunks of
// 64 against the list of already-kept boxes, which lives in shared memory.
#include "common.cuh"

#include <stdlib.h>
#include "sort.cuh"
#include <algorithm>

namespace ctx {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
// decode (box_utils.py:184-202) followed by the pixel scale of test.py:136
__device__ __forceinline__ float4 decode_box(float4 l, float4 p, float v0, float v1) {
  float cx = __fadd_rn(p.x, __fmul_rn(__fmul_rn(l.x, v0), p.z));
  float cy = __fadd_rn(p.y, __fmul_rn(__fmul_rn(l.y, v0), p.w));
  float w = __fmul_rn(p.z, expf(__fmul_rn(l.z, v1)));
  float h = __fmul_rn(p.w, expf(__fmul_rn(l.w, v1)));
  float x1 = __fsub_rn(cx, __fmul_rn(w, 0.5f));
  float y1 = __fsub_rn(cy, __fmul_rn(h, 0.5f));
  return make_float4(x1, y1, __fadd_rn(w, x1), __fadd_rn(h, y1));
}

// IoU with the +1 pixel convention, fp32, reference operation order (cpu_nms.pyx:24,56-64;
// nms_kernel.cu:24-32 computes the same expression).
__device__ __forceinline__ float area_p1(float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}
__device__ __forceinline__ bool iou_suppresses(float4 a, float area_a, float4 b, float area_b, float thresh,
                                               bool on_equal) {
  float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
  float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
  float inter = __fmul_rn(w, h);
  float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return on_equal ? (ovr >= thresh) : (ovr > thresh);
}

__device__ __forceinline__ float4 load_box(const float* base, int stride, uint32_t idx) {
  const float* p = base + (size_t)idx * stride;
  return make_float4(p[0], p[1], p[2], p[3]);
}

// ------------------------------------------------------------------------------------------------
// Detect.forward: boxes = decode(loc, priors), scores = cat(obj0, obj1 * conf)
//   algorithmic bytes / prior (fp32, C fg classes): read 16 + 4C + 8 (+16 priors, L2-resident),
//   write 16 + 4(C+1)  -> 204 B at C = 20.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
detect_forward_kernel(const float* __restrict__ loc, const float* __restrict__ conf, const float* __restrict__ obj,
                      const float* __restrict__ priors, int B, int P, int C, float v0, float v1,
                      float* __restrict__ boxes, float* __restrict__ scores) {
  const long long nbox = (long long)B * P;
  const long long nsc = nbox * (C + 1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = t0; i < nbox; i += stride) {
    int p = (int)(i % P);
    float4 l = reinterpret_cast<const float4*>(loc)[i];
    float4 pr = reinterpret_cast<const float4*>(priors)[p];
    reinterpret_cast<float4*>(boxes)[i] = decode_box(l, pr, v0, v1);
  }
  const int C1 = C + 1;
  for (long long i = t0; i < nsc; i += stride) {
    long long bp = i / C1;
    int c = (int)(i - bp * C1);
    float2 o = reinterpret_cast<const float2*>(obj)[bp];
    scores[i] = (c == 0) ? o.x : __fmul_rn(o.y, conf[bp * C + (c - 1)]);
  }
}

// ------------------------------------------------------------------------------------------------
// Candidate selection (test.py:136,142-151): decode + scale every prior once, append
// (score, prior) keys of every (prior, class) with score > thresh to the per-(image, class) list.
// One pass over loc/conf/obj, coalesced; list order is arbitrary (the NMS kernel sorts).
// ------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 128;
constexpr int kMaxClasses = 128;

__global__ void __launch_bounds__(kSelThreads)
select_candidates_kernel(const float* __restrict__ loc, const float* __restrict__ conf,
                         const float* __restrict__ obj, const float* __restrict__ priors,
                         const float* __restrict__ scale, int scale_per_image, int P, int C, float v0, float v1,
                         float thresh, float4* __restrict__ boxes_px, int* __restrict__ cand_count,
                         uint64_t* __restrict__ cand_keys, int key_stride) {
  __shared__ int s_cnt[kMaxClasses];
  __shared__ int s_base[kMaxClasses];
  const int b = blockIdx.y;
  const int p = blockIdx.x * kSelThreads + threadIdx.x;
  for (int j = threadIdx.x; j < C; j += kSelThreads) s_cnt[j] = 0;
  __syncthreads();
  float o1 = 0.f;
  const float* crow = nullptr;
  if (p < P) {
    const long long bp = (long long)b * P + p;
    float4 l = reinterpret_cast<const float4*>(loc)[bp];
    float4 pr = reinterpret_cast<const float4*>(priors)[p];
    float4 bx = decode_box(l, pr, v0, v1);
    const float* sc = scale + (scale_per_image ? 4 * b : 0);
    bx.x = __fmul_rn(bx.x, sc[0]); bx.y = __fmul_rn(bx.y, sc[1]);
    bx.z = __fmul_rn(bx.z, sc[2]); bx.w = __fmul_rn(bx.w, sc[3]);
    boxes_px[bp] = bx;
    o1 = obj[bp * 2 + 1];
    crow = conf + bp * C;
    for (int j = 0; j < C; ++j)
      if (__fmul_rn(o1, crow[j]) > thresh) atomicAdd(&s_cnt[j], 1);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < C; j += kSelThreads) {
    int c = s_cnt[j];
    s_base[j] = c ? atomicAdd(&cand_count[b * C + j], c) : 0;
    s_cnt[j] = 0;
  }
  __syncthreads();
  if (p < P) {
    for (int j = 0; j < C; ++j) {
      float s = __fmul_rn(o1, crow[j]);
      if (s > thresh) {
        int pos = s_base[j] + atomicAdd(&s_cnt[j], 1);
        cand_keys[((size_t)b * C + j) * key_stride + pos] = make_key(s, (uint32_t)p);
      }
    }
  }
}

constexpr int kNmsThreads = 512;
constexpr int kKeptSmem = 1024;      // kept boxes held in shared memory (16 KB); more spill to global

// ------------------------------------------------------------------------------------------------
// Greedy hard NMS of one sorted candidate list by one CTA (512 threads = 64 candidates x 8 slices).
// keys: sorted descending; boxes fetched as box_base[idx*box_stride + 0..3].
// Writes kept indices (and their scores) in keep order; returns the kept count (uniform).
// ------------------------------------------------------------------------------------------------
struct NmsSmem {
  uint64_t keys[kSortSmem];
  float4 kept[kKeptSmem];
  float4 cbox[64];
  uint64_t diag[64];
  uint32_t cidx[64];
  uint32_t sup[2];
  uint64_t keptmask;
  int scan[40];
};

__device__ int block_hard_nms(const uint64_t* keys, int n, const float* box_base, int box_stride,
                              const float* score_base, int score_stride, float thresh, bool on_equal,
                              float4* kept_spill, int* kept_idx, float* kept_score, NmsSmem& sm) {
  const int tid = threadIdx.x;
  const int c = tid & 63, s = tid >> 6;
  int nk = 0;
  for (int i0 = 0; i0 < n; i0 += 64) {
    if (tid < 64) {
      sm.diag[tid] = 0ull;
      if (i0 + tid < n) {
        uint32_t idx = key_index(keys[i0 + tid]);
        sm.cidx[tid] = idx;
        sm.cbox[tid] = load_box(box_base, box_stride, idx);
      }
      if (tid < 2) sm.sup[tid] = 0u;
    }
    __syncthreads();
    const bool valid = (i0 + c) < n;
    float4 mb = make_float4(0.f, 0.f, 0.f, 0.f);
    float ma = 0.f;
    if (valid) {
      mb = sm.cbox[c];
      ma = area_p1(mb);
      bool sup = false;
      for (int k = s; k < nk; k += 8) {               // against boxes kept in earlier chunks
        float4 kb = k < kKeptSmem ? sm.kept[k] : kept_spill[k];
        if (iou_suppresses(kb, area_p1(kb), mb, ma, thresh, on_equal)) { sup = true; break; }
      }
      if (sup) atomicOr(&sm.sup[c >> 5], 1u << (c & 31));
      uint64_t bits = 0ull;                            // intra-chunk: who would (c) suppress
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        int j = s * 8 + jj;
        if (j > c && i0 + j < n) {
          float4 jb = sm.cbox[j];
          if (iou_suppresses(mb, ma, jb, area_p1(jb), thresh, on_equal)) bits |= 1ull << j;
        }
      }
      if (bits) atomicOr(reinterpret_cast<unsigned long long*>(&sm.diag[c]), (unsigned long long)bits);
    }
    __syncthreads();
    if (tid == 0) {
      uint64_t cur = ((uint64_t)sm.sup[1] << 32) | sm.sup[0];
      int m = n - i0;
      if (m < 64) cur |= ~0ull << m;
      uint64_t km = 0ull;
      for (int bit = 0; bit < 64 && cur != ~0ull; ++bit)
        if (!((cur >> bit) & 1ull)) { km |= 1ull << bit; cur |= sm.diag[bit]; }
      sm.keptmask = km;
    }
    __syncthreads();
    const uint64_t km = sm.keptmask;
    if (s == 0 && ((km >> c) & 1ull)) {
      int pos = nk + __popcll(km & ((1ull << c) - 1ull));
      if (pos < kKeptSmem) sm.kept[pos] = mb; else kept_spill[pos] = mb;
      uint32_t idx = sm.cidx[c];
      kept_idx[pos] = (int)idx;
      if (kept_score) kept_score[pos] = score_base[(size_t)idx * score_stride];
    }
    nk += __popcll(km);
    __syncthreads();
  }
  return nk;
}

// ------------------------------------------------------------------------------------------------
// cpu_soft_nms (cpu_nms.pyx:70-163) by one CTA, exact positional semantics.
// State: box[pos], sc[pos], tag[pos] for pos < N.  Each outer iteration i: argmax over [i,N)
// (first maximum), swap into i, decay everything after i, then the reference's swap-with-last
// removal, which is equivalent to: holes (removed pos < N_final, ascending) are filled by the
// survivors at pos >= N_final taken in descending position order.
// The sub-expressions that Cython evaluates in double (literal "1" is emitted as 1.0) are
// evaluated in double here too.
// ------------------------------------------------------------------------------------------------
__device__ int block_soft_nms(float4* box, float* sc, int* tag, int* tmp, int n, float sigma, float Nt,
                              float threshold, unsigned method, int* scan_scratch) {
  __shared__ float s_best[32];
  __shared__ int s_bpos[32];
  __shared__ int s_maxpos;
  __shared__ int s_nrem;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  int N = n;
  for (int i = 0; i < N; ++i) {
    // 1. first maximum over [i, N)
    float best = -INFINITY; int bpos = 0x7fffffff;
    for (int pos = i + tid; pos < N; pos += T) {
      float v = sc[pos];
      if (bpos == 0x7fffffff || best < v) { best = v; bpos = pos; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int op = __shfl_xor_sync(0xffffffffu, bpos, o);
      if (op != 0x7fffffff && (bpos == 0x7fffffff || ov > best || (ov == best && op < bpos))) { best = ov; bpos = op; }
    }
    if (lane == 0) { s_best[warp] = best; s_bpos[warp] = bpos; }
    __syncthreads();
    if (warp == 0) {
      int nw = T >> 5;
      best = lane < nw ? s_best[lane] : -INFINITY;
      bpos = lane < nw ? s_bpos[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int op = __shfl_xor_sync(0xffffffffu, bpos, o);
        if (op != 0x7fffffff && (bpos == 0x7fffffff || ov > best || (ov == best && op < bpos))) { best = ov; bpos = op; }
      }
      if (lane == 0) {
        // the reference starts from maxscore = boxes[i,4] and only moves on a strictly larger score
        int mp = (sc[i] < best) ? bpos : i;
        s_maxpos = mp;
        if (mp != i) {
          float4 tb = box[i]; box[i] = box[mp]; box[mp] = tb;
          float ts = sc[i]; sc[i] = sc[mp]; sc[mp] = ts;
          if (tag) { int tt = tag[i]; tag[i] = tag[mp]; tag[mp] = tt; }
        }
        s_nrem = 0;
      }
    }
    __syncthreads();
    // 2. decay (i, N)
    const float4 t = box[i];
    int nrem_local = 0;
    for (int pos = i + 1 + tid; pos < N; pos += T) {
      float4 b = box[pos];
      int rem = 0;
      float area = (float)(((double)__fsub_rn(b.z, b.x) + 1.0) * ((double)__fsub_rn(b.w, b.y) + 1.0));
      float iw = (float)((double)__fsub_rn(fminf(t.z, b.z), fmaxf(t.x, b.x)) + 1.0);
      if (iw > 0.f) {
        float ih = (float)((double)__fsub_rn(fminf(t.w, b.w), fmaxf(t.y, b.y)) + 1.0);
        if (ih > 0.f) {
          float iwh = __fmul_rn(iw, ih);
          float ua = (float)(((((double)__fsub_rn(t.z, t.x) + 1.0) * ((double)__fsub_rn(t.w, t.y) + 1.0)) + (double)area)
                             - (double)iwh);
          float ov = __fdiv_rn(iwh, ua);
          float weight;
          if (method == 1) weight = ov > Nt ? __fsub_rn(1.0f, ov) : 1.0f;
          else if (method == 2) weight = (float)exp((double)__fdiv_rn(-__fmul_rn(ov, ov), sigma));
          else weight = ov > Nt ? 0.0f : 1.0f;
          float ns = __fmul_rn(weight, sc[pos]);
          sc[pos] = ns;
          if (ns < threshold) rem = 1;
        }
      }
      tmp[pos] = rem;                      // removal flag
      nrem_local += rem;
    }
    if (nrem_local) atomicAdd(&s_nrem, nrem_local);
    __syncthreads();
    const int nrem = s_nrem;
    if (nrem > 0) {
      // 3. swap-with-last compaction.  R(pos) = #removed in (i, pos).
      const int Nf = N - nrem;
      int running = 0;
      // pass A: record hole positions by rank, and filler -> hole rank (stored negative-coded)
      for (int base = i + 1; base < N; base += T) {
        int pos = base + tid;
        int rem = (pos < N) ? tmp[pos] : 0;
        int tot;
        int R = running + block_exclusive_scan(rem, scan_scratch, &tot);
        running += tot;
        if (pos < N) {
          if (rem && pos < Nf) tmp[N + R] = pos;                 // hole of rank R (needs tmp capacity 2N)
        }
      }
      __syncthreads();
      running = 0;
      for (int base = i + 1; base < N; base += T) {
        int pos = base + tid;
        int rem = (pos < N) ? tmp[pos] : 0;
        int tot;
        int R = running + block_exclusive_scan(rem, scan_scratch, &tot);
        running += tot;
        if (pos < N && pos >= Nf && !rem) {
          int rank_desc = (N - 1 - pos) - (nrem - R);
          int dst = tmp[N + rank_desc];
          box[dst] = box[pos]; sc[dst] = sc[pos];
          if (tag) tag[dst] = tag[pos];
        }
      }
      N = Nf;
      __syncthreads();
    }
  }
  return N;
}

// ------------------------------------------------------------------------------------------------
// cpu_soft_nms for lists of <= kSoftSmall candidates: the same algorithm and arithmetic as block_soft_nms above, with the
// whole state (box / score / prior per position) in SHARED memory, a small CTA (128 threads: block barriers cost tens of
// ns instead of the ~0.3 us of 512 threads going through global memory) and the swap-with-last compaction done by ONE warp
// from the short list of positions removed in this iteration (ballots instead of two block-wide scans).  The outer loop
// is inherently sequential (one iteration per surviving box), so the time of a class is iterations x latency: this cuts
// the latency of an iteration ~6x.  Longer lists keep the generic kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kSoftSmall = 2048;
constexpr int kSoftThreads = 128;     // lists <= split1; longer ones get 256, lists > split2 kSoftThreadsBig: the decay pass is (N / threads) deep
constexpr int kSoftThreadsBig = 512;
constexpr int kSoftSplit = 384, kSoftSplit2 = 1024;
struct SoftSmem {
  uint64_t keys[kSoftSmall];
  float4 box[kSoftSmall];
  float sc[kSoftSmall];
  float area[kSoftSmall];              // (x2 - x1 + 1)(y2 - y1 + 1) of the box at each position, evaluated once (it travels with the box)
  int tag[kSoftSmall];
  int removed[kSoftSmall];             // positions removed in the current iteration (unordered)
  int holes[32];
  float best[16];
  int bpos[16];
  int nrem, maxpos;
  int scan[40];
};

// first maximum (value, position) of two candidates: larger value wins, ties go to the smaller position
__device__ __forceinline__ void first_max(float& best, int& bpos, float ov, int op) {
  if (op != 0x7fffffff && (bpos == 0x7fffffff || ov > best || (ov == best && op < bpos))) { best = ov; bpos = op; }
}

// the same over a warp in two REDUX instructions instead of five shuffle rounds: scores map to unsigned keys that order like the
// floats (0 = no candidate), the maximum key is reduced, then the minimum position among the lanes that hold it.  All lanes
// receive the result.
__device__ __forceinline__ unsigned ordered_key(float v) {
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ void warp_first_max(float& best, int& bpos) {
  const unsigned key = bpos == 0x7fffffff ? 0u : ordered_key(best);
  const unsigned m = __reduce_max_sync(0xffffffffu, key);
  const unsigned c = (m != 0u && key == m) ? (unsigned)bpos : 0x7fffffffu;
  bpos = (int)__reduce_min_sync(0xffffffffu, c);
  best = m == 0u ? -INFINITY : __uint_as_float((m & 0x80000000u) ? (m & 0x7fffffffu) : ~m);
}

__device__ int block_soft_nms_small(SoftSmem& sm, int n, float sigma, float Nt, float threshold, unsigned method, int T) {
  // T = participating threads (the first T of the CTA; the others have exited)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
  float4* box = sm.box; float* sc = sm.sc; int* tag = sm.tag; float* ar = sm.area;
  int N = n;
  // partial first-maxima of this thread's strided positions in [lo, N) -> sm.best / sm.bpos per warp
  auto scan_max = [&](int lo) {
    float best = -INFINITY; int bpos = 0x7fffffff;
    for (int pos = lo + tid; pos < N; pos += T) {
      const float v = sc[pos];
      if (bpos == 0x7fffffff || best < v) { best = v; bpos = pos; }
    }
    warp_first_max(best, bpos);
    if (lane == 0) { sm.best[warp] = best; sm.bpos[warp] = bpos; }
  };
  // warp 0: combine the warps' partial maxima (all its lanes get them) ...
  auto combine = [&](float& best, int& bpos) {
    best = lane < nw ? sm.best[lane] : -INFINITY;
    bpos = lane < nw ? sm.bpos[lane] : 0x7fffffff;
    warp_first_max(best, bpos);
  };
  // ... and bring the winner over [i, N) to position i (the reference starts from maxscore = boxes[i,4] and only moves on a
  // strictly larger score)
  auto swap_in = [&](int i, float best, int bpos) {
    if (lane == 0) {
      if (i < N) {
        const int mp = (bpos != 0x7fffffff && sc[i] < best) ? bpos : i;
        if (mp != i) {
          const float4 tb = box[i]; box[i] = box[mp]; box[mp] = tb;
          const float ts = sc[i]; sc[i] = sc[mp]; sc[mp] = ts;
          const int tt = tag[i]; tag[i] = tag[mp]; tag[mp] = tt;
          const float ta = ar[i]; ar[i] = ar[mp]; ar[mp] = ta;
        }
      }
      sm.nrem = 0;
    }
  };
  auto select = [&](int i) {
    if (warp == 0) {
      float best; int bpos;
      combine(best, bpos);
      swap_in(i, best, bpos);
    }
  };
  scan_max(0);
  __syncthreads();
  select(0);
  __syncthreads();
  for (int i = 0; i < N; ++i) {
    // decay (i, N) against box i; removed positions are appended to sm.removed; the partial first-maxima of the UPDATED
    // scores over (i, N) — the next iteration's selection — are gathered in the same pass
    const float4 t = box[i];
    float best = -INFINITY; int bpos = 0x7fffffff;
    // one position: overlap test first (the reference computes the candidate's area before it, which has no effect when
    // the boxes are disjoint): iw = (float)((double)(min - max) + 1.0) > 0  <=>  min - max > -1 exactly
    // Arithmetic as cpu_nms.pyx:117-143 evaluates it (Cython promotes "+ 1" to double and rounds the assignment to float):
    // iw = (float)((double)dw + 1.0) equals the single-precision sum exactly (a correctly rounded sum of two floats is
    // immune to double rounding, 53 >= 2 * 24 + 2); the box areas are compound double expressions and come from sm.area.
    const double t_area = ((double)__fsub_rn(t.z, t.x) + 1.0) * ((double)__fsub_rn(t.w, t.y) + 1.0);
    auto decay_one = [&](int pos, const float4 b, float v) {
      const float dw = __fsub_rn(fminf(t.z, b.z), fmaxf(t.x, b.x)), dh = __fsub_rn(fminf(t.w, b.w), fmaxf(t.y, b.y));
      if (dw > -1.0f && dh > -1.0f) {
        const float iw = __fadd_rn(dw, 1.0f), ih = __fadd_rn(dh, 1.0f);
        if (iw > 0.f && ih > 0.f) {
          const float iwh = __fmul_rn(iw, ih);
          const float ua = (float)((t_area + (double)ar[pos]) - (double)iwh);
          const float ov = __fdiv_rn(iwh, ua);
          float weight;
          if (method == 1) weight = ov > Nt ? __fsub_rn(1.0f, ov) : 1.0f;
          else if (method == 2) weight = (float)exp((double)__fdiv_rn(-__fmul_rn(ov, ov), sigma));
          else weight = ov > Nt ? 0.0f : 1.0f;
          v = __fmul_rn(weight, v);
          sc[pos] = v;
          if (v < threshold) { sm.removed[atomicAdd(&sm.nrem, 1)] = pos; return; }   // leaves the list: not a candidate for the next maximum
        }
      }
      if (bpos == 0x7fffffff || best < v) { best = v; bpos = pos; }
    };
    int pos = i + 1 + tid;
    for (; pos + 3 * T < N; pos += 4 * T) {                    // four independent loads in flight: the pass is latency bound
      const float4 b0 = box[pos], b1 = box[pos + T], b2 = box[pos + 2 * T], b3 = box[pos + 3 * T];
      const float v0 = sc[pos], v1 = sc[pos + T], v2 = sc[pos + 2 * T], v3 = sc[pos + 3 * T];
      decay_one(pos, b0, v0); decay_one(pos + T, b1, v1); decay_one(pos + 2 * T, b2, v2); decay_one(pos + 3 * T, b3, v3);
    }
    for (; pos < N; pos += T) decay_one(pos, box[pos], sc[pos]);
    warp_first_max(best, bpos);
    if (lane == 0) { sm.best[warp] = best; sm.bpos[warp] = bpos; }
    __syncthreads();
    const int nrem = sm.nrem;
    if (nrem > 0 && nrem <= 32) {
      // The usual case, entirely in warp 0 between the same two barriers as an iteration without removals: swap-with-last
      // compaction (holes = removed positions < Nf, ascending, filled by the survivors at positions >= Nf in descending
      // position order), then the selection for the next iteration WITHOUT a second pass over the scores: removed boxes are
      // below the threshold and every survivor is not, so the maximum gathered above is still the maximum; only its position
      // can change — to the hole a filler with that score moved into.  The first position of the maximum afterwards is the
      // smaller of the old one (if it stayed, i.e. lies below Nf) and the destinations of the fillers that carry it.
      const int Nf = N - nrem;
      if (warp == 0) {
        float vbest; int vpos;
        combine(vbest, vpos);
        const int pr = lane < nrem ? sm.removed[lane] : 0x7fffffff;
        int rank = 0;                                 // rank of this lane's removed position among the holes
        bool taken = false;                           // is position N - 1 - lane (a filler candidate) itself removed?
        const int q = N - 1 - lane;
        for (int k = 0; k < nrem; ++k) {
          const int pk = __shfl_sync(0xffffffffu, pr, k);
          rank += (pk < pr && pk < Nf) ? 1 : 0;
          taken = taken || pk == q;
        }
        if (pr < Nf) sm.holes[rank] = pr;
        __syncwarp();
        const bool filler = lane < nrem && !taken;    // q >= Nf by construction
        const unsigned fm = __ballot_sync(0xffffffffu, filler);
        unsigned moved_to = 0x7fffffffu;
        if (filler) {
          const int dst = sm.holes[__popc(fm & ((1u << lane) - 1u))];
          const float qs = sc[q];
          box[dst] = box[q]; sc[dst] = qs; tag[dst] = tag[q]; ar[dst] = ar[q];
          if (qs == vbest) moved_to = (unsigned)dst;
        }
        const unsigned first_moved = __reduce_min_sync(0xffffffffu, moved_to);
        const unsigned stay = vpos < Nf ? (unsigned)vpos : 0x7fffffffu;
        vpos = (int)min(stay, first_moved);
        N = Nf;
        __syncwarp();
        swap_in(i + 1, vbest, vpos);
      }
      N = Nf;
      __syncthreads();
      continue;
    }
    if (nrem > 0) {
      // more than 32 removals at once (rare): the same compaction with flags + two block-wide scans, as in the generic kernel;
      // positions move, so the maxima are gathered again afterwards
      const int Nf = N - nrem;
      {
        int* flag = reinterpret_cast<int*>(sm.keys);            // the sort buffer is free by now: [0, N) flags, [N, 2N) holes
        for (int pos = i + 1 + tid; pos < N; pos += T) flag[pos] = 0;
        __syncthreads();
        for (int k = tid; k < nrem; k += T) flag[sm.removed[k]] = 1;
        __syncthreads();
        int running = 0;
        for (int base = i + 1; base < N; base += T) {
          const int pos = base + tid;
          const int rem = (pos < N) ? flag[pos] : 0;
          int tot;
          const int R = running + block_exclusive_scan_n(rem, sm.scan, &tot, nw);
          running += tot;
          if (pos < N && rem && pos < Nf) flag[N + R] = pos;
        }
        __syncthreads();
        running = 0;
        for (int base = i + 1; base < N; base += T) {
          const int pos = base + tid;
          const int rem = (pos < N) ? flag[pos] : 0;
          int tot;
          const int R = running + block_exclusive_scan_n(rem, sm.scan, &tot, nw);
          running += tot;
          if (pos < N && pos >= Nf && !rem) {
            const int dst = flag[N + (N - 1 - pos) - (nrem - R)];
            box[dst] = box[pos]; sc[dst] = sc[pos]; tag[dst] = tag[pos]; ar[dst] = ar[pos];
          }
        }
      }
      N = Nf;
      __syncthreads();
      scan_max(i + 1);
      __syncthreads();
    }
    select(i + 1);
    __syncthreads();
  }
  return N;
}

__global__ void __launch_bounds__(kSoftThreadsBig)
class_soft_nms_small_kernel(const float* __restrict__ conf, const float* __restrict__ obj, const float4* boxes_px,
                            const int* __restrict__ cand_count, const uint64_t* cand_keys, int key_stride, int P, int C,
                            float thresh, int method, float sigma, float soft_threshold,
                            float4* spill, int* kept_prior, float* kept_score, int* kept_count, int split1, int split2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SoftSmem& sm = *reinterpret_cast<SoftSmem*>(smem_raw);
  const int j = blockIdx.x, b = blockIdx.y;
  const size_t row = (size_t)b * C + j;
  const int n = cand_count[row];
  if (n > kSoftSmall) return;                                   // the generic kernel takes long lists
  if (n == 0) { if (threadIdx.x == 0) kept_count[row] = 0; return; }
  // short lists run on the first kSoftThreads threads only (cheaper barriers and reductions); the rest of the CTA leaves
  const int T = n > split2 ? kSoftThreadsBig : (n > split1 ? 256 : kSoftThreads);
  if ((int)threadIdx.x >= T) return;
  const float4* bpx = boxes_px + (size_t)b * P;
  // candidate order = ascending prior index (np.where order, test.py:143): sort keys (0xFFFFFFFF - prior) descending
  const uint64_t* gk = cand_keys + row * key_stride;
  int npad = 2;
  while (npad < n) npad <<= 1;
  for (int i = threadIdx.x; i < npad; i += T) sm.keys[i] = i < n ? (uint64_t)(0xFFFFFFFFu - key_index(gk[i])) + 1ull : 0ull;
  __syncthreads();
  for (int k = 2; k <= npad; k <<= 1)
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int t = threadIdx.x; t < (npad >> 1); t += T) {
        const int i = 2 * t - (t & (jj - 1));
        cmpxchg_desc(sm.keys, i, i + jj, (i & k) == 0);
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += T) {
    const int p = (int)(0xFFFFFFFFu - (uint32_t)(sm.keys[i] - 1ull));
    const size_t bp = (size_t)b * P + p;
    const float4 bx = bpx[p];
    sm.box[i] = bx;
    sm.area[i] = (float)(((double)__fsub_rn(bx.z, bx.x) + 1.0) * ((double)__fsub_rn(bx.w, bx.y) + 1.0));
    sm.sc[i] = __fmul_rn(obj[bp * 2 + 1], conf[bp * C + j]);
    sm.tag[i] = p;
  }
  __syncthreads();
  const int N = block_soft_nms_small(sm, n, sigma, thresh, soft_threshold, method == 3 ? 0u : (unsigned)method, T);
  float4* my_spill = spill + row * P;
  int* my_prior = kept_prior + row * P;
  float* my_score = kept_score + row * P;
  for (int i = threadIdx.x; i < N; i += T) { my_spill[i] = sm.box[i]; my_score[i] = sm.sc[i]; my_prior[i] = sm.tag[i]; }
  if (threadIdx.x == 0) kept_count[row] = N;
}

// ------------------------------------------------------------------------------------------------
// one CTA per (class, image): sort candidates, run NMS, emit the kept (prior, score) list
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNmsThreads)
class_nms_kernel(const float* __restrict__ conf, const float* __restrict__ obj, const float4* boxes_px,
                 const int* __restrict__ cand_count, uint64_t* cand_keys, int key_stride, int P, int C,
                 float thresh, int on_equal, int method, float sigma, float soft_threshold,
                 float4* spill, int* kept_prior, float* kept_score, int* kept_count, int* soft_tmp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NmsSmem& sm = *reinterpret_cast<NmsSmem*>(smem_raw);
  const int j = blockIdx.x, b = blockIdx.y;
  const size_t row = (size_t)b * C + j;
  const int n = cand_count[row];
  const float4* bpx = boxes_px + (size_t)b * P;
  float4* my_spill = spill + row * P;
  int* my_prior = kept_prior + row * P;
  float* my_score = kept_score + row * P;
  if (n == 0) { if (threadIdx.x == 0) kept_count[row] = 0; return; }
  uint64_t* gkeys = cand_keys + row * key_stride;
  if (method == 0) {
    const uint64_t* keys = block_sort_desc(gkeys, n, sm.keys);
    // score of prior p for class j is recomputed exactly as at selection time
    // (score_base trick: kept_score is filled below from the key order instead)
    int nk = block_hard_nms(keys, n, reinterpret_cast<const float*>(bpx), 4, nullptr, 0, thresh, on_equal != 0,
                            my_spill, my_prior, nullptr, sm);
    for (int k = threadIdx.x; k < nk; k += blockDim.x) {
      int p = my_prior[k];
      size_t bp = (size_t)b * P + p;
      my_score[k] = __fmul_rn(obj[bp * 2 + 1], conf[bp * C + j]);
    }
    if (threadIdx.x == 0) kept_count[row] = nk;
  } else {
    if (n <= kSoftSmall) return;              // class_soft_nms_small_kernel handles short lists
    // soft-NMS runs in candidate order = ascending prior index (np.where order, test.py:143):
    // rewrite keys so that the descending sort yields ascending prior index.
    for (int i = threadIdx.x; i < n; i += blockDim.x) gkeys[i] = (uint64_t)(0xFFFFFFFFu - key_index(gkeys[i]));
    __syncthreads();
    const uint64_t* keys = block_sort_desc(gkeys, n, sm.keys);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      int p = (int)key_index(keys[i]);
      size_t bp = (size_t)b * P + p;
      my_spill[i] = bpx[p];
      my_score[i] = __fmul_rn(obj[bp * 2 + 1], conf[bp * C + j]);
      my_prior[i] = p;
    }
    __syncthreads();
    int* tmp = soft_tmp + row * (size_t)(2 * P);
    int N = block_soft_nms(my_spill, my_score, my_prior, tmp, n, sigma, thresh, soft_threshold,
                           method == 3 ? 0u : (unsigned)method, sm.scan);
    if (threadIdx.x == 0) kept_count[row] = N;
  }
}

// ------------------------------------------------------------------------------------------------
// per-image cut to max_per_image (test.py:155-161) + record emission (class asc, keep order)
// ------------------------------------------------------------------------------------------------
constexpr int kSelImgThreads = 1024;

__global__ void __launch_bounds__(kSelImgThreads)
image_select_kernel(const float4* __restrict__ boxes_px, const int* __restrict__ kept_prior,
                    const float* __restrict__ kept_score, const int* __restrict__ kept_count, int P, int C,
                    int max_per_image, int max_out, float* __restrict__ records, int* __restrict__ counts,
                    int* __restrict__ prior_idx) {
  __shared__ int s_off[kMaxClasses + 1];
  __shared__ unsigned s_hist[256];
  __shared__ int s_scan[40];
  __shared__ unsigned s_prefix, s_remaining;
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    int acc = 0;
    for (int j = 0; j < C; ++j) { s_off[j] = acc; acc += kept_count[(size_t)b * C + j]; }
    s_off[C] = acc;
  }
  __syncthreads();
  const int total = s_off[C];
  auto locate = [&](int f, int& j, int& k) {       // flattened index -> (class, k)
    int lo = 0, hi = C;                            // largest j with s_off[j] <= f
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_off[mid] <= f) lo = mid; else hi = mid; }
    j = lo; k = f - s_off[lo];
  };
  unsigned thr_bits = 0u;                          // keep score_bits >= thr_bits
  if (max_per_image > 0 && total > max_per_image) {
    // radix select of the max_per_image-th largest score (np.sort(scores)[-max_per_image])
    if (tid == 0) { s_prefix = 0u; s_remaining = (unsigned)max_per_image; }
    for (int pass = 3; pass >= 0; --pass) {
      for (int i = tid; i < 256; i += blockDim.x) s_hist[i] = 0u;
      __syncthreads();
      const unsigned prefix = s_prefix;
      const unsigned himask = pass == 3 ? 0u : (0xFFFFFFFFu << ((pass + 1) * 8));
      for (int f = tid; f < total; f += blockDim.x) {
        int j, k; locate(f, j, k);
        unsigned bits = float_order_bits(kept_score[((size_t)b * C + j) * P + k]);
        if ((bits & himask) == prefix) atomicAdd(&s_hist[(bits >> (pass * 8)) & 255u], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        unsigned rem = s_remaining, acc = 0u; int bkt = 255;
        for (; bkt >= 0; --bkt) { if (acc + s_hist[bkt] >= rem) break; acc += s_hist[bkt]; }
        s_remaining = rem - acc;
        s_prefix = prefix | ((unsigned)bkt << (pass * 8));
      }
      __syncthreads();
    }
    thr_bits = s_prefix;
  }
  int running = 0;
  for (int base = 0; base < total; base += blockDim.x) {
    int f = base + tid;
    int j = 0, k = 0, flag = 0;
    float sc = 0.f; int p = 0;
    if (f < total) {
      locate(f, j, k);
      size_t at = ((size_t)b * C + j) * P + k;
      sc = kept_score[at]; p = kept_prior[at];
      flag = float_order_bits(sc) >= thr_bits;
    }
    int tot;
    int pos = running + block_exclusive_scan(flag, s_scan, &tot);
    running += tot;
    if (flag && pos < max_out) {
      float4 bx = boxes_px[(size_t)b * P + p];
      float* r = records + ((size_t)b * max_out + pos) * 6;
      r[0] = bx.x; r[1] = bx.y; r[2] = bx.z; r[3] = bx.w; r[4] = sc; r[5] = (float)(j + 1);
      if (prior_idx) prior_idx[(size_t)b * max_out + pos] = p;
    }
  }
  if (tid == 0) counts[b] = running;
}

// ------------------------------------------------------------------------------------------------
// stand-alone NMS over dets[n,5]
// ------------------------------------------------------------------------------------------------
__global__ void dets_keys_kernel(const float* __restrict__ dets, int n, int presorted, uint64_t* keys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = presorted ? (((uint64_t)(uint32_t)(n - i) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)i))
                                 : make_key(dets[(size_t)i * 5 + 4], (uint32_t)i);
}

__global__ void __launch_bounds__(kNmsThreads)
dets_nms_kernel(const float* dets, int n, int presorted, uint64_t* gkeys, float thresh, int on_equal,
                float4* spill, int* keep_out, int* num_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NmsSmem& sm = *reinterpret_cast<NmsSmem*>(smem_raw);
  const uint64_t* keys = gkeys;
  if (!presorted) keys = block_sort_desc(gkeys, n, sm.keys);
  int nk = block_hard_nms(keys, n, dets, 5, nullptr, 0, thresh, on_equal != 0, spill, keep_out, nullptr, sm);
  if (threadIdx.x == 0) *num_out = nk;
}

__global__ void __launch_bounds__(kNmsThreads)
dets_soft_nms_kernel(float* dets, int n, float sigma, float Nt, float threshold, unsigned method,
                     float4* box, float* sc, int* tmp, int* n_out) {
  __shared__ int scan[40];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float* d = dets + (size_t)i * 5;
    box[i] = make_float4(d[0], d[1], d[2], d[3]);
    sc[i] = d[4];
  }
  __syncthreads();
  int N = block_soft_nms(box, sc, nullptr, tmp, n, sigma, Nt, threshold, method, scan);
  __syncthreads();
  // the reference leaves rows >= N in an unspecified (partially overwritten) state; we rewrite the
  // first N rows only, which is all callers may read (keep = range(N)).
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float* d = dets + (size_t)i * 5;
    float4 b = box[i];
    d[0] = b.x; d[1] = b.y; d[2] = b.z; d[3] = b.w; d[4] = sc[i];
  }
  if (threadIdx.x == 0) *n_out = N;
}

struct PostWs {
  float4* boxes_px; int* cand_count; int* kept_count; uint64_t* cand_keys; int key_stride;
  int* kept_prior; float* kept_score; float4* spill; int* soft_tmp; size_t total;
};
static PostWs carve_post_ws(void* base, int B, int P, int C) {
  PostWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (char*)base + o; };
  w.key_stride = next_pow2(P);
  w.boxes_px = (float4*)take(sizeof(float4) * (size_t)B * P);
  w.cand_count = (int*)take(sizeof(int) * (size_t)B * C * 2);
  w.kept_count = w.cand_count + (size_t)B * C;
  w.cand_keys = (uint64_t*)take(sizeof(uint64_t) * (size_t)B * C * w.key_stride);
  w.kept_prior = (int*)take(sizeof(int) * (size_t)B * C * P);
  w.kept_score = (float*)take(sizeof(float) * (size_t)B * C * P);
  w.spill = (float4*)take(sizeof(float4) * (size_t)B * C * P);
  w.soft_tmp = (int*)take(sizeof(int) * (size_t)B * C * P * 2);
  w.total = off;
  return w;
}

}  // namespace ctx

using namespace ctx;

extern "C" int ctx_detect_forward(const float* loc, const float* conf, const float* obj, const float* priors,
                                  int batch, int num_priors, int num_fg_classes, float var0, float var1,
                                  float* boxes_out, float* scores_out, void* stream) {
  CTX_REQUIRE(loc && conf && obj && priors && boxes_out && scores_out, "ctx_detect_forward: null pointer");
  CTX_REQUIRE(batch >= 0 && num_priors >= 0 && num_fg_classes >= 1, "ctx_detect_forward: bad sizes");
  if (batch == 0 || num_priors == 0) return CTX_OK;
  long long work = (long long)batch * num_priors * (num_fg_classes + 1);
  int blocks = (int)std::min<long long>((work + 255) / 256, 148LL * 16);
  detect_forward_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(loc, conf, obj, priors, batch, num_priors,
                                                                  num_fg_classes, var0, var1, boxes_out, scores_out);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

extern "C" size_t ctx_postprocess_workspace_bytes(int batch, int num_priors, int num_fg_classes) {
  if (batch <= 0 || num_priors <= 0 || num_fg_classes <= 0) return 256;
  return carve_post_ws(nullptr, batch, num_priors, num_fg_classes).total;
}

extern "C" int ctx_detect_postprocess(const float* loc, const float* conf, const float* obj, const float* priors,
                                      const float* scale, const CtxPostParams* p, float* records, int* counts,
                                      int* prior_idx, void* workspace, size_t workspace_bytes, void* stream) {
  CTX_REQUIRE(p, "ctx_detect_postprocess: null params");
  CTX_REQUIRE(loc && conf && obj && priors && scale && records && counts, "ctx_detect_postprocess: null pointer");
  const int B = p->batch, P = p->num_priors, C = p->num_fg_classes;
  CTX_REQUIRE(B >= 0 && P >= 1 && C >= 1 && C <= kMaxClasses, "ctx_detect_postprocess: bad sizes (C <= %d)", kMaxClasses);
  CTX_REQUIRE(p->max_out >= 1, "ctx_detect_postprocess: max_out must be >= 1");
  CTX_REQUIRE(p->nms_method >= 0 && p->nms_method <= 3, "ctx_detect_postprocess: nms_method must be 0..3");
  if (B == 0) return CTX_OK;
  PostWs w = carve_post_ws(workspace, B, P, C);
  if (!workspace || workspace_bytes < w.total) {
    set_error("ctx_detect_postprocess: workspace %zu < required %zu", workspace_bytes, w.total);
    return CTX_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  CTX_CUDA_TRY(cudaMemsetAsync(w.cand_count, 0, sizeof(int) * (size_t)B * C * 2, st));
  dim3 g1(cdiv(P, kSelThreads), B);
  select_candidates_kernel<<<g1, kSelThreads, 0, st>>>(loc, conf, obj, priors, scale, p->scale_per_image, P, C,
                                                       p->var0, p->var1, p->score_thresh, w.boxes_px, w.cand_count,
                                                       w.cand_keys, w.key_stride);
  CTX_LAUNCH_CHECK();
  CTX_CUDA_TRY(cudaFuncSetAttribute(class_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NmsSmem)));
  if (p->nms_method != 0) {
    CTX_CUDA_TRY(cudaFuncSetAttribute(class_soft_nms_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SoftSmem)));
    static const int split1 = [] { const char* e = getenv("CTX_SOFT_SPLIT1"); return e ? atoi(e) : kSoftSplit; }();
    static const int split2 = [] { const char* e = getenv("CTX_SOFT_SPLIT2"); return e ? atoi(e) : kSoftSplit2; }();
    class_soft_nms_small_kernel<<<dim3(C, B), kSoftThreadsBig, sizeof(SoftSmem), st>>>(
        conf, obj, w.boxes_px, w.cand_count, w.cand_keys, w.key_stride, P, C, p->nms_thresh, p->nms_method, p->soft_sigma,
        p->soft_threshold, w.spill, w.kept_prior, w.kept_score, w.kept_count, split1, split2);
    CTX_LAUNCH_CHECK();
  }
  class_nms_kernel<<<dim3(C, B), kNmsThreads, sizeof(NmsSmem), st>>>(
      conf, obj, w.boxes_px, w.cand_count, w.cand_keys, w.key_stride, P, C, p->nms_thresh, p->suppress_on_equal,
      p->nms_method, p->soft_sigma, p->soft_threshold, w.spill, w.kept_prior, w.kept_score, w.kept_count, w.soft_tmp);
  CTX_LAUNCH_CHECK();
  image_select_kernel<<<B, kSelImgThreads, 0, st>>>(w.boxes_px, w.kept_prior, w.kept_score, w.kept_count, P, C,
                                                    p->max_per_image, p->max_out, records, counts, prior_idx);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

// ---- stand-alone NMS ---------------------------------------------------------------------------
namespace {
struct NmsWs { uint64_t* keys; float4* spill; float* sc; int* tmp; size_t total; };
NmsWs carve_nms_ws(void* base, int n) {
  NmsWs w; size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = ctx::align_up(off + bytes, 256); return (char*)base + o; };
  int cap = next_pow2(n > 1 ? n : 2);
  w.keys = (uint64_t*)take(sizeof(uint64_t) * cap);
  w.spill = (float4*)take(sizeof(float4) * (size_t)cap);
  w.sc = (float*)take(sizeof(float) * (size_t)cap);
  w.tmp = (int*)take(sizeof(int) * (size_t)cap * 2);
  w.total = off;
  return w;
}
int nms_device_impl(const float* dets, int n, float thresh, int on_equal, int presorted, int* keep_out, int* num_out,
                    void* workspace, size_t workspace_bytes, cudaStream_t st) {
  NmsWs w = carve_nms_ws(workspace, n);
  if (!workspace || workspace_bytes < w.total) {
    set_error("ctx_nms: workspace %zu < required %zu", workspace_bytes, w.total);
    return CTX_ERR_WORKSPACE;
  }
  if (n == 0) { CTX_CUDA_TRY(cudaMemsetAsync(num_out, 0, sizeof(int), st)); return CTX_OK; }
  dets_keys_kernel<<<cdiv(n, 256), 256, 0, st>>>(dets, n, presorted, w.keys);
  CTX_LAUNCH_CHECK();
  CTX_CUDA_TRY(cudaFuncSetAttribute(dets_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NmsSmem)));
  dets_nms_kernel<<<1, kNmsThreads, sizeof(NmsSmem), st>>>(dets, n, presorted, w.keys, thresh, on_equal, w.spill,
                                                           keep_out, num_out);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}
int nms_host_impl(const float* dets_host, int n, int dim, float thresh, int on_equal, int presorted, int* keep_host,
                  int* num_out_host, int device_id) {
  CTX_REQUIRE(n >= 0 && num_out_host, "ctx_nms_host: bad arguments");
  *num_out_host = 0;
  if (n == 0) return CTX_OK;
  CTX_REQUIRE(dets_host && keep_host, "ctx_nms_host: null pointer");
  CTX_REQUIRE(dim >= 4, "ctx_nms_host: boxes_dim must be >= 4");
  int cur = -1;
  CTX_CUDA_TRY(cudaGetDevice(&cur));
  if (cur != device_id) CTX_CUDA_TRY(cudaSetDevice(device_id));
  size_t wsb = ctx_nms_workspace_bytes(n);
  char* dev = nullptr;
  size_t dets_b = align_up(sizeof(float) * (size_t)n * 5, 256), keep_b = align_up(sizeof(int) * (size_t)(n + 1), 256);
  cudaError_t e = cudaMalloc(&dev, dets_b + keep_b + wsb);
  if (e != cudaSuccess) { set_error("ctx_nms_host: cudaMalloc failed: %s", cudaGetErrorString(e)); return CTX_ERR_CUDA; }
  float* d_dets = (float*)dev; int* d_keep = (int*)(dev + dets_b); int* d_num = d_keep + n; void* ws = dev + dets_b + keep_b;
  int rc = CTX_OK;
  // rows may be wider than 5 floats upstream (boxes_dim); repack to [n,5] (score column only matters unsorted)
  float* packed = nullptr;
  const float* src = dets_host;
  if (dim != 5) {
    packed = (float*)malloc(sizeof(float) * (size_t)n * 5);
    for (int i = 0; i < n; ++i) {
      for (int c = 0; c < 4; ++c) packed[i * 5 + c] = dets_host[(size_t)i * dim + c];
      packed[i * 5 + 4] = dim > 4 ? dets_host[(size_t)i * dim + 4] : 0.f;
    }
    src = packed;
  }
  cudaStream_t st = 0;
  do {
    e = cudaMemcpyAsync(d_dets, src, sizeof(float) * (size_t)n * 5, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { set_error("ctx_nms_host: H2D failed: %s", cudaGetErrorString(e)); rc = CTX_ERR_CUDA; break; }
    rc = nms_device_impl(d_dets, n, thresh, on_equal, presorted, d_keep, d_num, ws, wsb, st);
    if (rc) break;
    e = cudaMemcpyAsync(num_out_host, d_num, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("ctx_nms_host: D2H failed: %s", cudaGetErrorString(e)); rc = CTX_ERR_CUDA; break; }
    e = cudaMemcpy(keep_host, d_keep, sizeof(int) * (size_t)(*num_out_host), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("ctx_nms_host: D2H failed: %s", cudaGetErrorString(e)); rc = CTX_ERR_CUDA; break; }
  } while (0);
  free(packed);
  cudaFree(dev);
  return rc;
}
}  // namespace

extern "C" size_t ctx_nms_workspace_bytes(int n) { return carve_nms_ws(nullptr, n < 0 ? 0 : n).total; }

extern "C" int ctx_nms_device(const float* dets, int n, float thresh, int suppress_on_equal, int* keep_out, int* num_out,
                              void* workspace, size_t workspace_bytes, void* stream) {
  CTX_REQUIRE(n >= 0 && num_out, "ctx_nms_device: bad arguments");
  CTX_REQUIRE(n == 0 || (dets && keep_out), "ctx_nms_device: null pointer");
  return nms_device_impl(dets, n, thresh, suppress_on_equal, 0, keep_out, num_out, workspace, workspace_bytes,
                         (cudaStream_t)stream);
}

extern "C" int ctx_nms_host(const float* dets_host, int n, float thresh, int suppress_on_equal, int* keep_host,
                            int* num_out_host, int device_id) {
  return nms_host_impl(dets_host, n, 5, thresh, suppress_on_equal, 0, keep_host, num_out_host, device_id);
}

extern "C" void _nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
                     float nms_overlap_thresh, int device_id) {
  int n = 0;
  // gpu_nms convention: rows pre-sorted by the caller, suppress on ">" (nms_kernel.cu:71)
  int rc = nms_host_impl(boxes_host, boxes_num, boxes_dim, nms_overlap_thresh, 0, 1, keep_out, &n, device_id);
  if (num_out) *num_out = rc == CTX_OK ? n : 0;
}

extern "C" int ctx_soft_nms_host(float* boxes_host, int n, float sigma, float Nt, float threshold, unsigned method,
                                 int* n_out, int device_id) {
  CTX_REQUIRE(n >= 0 && n_out, "ctx_soft_nms_host: bad arguments");
  *n_out = 0;
  if (n == 0) return CTX_OK;
  CTX_REQUIRE(boxes_host, "ctx_soft_nms_host: null pointer");
  int cur = -1;
  CTX_CUDA_TRY(cudaGetDevice(&cur));
  if (cur != device_id) CTX_CUDA_TRY(cudaSetDevice(device_id));
  NmsWs w = carve_nms_ws(nullptr, n);
  size_t dets_b = align_up(sizeof(float) * (size_t)n * 5 + sizeof(int), 256);
  char* dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, dets_b + w.total);
  if (e != cudaSuccess) { set_error("ctx_soft_nms_host: cudaMalloc failed: %s", cudaGetErrorString(e)); return CTX_ERR_CUDA; }
  w = carve_nms_ws(dev + dets_b, n);
  float* d = (float*)dev; int* d_n = (int*)(dev + sizeof(float) * (size_t)n * 5);
  int rc = CTX_OK;
  do {
    e = cudaMemcpy(d, boxes_host, sizeof(float) * (size_t)n * 5, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("ctx_soft_nms_host: H2D failed: %s", cudaGetErrorString(e)); rc = CTX_ERR_CUDA; break; }
    dets_soft_nms_kernel<<<1, kNmsThreads>>>(d, n, sigma, Nt, threshold, method, w.spill, w.sc, w.tmp, d_n);
    count_launch();
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(n_out, d_n, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(boxes_host, d, sizeof(float) * (size_t)(*n_out) * 5, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("ctx_soft_nms_host: kernel/D2H failed: %s", cudaGetErrorString(e)); rc = CTX_ERR_CUDA; break; }
  } while (0);
  cudaFree(dev);
  return rc;
}
