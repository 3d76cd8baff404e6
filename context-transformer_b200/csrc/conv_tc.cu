// conv_tc.cu — implicit-GEMM convolution on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// Every conv of the detector (models/RFB_Net_vgg.py BasicConv :7-22, vgg() :323-343, RFB branches
// :26-112, heads :387-416) with 16-bit activations is one launch of this kernel:
//
//   D[m, co] = sum_{tap, ci} A[m, (tap, ci)] * Wt[co, (tap, ci)]          fp32 accumulate in TMEM
//   m  = flattened output pixel (n, oy, ox) over the WHOLE batch   -> no per-image tile waste
//   A  = the NHWC activation gathered on the fly (im2col never exists in HBM)
//
// CTA = 128 output pixels x BN output channels, K walked in steps of 64 channels of one filter tap.
// Warp roles (192 threads):
//   warps 0-3  im2col producers: each K-step they gather the 128 x 64 activation tile with 16-byte
//              cp.async (zero-fill for padding / channel tail) straight into the 128B-swizzled
//              K-major layout the UMMA descriptor expects (shared-memory im2col staging); after the
//              main loop the same warps run the epilogue: tcgen05.ld the accumulator, + bias
//              (BatchNorm folded) [+ residual] [ReLU], convert, vectorised NHWC store (or up to
//              three fp32 segments for the fused loc/conf/obj heads).
//   warp 4     TMA producer: one elected lane streams the BN x 64 weight tile of each K-step with
//              cp.async.bulk.tensor.2d (SWIZZLE_128B) onto the stage's mbarrier.
//   warp 5     TMEM allocator + MMA issuer: one elected lane issues 4 x tcgen05.mma
//              (M128 x N=BN x K16, kind::f16) per K-step and tcgen05.commit's the stage back to the
//              producers; a last commit hands the accumulator to the epilogue.
// A ring of S stages of {A tile 16 KB, B tile BN*128 B} with full/empty mbarriers decouples the three.
#include "tc_common.cuh"

#include <mutex>

namespace ctx {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 192;
constexpr int TC_A_STAGE = TC_BM * TC_BK * 2;     // 16 KB

struct TcParams {
  const void* in;
  const float* bias;
  const void* residual;
  int N, H, W, Cin, in_cstride, in_coffset;
  int Cout, KH, KW, stride, pad_h, pad_w, dil, Ho, Wo, relu;
  int M, cin_blocks, nk;
  int res_cstride, res_coffset, res_dtype;
  int is_bf16;
  int fast_out;                // single 16-bit segment, 8-channel aligned: vectorised stores
  SegTable segs;
};

// ---------------------------------------------------------------------------------------------------
template <int BN, int S>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const TcParams p) {
  constexpr int B_STAGE = BN * TC_BK * 2;
  constexpr int LOOKAHEAD = S - 2;     // stages whose gather may still be in flight when the next one is issued
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B needs 1024-B alignment
  const uint32_t sA = smem_base, sB = smem_base + S * TC_A_STAGE;
  const uint32_t bars = sB + S * B_STAGE;
  const uint32_t full0 = bars, empty0 = bars + 8 * S, accum_bar = bars + 16 * S, tmem_slot = bars + 16 * S + 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
  const int nk = p.nk;

  if (warp == 4 && lane == 0) tma_prefetch_desc(&tmap_w);
  if (warp == 5) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 4 + 1); mbar_init(empty0 + 8 * s, 1); }
      mbar_init(accum_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    // ================= im2col producers =================
    const int t = threadIdx.x;                 // 0..127
    const int chunk = t & 7;                   // 16-byte (8-channel) chunk of the 64-channel K-step
    const int row0 = t >> 3;                   // rows row0 + 16*i, i = 0..7
    const int HoWo = p.Ho * p.Wo;
    const uint16_t* in = reinterpret_cast<const uint16_t*>(p.in);
    long long base[8];
    uint32_t mask[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + row0 + 16 * i;
      mask[i] = 0u;
      base[i] = 0;
      if (m < p.M) {
        const int n = m / HoWo, rem = m - n * HoWo;
        const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
        const int iy0 = oy * p.stride - p.pad_h, ix0 = ox * p.stride - p.pad_w;
        base[i] = ((long long)(n * p.H + iy0) * p.W + ix0) * p.in_cstride + p.in_coffset + chunk * 8;
        for (int ky = 0; ky < p.KH; ++ky)
          for (int kx = 0; kx < p.KW; ++kx) {
            const int iy = iy0 + ky * p.dil, ix = ix0 + kx * p.dil;
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) mask[i] |= 1u << (ky * p.KW + kx);
          }
      }
    }
    int tap = 0, cc = 0, ky = 0, kx = 0;
    for (int it = 0; it < nk; ++it) {
      const int s = it % S;
      mbar_wait(empty0 + 8 * s, ((it / S) & 1) ^ 1);
      const long long tap_off = (long long)((ky * p.dil) * p.W + kx * p.dil) * p.in_cstride + cc * TC_BK;
      const bool ch_ok = (cc * TC_BK + chunk * 8) < p.Cin;
      const uint32_t dst_stage = sA + s * TC_A_STAGE;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = row0 + 16 * i;
        const bool ok = ch_ok && ((mask[i] >> tap) & 1u);
        const void* src = ok ? (const void*)(in + base[i] + tap_off) : (const void*)in;
        cp_async_16(dst_stage + r * 128 + ((chunk ^ (r & 7)) << 4), src, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (it >= LOOKAHEAD) {
        cp_async_wait<LOOKAHEAD>();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * ((it - LOOKAHEAD) % S));
      }
      if (++cc == p.cin_blocks) { cc = 0; ++tap; if (++kx == p.KW) { kx = 0; ++ky; } }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0)
      for (int it = (nk > LOOKAHEAD ? nk - LOOKAHEAD : 0); it < nk; ++it) mbar_arrive(full0 + 8 * (it % S));

    // ================= epilogue =================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int r = warp * 32 + lane;
    const int m = m0 + r;
    const bool row_ok = m < p.M;
    const int n_img = row_ok ? m / HoWo : 0;
    const int pix = row_ok ? m - n_img * HoWo : 0;
    const bool bf16 = p.is_bf16 != 0;
#pragma unroll 1
    for (int cb = 0; cb < BN / 32; ++cb) {
      const int c0 = n0 + cb * 32;
      if (c0 >= p.Cout) break;                                   // warp-uniform
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32), v);
      tmem_ld_wait();
      if (!row_ok) continue;
      if (p.fast_out) {
        const CtxOutSeg& sg = p.segs.seg[0];
        uint16_t* out = reinterpret_cast<uint16_t*>(sg.ptr) + (long long)n_img * sg.img_stride + (long long)pix * sg.pix_stride + sg.ch_offset;
        const uint16_t* res = p.residual ? reinterpret_cast<const uint16_t*>(p.residual) + (long long)m * p.res_cstride + p.res_coffset : nullptr;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c = c0 + g * 8;
          if (c >= p.Cout) break;
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4));
          float f[8] = {__uint_as_float(v[g * 8 + 0]) + b0.x, __uint_as_float(v[g * 8 + 1]) + b0.y,
                        __uint_as_float(v[g * 8 + 2]) + b0.z, __uint_as_float(v[g * 8 + 3]) + b0.w,
                        __uint_as_float(v[g * 8 + 4]) + b1.x, __uint_as_float(v[g * 8 + 5]) + b1.y,
                        __uint_as_float(v[g * 8 + 6]) + b1.z, __uint_as_float(v[g * 8 + 7]) + b1.w};
          if (res) {
            const uint4 rv = *reinterpret_cast<const uint4*>(res + c);
            const float2 r0 = unpack2(rv.x, bf16), r1 = unpack2(rv.y, bf16), r2 = unpack2(rv.z, bf16), r3 = unpack2(rv.w, bf16);
            f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y; f[4] += r2.x; f[5] += r2.y; f[6] += r3.x; f[7] += r3.y;
          }
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
          }
          uint4 o;
          o.x = pack2(f[0], f[1], bf16); o.y = pack2(f[2], f[3], bf16); o.z = pack2(f[4], f[5], bf16); o.w = pack2(f[6], f[7], bf16);
          *reinterpret_cast<uint4*>(out + c) = o;
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < 32; ++j) {
          const int c = c0 + j;
          if (c >= p.Cout) break;
          float f = __uint_as_float(v[j]) + (p.bias ? p.bias[c] : 0.f);
          if (p.residual) f += load_as(p.residual, (long long)m * p.res_cstride + p.res_coffset + c, p.res_dtype);
          if (p.relu) f = fmaxf(f, 0.f);
#pragma unroll
          for (int sgi = 0; sgi < 3; ++sgi) {
            if (sgi < p.segs.nseg && c >= p.segs.seg[sgi].c_begin && c < p.segs.seg[sgi].c_end) {
              const CtxOutSeg& sg = p.segs.seg[sgi];
              store_as(sg.ptr, (long long)n_img * sg.img_stride + (long long)pix * sg.pix_stride + sg.ch_offset + (c - sg.c_begin),
                       sg.dtype, f);
            }
          }
        }
      }
    }
  } else if (warp == 4) {
    // ================= weight-tile TMA producer =================
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it % S;
        mbar_wait(empty0 + 8 * s, ((it / S) & 1) ^ 1);
        mbar_arrive_expect_tx(full0 + 8 * s, B_STAGE);
        tma_load_2d(sB + s * B_STAGE, &tmap_w, it * TC_BK, n0, full0 + 8 * s);
      }
    }
  } else {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(p.is_bf16 != 0, TC_BM, BN);
      for (int it = 0; it < nk; ++it) {
        const int s = it % S;
        mbar_wait(full0 + 8 * s, (it / S) & 1);
        tc_fence_after();
        const uint32_t a = sA + s * TC_A_STAGE, b = sB + s * B_STAGE;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k)
          umma_f16(tmem_base, make_sw128_desc(a + k * 32), make_sw128_desc(b + k * 32), idesc, (it | k) ? 1u : 0u);
        umma_commit(empty0 + 8 * s);
      }
      umma_commit(accum_bar);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ---------------------------------------------------------------------------------------------------
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  });
  return fn;
}

int encode_2d_sw128(CUtensorMap* out, const void* base, bool bf16, unsigned long long rows, unsigned long long cols,
                    unsigned box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return CTX_ERR_CUDA; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {64u, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base),
                   gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return CTX_ERR_CUDA; }
  return CTX_OK;
}

struct TcPlan {
  CUtensorMap tmap;
  TcParams p;
  int bn, stages;
  dim3 grid;
  size_t smem;
};

static int tc_supported(const CtxConvParams* p) {
  if (!p) return 0;
  if (p->in_dtype != CTX_BF16 && p->in_dtype != CTX_F16) return 0;
  if (p->Cin % 8 || p->in_cstride % 8 || p->in_coffset % 8) return 0;
  if (p->KH * p->KW > 32) return 0;
  if (((uintptr_t)p->in) % 16 || ((uintptr_t)p->weight) % 16) return 0;
  return 1;
}

static void choose_tile(const CtxConvParams* p, int nk, int* bn, int* stages) {
  const int c = p->Cout;
  auto padded = [&](int b) { return (c + b - 1) / b * b; };
  int best = 256;
  if (padded(128) < padded(best)) best = 128;
  if (padded(64) < padded(best)) best = 64;
  *bn = best;
  if (best == 256) *stages = 4;
  else if (best == 128) *stages = nk <= 18 ? 3 : 6;
  else *stages = 4;
}

template <int BN, int S>
static int launch_tc(const TcPlan* pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CTX_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<BN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    attr_set = true;
  }
  conv_tc_kernel<BN, S><<<pl->grid, TC_THREADS, pl->smem, st>>>(pl->tmap, pl->p);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

}  // namespace ctx

using namespace ctx;

extern "C" int ctx_conv2d_tc_supported(const CtxConvParams* p) { return tc_supported(p); }

extern "C" int ctx_conv2d_tc_plan_create(const CtxConvParams* p, void** plan_out) {
  CTX_REQUIRE(p && plan_out, "ctx_conv2d_tc_plan_create: null argument");
  *plan_out = nullptr;
  if (!tc_supported(p)) { set_error("ctx_conv2d_tc: geometry / dtype not supported by the tcgen05 path"); return CTX_ERR_UNSUPPORTED; }
  CTX_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->Cin > 0 && p->Cout > 0 && p->stride > 0 && p->dil > 0, "conv_tc: bad dims");
  const int ho = (p->H + 2 * p->pad_h - p->dil * (p->KH - 1) - 1) / p->stride + 1;
  const int wo = (p->W + 2 * p->pad_w - p->dil * (p->KW - 1) - 1) / p->stride + 1;
  CTX_REQUIRE(ho == p->Ho && wo == p->Wo, "conv_tc: Ho/Wo inconsistent with geometry");
  CTX_REQUIRE(p->nseg >= 1 && p->nseg <= 3 && p->bias, "conv_tc: needs 1..3 output segments and a bias vector");
  CTX_REQUIRE((long long)p->N * p->Ho * p->Wo < (1ll << 31), "conv_tc: too many output pixels");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("ctx_conv2d_tc: cuTensorMapEncodeTiled not available from the driver"); return CTX_ERR_CUDA; }

  TcPlan* pl = new TcPlan();
  TcParams& t = pl->p;
  t.in = p->in; t.bias = p->bias; t.residual = p->residual;
  t.N = p->N; t.H = p->H; t.W = p->W; t.Cin = p->Cin; t.in_cstride = p->in_cstride; t.in_coffset = p->in_coffset;
  t.Cout = p->Cout; t.KH = p->KH; t.KW = p->KW; t.stride = p->stride; t.pad_h = p->pad_h; t.pad_w = p->pad_w; t.dil = p->dil;
  t.Ho = p->Ho; t.Wo = p->Wo; t.relu = p->relu;
  t.M = p->N * p->Ho * p->Wo;
  t.cin_blocks = (p->Cin + TC_BK - 1) / TC_BK;
  t.nk = p->KH * p->KW * t.cin_blocks;
  t.res_cstride = p->res_cstride; t.res_coffset = p->res_coffset; t.res_dtype = p->res_dtype;
  t.is_bf16 = p->in_dtype == CTX_BF16;
  t.segs.nseg = p->nseg;
  for (int s = 0; s < 3; ++s) t.segs.seg[s] = p->seg[s < p->nseg ? s : 0];
  const CtxOutSeg& s0 = p->seg[0];
  t.fast_out = p->nseg == 1 && s0.dtype == p->in_dtype && s0.c_begin == 0 && s0.c_end == p->Cout && p->Cout % 8 == 0 &&
               s0.pix_stride % 8 == 0 && s0.ch_offset % 8 == 0 && s0.img_stride % 8 == 0 && ((uintptr_t)s0.ptr) % 16 == 0 &&
               ((uintptr_t)p->bias) % 16 == 0 &&
               (!p->residual || (p->res_dtype == p->in_dtype && p->res_cstride % 8 == 0 && p->res_coffset % 8 == 0 &&
                                 ((uintptr_t)p->residual) % 16 == 0));
  choose_tile(p, t.nk, &pl->bn, &pl->stages);
  pl->grid = dim3((unsigned)cdiv(t.M, TC_BM), (unsigned)cdiv(p->Cout, pl->bn));
  pl->smem = (size_t)pl->stages * (TC_A_STAGE + pl->bn * TC_BK * 2) + 16 * pl->stages + 16 + 1024;

  // weights: [Cout_pad][KH*KW*Cin_pad] 16-bit, K-major; box = 64 (K) x BN (Cout), SWIZZLE_128B, OOB rows read as zero
  const cuuint64_t ktot = (cuuint64_t)p->KH * p->KW * t.cin_blocks * TC_BK;
  const cuuint64_t cout_pad = (cuuint64_t)((p->Cout + 15) / 16 * 16);
  int rc = encode_2d_sw128(&pl->tmap, p->weight, t.is_bf16 != 0, cout_pad, ktot, (unsigned)pl->bn);
  if (rc) { delete pl; return rc; }
  *plan_out = pl;
  return CTX_OK;
}

extern "C" int ctx_conv2d_tc_plan_run(void* plan, void* stream) {
  CTX_REQUIRE(plan, "ctx_conv2d_tc_plan_run: null plan");
  const TcPlan* pl = (const TcPlan*)plan;
  cudaStream_t st = (cudaStream_t)stream;
  if (pl->bn == 256) return launch_tc<256, 4>(pl, st);
  if (pl->bn == 128) return pl->stages == 3 ? launch_tc<128, 3>(pl, st) : launch_tc<128, 6>(pl, st);
  return launch_tc<64, 4>(pl, st);
}

extern "C" void ctx_conv2d_tc_plan_destroy(void* plan) { delete (TcPlan*)plan; }
