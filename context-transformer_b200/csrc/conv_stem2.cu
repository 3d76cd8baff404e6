// conv_stem2.cu — conv1_1 -> ReLU -> conv1_2 -> ReLU [-> MaxPool2d(2,2)] of the VGG trunk (models/RFB_Net_vgg.py vgg()
// :323-343, base.0 .. base.4) as ONE persistent tcgen05 kernel: the 64-channel full-resolution activation between the two
// convs (368 MB per batch of 32 at 300x300, written once and read once by the unfused pair) never exists in HBM.
//
// Per 8 x 16 output tile of conv1_2 (the HALO tiling of conv_tc.cu):
//   warp 0        TMA: the tile's raw fp32 NCHW input neighbourhood, 3 channels x 20 x 16 pixels (the 18 x 10 conv1_1 outputs
//                 conv1_2 needs + their own halo; out-of-image pixels are zero-filled by the TMA unit = conv1_1's padding), and,
//                 once per CTA, both weight tensors (conv1_2: nine 64-channel taps, 72 KB resident; conv1_1: [64][27 -> 64]).
//   6 builders    one conv1_1 output pixel per thread: 27 shared-memory loads -> one 32-value K row (27 + 5 zeros) of the
//                 SWIZZLE_128B operand A1 (180 rows = two M tiles, the second one partly unused).
//   warp 1        MMA issuer.  Stem: S[256 x 64] = A1 x W1^T (two M tiles x two K = 16 steps) into TMEM; main: the nine taps of
//                 conv1_2 read the staged 18 x 10 x 64 activation patch A2 through nine shifted descriptor windows (36 MMAs,
//                 M128 x N64 x K16).  The stem of tile j + 1 is issued BEFORE the main MMAs of tile j, so the mid warps
//                 convert it while the tensor pipe works on tile j.
//   6 mid warps   S -> + bias, ReLU, 16-bit -> A2 rows; rows whose pixel lies outside the image are written as ZERO (conv1_2
//                 pads its input, i.e. conv1_1's output map, with zeros — not with conv1_1 evaluated on padding).
//   2 x 4 epilogue warps  the shared conv epilogue (conv_tc_epilogue.cuh): + bias, ReLU, 2 x 2 max-pool, NHWC store.
// Arithmetic is the unfused pair's: 16-bit operands, fp32 accumulation in TMEM, conv1_1's output rounded to 16 bits before
// conv1_2 reads it — the results are bit-identical to conv_tc_kernel (STEM mode) followed by conv_halo_kernel.
#include "conv_tc_epilogue.cuh"

namespace ctx {

constexpr int S2_THREADS = 704;                    // 22 warps
constexpr int S2_PW = 10, S2_PH = 18, S2_NPIX = S2_PW * S2_PH;        // conv1_1 outputs staged per tile
// raw input box: rows y0 - 2 .. y0 + 17, columns x0 - 4 .. x0 + 11 (the patch needs x0 - 2 .. x0 + 9, but a TMA box must start on a
// 16-byte boundary of the global row: the innermost coordinate of an fp32 tensor has to be a multiple of 4 — a start at x0 - 2
// is an illegal instruction, profiles/microbench/tma_raw_probe.cu)
constexpr int S2_RW = 16, S2_RH = 20, S2_RX = 2, S2_RAW_BYTES = S2_RW * S2_RH * 3 * 4, S2_RAW_SLOT = 4096;
constexpr int S2_A_SLOT = 23552;                   // 180 rows x 128 B, rounded up to 1024
constexpr int S2_SA2 = 3;                          // A2 ring
constexpr int S2_BUILDERS = 6, S2_MIDS = 6;
constexpr uint32_t S2_D_COLS = 0, S2_S_COLS = 128; // TMEM: conv1_2 accumulators 2 x 64 columns, stem results 2 x (2 x 64)

__global__ void __launch_bounds__(S2_THREADS, 1)
conv_stem2_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_w1,
                  const __grid_constant__ CUtensorMap tmap_raw, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int BN = p.bn;
  const uint32_t B_TAP = (uint32_t)BN * TC_BK * 2;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW2 = smem_base, sW1 = sW2 + 9u * B_TAP, sA1 = sW1 + 8192u, sA2 = sA1 + 2u * S2_A_SLOT,
                 sRaw = sA2 + (uint32_t)S2_SA2 * S2_A_SLOT, bars = sRaw + 2u * S2_RAW_SLOT;
  const uint32_t raw_full0 = bars, raw_empty0 = bars + 16, a1_full0 = bars + 32, a1_empty0 = bars + 48, s_full0 = bars + 64,
                 s_empty0 = bars + 80, a2_full0 = bars + 96, a2_empty0 = a2_full0 + 8 * S2_SA2, accf0 = a2_empty0 + 8 * S2_SA2,
                 acce0 = accf0 + 16, w_full = acce0 + 16, tmem_slot = w_full + 8;
  float* s_bias = reinterpret_cast<float*>(smem_raw + (tmem_slot + 8 - smem_u32(smem_raw)));      // conv1_2: [ceil32(Cout) + 32]
  float* s_bias1 = s_bias + ((p.Cout + 31) & ~31) + 32;                                            // conv1_1: [64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group0 = blockIdx.x, ngroups = gridDim.x;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_w); tma_prefetch_desc(&tmap_w1); tma_prefetch_desc(&tmap_raw); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(raw_full0 + 8 * s, 1); mbar_init(raw_empty0 + 8 * s, S2_BUILDERS);
        mbar_init(a1_full0 + 8 * s, S2_BUILDERS); mbar_init(a1_empty0 + 8 * s, 1);
        mbar_init(s_full0 + 8 * s, 1); mbar_init(s_empty0 + 8 * s, S2_MIDS);
        mbar_init(accf0 + 8 * s, 1); mbar_init(acce0 + 8 * s, 4);
      }
      for (int s = 0; s < S2_SA2; ++s) { mbar_init(a2_full0 + 8 * s, S2_MIDS); mbar_init(a2_empty0 + 8 * s, 1); }
      mbar_init(w_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512); tmem_relinquish();
  }
  for (int c = threadIdx.x; c < ((p.Cout + 31) & ~31) + 32; c += S2_THREADS) s_bias[c] = c < p.Cout ? p.bias[c] : 0.f;
  if (threadIdx.x < 64) s_bias1[threadIdx.x] = p.bias1[threadIdx.x];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");          // programmatic dependent launch (see conv_tc_kernel)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int per_img = p.tiles_x * p.tiles_y;
  auto tile_origin = [&](int tile, int& n_img, int& x0, int& y0) {
    n_img = tile / per_img;
    const int t = tile - n_img * per_img, ty = t / p.tiles_x;
    x0 = (t - ty * p.tiles_x) * p.TW; y0 = ty * p.TH;
  };

  if (warp == 0) {
    // ================= TMA producer: weights once, then one raw input patch per tile =================
    if (elect_one() && group0 < p.num_tiles) {
      mbar_arrive_expect_tx(w_full, 9u * B_TAP + 8192u);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(sW2 + (uint32_t)tap * B_TAP, &tmap_w, tap * TC_BK, 0, w_full);
      tma_load_2d(sW1, &tmap_w1, 0, 0, w_full);
      uint32_t slot = 0, ph = 1;
      for (int tile = group0; tile < p.num_tiles; tile += ngroups) {
        int n_img, x0, y0;
        tile_origin(tile, n_img, x0, y0);
        mbar_wait(raw_empty0 + 8 * slot, ph);
        mbar_arrive_expect_tx(raw_full0 + 8 * slot, (uint32_t)S2_RAW_BYTES);
        tma_load_4d(sRaw + slot * S2_RAW_SLOT, &tmap_raw, x0 - 2 - S2_RX, y0 - 2, 0, n_img, raw_full0 + 8 * slot);
        if ((slot ^= 1u) == 0u) ph ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one() && group0 < p.num_tiles) {
      const uint32_t idesc = make_idesc_f16(p.is_bf16 != 0, TC_BM, BN), idesc1 = make_idesc_f16(p.is_bf16 != 0, TC_BM, 64);
      const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)((uint32_t)(S2_PW * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t bdesc0 = make_sw128_desc(sW2), b_tap = (uint64_t)(B_TAP >> 4), w1desc = make_sw128_desc(sW1);
      const int n_local = (p.num_tiles - group0 + ngroups - 1) / ngroups;
      mbar_wait(w_full, 0);
      auto stem = [&](int j) {
        const uint32_t slot = (uint32_t)j & 1u, par = ((uint32_t)j >> 1) & 1u;
        mbar_wait(s_empty0 + 8 * slot, par ^ 1u);
        mbar_wait(a1_full0 + 8 * slot, par);
        tc_fence_after();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint64_t ad = make_sw128_desc(sA1 + slot * S2_A_SLOT + (uint32_t)mt * 16384u);
          const uint32_t d = tmem_base + S2_S_COLS + slot * 128u + (uint32_t)mt * 64u;
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_f16(d, ad + 2 * k, w1desc + 2 * k, idesc1, k ? 1u : 0u);
        }
        umma_commit(a1_empty0 + 8 * slot);
        umma_commit(s_full0 + 8 * slot);
      };
      stem(0);
      uint32_t a2slot = 0, a2par = 0;
      for (int j = 0; j < n_local; ++j) {
        if (j + 1 < n_local) stem(j + 1);
        const uint32_t buf = (uint32_t)j & 1u;
        mbar_wait(acce0 + 8 * buf, (((uint32_t)j >> 1) & 1u) ^ 1u);
        mbar_wait(a2_full0 + 8 * a2slot, a2par);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + S2_D_COLS + buf * (uint32_t)p.acc_stride;
        uint32_t a_tap = ((sA2 + a2slot * S2_A_SLOT) >> 4) & 0x3FFFu;
        uint64_t bt = bdesc0;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap, bt += b_tap) {
          const uint64_t ad = desc_hi | (uint64_t)a_tap;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) umma_f16(tmem_d, ad + 2 * k, bt + 2 * k, idesc, (tap | k) ? 1u : 0u);
          a_tap += (tap % 3 == 2) ? (uint32_t)((S2_PW - 2) * 128) >> 4 : 8u;
        }
        umma_commit(a2_empty0 + 8 * a2slot);
        umma_commit(accf0 + 8 * buf);
        if (++a2slot == (uint32_t)S2_SA2) { a2slot = 0; a2par ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp >= 12 && warp < 20) {
    // single 16-bit segment, no residual (checked by the plan): the lean epilogue, with or without the fused 2 x 2 pool
    if (p.pool2) epilogue_fast_role<1, true, false, false>(p, s_bias, tmem_base + S2_D_COLS, accf0, acce0, warp, lane, (uint32_t)(warp - 12) >> 2, 0, group0, ngroups, 0u);
    else epilogue_fast_role<1, false, false, false>(p, s_bias, tmem_base + S2_D_COLS, accf0, acce0, warp, lane, (uint32_t)(warp - 12) >> 2, 0, group0, ngroups, 0u);
  } else if (warp >= 4 && warp < 10) {
    // ================= mid warps: stem accumulator -> 16-bit activation patch (operand A2) =================
    const int mt = warp >= 8 ? 1 : 0, q = warp & 3;
    const int r = mt * 128 + q * 32 + lane;                       // patch pixel = A2 row
    const int py = r / S2_PW, px = r - py * S2_PW;
    const bool bf16 = p.is_bf16 != 0;
    const uint32_t row_off = (uint32_t)r * 128u, sw = (uint32_t)(r & 7);
    uint32_t j = 0, a2slot = 0, a2par = 0;
    for (int tile = group0; tile < p.num_tiles; tile += ngroups, ++j) {
      int n_img, x0, y0;
      tile_origin(tile, n_img, x0, y0);
      const int y = y0 - 1 + py, x = x0 - 1 + px;
      const bool inside = r < S2_NPIX && y >= 0 && y < p.H && x >= 0 && x < p.W;
      const uint32_t slot = j & 1u, par = (j >> 1) & 1u;
      mbar_wait(a2_empty0 + 8 * a2slot, a2par ^ 1u);
      mbar_wait(s_full0 + 8 * slot, par);
      tc_fence_after();
      const uint32_t taddr = tmem_base + S2_S_COLS + slot * 128u + (uint32_t)mt * 64u + ((uint32_t)(q * 32) << 16);
      const uint32_t row = sA2 + a2slot * S2_A_SLOT + row_off;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)h * 32u, v);
        tmem_ld_wait();
        if (h == 1) {                                             // the stem accumulator is free again
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty0 + 8 * slot);
        }
        if (r < S2_NPIX) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = h * 32 + g * 8 + e * 2;
              const float f0 = fmaxf(__uint_as_float(v[g * 8 + e * 2]) + s_bias1[c], 0.f);
              const float f1 = fmaxf(__uint_as_float(v[g * 8 + e * 2 + 1]) + s_bias1[c + 1], 0.f);
              w[e] = inside ? pack2(f0, f1, bf16) : 0u;
            }
            const uint32_t chunk = (uint32_t)(h * 4 + g);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((chunk ^ sw) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2_full0 + 8 * a2slot);
      if (++a2slot == (uint32_t)S2_SA2) { a2slot = 0; a2par ^= 1u; }
    }
  } else {
    // ================= builders: raw fp32 patch -> 27-value K rows of conv1_1 (operand A1) =================
    const int bi = warp < 4 ? warp - 2 : (warp < 12 ? warp - 8 : warp - 16);      // warps 2, 3, 10, 11, 20, 21 -> 0..5
    const int r = bi * 32 + lane;
    const int py = r / S2_PW, px = r - py * S2_PW;
    const bool bf16 = p.is_bf16 != 0;
    const uint32_t row_off = (uint32_t)r * 128u, sw = (uint32_t)(r & 7);
    const uint32_t src_off = (uint32_t)(py * S2_RW + px + S2_RX) * 4u;
    uint32_t j = 0;
    for (int tile = group0; tile < p.num_tiles; tile += ngroups, ++j) {
      const uint32_t slot = j & 1u, par = (j >> 1) & 1u;
      float v[32];
#pragma unroll
      for (int e = 27; e < 32; ++e) v[e] = 0.f;
      mbar_wait(raw_full0 + 8 * slot, par);
      if (r < S2_NPIX) {
        const uint32_t src = sRaw + slot * S2_RAW_SLOT + src_off;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[(ky * 3 + kx) * 3 + c]) : "r"(src + (uint32_t)((c * S2_RH + ky) * S2_RW + kx) * 4u));
      }
      uint32_t w[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) w[e] = (r < S2_NPIX) ? pack2(v[2 * e], v[2 * e + 1], bf16) : 0u;
      __syncwarp();
      if (lane == 0) mbar_arrive(raw_empty0 + 8 * slot);          // the raw patch is in registers
      mbar_wait(a1_empty0 + 8 * slot, par ^ 1u);
      if (r < S2_NPIX) {
        const uint32_t row = sA1 + slot * S2_A_SLOT + row_off;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((uint32_t)ch ^ sw) << 4)), "r"(w[ch * 4]), "r"(w[ch * 4 + 1]),
                       "r"(w[ch * 4 + 2]), "r"(w[ch * 4 + 3]) : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a1_full0 + 8 * slot);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int launch_stem2(const TcPlan* pl, cudaStream_t st) {
  CTX_CUDA_TRY(cudaFuncSetAttribute(conv_stem2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)pl->grid);
  cfg.blockDim = dim3(S2_THREADS);
  cfg.dynamicSmemBytes = pl->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CTX_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_stem2_kernel, pl->tmap_w, pl->tmap_w1, pl->tmap_raw, pl->p));
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

// raw network input [N][3][H][W] fp32 as a 4-D tensor (W innermost); box = 16 x 20 pixels x 3 channels of one image, no swizzle;
// out-of-image pixels read as zero (= the zero padding of conv1_1)
static int encode_raw_nchw(CUtensorMap* out, const float* base, int N, int H, int W) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return CTX_ERR_CUDA; }
  cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, 3u, (cuuint64_t)N};
  cuuint64_t gstride[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
  cuuint32_t box[4] = {(cuuint32_t)S2_RW, (cuuint32_t)S2_RH, 3u, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (raw NCHW input) failed (CUresult %d)", (int)r); return CTX_ERR_CUDA; }
  return CTX_OK;
}

}  // namespace ctx

using namespace ctx;

// conv1_2 as the caller would hand it to ctx_conv2d_tc_plan_create: 3x3 / stride 1 / pad 1 / dilation 1, 64 input channels
// (= conv1_1's outputs), at most 64 output channels in one 16-bit segment, optional fused 2x2 pooling
extern "C" int ctx_conv2d_stem2_supported(const CtxConvParams* c) {
  if (!c || c->in_nchw || c->split) return 0;
  if (c->in_dtype != CTX_BF16 && c->in_dtype != CTX_F16) return 0;
  if (c->Cin != 64 || c->Cout > 64 || c->Cout % 16 || c->KH != 3 || c->KW != 3 || c->stride != 1 || c->pad_h != 1 || c->pad_w != 1 || c->dil != 1) return 0;
  if (c->W % 4 || c->residual || c->nseg != 1 || c->seg[0].dtype != c->in_dtype || !c->relu) return 0;
  if (c->pool2 && ((c->H | c->W) & 1)) return 0;
  return 1;
}

extern "C" int ctx_conv2d_stem2_plan_create(const CtxConvParams* conv, const float* stem_in, const void* stem_weight, const float* stem_bias,
                                            void** plan_out) {
  CTX_REQUIRE(conv && stem_in && stem_weight && stem_bias && plan_out, "ctx_conv2d_stem2_plan_create: null argument");
  *plan_out = nullptr;
  if (!ctx_conv2d_stem2_supported(conv)) { set_error("ctx_conv2d_stem2: geometry / dtype not supported by the fused conv1_1 + conv1_2 kernel"); return CTX_ERR_UNSUPPORTED; }
  CTX_REQUIRE(((uintptr_t)stem_in) % 16 == 0 && ((uintptr_t)stem_weight) % 16 == 0, "ctx_conv2d_stem2: stem input / weights must be 16-byte aligned");
  // tiling, epilogue flags and the conv1_2 weight descriptor come from the HALO plan with resident weights; its activation
  // descriptor is never used (the activation is produced on chip), so any aligned pointer stands in for `in`
  CtxConvParams c = *conv;
  c.in = conv->weight;
  void* plan = nullptr;
  int rc = ctx_conv2d_tc_plan_create_tuned(&c, 0, 1, 4, 0, &plan);
  if (rc) return rc;
  TcPlan* pl = (TcPlan*)plan;
  TcParams& t = pl->p;
  if (t.a_mode != A_HALO || !t.resident || t.n_tiles_n != 1 || t.cin_blocks != 1 || t.TW != 8 || t.TH != 16 || !t.fast_out) {
    ctx_conv2d_tc_plan_destroy(plan);
    set_error("ctx_conv2d_stem2: conv1_2 does not plan as a single resident HALO tile");
    return CTX_ERR_UNSUPPORTED;
  }
  t.a_mode = A_STEM2;
  t.in = nullptr;
  t.bias1 = stem_bias;
  t.occ = 1; t.acc_stride = 64;
  const size_t bias_bytes = 4 * (((size_t)conv->Cout + 31) / 32 * 32 + 32) + 4 * 64;
  pl->smem = 1024 + 9 * (size_t)t.bn * TC_BK * 2 + 8192 + (2 + S2_SA2) * (size_t)S2_A_SLOT + 2 * S2_RAW_SLOT + 256 /* barriers */ + bias_bytes;
  pl->grid = std::min(t.num_tiles, num_sms());
  rc = encode_2d_sw128(&pl->tmap_w1, stem_weight, t.is_bf16 != 0, 64, 64, 64u);
  if (!rc) rc = encode_raw_nchw(&pl->tmap_raw, stem_in, conv->N, conv->H, conv->W);
  if (rc) { ctx_conv2d_tc_plan_destroy(plan); return rc; }
  *plan_out = plan;
  return CTX_OK;
}
