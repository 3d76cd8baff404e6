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
//   warps 1, 2    MMA issuers (even / odd tiles).  Stem: S[256 x 64] = A1 x W1^T (two M tiles x two K = 16 steps) into TMEM; main: the nine taps of
//                 conv1_2 read the staged 18 x 10 x 64 activation patch A2 through nine shifted descriptor windows (36 MMAs,
//                 M128 x N64 x K16).  An issuer queues the stem of its next tile right behind the main MMAs of the current
//                 one, so the mid warps convert it while the tensor pipe works on the other issuer's tile.
//   6 mid warps   S -> + bias, ReLU, 16-bit -> A2 rows; rows whose pixel lies outside the image are written as ZERO (conv1_2
//                 pads its input, i.e. conv1_1's output map, with zeros — not with conv1_1 evaluated on padding).
//   2 x 4 epilogue warps  the shared conv epilogue (conv_tc_epilogue.cuh): + bias, ReLU, 2 x 2 max-pool, NHWC store.
// Arithmetic is the unfused pair's: 16-bit operands, fp32 accumulation in TMEM, conv1_1's output rounded to 16 bits before
// conv1_2 reads it — the results are bit-identical to conv_tc_kernel (STEM mode) followed by conv_halo_kernel.
#include "conv_tc_epilogue.cuh"

#include <stdlib.h>

namespace ctx {

constexpr int S2_THREADS = 736;                    // 23 warps
constexpr int S2_PW = 10, S2_PH = 18, S2_NPIX = S2_PW * S2_PH;        // conv1_1 outputs staged per tile
// raw input box: rows y0 - 2 .. y0 + 17, columns x0 - 4 .. x0 + 11 (the patch needs x0 - 2 .. x0 + 9, but a TMA box must start on a
// 16-byte boundary of the global row: the innermost coordinate of an fp32 tensor has to be a multiple of 4 — a start at x0 - 2
// is an illegal instruction, profiles/microbench/tma_raw_probe.cu)
constexpr int S2_RW = 16, S2_RH = 20, S2_RX = 2, S2_RAW_BYTES = S2_RW * S2_RH * 3 * 4, S2_RAW_SLOT = S2_RAW_BYTES;
// Operand slots hold exactly 180 rows of 128 B.  They are NOT multiples of 1024 B: SWIZZLE_128B XORs the 16-byte chunk index with
// bits 7..9 of the absolute shared-memory address, so writers swizzle with (slot_base / 128 + row) & 7 and the descriptors may
// start at any 128-byte row (profiles/r1_mma_probe.txt).
constexpr int S2_A_SLOT = S2_NPIX * 128;
constexpr int S2_SA2 = 4;                          // A2 ring (two slots per issuer)
constexpr int S2_BUILDERS = 6, S2_MIDS = 6;
constexpr uint32_t S2_D_COLS = 0, S2_S_COLS = 256; // TMEM: conv1_2 accumulators 4 x 64 columns, stem results 2 x (2 x 64)

struct Bias64 { float v[64]; };                    // conv1_1's bias as a kernel parameter: FADD takes constant-bank operands directly

// Local tile j uses raw slot / A1 / S number j & 1, MMA issuer j & 1, and A2 slot / accumulator D number j & 3.
// Hand-offs of tile j (barriers indexed like their buffer; phase = use count of the buffer & 1):
//   raw_full  TMA -> builders            raw_empty  builders -> TMA (patch is in registers)
//   a1_full   builders -> issuer         s_full     stem MMAs done: S ready for the mid warps AND A1 free for the builders (tile j + 2)
//   a2_full   mid warps -> issuer (also: S free again)          a2_empty  main MMAs done, A2 slot free for the mid warps (tile j + 4)
//   accf      main MMAs done: D ready for the epilogue          acce      epilogue -> issuer
// s_full frees A1 for tile j + 2 and cannot run a phase ahead of the builders waiting on it: its next completion needs tile j + 2.
template <bool BF16>
__global__ void __launch_bounds__(S2_THREADS, 1)
conv_stem2_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_w1,
                  const __grid_constant__ CUtensorMap tmap_raw, const TcParams p, const __grid_constant__ Bias64 bias1,
                  const __grid_constant__ Bias64 bias2) {
  extern __shared__ uint8_t smem_raw[];
  const int BN = p.bn;
  const uint32_t B_TAP = (uint32_t)BN * TC_BK * 2;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW2 = smem_base, sW1 = sW2 + 9u * B_TAP, sA1 = sW1 + 8192u, sA2 = sA1 + 2u * S2_A_SLOT,
                 sRaw = sA2 + (uint32_t)S2_SA2 * S2_A_SLOT, bars = sRaw + 2u * S2_RAW_SLOT;
  const uint32_t raw_full0 = bars, raw_empty0 = bars + 16, a1_full0 = bars + 32, s_full0 = bars + 48, accf0 = bars + 64, acce0 = bars + 96,
                 a2_full0 = bars + 128, a2_empty0 = bars + 160, w_full = bars + 192, tmem_slot = bars + 200;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group0 = blockIdx.x, ngroups = gridDim.x;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_w); tma_prefetch_desc(&tmap_w1); tma_prefetch_desc(&tmap_raw); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(raw_full0 + 8 * s, 1); mbar_init(raw_empty0 + 8 * s, S2_BUILDERS);
        mbar_init(a1_full0 + 8 * s, S2_BUILDERS); mbar_init(s_full0 + 8 * s, 1);
      }
      for (int s = 0; s < S2_SA2; ++s) {
        mbar_init(a2_full0 + 8 * s, S2_MIDS); mbar_init(a2_empty0 + 8 * s, 1);
        mbar_init(accf0 + 8 * s, 1); mbar_init(acce0 + 8 * s, 4);          // (the epilogue sets take alternate tiles)
      }
      mbar_init(w_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512); tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");          // programmatic dependent launch (see conv_tc_kernel)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================= TMA producer: weights once, then one raw input patch per tile =================
    if (elect_one() && group0 < p.num_tiles) {
      mbar_arrive_expect_tx(w_full, 9u * B_TAP + 8192u);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(sW2 + (uint32_t)tap * B_TAP, &tmap_w, tap * TC_BK, 0, w_full);
      tma_load_2d(sW1, &tmap_w1, 0, 0, w_full);
      uint32_t slot = 0, ph = 1;
      for (int tile = group0; tile < p.num_tiles; tile += ngroups) {
        const int n_img = tile / per_img, t = tile - n_img * per_img, ty = t / p.tiles_x;
        const int x0 = (t - ty * p.tiles_x) * p.TW, y0 = ty * p.TH;
        mbar_wait(raw_empty0 + 8 * slot, ph);
        dbg_stamp(p, 0, (uint32_t)((tile - group0) / ngroups), 0);
        mbar_arrive_expect_tx(raw_full0 + 8 * slot, (uint32_t)S2_RAW_BYTES);
        tma_load_4d(sRaw + slot * S2_RAW_SLOT, &tmap_raw, x0 - 2 - S2_RX, y0 - 2, 0, n_img, raw_full0 + 8 * slot);
        if ((slot ^= 1u) == 0u) ph ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2) {
    // ================= two MMA issuers: warp 1 takes the even local tiles, warp 2 the odd ones =================
    // Separate accumulators, operand buffers and barriers per issuer: while one waits for its tile's activation patch, the
    // other keeps the tensor pipe busy.  The stem of an issuer's NEXT tile is queued before the main MMAs of the current one,
    // so the mid warps convert it while the tensor pipe runs the 36 main MMAs of this tile and of the other issuer's.
    const uint32_t w = (uint32_t)(warp - 1);
    const int n_local = group0 < p.num_tiles ? (p.num_tiles - group0 + ngroups - 1) / ngroups : 0;
    if (elect_one() && (int)w < n_local) {
      const uint32_t idesc = make_idesc_f16(BF16, TC_BM, BN), idesc1 = make_idesc_f16(BF16, TC_BM, 64);
      const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)((uint32_t)(S2_PW * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t bdesc0 = make_sw128_desc(sW2), b_tap = (uint64_t)(B_TAP >> 4), w1desc = make_sw128_desc(sW1);
      const uint64_t a1d0 = make_sw128_desc(sA1 + w * S2_A_SLOT), a1d1 = make_sw128_desc(sA1 + w * S2_A_SLOT + 16384u);
      const uint32_t s_tmem = tmem_base + S2_S_COLS + w * 128u;
      const uint32_t b_a1 = a1_full0 + 8 * w, b_s = s_full0 + 8 * w;
      mbar_wait(w_full, 0);
      auto stem = [&](uint32_t par, uint32_t jj) {
        mbar_wait(b_a1, par);
        tc_fence_after();
        dbg_stamp(p, 2 + (int)w, jj, 3);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(s_tmem, a1d0 + 2 * k, w1desc + 2 * k, idesc1, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16(s_tmem + 64u, a1d1 + 2 * k, w1desc + 2 * k, idesc1, k ? 1u : 0u);
        umma_commit(b_s);
        dbg_stamp(p, 2 + (int)w, jj, 4);
      };
      stem(0, w);
      uint32_t par = 0, half = 0, par2 = 0;                        // half: which of the issuer's two A2 slots; par2: that slot's phase
      for (int j = (int)w; j < n_local; j += 2, par ^= 1u) {
        const uint32_t slot = w + 2u * half;
        mbar_wait(a2_full0 + 8 * slot, par2);                     // A2 ready; the mid warps are done with S
        dbg_stamp(p, 2 + (int)w, (uint32_t)j, 1);
        if (j + 2 < n_local) stem(par ^ 1u, (uint32_t)j + 2u);
        mbar_wait(acce0 + 8 * slot, par2 ^ 1u);                    // the epilogue has drained this accumulator (tile j - 4)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + S2_D_COLS + slot * (uint32_t)p.acc_stride;
        dbg_stamp(p, 2 + (int)w, (uint32_t)j, 0);
        uint32_t a_tap = ((sA2 + slot * S2_A_SLOT) >> 4) & 0x3FFFu;
        uint64_t bt = bdesc0;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap, bt += b_tap) {
          const uint64_t ad = desc_hi | (uint64_t)a_tap;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) umma_f16(d_tmem, ad + 2 * k, bt + 2 * k, idesc, (tap | k) ? 1u : 0u);
          a_tap += (tap % 3 == 2) ? (uint32_t)((S2_PW - 2) * 128) >> 4 : 8u;
        }
        umma_commit(accf0 + 8 * slot);
        umma_commit(a2_empty0 + 8 * slot);
        dbg_stamp(p, 2 + (int)w, (uint32_t)j, 2);
        if ((half ^= 1u) == 0u) par2 ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp >= 12 && warp < 20) {
    // single 16-bit segment, no residual (checked by the plan): the lean epilogue, with or without the fused 2 x 2 pool
    if (p.pool2) epilogue_fast_role<1, true, false, false, BF16 ? 1 : 0, S2_SA2, 64, false, false>(p, bias2.v, tmem_base + S2_D_COLS, accf0, acce0, warp, lane, (uint32_t)(warp - 12) >> 2, 0, group0, ngroups, 0u);
    else epilogue_fast_role<1, false, false, false, BF16 ? 1 : 0, S2_SA2, 64, false, false>(p, bias2.v, tmem_base + S2_D_COLS, accf0, acce0, warp, lane, (uint32_t)(warp - 12) >> 2, 0, group0, ngroups, 0u);
  } else if (warp >= 4 && warp < 10) {
    // ================= mid warps: stem accumulator -> 16-bit activation patch (operand A2) =================
    const int mt = warp >= 8 ? 1 : 0, q = warp & 3;
    const int r = mt * 128 + q * 32 + lane;                       // patch pixel = A2 row
    const int py = r / S2_PW, px = r - py * S2_PW;
    constexpr bool bf16 = BF16;
    // tile cursor advanced incrementally (no division per tile): tile = group0 + k * ngroups -> (tx, ty) inside its image
    const int t0 = group0 % per_img, dt = ngroups % per_img;
    int ty = t0 / p.tiles_x, tx = t0 - ty * p.tiles_x;
    const int dty = dt / p.tiles_x, dtx = dt - dty * p.tiles_x;
    uint32_t j = 0;
    for (int tile = group0; tile < p.num_tiles; tile += ngroups, ++j) {
      const int y = ty * p.TH - 1 + py, x = tx * p.TW - 1 + px;
      const bool inside = r < S2_NPIX && y >= 0 && y < p.H && x >= 0 && x < p.W;
      tx += dtx; ty += dty;
      if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
      if (ty >= p.tiles_y) ty -= p.tiles_y;
      const uint32_t par = (j >> 1) & 1u, slot = j & 3u;
      mbar_wait(a2_empty0 + 8 * slot, ((j >> 2) & 1u) ^ 1u);      // main MMAs of tile j - 4 have read this slot
      if (warp == 4 && lane == 0) dbg_stamp(p, 4, j, 0);
      mbar_wait(s_full0 + 8 * (j & 1u), par);
      tc_fence_after();
      if (warp == 4 && lane == 0) dbg_stamp(p, 4, j, 1);
      const uint32_t taddr = tmem_base + S2_S_COLS + (j & 1u) * 128u + (uint32_t)mt * 64u + ((uint32_t)(q * 32) << 16);
      const uint32_t slot_base = sA2 + slot * S2_A_SLOT;
      const uint32_t row = slot_base + (uint32_t)r * 128u, sw = ((slot_base >> 7) + (uint32_t)r) & 7u;
      uint32_t v[64];
      tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(taddr + 32u, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
      if (warp == 4 && lane == 0) dbg_stamp(p, 4, j, 3);
      if (inside) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t w0 = pack2_relu(__uint_as_float(v[g * 8 + 0]) + bias1.v[g * 8 + 0], __uint_as_float(v[g * 8 + 1]) + bias1.v[g * 8 + 1], bf16);
          const uint32_t w1 = pack2_relu(__uint_as_float(v[g * 8 + 2]) + bias1.v[g * 8 + 2], __uint_as_float(v[g * 8 + 3]) + bias1.v[g * 8 + 3], bf16);
          const uint32_t w2 = pack2_relu(__uint_as_float(v[g * 8 + 4]) + bias1.v[g * 8 + 4], __uint_as_float(v[g * 8 + 5]) + bias1.v[g * 8 + 5], bf16);
          const uint32_t w3 = pack2_relu(__uint_as_float(v[g * 8 + 6]) + bias1.v[g * 8 + 6], __uint_as_float(v[g * 8 + 7]) + bias1.v[g * 8 + 7], bf16);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((uint32_t)g ^ sw) << 4)), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
        }
      } else if (r < S2_NPIX) {                                   // conv1_2's zero padding
#pragma unroll
        for (int g = 0; g < 8; ++g)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(row + ((uint32_t)g << 4)), "r"(0u) : "memory");
      }
      if (warp == 4 && lane == 0) dbg_stamp(p, 4, j, 4);
      tc_fence_before();                                          // S has been read (the issuer overwrites it after a2_full)
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2_full0 + 8 * slot);
      if (warp == 4 && lane == 0) dbg_stamp(p, 4, j, 2);
    }
  } else {
    // ================= builders: raw fp32 patch -> 27-value K rows of conv1_1 (operand A1) =================
    const int bi = warp == 3 ? 0 : (warp < 12 ? warp - 9 : warp - 17);      // warps 3, 10, 11, 20, 21, 22 -> 0..5
    const int r = bi * 32 + lane;
    const int py = r / S2_PW, px = r - py * S2_PW;
    constexpr bool bf16 = BF16;
    const uint32_t src_off = (uint32_t)(py * S2_RW + px + S2_RX) * 4u;
    uint32_t j = 0;
    for (int tile = group0; tile < p.num_tiles; tile += ngroups, ++j) {
      const uint32_t slot = j & 1u, par = (j >> 1) & 1u;
      float v[32];
#pragma unroll
      for (int e = 27; e < 32; ++e) v[e] = 0.f;
      mbar_wait(raw_full0 + 8 * slot, par);
      if (warp == 3 && lane == 0) dbg_stamp(p, 1, j, 0);
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
      uint32_t wd[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) wd[e] = (r < S2_NPIX) ? pack2(v[2 * e], v[2 * e + 1], bf16) : 0u;
      __syncwarp();
      if (lane == 0) mbar_arrive(raw_empty0 + 8 * slot);          // the raw patch is in registers
      mbar_wait(s_full0 + 8 * slot, par ^ 1u);                    // stem MMAs of tile j - 2 have read A1[slot]
      if (warp == 3 && lane == 0) dbg_stamp(p, 1, j, 1);
      if (r < S2_NPIX) {
        const uint32_t slot_base = sA1 + slot * S2_A_SLOT;
        const uint32_t row = slot_base + (uint32_t)r * 128u, sw = ((slot_base >> 7) + (uint32_t)r) & 7u;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((uint32_t)ch ^ sw) << 4)), "r"(wd[ch * 4]), "r"(wd[ch * 4 + 1]),
                       "r"(wd[ch * 4 + 2]), "r"(wd[ch * 4 + 3]) : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a1_full0 + 8 * slot);
      if (warp == 3 && lane == 0) dbg_stamp(p, 1, j, 2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static long long* g_conv_timeline = nullptr;
long long* conv_timeline_ptr() { return g_conv_timeline; }

int launch_stem2(const TcPlan* pl, cudaStream_t st) {
  auto kernel = pl->p.is_bf16 ? conv_stem2_kernel<true> : conv_stem2_kernel<false>;
  CTX_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)pl->grid);
  cfg.blockDim = dim3(S2_THREADS);
  cfg.dynamicSmemBytes = pl->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int pdl = [] { const char* e = getenv("CTX_CONV_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
  attr[0].val.programmaticStreamSerializationAllowed = pdl;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TcParams prm = pl->p;
  prm.dbg = g_conv_timeline;
  Bias64 b1, b2;
  memcpy(b1.v, pl->bias1_host, sizeof b1.v);
  memcpy(b2.v, pl->bias2_host, sizeof b2.v);
  CTX_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, pl->tmap_w, pl->tmap_w1, pl->tmap_raw, prm, b1, b2));
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

extern "C" void ctx_debug_set_conv_timeline(void* device_buffer) { g_conv_timeline = (long long*)device_buffer; }

// conv1_2 as the caller would hand it to ctx_conv2d_tc_plan_create: 3x3 / stride 1 / pad 1 / dilation 1, 64 input channels
// (= conv1_1's outputs), 64 output channels in one 16-bit segment, optional fused 2x2 pooling
extern "C" int ctx_conv2d_stem2_supported(const CtxConvParams* c) {
  if (!c || c->in_nchw || c->split) return 0;
  if (c->in_dtype != CTX_BF16 && c->in_dtype != CTX_F16) return 0;
  if (c->Cin != 64 || c->Cout != 64 || c->KH != 3 || c->KW != 3 || c->stride != 1 || c->pad_h != 1 || c->pad_w != 1 || c->dil != 1) return 0;
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
  // both bias vectors travel as kernel parameters (constant-bank operands: no shared-memory traffic, no load latency): fetch them once, after whatever
  // is still writing it has finished
  CTX_CUDA_TRY(cudaDeviceSynchronize());
  CTX_CUDA_TRY(cudaMemcpy(pl->bias1_host, stem_bias, 64 * sizeof(float), cudaMemcpyDeviceToHost));
  memset(pl->bias2_host, 0, sizeof pl->bias2_host);
  CTX_CUDA_TRY(cudaMemcpy(pl->bias2_host, conv->bias, (size_t)conv->Cout * sizeof(float), cudaMemcpyDeviceToHost));
  t.occ = 1; t.acc_stride = 64;
  pl->smem = 1024 + 9 * (size_t)t.bn * TC_BK * 2 + 8192 + (2 + S2_SA2) * (size_t)S2_A_SLOT + 2 * S2_RAW_SLOT + 256 /* barriers */;
  pl->grid = std::min(t.num_tiles, num_sms());
  rc = encode_2d_sw128(&pl->tmap_w1, stem_weight, t.is_bf16 != 0, 64, 64, 64u);
  if (!rc) rc = encode_raw_nchw(&pl->tmap_raw, stem_in, conv->N, conv->H, conv->W);
  if (rc) { ctx_conv2d_tc_plan_destroy(plan); return rc; }
  *plan_out = plan;
  return CTX_OK;
}
