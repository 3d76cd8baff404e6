// conv_tc_epilogue.cuh — epilogue warps of the tensor-core conv kernels (conv_tc.cu: conv_tc_kernel / conv_halo_kernel;
// conv_stem2.cu: the fused conv1_1 + conv1_2 kernel): TMEM accumulator -> bias / residual / ReLU / 2x2 max-pool -> NHWC stores.
#pragma once
#include "conv_tc.cuh"

namespace ctx {

// ---------------------------------------------------------------------------------------------------
// Lean epilogue for the common case (one 16-bit output segment, 8-channel aligned: TcParams::fast_out), specialised at
// compile time on fused pooling / residual / bulk-copy output so that the per-chunk code is branch-free: on the layers
// with little K (the stem, conv1_2) the epilogue warps, not the MMAs, set the pace (ncu source view: ~320 warp
// instructions per 32-column chunk in the generic path below, most of them flag tests and 64-bit index arithmetic).
// 16-bit pair maximum (2 x 2 pooling / ReLU on packed values: rounding is monotonic, so max(round(a), round(b)) == round(max(a, b)))
__device__ __forceinline__ uint32_t max2_packed(uint32_t a, uint32_t b, bool bf16) {
  if (bf16) {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// DT: operand type known at compile time (1 bf16, 0 fp16) or -1; NBUF: accumulator buffers (and accf / acce barrier pairs) in rotation;
// CBN > 0: the layer is ONE N tile of exactly CBN channels (chunk loop unrolled: with `s_bias` in kernel-parameter space the bias
// becomes constant-bank operands of the FADDs)
// BPRE: the bias of a 32-column chunk is fetched while the chunk's TMEM load is in flight (the first one before the accumulator
// is complete): a shared-memory load issued right before its FADDs waits behind the tensor cores' operand fetch, which keeps the
// shared-memory pipe busy in exactly the layers whose epilogue sets the pace
template <int CL, bool POOL, bool RES, bool BULK, int DT = -1, int NBUF = 2, int CBN = 0, bool BPRE = false, bool SPLIT_ = true>
__device__ __forceinline__ void epilogue_fast_role(const TcParams& p, const float* s_bias, uint32_t tmem_base, uint32_t accf0, uint32_t acce0,
                                                   int warp, int lane, uint32_t eset, int rank, int group0, int ngroups, uint32_t stage_smem) {
  const int q = warp & 3;
  const int BN = p.bn, Cout = p.Cout, relu_cend = p.relu ? p.relu_cend : 0;     // ReLU on channels < relu_cend (multiple of 8)
  const int r = q * 32 + lane;
  const bool bf16 = DT < 0 ? p.is_bf16 != 0 : DT == 1;
  uint16_t* const seg_ptr = reinterpret_cast<uint16_t*>(p.segs.seg[0].ptr) + p.segs.seg[0].ch_offset;
  const long long img_stride = p.segs.seg[0].img_stride;
  const int pix_stride = p.segs.seg[0].pix_stride;
  const uint32_t stage_row = stage_smem + (uint32_t)(lane * Cout) * 2u;
  // Both epilogue sets drain EVERY tile, half of its 32-column chunks each (SPLIT): the epilogue of a tile then takes half as long,
  // which is what a kernel with one or two tiles per CTA — most of the pyramid — pays in full at its end (with the sets on
  // alternate tiles one of them idled through the last tile).  The bulk-copy output path needs whole rows per warp and keeps
  // the alternate-tile scheme.
  const bool SPLIT = SPLIT_ && !BULK && p.epi_split;      // (SPLIT_ = false: kernels with many tiles per CTA and four accumulators measured better with the sets on alternate tiles)
  uint32_t lt = SPLIT ? 0u : eset;
  for (int tile = group0 + (SPLIT ? 0 : (int)eset * ngroups); tile < p.num_tiles; tile += (SPLIT ? 1 : 2) * ngroups, lt += (SPLIT ? 1u : 2u)) {
    const uint32_t buf = lt & (uint32_t)(NBUF - 1);
    const int mg = CBN ? tile : tile / p.n_tiles_n, n0 = CBN ? 0 : (tile - mg * p.n_tiles_n) * BN, mt = mg * CL + rank;
    int n_img, pix;
    bool row_ok;
    if (BULK) {            // pixel-linear tile of a dense map: the batch is one long pixel row (no per-tile division)
      n_img = 0; pix = mt * TC_BM + r; row_ok = pix < p.M;
    } else {
      row_ok = tile_row_pixel(p, mt, r, n_img, pix);
    }
    long long opix = pix;
    bool store = row_ok;
    if (POOL) {            // row r = pixel (r / TW, r % TW) of the patch: 2 x 2 partners are lanes ^1 and ^TW (all four store a piece)
      const int oy = pix / p.Wo, ox = pix - oy * p.Wo;
      opix = (long long)(oy >> 1) * (p.Wo >> 1) + (ox >> 1);
    }
    uint16_t* const out = seg_ptr + (long long)n_img * img_stride + opix * pix_stride;
    const uint16_t* res = nullptr;
    if (RES) res = reinterpret_cast<const uint16_t*>(p.residual) + ((long long)n_img * p.Ho * p.Wo + pix) * p.res_cstride + p.res_coffset;
    const int c_end = CBN ? CBN : min(Cout, n0 + BN);
    // this set's chunks: [c_lo, c_hi)
    const int nchunk = (c_end - n0 + 31) >> 5, first = (nchunk + 1) >> 1;
    const int c_lo = SPLIT && eset ? n0 + first * 32 : n0, c_hi = SPLIT && !eset ? min(c_end, n0 + first * 32) : c_end;
    // Transposed stores (plain case): a thread owns one output row, so its 16-byte pieces of a 32-column chunk would go out
    // as four requests touching 32 half-written sectors each.  The four lanes of a quad exchange pieces (4 x 4 transpose, 16
    // shuffles per chunk) so that a request writes, per row, the 64 contiguous bytes of the chunk from four adjacent lanes:
    // whole 32-byte sectors, half as many of them.
    constexpr bool TR = !POOL && !BULK;
    unsigned long long outq[4] = {0ull, 0ull, 0ull, 0ull};
    unsigned okbits = 0u;
    if (TR) {
      okbits = __ballot_sync(0xffffffffu, row_ok);
#pragma unroll
      for (int k = 0; k < 4; ++k) outq[k] = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)out, (lane & ~3) + k);
    }
    // residual (RFB shortcut): the 64 bytes a thread needs per 32-column chunk are fetched one chunk ahead — the first
    // chunk before the accumulator is even complete — so their latency hides behind the MMAs / the previous chunk
    // (ncu what-if: without this the ConvLinear layers spent 70 % of their time waiting on these loads)
    uint4 rnext[4];
    auto load_res = [&](int c0) {
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) {
        const int c = c0 + gq * 8;
        rnext[gq] = (row_ok && c < c_end) ? __ldg(reinterpret_cast<const uint4*>(res + c)) : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    if (RES) load_res(c_lo);
    float4 bnext[8];
    auto load_bias = [&](int c0) {                                // s_bias is zero-padded to a whole chunk past Cout
#pragma unroll
      for (int k = 0; k < 8; ++k) bnext[k] = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * k);
    };
    if (BPRE && !CBN) load_bias(c_lo);
    mbar_wait(accf0 + 8 * buf, (lt / (uint32_t)NBUF) & 1);
    tc_fence_after();
    if (q == 0 && lane == 0) dbg_stamp(p, 5 + (int)eset, lt, 0);
    const uint32_t tmem_d = tmem_base + buf * (uint32_t)p.acc_stride + ((uint32_t)(q * 32) << 16);
    if (BULK) {                                                   // the previous tile's bulk store has read the staging rows
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
    auto chunk = [&](const int c0) {
      uint32_t v[32];
      tmem_ld32(tmem_d + (uint32_t)(c0 - n0), v);
      uint4 rcur[4];
      if (RES) {
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) rcur[gq] = rnext[gq];
        if (c0 + 32 < c_hi) load_res(c0 + 32);
      }
      if (BPRE && !CBN && c0 != c_lo) load_bias(c0);                 // (the first chunk's was issued before the accumulator wait)
      tmem_ld_wait();
      if (q == 0 && lane == 0) dbg_stamp(p, 5 + (int)eset, lt, c0 == n0 ? 2 : 4);
      if (BULK && !row_ok) return;
      uint4 o[4];
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) {
        const int c = c0 + gq * 8;
        o[gq] = make_uint4(0u, 0u, 0u, 0u);
        if (c < c_end) {
          const float4 b0 = (BPRE && !CBN) ? bnext[2 * gq] : *reinterpret_cast<const float4*>(s_bias + c);
          const float4 b1 = (BPRE && !CBN) ? bnext[2 * gq + 1] : *reinterpret_cast<const float4*>(s_bias + c + 4);
          float f[8] = {__uint_as_float(v[gq * 8 + 0]) + b0.x, __uint_as_float(v[gq * 8 + 1]) + b0.y,
                        __uint_as_float(v[gq * 8 + 2]) + b0.z, __uint_as_float(v[gq * 8 + 3]) + b0.w,
                        __uint_as_float(v[gq * 8 + 4]) + b1.x, __uint_as_float(v[gq * 8 + 5]) + b1.y,
                        __uint_as_float(v[gq * 8 + 6]) + b1.z, __uint_as_float(v[gq * 8 + 7]) + b1.w};
          if (RES) {
            if (row_ok) {
              const uint4 rv = rcur[gq];
              const float2 r0 = unpack2(rv.x, bf16), r1 = unpack2(rv.y, bf16), r2 = unpack2(rv.z, bf16), r3 = unpack2(rv.w, bf16);
              f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y; f[4] += r2.x; f[5] += r2.y; f[6] += r3.x; f[7] += r3.y;
            }
          }
          if (POOL) {
            // ReLU and the 2 x 2 maximum on the PACKED values: rounding to 16 bits is monotonic, so this is bit for bit the
            // fp32 max followed by the conversion
            // ReLU rides on the conversion (cvt.rn.relu); the window maximum follows below, for the whole chunk, on packed values
            if (c < relu_cend) o[gq] = make_uint4(pack2_relu(f[0], f[1], bf16), pack2_relu(f[2], f[3], bf16), pack2_relu(f[4], f[5], bf16), pack2_relu(f[6], f[7], bf16));
            else o[gq] = make_uint4(pack2(f[0], f[1], bf16), pack2(f[2], f[3], bf16), pack2(f[4], f[5], bf16), pack2(f[6], f[7], bf16));
          } else {
            if (c < relu_cend)                                      // warp-uniform; ReLU rides on the conversion
              o[gq] = make_uint4(pack2_relu(f[0], f[1], bf16), pack2_relu(f[2], f[3], bf16), pack2_relu(f[4], f[5], bf16), pack2_relu(f[6], f[7], bf16));
            else
              o[gq] = make_uint4(pack2(f[0], f[1], bf16), pack2(f[2], f[3], bf16), pack2(f[4], f[5], bf16), pack2(f[6], f[7], bf16));
          }
          if (BULK && store) {
            if (BULK)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + (uint32_t)c * 2u), "r"(o[gq].x), "r"(o[gq].y), "r"(o[gq].z),
                           "r"(o[gq].w) : "memory");
            else
              *reinterpret_cast<uint4*>(out + c) = o[gq];
          }
        }
      }
      if (POOL) {
        // 2 x 2 maximum over lanes ^1 (x) and ^TW (y) as a butterfly that HALVES the data at each step: the even-x lane keeps
        // the first two 8-channel pieces of the chunk and receives its partner's copies of them, the odd-x lane the last two
        // (8 shuffles); the same between the two rows leaves every lane of the window with ONE fully pooled piece (4 shuffles):
        // 12 shuffles per chunk instead of 32, and each lane stores one 16-byte piece — four adjacent lanes write the 64
        // contiguous bytes of the pooled pixel's chunk.  (Shuffles share the shared-memory datapath with the tensor cores'
        // operand fetch, which is what bounds the 64-channel full-resolution layers.)
        const bool odd = (lane & 1) != 0, up = (lane & p.TW) != 0;
        uint4 x2[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint4 send = odd ? o[k] : o[k + 2], keep = odd ? o[k + 2] : o[k];
          x2[k].x = max2_packed(keep.x, __shfl_xor_sync(0xffffffffu, send.x, 1), bf16);
          x2[k].y = max2_packed(keep.y, __shfl_xor_sync(0xffffffffu, send.y, 1), bf16);
          x2[k].z = max2_packed(keep.z, __shfl_xor_sync(0xffffffffu, send.z, 1), bf16);
          x2[k].w = max2_packed(keep.w, __shfl_xor_sync(0xffffffffu, send.w, 1), bf16);
        }
        const uint4 send = up ? x2[0] : x2[1], keep = up ? x2[1] : x2[0];
        uint4 y1;
        y1.x = max2_packed(keep.x, __shfl_xor_sync(0xffffffffu, send.x, p.TW), bf16);
        y1.y = max2_packed(keep.y, __shfl_xor_sync(0xffffffffu, send.y, p.TW), bf16);
        y1.z = max2_packed(keep.z, __shfl_xor_sync(0xffffffffu, send.z, p.TW), bf16);
        y1.w = max2_packed(keep.w, __shfl_xor_sync(0xffffffffu, send.w, p.TW), bf16);
        const int c = c0 + ((odd ? 2 : 0) + (up ? 1 : 0)) * 8;         // the piece this lane ended up with
        if (row_ok && c < c_end) *reinterpret_cast<uint4*>(out + c) = y1;
      }
      if (TR) {
        // 4 x 4 transpose of the 16-byte pieces inside each quad: afterwards lane j of the quad holds piece j of rows 0..3
        const bool hi2 = (lane & 2) != 0, hi1 = (lane & 1) != 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          uint4 t = hi2 ? o[k] : o[k + 2];
          t.x = __shfl_xor_sync(0xffffffffu, t.x, 2); t.y = __shfl_xor_sync(0xffffffffu, t.y, 2);
          t.z = __shfl_xor_sync(0xffffffffu, t.z, 2); t.w = __shfl_xor_sync(0xffffffffu, t.w, 2);
          if (hi2) o[k] = t; else o[k + 2] = t;
        }
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
          uint4 t = hi1 ? o[k] : o[k + 1];
          t.x = __shfl_xor_sync(0xffffffffu, t.x, 1); t.y = __shfl_xor_sync(0xffffffffu, t.y, 1);
          t.z = __shfl_xor_sync(0xffffffffu, t.z, 1); t.w = __shfl_xor_sync(0xffffffffu, t.w, 1);
          if (hi1) o[k] = t; else o[k + 1] = t;
        }
        const int c = c0 + (lane & 3) * 8;
        if (c < c_end) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((okbits >> ((lane & ~3) + k)) & 1u) *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>((uintptr_t)outq[k]) + c) = o[k];
        }
      }
      if (q == 0 && lane == 0) dbg_stamp(p, 5 + (int)eset, lt, c0 == n0 ? 3 : 5);
    };
    if constexpr (CBN > 0) {                                      // tile width known at compile time: chunks unrolled, bias indices constant
#pragma unroll
      for (int c0 = 0; c0 < CBN; c0 += 32)
        if (c0 >= c_lo && c0 < c_hi) chunk(c0);
    } else {
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) chunk(c0);
    }
    if (BULK) {
      // rows of this warp are 32 consecutive pixels of a dense NHWC map: one contiguous block (valid rows are a prefix)
      const uint32_t nrow = (uint32_t)__popc(__ballot_sync(0xffffffffu, row_ok));
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && nrow) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out), "r"(stage_smem), "r"(nrow * (uint32_t)Cout * 2u) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (CL == 2 && rank == 1) mbar_arrive_remote(acce0 + 8 * buf, 0);     // the leader's MMA owns the accumulator hand-off
      else mbar_arrive(acce0 + 8 * buf);
      if (q == 0) dbg_stamp(p, 5 + (int)eset, lt, 1);
    }
  }
  if (BULK && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------
// Epilogue role (shared by the kernels below): epilogue set `eset` (four warps, one per TMEM lane quarter) drains
// accumulator `eset` = local tiles eset, eset + 2, ...: tcgen05.ld, + bias (BatchNorm folded) [+ residual] [ReLU]
// [2x2 max-pool], convert, vectorised NHWC store(s).
template <int CL, bool BPRE = true, int NBUF = 2>       // NBUF = 4 (fast_out only): four accumulators in rotation
__device__ __forceinline__ void epilogue_role(const TcParams& p, const float* s_bias, uint32_t tmem_base, uint32_t accf0, uint32_t acce0,
                                              int warp, int lane, uint32_t eset, int rank, int group0, int ngroups, uint32_t stage_smem = 0u) {
    if (p.fast_out) {
#define CTX_EPI_FAST(POOL, RES, BULK) epilogue_fast_role<CL, POOL, RES, BULK, -1, NBUF, 0, BPRE>(p, s_bias, tmem_base, accf0, acce0, warp, lane, eset, rank, group0, ngroups, stage_smem)
      const bool res = p.residual != nullptr;
      if (p.bulk_out) { if (res) CTX_EPI_FAST(false, true, true); else CTX_EPI_FAST(false, false, true); }
      else if (p.pool2) CTX_EPI_FAST(true, false, false);           // fused pooling never has a residual (tc_supported)
      else if (res) CTX_EPI_FAST(false, true, false);
      else CTX_EPI_FAST(false, false, false);
#undef CTX_EPI_FAST
      return;
    }
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int BN = p.bn;
    const int r = q * 32 + lane;
    const bool bf16 = p.is_bf16 != 0;
    const bool split = !p.bulk_out && p.epi_split;                    // both sets on every tile, half of the chunks each (see above)
    uint32_t lt = split ? 0u : eset;
    for (int tile = group0 + (split ? 0 : (int)eset * ngroups); tile < p.num_tiles; tile += (split ? 1 : 2) * ngroups, lt += (split ? 1u : 2u)) {
      const uint32_t buf = lt & 1;
      const int mg = tile / p.n_tiles_n, n0 = (tile - mg * p.n_tiles_n) * BN, mt = mg * CL + rank;
      int n_img, pix;
      const bool row_ok = tile_row_pixel(p, mt, r, n_img, pix);
      const long long m_lin = (long long)n_img * p.Ho * p.Wo + pix;
      mbar_wait(accf0 + 8 * buf, (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * (uint32_t)p.acc_stride + ((uint32_t)(q * 32) << 16);
      if (p.bulk_out) {                                             // the previous tile's bulk store has read the staging rows
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
      }
      const int nch = (min(p.Cout - n0, BN) + 31) >> 5, first = (nch + 1) >> 1;
      const int cb_lo = split && eset ? first : 0, cb_hi = split && !eset ? first : nch;
#pragma unroll 1
      for (int cb = cb_lo; cb < cb_hi; ++cb) {
        const int c0 = n0 + cb * 32;
        if (c0 >= p.Cout) break;                                   // warp-uniform
        const int lim = BN - cb * 32;                              // columns of this chunk that belong to the tile
        uint32_t v[32];
        tmem_ld32(tmem_d + (uint32_t)(cb * 32), v);
        tmem_ld_wait();
        if (!row_ok && !p.pool2) continue;
        if (p.fast_out) {
          const CtxOutSeg& sg = p.segs.seg[0];
          // pool2: row r of the tile is pixel (r / TW, r % TW) of a TW x 128/TW patch, TW = 16 or 8, so the 2 x 2 window
          // partners are lanes ^1 (x) and ^TW (y) of the same warp; the even/even lane stores the pooled pixel
          long long opix = pix;
          bool store = row_ok;
          if (p.pool2) {
            const int oy = pix / p.Wo, ox = pix - oy * p.Wo;
            opix = (long long)(oy >> 1) * (p.Wo >> 1) + (ox >> 1);
            store = row_ok && !(lane & (1 | p.TW));
          }
          uint16_t* out = reinterpret_cast<uint16_t*>(sg.ptr) + (long long)n_img * sg.img_stride + opix * sg.pix_stride + sg.ch_offset;
          const uint16_t* res = (p.residual && row_ok) ? reinterpret_cast<const uint16_t*>(p.residual) + m_lin * p.res_cstride + p.res_coffset : nullptr;
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const int c = c0 + gq * 8;
            if (c < p.Cout && gq * 8 < lim) {
              const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c);
              const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c + 4);
              float f[8] = {__uint_as_float(v[gq * 8 + 0]) + b0.x, __uint_as_float(v[gq * 8 + 1]) + b0.y,
                            __uint_as_float(v[gq * 8 + 2]) + b0.z, __uint_as_float(v[gq * 8 + 3]) + b0.w,
                            __uint_as_float(v[gq * 8 + 4]) + b1.x, __uint_as_float(v[gq * 8 + 5]) + b1.y,
                            __uint_as_float(v[gq * 8 + 6]) + b1.z, __uint_as_float(v[gq * 8 + 7]) + b1.w};
              if (res) {
                const uint4 rv = *reinterpret_cast<const uint4*>(res + c);
                const float2 r0 = unpack2(rv.x, bf16), r1 = unpack2(rv.y, bf16), r2 = unpack2(rv.z, bf16), r3 = unpack2(rv.w, bf16);
                f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y; f[4] += r2.x; f[5] += r2.y; f[6] += r3.x; f[7] += r3.y;
              }
              if (p.relu && c < p.relu_cend) {           // relu_cend is a multiple of 8 on this path
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
              }
              if (p.pool2) {                             // max is exact in any precision: pool the fp32 values
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  f[e] = fmaxf(f[e], __shfl_xor_sync(0xffffffffu, f[e], 1));
                  f[e] = fmaxf(f[e], __shfl_xor_sync(0xffffffffu, f[e], p.TW));
                }
              }
              if (store) {
                uint4 o;
                o.x = pack2(f[0], f[1], bf16); o.y = pack2(f[2], f[3], bf16); o.z = pack2(f[4], f[5], bf16); o.w = pack2(f[6], f[7], bf16);
                if (p.bulk_out)
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_smem + (uint32_t)(lane * p.Cout + c) * 2u), "r"(o.x), "r"(o.y),
                               "r"(o.z), "r"(o.w) : "memory");
                else
                  *reinterpret_cast<uint4*>(out + c) = o;
              }
            }
          }
        } else if (!row_ok) {
          continue;
        } else if (p.vec_f32) {
          // fp32 segments whose boundaries, strides and offsets are multiples of 4 channels (the fused heads)
#pragma unroll
          for (int gq = 0; gq < 8; ++gq) {
            const int c = c0 + gq * 4;
            if (c < p.Cout && gq * 4 < lim) {
              const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c);
              float4 f = make_float4(__uint_as_float(v[gq * 4 + 0]) + b0.x, __uint_as_float(v[gq * 4 + 1]) + b0.y,
                                     __uint_as_float(v[gq * 4 + 2]) + b0.z, __uint_as_float(v[gq * 4 + 3]) + b0.w);
              if (p.relu && c < p.relu_cend) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
              int sgi = 0;
              if (p.segs.nseg > 1 && c >= p.segs.seg[1].c_begin) sgi = 1;
              if (p.segs.nseg > 2 && c >= p.segs.seg[2].c_begin) sgi = 2;
              const CtxOutSeg& sg = p.segs.seg[sgi];
              float* out = reinterpret_cast<float*>(sg.ptr) + (long long)n_img * sg.img_stride + (long long)pix * sg.pix_stride +
                           sg.ch_offset + (c - sg.c_begin);
              *reinterpret_cast<float4*>(out) = f;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = c0 + j;
            if (c < p.Cout && j < lim) {
              float f = __uint_as_float(v[j]) + s_bias[c];
              if (p.residual) f += load_as(p.residual, m_lin * p.res_cstride + p.res_coffset + c, p.res_dtype);
              if (p.relu && c < p.relu_cend) f = fmaxf(f, 0.f);
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
      }
      if (p.bulk_out) {
        // rows of this warp are 32 consecutive pixels of a dense NHWC map: one contiguous block (valid rows are a prefix)
        const uint32_t nrow = (uint32_t)__popc(__ballot_sync(0xffffffffu, row_ok));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && nrow) {
          const CtxOutSeg& sg = p.segs.seg[0];
          uint16_t* dst = reinterpret_cast<uint16_t*>(sg.ptr) + (long long)n_img * sg.img_stride + (long long)pix * sg.pix_stride + sg.ch_offset;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(stage_smem), "r"(nrow * (uint32_t)p.Cout * 2u)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL == 2 && rank == 1) mbar_arrive_remote(acce0 + 8 * buf, 0);     // the leader's MMA owns the accumulator hand-off
        else mbar_arrive(acce0 + 8 * buf);
      }
    }
    if (p.bulk_out && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace ctx
