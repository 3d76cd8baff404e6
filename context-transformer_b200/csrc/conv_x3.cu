// conv_x3.cu — fp32 convolution emulated on the sm_100a tensor cores (precision 'fp32x3'): the parity-grade mode of the
// conv stack (models/RFB_Net_vgg.py BasicConv :7-22, vgg() :323-343, RFB blocks :26-112, heads :387-416 compute in fp32).
//
// Representation.  Every activation and weight is a pair of fp16 planes, v = hi + lo with hi = fp16(v), lo = fp16(v - hi):
// 22 significand bits.  Weights are scaled per output channel by a power of two (max |w| -> [2048, 4096)) so that their lo
// plane stays in fp16's normal range; the epilogue multiplies the sum by the inverse (exact).
//   a * w  ~=  a_lo * w_hi + a_hi * w_lo + a_hi * w_hi           (the dropped a_lo * w_lo is 2^-22 relative)
//
// Accumulation.  tcgen05.mma's fp32 accumulator TRUNCATES after every K = 16 instruction (profiles/r2_acc_probe.txt: the
// relative error of an all-positive dot product grows as -4e-8 per instruction, -1.5e-5 at K = 4608), which by itself costs
// more accuracy than the 16-bit operands do.  So the tensor core only ever sums SHORT chains: per K-step (64 input channels
// of one filter tap) the three products — small ones first — are chained into a fresh TMEM accumulator (12 instructions),
// and the flush warps add that partial sum to an fp32 running sum held in registers with round-to-nearest adds.  The chain
// of K-steps therefore accumulates exactly like a blocked fp32 summation; CPU emulation of this scheme over the whole
// network (seeded weights) stays within 4e-5 of the fp32 reference on every output.
//   What is left is the truncation INSIDE a chain: the four hi * hi instructions each round the growing partial sum
// towards zero, a systematic shrink of every output by 7e-8 .. 1e-7 (measured, tests/test_gpu_net.py::test_conv_x3_kernel_vs_fp64
// with CTX_X3_COMP=0) — below fp32's resolution per layer but it adds up linearly over the 20+ layers of the trunk (1.4e-6 at
// the heads, 1e-4 absolute on loc values of ~50).  Model: an instruction that brings the partial sum to s loses 3.4e-8 s on
// average (half an ulp, averaged over the mantissa), so a chain of m equal blocks loses 3.4e-8 (m + 1) / 2 of its sum (m = 4
// for a full 64-channel K-step, fewer for a channel tail); the residual bias measured with it is within +-2.5e-8.
// The epilogue undoes the expected loss: sum * (1 + trunc_comp) * scale + bias is rounded ONCE (error-free fp32 arithmetic,
// see finish()) — an unbiased estimate of the exact sum instead of a biased one.
//
// Kernel (448 threads, persistent, one CTA per SM, 128 pixels x BN <= 128 channels per tile):
//   warps 0-3  A producers in modes GATHER (16-byte cp.async im2col, both planes) and STEM (raw fp32 NCHW image -> 27-value
//              patches split into hi / lo rows)
//   warp 4     TMA producer: per K-step one ring slot {A_hi, A_lo, W_hi, W_lo} (mode TMA: activation patches by 4-D TMA)
//   warp 5     MMA issuer: 12 x tcgen05.mma per slot into accumulator (chain & 1), ONE tcgen05.commit per chain
//   warps 6-13 flush / epilogue: warp (q, h) owns TMEM lanes 32q.. and columns 64h..: tcgen05.ld the finished chain, hand the
//              accumulator back, add into 64 registers; after the last chain: * scale + bias [+ residual] [ReLU], split
//              into hi / lo fp16 planes (or fp32 head segments), store.  One of them releases the ring slot to the producers
//              (the commit that says "chain done" also says "slot consumed").
#include "conv_tc.cuh"

#include <stdlib.h>

namespace ctx {

constexpr int X3_ACC_COLS = 128;       // TMEM columns per accumulator (two accumulators)

__device__ __forceinline__ void split_store8(uint16_t* hi, uint16_t* lo, const float (&f)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    // saturate to the fp16 range (a value beyond it has no split representation)
    const float a = fminf(fmaxf(f[2 * e], -65504.f), 65504.f), b = fminf(fmaxf(f[2 * e + 1], -65504.f), 65504.f);
    const __half2 hh = __floats2half2_rn(a, b);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(a - back.x, b - back.y);
    h[e] = *reinterpret_cast<const uint32_t*>(&hh);
    l[e] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int S>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_x3_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a_lo,
               const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int BN = p.bn;
  const uint32_t B_TILE = (uint32_t)BN * TC_BK * 2;
  const uint32_t STAGE = 2u * TC_A_STAGE + 2u * B_TILE;            // {A_hi, A_lo, W_hi, W_lo}
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + S * STAGE;
  const uint32_t full0 = bars, empty0 = bars + 8 * S, accf0 = bars + 16 * S, acce0 = accf0 + 16, tmem_slot = acce0 + 16;
  float* s_bias = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - smem_u32(smem_raw)));       // [cpad] bias, then [cpad] scale
  const int cpad = ((p.Cout + 31) & ~31) + 32;
  float* s_scale = s_bias + cpad;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = p.nk;
  const int group0 = blockIdx.x, ngroups = gridDim.x;

  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tmap_w); if (p.a_mode == A_TMA) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_a_lo); } }
  if (warp == 5) {
    if (lane == 0) {
      const uint32_t full_count = p.a_mode == A_TMA ? 1u : 5u;     // the TMA thread (+ four producer warps)
      for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, full_count); mbar_init(empty0 + 8 * s, 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(accf0 + 8 * b, 1); mbar_init(acce0 + 8 * b, 8); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 2 * X3_ACC_COLS); tmem_relinquish();
  }
  for (int c = threadIdx.x; c < cpad; c += TC_THREADS) {
    s_bias[c] = c < p.Cout ? p.bias[c] : 0.f;
    s_scale[c] = c < p.Cout ? p.out_scale[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");          // programmatic dependent launch (see conv_tc_kernel)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    if (p.a_mode == A_GATHER) {
      // ================= im2col gather, both planes =================
      const int t = threadIdx.x;                 // 0..127
      const int chunk = t & 7, row0 = t >> 3;    // 16-byte chunk of the K-step, rows row0 + 16 i
      const int HoWo = p.Ho * p.Wo;
      const uint16_t* in_hi = reinterpret_cast<const uint16_t*>(p.in);
      const uint16_t* in_lo = reinterpret_cast<const uint16_t*>(p.in_lo);
      uint32_t g = 0;
      for (int tile = group0; tile < p.num_tiles; tile += ngroups) {
        const int m0 = (tile / p.n_tiles_n) * TC_BM;
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
        for (int it = 0; it < nk; ++it, ++g) {
          const uint32_t s = g % S;
          mbar_wait(empty0 + 8 * s, ((g / S) & 1) ^ 1);
          const long long tap_off = (long long)((ky * p.dil) * p.W + kx * p.dil) * p.in_cstride + cc * TC_BK;
          const bool ch_ok = (cc * TC_BK + chunk * 8) < p.Cin;
          const uint32_t dst = smem_base + s * STAGE;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = row0 + 16 * i;
            const bool ok = ch_ok && ((mask[i] >> tap) & 1u);
            const uint32_t off = r * 128 + ((chunk ^ (r & 7)) << 4);
            cp_async_16(dst + off, ok ? (const void*)(in_hi + base[i] + tap_off) : (const void*)in_hi, ok ? 16u : 0u);
            cp_async_16(dst + TC_A_STAGE + off, ok ? (const void*)(in_lo + base[i] + tap_off) : (const void*)in_lo, ok ? 16u : 0u);
          }
          cp_async_commit();
          // one K-step of look-ahead: signal slot g - 1 once its copies have landed
          if (g >= 1u) {
            cp_async_wait<1>();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * ((g - 1) % S));
          }
          ++tap; if (++kx == p.KW) { kx = 0; if (++ky == p.KH) { ky = 0; tap = 0; ++cc; } }
        }
      }
      cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && g >= 1u) mbar_arrive(full0 + 8 * ((g - 1) % S));
    } else if (p.a_mode == A_STEM) {
      // ================= stem: one output pixel per thread, 27 fp32 loads -> hi row + lo row =================
      const int r = threadIdx.x;
      const int HW = p.H * p.W;
      const float* img = reinterpret_cast<const float*>(p.in);
      uint32_t g = 0;
      for (int tile = group0; tile < p.num_tiles; tile += ngroups, ++g) {
        const long long m = (long long)(tile / p.n_tiles_n) * TC_BM + r;
        float v[28];
#pragma unroll
        for (int e = 0; e < 28; ++e) v[e] = 0.f;
        if (m < (long long)p.M) {
          const int n = (int)(m / HW), rem = (int)(m - (long long)n * HW), y = rem / p.W, x = rem - y * p.W;
          const float* base = img + (long long)n * 3 * HW + y * p.W + x;
          const bool yo[3] = {y > 0, true, y + 1 < p.H}, xo[3] = {x > 0, true, x + 1 < p.W};
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              if (yo[ky] && xo[kx]) {
#pragma unroll
                for (int c = 0; c < 3; ++c) v[(ky * 3 + kx) * 3 + c] = __ldg(base + c * HW + (ky - 1) * p.W + (kx - 1));
              }
        }
        const uint32_t s = g % S;
        mbar_wait(empty0 + 8 * s, ((g / S) & 1) ^ 1);
        const uint32_t row = smem_base + s * STAGE + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint32_t h[4] = {0u, 0u, 0u, 0u}, l[4] = {0u, 0u, 0u, 0u};
          if (ch < 4) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (ch * 8 + 2 * e < 28) {
                const float a = v[ch * 8 + 2 * e], b = (ch * 8 + 2 * e + 1 < 28) ? v[ch * 8 + 2 * e + 1] : 0.f;
                const __half2 hh = __floats2half2_rn(a, b);
                const float2 back = __half22float2(hh);
                const __half2 ll = __floats2half2_rn(a - back.x, b - back.y);
                h[e] = *reinterpret_cast<const uint32_t*>(&hh);
                l[e] = *reinterpret_cast<const uint32_t*>(&ll);
              }
            }
          }
          const uint32_t off = (uint32_t)((ch ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + TC_A_STAGE + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * s);
      }
    }
  } else if (warp == 4) {
    // ================= TMA producer: weights (both planes); activations too in mode TMA =================
    uint32_t g = 0;
    const uint32_t tx_bytes = 2u * B_TILE + (p.a_mode == A_TMA ? 2u * (uint32_t)(p.TW * p.TH * 128) : 0u);
    const int lo_k = p.nk * TC_BK;                          // K offset of the lo weight plane
    for (int tile = group0; tile < p.num_tiles; tile += ngroups) {
      const int mt = tile / p.n_tiles_n, n0 = (tile - mt * p.n_tiles_n) * BN;
      int n_img = 0, x0 = 0, y0 = 0;
      if (p.a_mode == A_TMA) {
        const int per_img = p.tiles_x * p.tiles_y;
        n_img = mt / per_img;
        const int t = mt - n_img * per_img;
        y0 = (t / p.tiles_x) * p.TH * p.stride - p.pad_h;       // input coordinates of the patch's first output pixel
        x0 = (t % p.tiles_x) * p.TW * p.stride - p.pad_w;
      }
      int cc = 0, ky = 0, kx = 0;
      for (int it = 0; it < nk; ++it, ++g) {
        const uint32_t s = g % S;
        mbar_wait(empty0 + 8 * s, ((g / S) & 1) ^ 1);
        if (elect_one()) {
          const uint32_t dst = smem_base + s * STAGE, bar = full0 + 8 * s;
          const int kw = ((ky * p.KW + kx) * p.cin_blocks + cc) * TC_BK;
          mbar_arrive_expect_tx(bar, tx_bytes);
          tma_load_2d(dst + 2 * TC_A_STAGE, &tmap_w, kw, n0, bar);
          tma_load_2d(dst + 2 * TC_A_STAGE + B_TILE, &tmap_w, lo_k + kw, n0, bar);
          if (p.a_mode == A_TMA) {
            tma_load_4d(dst, &tmap_a, p.in_coffset + cc * TC_BK, x0 + kx * p.dil, y0 + ky * p.dil, n_img, bar);
            tma_load_4d(dst + TC_A_STAGE, &tmap_a_lo, p.in_coffset + cc * TC_BK, x0 + kx * p.dil, y0 + ky * p.dil, n_img, bar);
          }
        }
        __syncwarp();
        if (++kx == p.KW) { kx = 0; if (++ky == p.KH) { ky = 0; ++cc; } }
      }
    }
  } else if (warp == 5) {
    // ================= MMA issuer: one 12-instruction chain per K-step into accumulator (chain & 1) =================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(false, TC_BM, BN);
      const uint64_t a_hi0 = make_sw128_desc(smem_base), stage_step = (uint64_t)(STAGE >> 4);
      const uint64_t lo_off = (uint64_t)(TC_A_STAGE >> 4), w_off = (uint64_t)((2 * TC_A_STAGE) >> 4), wlo_off = w_off + (uint64_t)(B_TILE >> 4);
      uint64_t ad = a_hi0;
      uint32_t s = 0, ph = 0, chain = 0;
      for (int tile = group0; tile < p.num_tiles; tile += ngroups) {
        for (int it = 0; it < nk; ++it, ++chain) {
          const uint32_t buf = chain & 1;
          mbar_wait(acce0 + 8 * buf, ((chain >> 1) & 1) ^ 1);          // the flush warps have read chain - 2
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * (uint32_t)X3_ACC_COLS;
          const uint64_t a_hi = ad, a_lo = ad + lo_off, w_hi = ad + w_off, w_lo = ad + wlo_off;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) umma_f16(tmem_d, a_lo + 2 * k, w_hi + 2 * k, idesc, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) umma_f16(tmem_d, a_hi + 2 * k, w_lo + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) umma_f16(tmem_d, a_hi + 2 * k, w_hi + 2 * k, idesc, 1u);
          umma_commit(accf0 + 8 * buf);                                // chain complete == ring slot consumed
          ad += stage_step;
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; ad = a_hi0; }
        }
      }
    }
    __syncwarp();
  } else {
    // ================= flush + epilogue warps =================
    const int e = warp - 6, q = warp & 3, h = e >> 2;        // TMEM lane quarter (hardware: warp % 4), column half
    const int r = q * 32 + lane;
    const bool releaser = e == 0;                             // forwards "slot consumed" to the producers
    uint32_t chain = 0, s = 0;
    for (int tile = group0; tile < p.num_tiles; tile += ngroups) {
      const int mt = tile / p.n_tiles_n, n0 = (tile - mt * p.n_tiles_n) * BN;
      // the two warps of a TMEM lane quarter split the tile's columns: 64 + rest for wide tiles, 32 + rest for tiles <= 64 wide
      // (so that all eight warps share the flush and the epilogue of the narrow stem / conv1_2 tiles)
      const int cs = BN <= 64 ? 32 : 64;
      const int cbase = n0 + h * cs;                         // first channel of this warp's columns
      const int c_end = min(p.Cout, n0 + BN);
      const bool active = h * cs < BN && cbase < c_end;      // warp-uniform
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.f;
      for (int it = 0; it < nk; ++it, ++chain) {
        const uint32_t buf = chain & 1;
        mbar_wait(accf0 + 8 * buf, (chain >> 1) & 1);
        tc_fence_after();
        if (releaser && lane == 0) mbar_arrive(empty0 + 8 * s);
        if (++s == (uint32_t)S) s = 0;
        uint32_t v[32];
        const uint32_t taddr = tmem_base + buf * (uint32_t)X3_ACC_COLS + (uint32_t)(h * cs) + ((uint32_t)(q * 32) << 16);
        const bool second = active && cs == 64;
        if (active) {
          tmem_ld32(taddr, v);
          tmem_ld_wait();
          if (second) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(v[i]);
            tmem_ld32(taddr + 32u, v);
            tmem_ld_wait();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acce0 + 8 * buf);
        if (active) {
          if (second) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[32 + i] += __uint_as_float(v[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(v[i]);
          }
        }
      }
      if (!active) continue;
      int n_img, pix;
      const bool row_ok = tile_row_pixel(p, mt, r, n_img, pix);
      if (!row_ok) continue;
      const long long m_lin = (long long)n_img * p.Ho * p.Wo + pix;
      const int relu_cend = p.relu ? p.relu_cend : 0;
      // sum * (1 + comp) * scale + bias rounded ONCE, in fp32 only (fp64 conversions run on the quarter-rate XU pipe: with them
      // the stem's epilogue ran that pipe at 120 % — ncu): scale is a power of two, so a = sum * scale is exact; the compensation
      // t = a * comp is far below an ulp of the result and rides on the rounding error of a + bias (TwoSum), which is added back
      // before the final rounding.
      const float comp = p.trunc_comp;
      auto finish = [&](float acc_v, int c) {
        const float a = __fmul_rn(acc_v, s_scale[c]), b = s_bias[c];
        const float s = __fadd_rn(a, b);
        const float bb = __fsub_rn(s, a);
        const float err = __fadd_rn(__fsub_rn(a, __fsub_rn(s, bb)), __fsub_rn(b, bb));
        return __fadd_rn(s, __fadd_rn(err, __fmul_rn(a, comp)));
      };
      if (p.fast_out) {
        // one 16-bit output tensor: hi / lo planes, 8 channels (16 bytes) per store
        const CtxOutSeg& sg = p.segs.seg[0];
        const long long o = (long long)n_img * sg.img_stride + (long long)pix * sg.pix_stride + sg.ch_offset;
        uint16_t* out_hi = reinterpret_cast<uint16_t*>(sg.ptr) + o;
        uint16_t* out_lo = reinterpret_cast<uint16_t*>(p.out_lo) + o;
        const uint16_t* res_hi = p.residual ? reinterpret_cast<const uint16_t*>(p.residual) + m_lin * p.res_cstride + p.res_coffset : nullptr;
        const uint16_t* res_lo = p.residual ? reinterpret_cast<const uint16_t*>(p.residual_lo) + m_lin * p.res_cstride + p.res_coffset : nullptr;
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8) {
          const int c = cbase + g8 * 8;
          if (g8 * 8 < cs && c < c_end) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = finish(acc[g8 * 8 + j], c + j);
            if (res_hi) {
              const uint4 rh = __ldg(reinterpret_cast<const uint4*>(res_hi + c)), rl = __ldg(reinterpret_cast<const uint4*>(res_lo + c));
              const uint32_t hw[4] = {rh.x, rh.y, rh.z, rh.w}, lw[4] = {rl.x, rl.y, rl.z, rl.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[j])), b = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
                f[2 * j] += a.x + b.x; f[2 * j + 1] += a.y + b.y;
              }
            }
            if (c < relu_cend) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            split_store8(out_hi + c, out_lo + c, f);
          }
        }
      } else if (!p.vec_f32) {
        // fp32 segments of any alignment: one channel at a time
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const int c = cbase + j;
          if (j < cs && c < c_end) {
            float f = finish(acc[j], c);
            if (c < relu_cend) f = fmaxf(f, 0.f);
            int sgi = 0;
            if (p.segs.nseg > 1 && c >= p.segs.seg[1].c_begin) sgi = 1;
            if (p.segs.nseg > 2 && c >= p.segs.seg[2].c_begin) sgi = 2;
            const CtxOutSeg& sg = p.segs.seg[sgi];
            reinterpret_cast<float*>(sg.ptr)[(long long)n_img * sg.img_stride + (long long)pix * sg.pix_stride + sg.ch_offset + (c - sg.c_begin)] = f;
          }
        }
      } else {
        // fp32 segments, 4-channel aligned (the fused loc / conf / obj heads write into the concatenated [B,P,*] buffers)
#pragma unroll
        for (int g4 = 0; g4 < 16; ++g4) {
          const int c = cbase + g4 * 4;
          if (g4 * 4 < cs && c < c_end) {
            float4 f = make_float4(finish(acc[g4 * 4 + 0], c + 0), finish(acc[g4 * 4 + 1], c + 1), finish(acc[g4 * 4 + 2], c + 2), finish(acc[g4 * 4 + 3], c + 3));
            if (c < relu_cend) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
            int sgi = 0;
            if (p.segs.nseg > 1 && c >= p.segs.seg[1].c_begin) sgi = 1;
            if (p.segs.nseg > 2 && c >= p.segs.seg[2].c_begin) sgi = 2;
            const CtxOutSeg& sg = p.segs.seg[sgi];
            float* out = reinterpret_cast<float*>(sg.ptr) + (long long)n_img * sg.img_stride + (long long)pix * sg.pix_stride + sg.ch_offset + (c - sg.c_begin);
            *reinterpret_cast<float4*>(out) = f;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * X3_ACC_COLS); }
}

struct X3Plan {
  CUtensorMap tmap_w, tmap_a, tmap_a_lo;
  TcParams p;
  int stages, grid;
  size_t smem;
};

static int x3_supported(const CtxConvParams* p) {
  if (!p || !p->split || !p->out_scale || !p->bias) return 0;
  if (p->pool2 || p->KH * p->KW > 32 || p->nseg < 1 || p->nseg > 3) return 0;
  if (((uintptr_t)p->weight) % 16 || ((uintptr_t)p->bias) % 16) return 0;
  if (p->in_nchw) {
    if (!(p->in_dtype == CTX_F32 && p->Cin == 3 && p->KH == 3 && p->KW == 3 && p->stride == 1 && p->pad_h == 1 && p->pad_w == 1 && p->dil == 1)) return 0;
  } else {
    if (p->in_dtype != CTX_F16 || !p->in_lo || p->Cin % 8 || p->in_cstride % 8 || p->in_coffset % 8) return 0;
    if (((uintptr_t)p->in) % 16 || ((uintptr_t)p->in_lo) % 16) return 0;
  }
  const CtxOutSeg& s0 = p->seg[0];
  const int relu_cend = p->relu_channels > 0 ? p->relu_channels : p->Cout;
  const bool half_out = p->nseg == 1 && s0.dtype == CTX_F16 && p->out_lo && s0.c_begin == 0 && s0.c_end == p->Cout && p->Cout % 8 == 0 &&
                        relu_cend % 8 == 0 && s0.pix_stride % 8 == 0 && s0.ch_offset % 8 == 0 && s0.img_stride % 8 == 0 &&
                        ((uintptr_t)s0.ptr) % 16 == 0 && ((uintptr_t)p->out_lo) % 16 == 0 &&
                        (!p->residual || (p->residual_lo && p->res_dtype == CTX_F16 && p->res_cstride % 8 == 0 && p->res_coffset % 8 == 0 &&
                                          ((uintptr_t)p->residual) % 16 == 0 && ((uintptr_t)p->residual_lo) % 16 == 0));
  if (half_out) return 1;
  if (p->residual) return 0;
  int expect = 0;
  for (int s = 0; s < p->nseg; ++s) {                       // fp32 segments covering [0, Cout) in order
    const CtxOutSeg& sg = p->seg[s];
    if (!(sg.dtype == CTX_F32 && sg.c_begin == expect && sg.c_end > sg.c_begin && sg.ptr)) return 0;
    expect = sg.c_end;
  }
  return expect == p->Cout;
}

// fp32 segments whose boundaries, strides and base addresses allow 128-bit stores
static bool x3_vec_f32(const CtxConvParams* p) {
  const int relu_cend = p->relu_channels > 0 ? p->relu_channels : p->Cout;
  if (relu_cend % 4 || p->Cout % 4) return false;
  for (int s = 0; s < p->nseg; ++s) {
    const CtxOutSeg& sg = p->seg[s];
    if (sg.c_begin % 4 || sg.c_end % 4 || sg.pix_stride % 4 || sg.ch_offset % 4 || sg.img_stride % 4 || ((uintptr_t)sg.ptr) % 16) return false;
  }
  return true;
}

template <int S>
static int launch_x3(const X3Plan* pl, cudaStream_t st) {
  CTX_CUDA_TRY(cudaFuncSetAttribute(conv_x3_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)pl->grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = pl->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int pdl = [] { const char* e = getenv("CTX_CONV_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
  attr[0].val.programmaticStreamSerializationAllowed = pdl;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CTX_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_x3_kernel<S>, pl->tmap_w, pl->tmap_a, pl->tmap_a_lo, pl->p));
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

}  // namespace ctx

using namespace ctx;

extern "C" int ctx_conv2d_x3_supported(const CtxConvParams* p) { return x3_supported(p); }

extern "C" int ctx_conv2d_x3_plan_create(const CtxConvParams* p, int n_tiles_n, void** plan_out) {
  CTX_REQUIRE(p && plan_out, "ctx_conv2d_x3_plan_create: null argument");
  *plan_out = nullptr;
  if (!x3_supported(p)) { set_error("ctx_conv2d_x3: geometry / layout not supported by the split-precision tcgen05 path"); return CTX_ERR_UNSUPPORTED; }
  CTX_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->Cin > 0 && p->Cout > 0 && p->stride > 0 && p->dil > 0, "conv_x3: bad dims");
  const int ho = (p->H + 2 * p->pad_h - p->dil * (p->KH - 1) - 1) / p->stride + 1;
  const int wo = (p->W + 2 * p->pad_w - p->dil * (p->KW - 1) - 1) / p->stride + 1;
  CTX_REQUIRE(ho == p->Ho && wo == p->Wo, "conv_x3: Ho/Wo inconsistent with geometry");
  CTX_REQUIRE((long long)p->N * p->Ho * p->Wo < (1ll << 31), "conv_x3: too many output pixels");

  X3Plan* pl = new X3Plan();
  memset(&pl->tmap_a, 0, sizeof pl->tmap_a);
  memset(&pl->tmap_a_lo, 0, sizeof pl->tmap_a_lo);
  TcParams& t = pl->p;
  memset(&t, 0, sizeof t);
  t.in = p->in; t.in_lo = p->in_lo; t.bias = p->bias; t.residual = p->residual; t.residual_lo = p->residual_lo;
  t.out_lo = p->out_lo; t.out_scale = p->out_scale;
  t.N = p->N; t.H = p->H; t.W = p->W; t.Cin = p->Cin; t.in_cstride = p->in_cstride; t.in_coffset = p->in_coffset;
  t.Cout = p->Cout; t.KH = p->KH; t.KW = p->KW; t.stride = p->stride; t.pad_h = p->pad_h; t.pad_w = p->pad_w; t.dil = p->dil;
  t.Ho = p->Ho; t.Wo = p->Wo; t.relu = p->relu;
  t.relu_cend = p->relu_channels > 0 ? p->relu_channels : p->Cout;
  t.M = p->N * p->Ho * p->Wo;
  t.cin_blocks = (p->Cin + TC_BK - 1) / TC_BK;
  t.nk = p->in_nchw ? 1 : p->KH * p->KW * t.cin_blocks;
  t.res_cstride = p->res_cstride; t.res_coffset = p->res_coffset; t.res_dtype = p->res_dtype;
  t.segs.nseg = p->nseg;
  for (int s = 0; s < 3; ++s) t.segs.seg[s] = p->seg[s < p->nseg ? s : 0];
  t.fast_out = p->seg[0].dtype == CTX_F16;
  t.vec_f32 = !t.fast_out && x3_vec_f32(p);
  t.n_tiles_n = std::max(cdiv(p->Cout, X3_ACC_COLS), std::min(n_tiles_n, cdiv(p->Cout, 16)));
  t.bn = (cdiv(p->Cout, t.n_tiles_n) + 15) / 16 * 16;
  t.n_tiles_n = cdiv(p->Cout, t.bn);
  t.cluster = 1; t.occ = 1; t.acc_stride = X3_ACC_COLS;
  {
    // expected truncation loss per chain, averaged over the channel blocks weighted by their real channels
    const int cin = p->in_nchw ? 27 : p->Cin;
    double w = 0.0;
    for (int c0 = 0; c0 < cin; c0 += TC_BK) {
      const int real = std::min(TC_BK, cin - c0), m = (real + 15) / 16;
      w += real * 0.5 * (m + 1);
    }
    const char* e = getenv("CTX_X3_COMP");              // development aid: CTX_X3_COMP=0 disables the compensation
    t.trunc_comp = (e && e[0] == '0') ? 0.f : (float)(3.4e-8 * w / cin);
  }

  int tw = 0, th = 0;
  if (p->in_nchw) t.a_mode = A_STEM;
  else if (flat_eligible(p)) { t.a_mode = A_TMA; tw = 128; th = 1; }
  else t.a_mode = choose_patch(p, &tw, &th, 150, 2) ? A_TMA : A_GATHER;      // (stride 2: the tensor maps walk the input with a traversal stride)
  t.flat = t.a_mode == A_TMA && flat_eligible(p);
  t.TW = tw; t.TH = th;
  const bool patches = t.a_mode == A_TMA;
  t.tiles_x = patches ? (t.flat ? cdiv(t.M, 128) : cdiv(p->Wo, tw)) : 0;
  t.tiles_y = patches ? (t.flat ? 1 : cdiv(p->Ho, th)) : 0;
  t.m_tiles = patches ? (t.flat ? t.tiles_x : p->N * t.tiles_x * t.tiles_y) : cdiv(t.M, TC_BM);
  t.num_tiles = t.m_tiles * t.n_tiles_n;

  const size_t stage = 2 * (size_t)TC_A_STAGE + 2 * (size_t)t.bn * TC_BK * 2;
  const size_t fixed = 16 * 6 + 64 + 2 * 4 * (((size_t)p->Cout + 31) / 32 * 32 + 32) + 1024;
  pl->stages = (int)std::min<size_t>(6, (232448 - fixed) / stage);
  if (pl->stages == 5) pl->stages = 4;
  if (pl->stages < 2) { delete pl; set_error("conv_x3: ring does not fit shared memory (Cout %d)", p->Cout); return CTX_ERR_UNSUPPORTED; }
  pl->smem = (size_t)pl->stages * stage + fixed;
  pl->grid = std::min(t.num_tiles, num_sms());

  // weights: [Cout_pad][2][KH*KW*Cin_pad] fp16 (hi plane, lo plane), K-major
  const unsigned long long ktot = 2ull * (unsigned long long)t.nk * TC_BK;
  const unsigned long long cout_pad = (unsigned long long)((p->Cout + 15) / 16 * 16);
  int rc = encode_2d_sw128(&pl->tmap_w, p->weight, false, cout_pad, ktot, (unsigned)t.bn);
  if (!rc && t.a_mode == A_TMA) {
    if (t.flat) {
      rc = encode_nhwc_sw128(&pl->tmap_a, p->in, false, 1, 1, t.M, p->in_cstride, p->in_coffset + p->Cin, 128u, 1u);
      if (!rc) rc = encode_nhwc_sw128(&pl->tmap_a_lo, p->in_lo, false, 1, 1, t.M, p->in_cstride, p->in_coffset + p->Cin, 128u, 1u);
    } else {
      rc = encode_nhwc_sw128(&pl->tmap_a, p->in, false, p->N, p->H, p->W, p->in_cstride, p->in_coffset + p->Cin, (unsigned)tw, (unsigned)th, (unsigned)p->stride);
      if (!rc) rc = encode_nhwc_sw128(&pl->tmap_a_lo, p->in_lo, false, p->N, p->H, p->W, p->in_cstride, p->in_coffset + p->Cin, (unsigned)tw, (unsigned)th,
                                      (unsigned)p->stride);
    }
  }
  if (rc) { delete pl; return rc; }
  *plan_out = pl;
  return CTX_OK;
}

extern "C" int ctx_conv2d_x3_plan_info(void* plan, int* info8) {
  CTX_REQUIRE(plan && info8, "ctx_conv2d_x3_plan_info: null argument");
  const X3Plan* pl = (const X3Plan*)plan;
  info8[0] = pl->p.bn; info8[1] = pl->p.n_tiles_n; info8[2] = 1; info8[3] = pl->p.a_mode; info8[4] = pl->stages; info8[5] = pl->grid;
  info8[6] = 1; info8[7] = pl->p.TW * 1000 + pl->p.TH;
  return CTX_OK;
}

extern "C" int ctx_conv2d_x3_plan_run(void* plan, void* stream) {
  CTX_REQUIRE(plan, "ctx_conv2d_x3_plan_run: null plan");
  const X3Plan* pl = (const X3Plan*)plan;
  cudaStream_t st = (cudaStream_t)stream;
  switch (pl->stages) {
    case 6: return launch_x3<6>(pl, st);
    case 4: return launch_x3<4>(pl, st);
    case 3: return launch_x3<3>(pl, st);
    default: return launch_x3<2>(pl, st);
  }
}

extern "C" void ctx_conv2d_x3_plan_destroy(void* plan) { delete (X3Plan*)plan; }
