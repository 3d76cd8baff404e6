// conv_tc.cu — persistent implicit-GEMM convolution on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// Every conv of the detector (models/RFB_Net_vgg.py BasicConv :7-22, vgg() :323-343, RFB branches
// :26-112, heads :387-416) with 16-bit activations is one launch of this kernel:
//
//   D[m, co] = sum_{tap, ci} A[m, (tap, ci)] * Wt[co, (tap, ci)]          fp32 accumulate in TMEM
//   A  = the NHWC activation gathered on the fly (im2col never exists in HBM)
//
// conv_tc_kernel: one CTA per SM walks a static round-robin list of (128 output pixels) x (BN output channels) tiles;
// K is walked channel block by channel block, taps innermost, in steps of 64 channels of one filter tap through a
// ring of S shared-memory stages {A tile 16 KB, B tile BN*128 B} released in commit groups.  Warp roles (448 threads):
//   warps 0-3  im2col gather producers (mode GATHER): 16-byte cp.async with zero-fill for padding /
//              channel tails, written straight into the 128B-swizzled K-major layout of the UMMA
//              descriptor.  M is the flattened (n, oy, ox) pixel index of the whole batch: no tile waste
//              on the small pyramid levels, any stride / dilation / kernel shape.
//              Mode STEM (Cin = 3 first layer): the same warps read the raw fp32 NCHW image, build each pixel's
//              27-value 3x3x3 patch in registers and st.shared it as one 64-channel K-step row, so conv1_1
//              runs on the tensor cores without an im2col tensor in HBM.
//   warp 4     TMA producer: the BN x 64 weight tile of each K-step (cp.async.bulk.tensor.2d), and in
//              mode TMA also the activation tile: the output tile is a TW x TH <= 128 pixel patch of one image
//              and the A tile of tap (ky, kx) is the 4-D box {64 ch, TW, TH, 1} at
//              (x0 + kx*dil - pad, y0 + ky*dil - pad); out-of-image pixels and channels past the slice are
//              zero-filled by the TMA unit, i.e. conv padding costs nothing and no thread touches an address
//              (stride-1 convs; 1x1 convs see the whole batch as one pixel row: "flat" 128-pixel runs).
//   warp 5     TMEM allocator + MMA issuer: one elected lane issues 4 x tcgen05.mma (M128 x N=BN x K16,
//              kind::f16) per K-step, tcgen05.commit's ring slots back to the producers and, after the
//              last K-step of a tile, the accumulator to the epilogue.  Two accumulators (2 x 256 TMEM
//              columns) let tile i+1 start while tile i drains.  CL = 2: CTA pairs, one cta_group::2 MMA (M = 256)
//              issued by the leader, each CTA staging half of the weight tile.
//   warps 6-13 epilogue, two sets of four warps (set 0 drains accumulator 0 = even tiles, set 1 accumulator 1 =
//              odd tiles, so two tiles drain concurrently): tcgen05.ld the accumulator, + bias (BatchNorm folded)
//              [+ residual] [ReLU] [2x2 max-pool], convert, NHWC store — compile-time-specialised for the 16-bit
//              single-segment case (residual prefetch, sector-aligned transposed stores, bulk-copy rows), generic
//              for the up to three fp32 segments of the fused loc / conf / obj heads (writes land directly in the
//              concatenated [B,P,*] buffers).
// conv_halo_kernel (3x3 stride-1 convs): the tile's input neighbourhood is staged once and read through nine shifted
// descriptor windows; optionally two CTAs per SM.  See its own header below.
// Which kernel / tiling a layer uses is decided per layer by measurement (ctx_prog_autotune); every choice gives
// bit-identical results.
#include "conv_tc_epilogue.cuh"

#include <stdlib.h>

namespace ctx {


// ---------------------------------------------------------------------------------------------------
template <int S, int CL>       // ring stages; CTAs per MMA (1, or 2 = cta_group::2 pairs in a cluster of two)
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a, const TcParams p) {
  constexpr int LOOKAHEAD = S - 2;      // gathers that may still be in flight when the next one is issued
  extern __shared__ uint8_t smem_raw[];
  const int BN = p.bn;
  const int B_STAGE = (BN / CL) * TC_BK * 2;           // CTA pairs: each CTA stages its own N-half of the weight tile
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B needs 1024-B alignment
  const uint32_t sA = smem_base, sB = smem_base + S * TC_A_STAGE;
  const uint32_t bars = sB + S * B_STAGE;
  const uint32_t full0 = bars, empty0 = bars + 8 * S, accf0 = bars + 24 * S, acce0 = accf0 + 16,
                 tmem_slot = acce0 + 16;
  // per-channel epilogue vector (bias with BatchNorm folded), staged once per CTA: [ceil32(Cout) + 32] floats
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 24 * S + 64 - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = p.nk;
  // A tcgen05.commit costs the issuing thread ~350 clk in steady state ({4 MMA + commit} = 546 clk whatever N,
  // profiles/r1_mma_probe.txt): with narrow N tiles one commit per K-step would bound the kernel, so ring slots are
  // released in groups of 1 << clog K-steps (one commit, one EMPTY barrier per group).
  const uint32_t cmask = (1u << p.clog) - 1u;
  const int rank = CL == 2 ? (int)cluster_ctarank() : 0;
  const int group0 = blockIdx.x / CL, ngroups = gridDim.x / CL;   // persistent walk over tile groups

  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tmap_w); if (p.a_mode == A_TMA) tma_prefetch_desc(&tmap_a); }
  if (warp == 5) {
    if (lane == 0) {
      // arrivals per phase: the (leader's) TMA thread, plus four producer warps per CTA unless the TMA unit stages A too;
      // in pair mode every producer of either CTA signals the LEADER's barrier (the peer's own full barriers stay unused)
      const uint32_t full_count = p.a_mode == A_TMA ? 1u : 1u + 4u * (uint32_t)CL;
      for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, full_count); mbar_init(empty0 + 8 * s, 1); }
      // pair mode: the leader's MMA also waits for the peer's four epilogue warps (remote arrivals)
      // (accumulator drained: both epilogue sets work on every tile, except with the bulk-copy output path — conv_tc_epilogue.cuh)
      for (int b = 0; b < 2; ++b) { mbar_init(accf0 + 8 * b, 1); mbar_init(acce0 + 8 * b, (p.bulk_out || !p.epi_split ? 4u : 8u) * (uint32_t)CL); }
      fence_barrier_init();
    }
    __syncwarp();
    if (CL == 2) { tmem_alloc_2cta(tmem_slot, 512); tmem_relinquish_2cta(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  for (int c = threadIdx.x; c < ((p.Cout + 31) & ~31) + 32; c += TC_THREADS) s_bias[c] = c < p.Cout ? p.bias[c] : 0.f;
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();                          // peer barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch, bias staging)
  // overlapped the tail of the previous kernel in the stream; its outputs are visible after the wait.  Let the next
  // kernel start its own prologue as soon as SMs drain.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto arrive_full = [&](uint32_t bar) {            // producer-warp arrival on the barrier the MMA issuer waits on
    if (CL == 2 && rank == 1) mbar_arrive_remote(bar, 0); else mbar_arrive(bar);
  };

  if (warp < 4) {
    // ================= im2col gather producers =================
    if (p.a_mode == A_GATHER) {
      const int t = threadIdx.x;                 // 0..127
      const int chunk = t & 7;                   // 16-byte (8-channel) chunk of the 64-channel K-step
      const int row0 = t >> 3;                   // rows row0 + 16*i, i = 0..7
      const int HoWo = p.Ho * p.Wo;
      const uint16_t* in = reinterpret_cast<const uint16_t*>(p.in);
      uint32_t g = 0;                            // K-steps issued so far (across tiles)
      for (int tile = group0; tile < p.num_tiles; tile += ngroups) {
        const int m0 = ((tile / p.n_tiles_n) * CL + rank) * TC_BM;
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
          if ((s & cmask) == 0) mbar_wait(empty0 + 8 * (s >> p.clog), ((g / S) & 1) ^ 1);
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
          if (g >= (uint32_t)LOOKAHEAD) {
            cp_async_wait<LOOKAHEAD>();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) arrive_full(full0 + 8 * ((g - LOOKAHEAD) % S));
          }
          ++tap; if (++kx == p.KW) { kx = 0; if (++ky == p.KH) { ky = 0; tap = 0; ++cc; } }      // taps innermost, then channel blocks
        }
      }
      cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0)
        for (uint32_t q = (g > (uint32_t)LOOKAHEAD ? g - LOOKAHEAD : 0u); q < g; ++q) arrive_full(full0 + 8 * (q % S));
    } else if (p.a_mode == A_STEM) {
      // one output pixel per thread: 27 coalesced fp32 loads of the 3x3x3 neighbourhood -> one 128-byte K-step row
      const int r = threadIdx.x;                 // 0..127
      const int HW = p.H * p.W;
      const float* img = reinterpret_cast<const float*>(p.in);
      const bool bf16 = p.is_bf16 != 0;
      uint32_t g = 0;
      // Software pipelining: the 27 loads of the tile after next are issued before the current tile is packed and stored
      // (with one pixel per thread there is no other memory-level parallelism in this role).  The role is paced by its own
      // instruction stream, so the loop is kept lean: the pixel cursor (n, y, x) advances incrementally (no division per
      // tile), padding is decided by six predicates per pixel, and three register buffers rotate by unrolling (no copies).
      // (STEM tiles are never split over N or CTA pairs: tile == M tile.)
      const int step = ngroups * TC_BM, step_rows = step / p.W, step_cols = step - step_rows * p.W;
      long long m_cur = (long long)group0 * TC_BM + r;            // pixel the NEXT load_patch call fetches
      int cn = (int)(m_cur / HW), cy = (int)((m_cur - (long long)cn * HW) / p.W), cx = (int)(m_cur - (long long)cn * HW - (long long)cy * p.W);
      auto load_patch = [&](float (&v)[28]) {
#pragma unroll
        for (int e = 0; e < 28; ++e) v[e] = 0.f;
        if (m_cur < (long long)p.M) {
          const float* base = img + (long long)cn * 3 * HW + cy * p.W + cx;
          const bool yo[3] = {cy > 0, true, cy + 1 < p.H}, xo[3] = {cx > 0, true, cx + 1 < p.W};
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              if (yo[ky] && xo[kx]) {
#pragma unroll
                for (int c = 0; c < 3; ++c) v[(ky * 3 + kx) * 3 + c] = __ldg(base + c * HW + (ky - 1) * p.W + (kx - 1));
              }
        }
        m_cur += step;
        cx += step_cols; cy += step_rows;
        if (cx >= p.W) { cx -= p.W; ++cy; }
        while (cy >= p.H) { cy -= p.H; ++cn; }
      };
      auto emit = [&](const float (&v)[28]) {
        const uint32_t s = g % S;
        if ((s & cmask) == 0) mbar_wait(empty0 + 8 * (s >> p.clog), ((g / S) & 1) ^ 1);
        const uint32_t row = sA + s * TC_A_STAGE + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
          if (ch < 4) {
            w0 = pack2(v[ch * 8 + 0], v[ch * 8 + 1], bf16); w1 = pack2(v[ch * 8 + 2], v[ch * 8 + 3], bf16);
            if (ch < 3) { w2 = pack2(v[ch * 8 + 4], v[ch * 8 + 5], bf16); w3 = pack2(v[ch * 8 + 6], v[ch * 8 + 7], bf16); }
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((ch ^ (r & 7)) << 4)), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) arrive_full(full0 + 8 * s);
        ++g;
      };
      float va[28], vb[28], vc[28];                // two tiles of loads in flight: a tile period is shorter than the load latency
      load_patch(va);
      load_patch(vb);
      for (int tile = group0; tile < p.num_tiles;) {
        load_patch(vc); emit(va); tile += ngroups;
        if (tile >= p.num_tiles) break;
        load_patch(va); emit(vb); tile += ngroups;
        if (tile >= p.num_tiles) break;
        load_patch(vb); emit(vc); tile += ngroups;
      }
    }
  } else if (warp == 4) {
    // ================= TMA producer (weights; activations too in mode TMA) =================
    uint32_t g = 0;
    // pair mode: the leader's barrier counts the bytes landing in BOTH CTAs
    const uint32_t tx_bytes = (uint32_t)CL * ((uint32_t)B_STAGE + (p.a_mode == A_TMA ? (uint32_t)(p.TW * p.TH * 128) : 0u));
    for (int tile = group0; tile < p.num_tiles; tile += ngroups) {
      const int mg = tile / p.n_tiles_n, n0 = (tile - mg * p.n_tiles_n) * BN, mt = mg * CL + rank;
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
        if ((s & cmask) == 0) mbar_wait(empty0 + 8 * (s >> p.clog), ((g / S) & 1) ^ 1);
        if (elect_one()) {
          // this CTA stages rows [rank * BN/CL, +BN/CL) of the weight tile (a 2-CTA MMA reads both halves)
          if (CL == 2 && rank == 1) {
            tma_load_2d_2cta(sB + s * B_STAGE, &tmap_w, ((ky * p.KW + kx) * p.cin_blocks + cc) * TC_BK, n0 + (BN / CL), full0 + 8 * s);
            if (p.a_mode == A_TMA)
              tma_load_4d_2cta(sA + s * TC_A_STAGE, &tmap_a, p.in_coffset + cc * TC_BK, x0 + kx * p.dil, y0 + ky * p.dil, n_img,
                               full0 + 8 * s);
          } else {
            mbar_arrive_expect_tx(full0 + 8 * s, tx_bytes);
            tma_load_2d(sB + s * B_STAGE, &tmap_w, ((ky * p.KW + kx) * p.cin_blocks + cc) * TC_BK, n0, full0 + 8 * s);
            if (p.a_mode == A_TMA)
              tma_load_4d(sA + s * TC_A_STAGE, &tmap_a, p.in_coffset + cc * TC_BK, x0 + kx * p.dil, y0 + ky * p.dil, n_img,
                          full0 + 8 * s);
          }
        }
        __syncwarp();
        if (++kx == p.KW) { kx = 0; if (++ky == p.KH) { ky = 0; ++cc; } }
      }
    }
  } else if (warp == 5) {
    // ================= MMA issuer =================
    // Pair mode: only the leader CTA (rank 0) issues —
    // one tcgen05.mma.cta_group::2 drives both SMs' tensor cores (M = 256: 128 rows from each CTA, B halves from each
    // CTA's shared memory); the peer's producers and TMA loads signal the leader's full barriers directly.
    // The issue rate of this one thread bounds the kernel whenever a K-step holds little math (narrow N tiles: a
    // tcgen05.mma occupies the issuing thread ~48 clk, every other instruction of the loop adds to that —
    // profiles/r1_mma_probe.txt), so a single elected lane runs the whole role with ring slot, phase and descriptors
    // tracked incrementally.
    if (!(CL == 2 && rank == 1) && elect_one()) {
      const uint32_t idesc = make_idesc_f16(p.is_bf16 != 0, TC_BM * CL, BN);
      const uint64_t adesc0 = make_sw128_desc(sA), bdesc0 = make_sw128_desc(sB);
      const uint64_t a_step = (uint64_t)(TC_A_STAGE >> 4), b_step = (uint64_t)(B_STAGE >> 4);
      uint64_t ad = adesc0, bd = bdesc0;
      uint32_t s = 0, ph = 0, lt = 0;
      for (int tile = group0; tile < p.num_tiles; tile += ngroups, ++lt) {
        const uint32_t buf = lt & 1;
        mbar_wait(acce0 + 8 * buf, ((lt >> 1) & 1) ^ 1);            // (pair mode: both CTAs' epilogues) have drained it
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * (uint32_t)p.acc_stride;
        for (int it = 0; it < nk; ++it) {
          mbar_wait(full0 + 8 * s, ph);                            // pair mode: arrivals come from both CTAs
          tc_fence_after();
          if (CL == 2) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_f16_2cta(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (it | k) ? 1u : 0u);
            if ((s & cmask) == cmask) umma_commit_2cta(empty0 + 8 * (s >> p.clog), (uint16_t)3);
          } else {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_f16(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (it | k) ? 1u : 0u);
            if ((s & cmask) == cmask) umma_commit(empty0 + 8 * (s >> p.clog));
          }
          ad += a_step; bd += b_step;
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; ad = adesc0; bd = bdesc0; }
        }
        if (CL == 2) umma_commit_2cta(accf0 + 8 * buf, (uint16_t)3);
        else umma_commit(accf0 + 8 * buf);
      }
    }
    __syncwarp();
  } else {
    // bulk_out: 8 staging blocks of 32 rows x Cout x 2 B behind the bias vector
    const uint32_t stage0 = (smem_u32(s_bias) + 4u * (uint32_t)(((p.Cout + 31) & ~31) + 32) + 127u) & ~127u;
    epilogue_role<CL>(p, s_bias, tmem_base, accf0, acce0, warp, lane, (uint32_t)(warp - 6) >> 2, rank, group0, ngroups,
                      stage0 + (uint32_t)(warp - 6) * 64u * (uint32_t)p.Cout);
  }

  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();                          // no CTA exits while its peer can still multicast into it
  if (warp == 5) {
    tc_fence_after();
    if (CL == 2) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------
// HALO mode: 3x3 (dilated) stride-1 convs.  The output tile is an 8 x 16 pixel patch; its (8 + 2 dil) x (16 + 2 dil)
// input neighbourhood of one 64-channel block is staged ONCE (one 4-D TMA box, zero-filled outside the image) and
// the nine taps read it through nine shifted UMMA descriptors: with 8-pixel tile rows every 8-row group of the
// M = 128 operand is one patch row, so tap (ky, kx) is the same SWIZZLE_128B K-major matrix started
// (ky dil PW + kx dil) rows later with PW rows between groups (profiles/r1_mma_probe.txt: exact for any start row and
// group stride).  Against the per-tap patches of mode TMA this cuts the activation traffic L2 -> shared memory about
// 6x — those layers sit at the L2 bandwidth — and needs one A barrier per channel block instead of one per tap.
// Weights stream through their own ring, one BN x 64 tile per (channel block, tap), released in commit groups.
// Warp roles (352 threads): warp 0 patch producer (runs a full patch ring ahead: a patch load has several channel blocks
// of MMA work to hide behind), warp 1 weight producer, warp 2 MMA issuer, warps 3-10 two epilogue sets.
// OCC = 2 (tiles <= 128 wide): two CTAs per SM, each with half of TMEM and of shared memory — two MMA-issuing threads per
// SM, because with narrow tiles the issue rate of ONE thread (~50-90 clk per tcgen05.mma against a tensor-pipe floor of
// 32 / 64 clk for N = 64 / 128) is what bounds the kernel.
template <int OCC, int NBUF>      // CTAs per SM; accumulators in rotation (4 for tiles <= 128 wide with the lean epilogue: the epilogue of a
                                  // tile takes longer than its MMAs there, and with two accumulators the MMA thread waited for it every tile —
                                  // clock64 timeline, profiles/dev/halo_timeline.py)
__global__ void __launch_bounds__(HALO_THREADS, OCC)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int BN = p.bn, SA = p.sa, SB = p.sb;
  const uint32_t B_TAP = (uint32_t)BN * TC_BK * 2, B_STAGE = B_TAP * (uint32_t)p.tps;     // a slot holds the weight tiles of `tps` taps
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base, sB = smem_base + (uint32_t)(SA * p.a_slot);
  const uint32_t bars = sB + (uint32_t)SB * B_STAGE;
  const uint32_t fullA0 = bars, emptyA0 = bars + 8 * SA, fullB0 = bars + 16 * SA, emptyB0 = fullB0 + 8 * SB, accf0 = emptyB0 + 8 * SB,
                 acce0 = accf0 + 8 * NBUF, tmem_slot = acce0 + 8 * NBUF;
  float* s_bias = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group0 = blockIdx.x, ngroups = gridDim.x;
  const uint32_t cmask = (1u << p.clog) - 1u;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_a);
  if (warp == 1 && lane == 0) tma_prefetch_desc(&tmap_w);
  if (warp == 2) {
    if (lane == 0) {
      for (int s = 0; s < SA; ++s) { mbar_init(fullA0 + 8 * s, 1); mbar_init(emptyA0 + 8 * s, 1); }
      for (int s = 0; s < SB; ++s) { mbar_init(fullB0 + 8 * s, 1); mbar_init(emptyB0 + 8 * s, 1); }
      for (int b = 0; b < NBUF; ++b) { mbar_init(accf0 + 8 * b, 1); mbar_init(acce0 + 8 * b, p.epi_split ? 8 : 4); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512 / OCC); tmem_relinquish();
  }
  for (int c = threadIdx.x; c < ((p.Cout + 31) & ~31) + 32; c += HALO_THREADS) s_bias[c] = c < p.Cout ? p.bias[c] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");          // programmatic dependent launch (see conv_tc_kernel)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // Each of the three issue roles is one elected lane; ring slots and phases are tracked incrementally (no division in
  // the loops: the issue rate of these threads is the kernel's pace).
  const int per_img = p.tiles_x * p.tiles_y;
  if (warp == 0) {
    // ================= patch producer =================
    if (elect_one()) {
      const uint32_t a_bytes = (uint32_t)(p.PW * p.PH * 128);
      uint32_t slot = 0, ph = 1, dst = sA;
      for (int tile = group0; tile < p.num_tiles; tile += ngroups) {
        const int mt = tile / p.n_tiles_n;
        const int n_img = mt / per_img, t = mt - n_img * per_img;
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        const int x0 = tx * p.TW - p.pad_w, y0 = ty * p.TH - p.pad_h;
        for (int cc = 0; cc < p.cin_blocks; ++cc) {
          mbar_wait(emptyA0 + 8 * slot, ph);
          dbg_stamp(p, 0, (uint32_t)((tile - group0) / ngroups), cc == 0 ? 0 : 1);
          mbar_arrive_expect_tx(fullA0 + 8 * slot, a_bytes);
          tma_load_4d(dst, &tmap_a, p.in_coffset + cc * TC_BK, x0, y0, n_img, fullA0 + 8 * slot);
          dst += (uint32_t)p.a_slot;
          if (++slot == (uint32_t)SA) { slot = 0; ph ^= 1u; dst = sA; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= weight producer =================
    if (p.resident) {
      // small layers (conv1_2: 72 KB, conv2_1: 144 KB): all nine taps of every channel block are loaded ONCE per CTA and
      // stay; the MMA thread then has a single barrier to wait for per channel block (the patch)
      if (elect_one() && group0 < p.num_tiles) {
        mbar_arrive_expect_tx(fullB0, 9u * (uint32_t)p.cin_blocks * B_TAP);
        for (int cc = 0; cc < p.cin_blocks; ++cc)
          for (int tap = 0; tap < 9; ++tap)
            tma_load_2d(sB + (uint32_t)(cc * 9 + tap) * B_TAP, &tmap_w, (tap * p.cin_blocks + cc) * TC_BK, 0, fullB0);
      }
    } else
    if (elect_one()) {
      uint32_t s = 0, ph = 1, dst = sB;
      for (int tile = group0; tile < p.num_tiles; tile += ngroups) {
        const int n0 = (tile % p.n_tiles_n) * BN;
        for (int cc = 0; cc < p.cin_blocks; ++cc) {
          int kcoord = cc * TC_BK;                              // weight K index of (tap, cc) = (tap * cin_blocks + cc) * 64
          for (int tap0 = 0; tap0 < 9; tap0 += p.tps) {
            if ((s & cmask) == 0) mbar_wait(emptyB0 + 8 * (s >> p.clog), ph);
            mbar_arrive_expect_tx(fullB0 + 8 * s, B_STAGE);
            for (int t = 0; t < p.tps; ++t, kcoord += p.cin_blocks * TC_BK) tma_load_2d(dst + t * B_TAP, &tmap_w, kcoord, n0, fullB0 + 8 * s);
            dst += B_STAGE;
            if (++s == (uint32_t)SB) { s = 0; ph ^= 1u; dst = sB; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ================= MMA issuer =================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(p.is_bf16 != 0, TC_BM, BN);
      const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)((uint32_t)(p.PW * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t bdesc0 = make_sw128_desc(sB), b_step = (uint64_t)(B_STAGE >> 4), b_tap = (uint64_t)(B_TAP >> 4);
      const uint32_t row_step = (uint32_t)(p.dil * 128) >> 4, line_step = (uint32_t)((p.PW - 2) * p.dil * 128) >> 4;   // descriptor units (16 B)
      uint32_t slot = 0, pha = 0, s = 0, phb = 0, lt = 0, a_lo = (sA >> 4) & 0x3FFFu;
      uint64_t bd = bdesc0;
      if (p.resident && group0 < p.num_tiles) mbar_wait(fullB0, 0);
      for (int tile = group0; tile < p.num_tiles; tile += ngroups, ++lt) {
        const uint32_t buf = lt & (uint32_t)(NBUF - 1);
        mbar_wait(acce0 + 8 * buf, ((lt / (uint32_t)NBUF) & 1) ^ 1);
        tc_fence_after();
        dbg_stamp(p, 2, lt, 0);
        const uint32_t tmem_d = tmem_base + buf * (uint32_t)p.acc_stride;
        for (int cc = 0; cc < p.cin_blocks; ++cc) {
          mbar_wait(fullA0 + 8 * slot, pha);
          if (cc == 0) dbg_stamp(p, 2, lt, 1);
          uint32_t a_tap = a_lo;                                // window of tap (0, 0); +dil rows per kx, +dil patch lines per ky
          int kx = 0;
          if (p.resident) {
            tc_fence_after();
            uint64_t bt = bdesc0 + (uint64_t)(cc * 9) * b_tap;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap, bt += b_tap) {
              const uint64_t ad = desc_hi | (uint64_t)a_tap;
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) umma_f16(tmem_d, ad + 2 * k, bt + 2 * k, idesc, (cc | tap | k) ? 1u : 0u);
              a_tap += (tap % 3 == 2) ? line_step : row_step;
            }
          } else
          for (int tap0 = 0; tap0 < 9; tap0 += p.tps) {
            mbar_wait(fullB0 + 8 * s, phb);
            tc_fence_after();
            uint64_t bt = bd;
            for (int t = 0; t < p.tps; ++t, bt += b_tap) {
              const uint64_t ad = desc_hi | (uint64_t)a_tap;
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) umma_f16(tmem_d, ad + 2 * k, bt + 2 * k, idesc, (cc | tap0 | t | k) ? 1u : 0u);
              if (++kx == 3) { kx = 0; a_tap += line_step; } else a_tap += row_step;
            }
            if ((s & cmask) == cmask) umma_commit(emptyB0 + 8 * (s >> p.clog));
            bd += b_step;
            if (++s == (uint32_t)SB) { s = 0; phb ^= 1u; bd = bdesc0; }
          }
          umma_commit(emptyA0 + 8 * slot);
          a_lo += (uint32_t)p.a_slot >> 4;
          if (++slot == (uint32_t)SA) { slot = 0; pha ^= 1u; a_lo = (sA >> 4) & 0x3FFFu; }
        }
        umma_commit(accf0 + 8 * buf);
        dbg_stamp(p, 2, lt, 2);
      }
    }
    __syncwarp();
  } else {
    epilogue_role<1, OCC == 1, NBUF>(p, s_bias, tmem_base, accf0, acce0, warp, lane, (uint32_t)(warp - 3) >> 2, 0, group0, ngroups);   // (OCC = 2: 80 registers)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512 / OCC); }
}

// ---------------------------------------------------------------------------------------------------
// HALO mode on CTA pairs (tune a_mode 5): two CTAs of a cluster take two adjacent 8 x 16 tiles and ONE tcgen05.mma.cta_group::2
// (M = 256) per K = 16 step serves both.  Each CTA keeps HALF of the layer's weight rows resident in its shared memory (the
// pair instruction reads N / 2 rows of B from each CTA), so
//   * layers whose weights are too large for one CTA's shared memory (conv2_2: 288 KB) become resident (144 KB per CTA) instead
//     of streaming every tap through a ring once per tile — that stream, not the math, bounded them (shared-memory write +
//     read bandwidth), and
//   * the operand fetch per MMA drops from A + B to A + B / 2 bytes per CTA: for N <= 128 the single-CTA instruction needs the
//     full 128 B / clk of shared-memory bandwidth, which it has to share with the patch loads and the epilogue.
// Hand-offs as in conv_tc_kernel<S, 2>: the peer's TMA loads complete on the LEADER's full barriers, the leader's commits are
// multicast to both CTAs, the peer's epilogue warps arrive on the leader's accumulator-empty barriers.
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_pair_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int BN = p.bn, SA = p.sa;
  const uint32_t B_TAP = (uint32_t)(BN / 2) * TC_BK * 2;       // this CTA's half of one tap's weight tile
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base, sB = smem_base + (uint32_t)(SA * p.a_slot);
  const uint32_t bars = sB + 9u * (uint32_t)p.cin_blocks * B_TAP;
  constexpr int NBUF = 4;                                   // accumulators in rotation (BN <= 128: 4 x 128 TMEM columns)
  const uint32_t fullA0 = bars, emptyA0 = bars + 8 * SA, fullB = bars + 16 * SA, accf0 = fullB + 16, acce0 = accf0 + 8 * NBUF, tmem_slot = acce0 + 8 * NBUF;   // (s_bias stays 16-byte aligned)
  float* s_bias = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int group0 = blockIdx.x / 2, ngroups = gridDim.x / 2;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_a);
  if (warp == 1 && lane == 0) tma_prefetch_desc(&tmap_w);
  if (warp == 2) {
    if (lane == 0) {
      for (int s = 0; s < SA; ++s) { mbar_init(fullA0 + 8 * s, 1); mbar_init(emptyA0 + 8 * s, 1); }
      mbar_init(fullB, 1);
      for (int b = 0; b < NBUF; ++b) { mbar_init(accf0 + 8 * b, 1); mbar_init(acce0 + 8 * b, p.epi_split ? 16 : 8); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2cta(tmem_slot, 512); tmem_relinquish_2cta();
  }
  for (int c = threadIdx.x; c < ((p.Cout + 31) & ~31) + 32; c += HALO_THREADS) s_bias[c] = c < p.Cout ? p.bias[c] : 0.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // the peer's barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int per_img = p.tiles_x * p.tiles_y;
  if (warp == 0) {
    // ================= patch producer: this CTA's tile (2 g + rank) of every pair =================
    if (elect_one()) {
      const uint32_t a_bytes = (uint32_t)(p.PW * p.PH * 128);
      uint32_t slot = 0, ph = 1, dst = sA;
      for (int tile = group0; tile < p.num_tiles; tile += ngroups) {
        const int mt = tile * 2 + rank;                      // (one N tile: tile == pair index); past the last tile: image index >= N reads as zero
        const int n_img = mt / per_img, t = mt - n_img * per_img;
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        const int x0 = tx * p.TW - p.pad_w, y0 = ty * p.TH - p.pad_h;
        for (int cc = 0; cc < p.cin_blocks; ++cc) {
          mbar_wait(emptyA0 + 8 * slot, ph);
          if (rank == 0) {
            mbar_arrive_expect_tx(fullA0 + 8 * slot, 2u * a_bytes);     // the leader's barrier counts the bytes landing in both CTAs
            tma_load_4d(dst, &tmap_a, p.in_coffset + cc * TC_BK, x0, y0, n_img, fullA0 + 8 * slot);
          } else {
            tma_load_4d_2cta(dst, &tmap_a, p.in_coffset + cc * TC_BK, x0, y0, n_img, fullA0 + 8 * slot);
          }
          dst += (uint32_t)p.a_slot;
          if (++slot == (uint32_t)SA) { slot = 0; ph ^= 1u; dst = sA; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= weights: rows [rank BN / 2, +BN / 2) of every (channel block, tap), once =================
    if (elect_one() && group0 < p.num_tiles) {
      if (rank == 0) mbar_arrive_expect_tx(fullB, 2u * 9u * (uint32_t)p.cin_blocks * B_TAP);
      for (int cc = 0; cc < p.cin_blocks; ++cc)
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t dst = sB + (uint32_t)(cc * 9 + tap) * B_TAP;
          if (rank == 0) tma_load_2d(dst, &tmap_w, (tap * p.cin_blocks + cc) * TC_BK, 0, fullB);
          else tma_load_2d_2cta(dst, &tmap_w, (tap * p.cin_blocks + cc) * TC_BK, BN / 2, fullB);
        }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc_f16(p.is_bf16 != 0, TC_BM * 2, BN);
      const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)((uint32_t)(p.PW * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t bdesc0 = make_sw128_desc(sB), b_tap = (uint64_t)(B_TAP >> 4);
      const uint32_t row_step = (uint32_t)(p.dil * 128) >> 4, line_step = (uint32_t)((p.PW - 2) * p.dil * 128) >> 4;
      uint32_t slot = 0, pha = 0, lt = 0, a_lo = (sA >> 4) & 0x3FFFu;
      if (group0 < p.num_tiles) mbar_wait(fullB, 0);
      for (int tile = group0; tile < p.num_tiles; tile += ngroups, ++lt) {
        const uint32_t buf = lt & (uint32_t)(NBUF - 1);
        mbar_wait(acce0 + 8 * buf, ((lt / (uint32_t)NBUF) & 1) ^ 1);            // both CTAs' epilogues have drained it
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * (uint32_t)p.acc_stride;
        for (int cc = 0; cc < p.cin_blocks; ++cc) {
          mbar_wait(fullA0 + 8 * slot, pha);
          tc_fence_after();
          uint32_t a_tap = a_lo;
          uint64_t bt = bdesc0 + (uint64_t)(cc * 9) * b_tap;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap, bt += b_tap) {
            const uint64_t ad = desc_hi | (uint64_t)a_tap;
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_f16_2cta(tmem_d, ad + 2 * k, bt + 2 * k, idesc, (cc | tap | k) ? 1u : 0u);
            a_tap += (tap % 3 == 2) ? line_step : row_step;
          }
          umma_commit_2cta(emptyA0 + 8 * slot, (uint16_t)3);
          a_lo += (uint32_t)p.a_slot >> 4;
          if (++slot == (uint32_t)SA) { slot = 0; pha ^= 1u; a_lo = (sA >> 4) & 0x3FFFu; }
        }
        umma_commit_2cta(accf0 + 8 * buf, (uint16_t)3);
      }
    }
    __syncwarp();
  } else {
    epilogue_role<2, true, NBUF>(p, s_bias, tmem_base, accf0, acce0, warp, lane, (uint32_t)(warp - 3) >> 2, rank, group0, ngroups);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // no CTA exits while its peer can still signal it
  if (warp == 2) { tc_fence_after(); tmem_dealloc_2cta(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------
static int tc_supported(const CtxConvParams* p) {
  if (!p) return 0;
  if (p->in_nchw)       // stem mode: raw fp32 NCHW input, 3x3 / s1 / p1, 16-bit output
    return p->in_dtype == CTX_F32 && p->Cin == 3 && p->KH == 3 && p->KW == 3 && p->stride == 1 && p->pad_h == 1 && p->pad_w == 1 &&
           p->dil == 1 && p->nseg >= 1 && (p->seg[0].dtype == CTX_BF16 || p->seg[0].dtype == CTX_F16) &&
           ((uintptr_t)p->weight) % 16 == 0;
  if (p->in_dtype != CTX_BF16 && p->in_dtype != CTX_F16) return 0;
  if (p->Cin % 8 || p->in_cstride % 8 || p->in_coffset % 8) return 0;
  if (p->KH * p->KW > 32) return 0;
  if (((uintptr_t)p->in) % 16 || ((uintptr_t)p->weight) % 16) return 0;
  if (p->pool2) {                        // fused 2x2 pooling: TMA patch mode, one 16-bit output segment
    int tw, th;
    if (!choose_patch(p, &tw, &th) || p->Cin % 64 || p->nseg != 1 || p->residual || p->seg[0].dtype != p->in_dtype || p->Cout % 8) return 0;
  }
  return 1;
}

template <int S, int CL>
static int launch_tc(const TcPlan* pl, cudaStream_t st) {
  CTX_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<S, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)pl->grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = pl->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int pdl = [] { const char* e = getenv("CTX_CONV_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
  attr[0].val.programmaticStreamSerializationAllowed = pdl;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = (unsigned)pl->p.cluster;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pl->p.cluster > 1 ? 2 : 1;
  CTX_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_tc_kernel<S, CL>, pl->tmap_w, pl->tmap_a, pl->p));
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

template <int OCC, int NBUF>
static int launch_halo(const TcPlan* pl, cudaStream_t st) {
  CTX_CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<OCC, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)pl->grid);
  cfg.blockDim = dim3(HALO_THREADS);
  cfg.dynamicSmemBytes = pl->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int pdl = [] { const char* e = getenv("CTX_CONV_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
  attr[0].val.programmaticStreamSerializationAllowed = pdl;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TcParams prm = pl->p;
  prm.dbg = conv_timeline_ptr();
  CTX_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_halo_kernel<OCC, NBUF>, pl->tmap_w, pl->tmap_a, prm));
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

static int launch_halo_pair(const TcPlan* pl, cudaStream_t st) {
  CTX_CUDA_TRY(cudaFuncSetAttribute(conv_halo_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)pl->grid);
  cfg.blockDim = dim3(HALO_THREADS);
  cfg.dynamicSmemBytes = pl->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int pdl = [] { const char* e = getenv("CTX_CONV_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
  attr[0].val.programmaticStreamSerializationAllowed = pdl;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  CTX_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_halo_pair_kernel, pl->tmap_w, pl->tmap_a, pl->p));
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

}  // namespace ctx

using namespace ctx;

extern "C" int ctx_conv2d_tc_supported(const CtxConvParams* p) { return tc_supported(p); }

// tune_n: number of N tiles (0 = ceil(Cout / 256)); tune_cluster: 1 / 2 CTAs per MMA (0 = default); tune_amode: -1 = rule
// of thumb (TMA patches when they tile the map with <= 10 % waste), 0 = im2col gather, 1 = TMA patches whenever the
// geometry allows them (any waste).  Results are bit-identical across all settings: only the tiling changes.
static int plan_create(const CtxConvParams* p, int tune_n, int tune_cluster, int tune_amode, int tune_commit, void** plan_out) {
  CTX_REQUIRE(p && plan_out, "ctx_conv2d_tc_plan_create: null argument");
  *plan_out = nullptr;
  if (!tc_supported(p)) { set_error("ctx_conv2d_tc: geometry / dtype not supported by the tcgen05 path"); return CTX_ERR_UNSUPPORTED; }
  CTX_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->Cin > 0 && p->Cout > 0 && p->stride > 0 && p->dil > 0, "conv_tc: bad dims");
  const int ho = (p->H + 2 * p->pad_h - p->dil * (p->KH - 1) - 1) / p->stride + 1;
  const int wo = (p->W + 2 * p->pad_w - p->dil * (p->KW - 1) - 1) / p->stride + 1;
  CTX_REQUIRE(ho == p->Ho && wo == p->Wo, "conv_tc: Ho/Wo inconsistent with geometry");
  CTX_REQUIRE(p->nseg >= 1 && p->nseg <= 3 && p->bias, "conv_tc: needs 1..3 output segments and a bias vector");
  CTX_REQUIRE((long long)p->N * p->Ho * p->Wo < (1ll << 31), "conv_tc: too many output pixels");

  TcPlan* pl = new TcPlan();
  memset(&pl->tmap_a, 0, sizeof pl->tmap_a);
  TcParams& t = pl->p;
  t.in = p->in; t.bias = p->bias; t.residual = p->residual;
  t.N = p->N; t.H = p->H; t.W = p->W; t.Cin = p->Cin; t.in_cstride = p->in_cstride; t.in_coffset = p->in_coffset;
  t.Cout = p->Cout; t.KH = p->KH; t.KW = p->KW; t.stride = p->stride; t.pad_h = p->pad_h; t.pad_w = p->pad_w; t.dil = p->dil;
  t.Ho = p->Ho; t.Wo = p->Wo; t.relu = p->relu;
  t.relu_cend = p->relu_channels > 0 ? p->relu_channels : p->Cout;
  t.pool2 = p->pool2 != 0;
  { const char* e = getenv("CTX_EPI_SPLIT"); t.epi_split = (e && e[0] == '0') ? 0 : 1; }
  t.M = p->N * p->Ho * p->Wo;
  t.cin_blocks = (p->Cin + TC_BK - 1) / TC_BK;
  t.nk = p->in_nchw ? 1 : p->KH * p->KW * t.cin_blocks;       // stem: all 27 taps*channels in one 64-wide K-step
  t.res_cstride = p->res_cstride; t.res_coffset = p->res_coffset; t.res_dtype = p->res_dtype;
  const int op_dtype = p->in_nchw ? p->seg[0].dtype : p->in_dtype;     // operand (and 16-bit output) type
  t.is_bf16 = op_dtype == CTX_BF16;
  t.segs.nseg = p->nseg;
  for (int s = 0; s < 3; ++s) t.segs.seg[s] = p->seg[s < p->nseg ? s : 0];
  const CtxOutSeg& s0 = p->seg[0];
  t.fast_out = t.relu_cend % 8 == 0 && p->nseg == 1 && s0.dtype == op_dtype && s0.c_begin == 0 && s0.c_end == p->Cout && p->Cout % 8 == 0 &&
               s0.pix_stride % 8 == 0 && s0.ch_offset % 8 == 0 && s0.img_stride % 8 == 0 && ((uintptr_t)s0.ptr) % 16 == 0 &&
               ((uintptr_t)p->bias) % 16 == 0 &&
               (!p->residual || (p->res_dtype == op_dtype && p->res_cstride % 8 == 0 && p->res_coffset % 8 == 0 &&
                                 ((uintptr_t)p->residual) % 16 == 0));
  t.bulk_out = 0;
  t.vec_f32 = 0;
  if (!t.fast_out && !p->residual && t.relu_cend % 4 == 0 && p->Cout % 4 == 0 && ((uintptr_t)p->bias) % 16 == 0) {
    bool ok = true;
    int expect = 0;
    for (int s = 0; s < p->nseg; ++s) {
      const CtxOutSeg& sg = p->seg[s];
      ok = ok && sg.dtype == CTX_F32 && sg.c_begin == expect && sg.c_begin % 4 == 0 && sg.c_end % 4 == 0 && sg.pix_stride % 4 == 0 &&
           sg.ch_offset % 4 == 0 && sg.img_stride % 4 == 0 && ((uintptr_t)sg.ptr) % 16 == 0;
      expect = sg.c_end;
    }
    t.vec_f32 = ok && expect == p->Cout;
  }

  // N tile: split Cout evenly over ceil(Cout / 256) tiles, rounded up to the UMMA granularity of 16
  t.cluster_req = tune_cluster;
  t.n_tiles_n = std::max(cdiv(p->Cout, 256), std::min(tune_n, cdiv(p->Cout, 16)));
  t.bn = (cdiv(p->Cout, t.n_tiles_n) + 15) / 16 * 16;
  // CTA pairs split the weight tile in two halves of a multiple of 16 rows: an explicitly requested pair rounds the tile
  // up to a multiple of 32 (the fused heads: 396 channels = 2 x 208 -> 2 x 224) when that does not cost another tile
  if (tune_cluster == 2 && t.bn % 32 && t.bn + 16 <= 256) t.bn += 16;
  t.n_tiles_n = cdiv(p->Cout, t.bn);

  int tw = 0, th = 0;
  const bool halo_ok = tma_eligible(p) && p->KH == 3 && p->KW == 3 && p->pad_h == p->pad_w && t.cluster_req != 2 &&
                       (!p->pool2 || (((p->Ho | p->Wo) & 1) == 0 && p->Cin % 64 == 0));
  t.occ = 1; t.acc_stride = 256;
  t.resident = 0;
  if ((tune_amode == 2 || tune_amode == 3 || tune_amode == 4 || tune_amode == 5) && halo_ok) { t.a_mode = A_HALO; tw = 8; th = 16; }
  else if (p->in_nchw) t.a_mode = A_STEM;
  else if (tune_amode == 0 && !p->pool2) t.a_mode = A_GATHER;
  else if (flat_eligible(p)) { t.a_mode = A_TMA; tw = 128; th = 1; }
  else t.a_mode = choose_patch(p, &tw, &th, tune_amode == 1 ? 100000 : 150, 2) ? A_TMA : A_GATHER;
  t.flat = t.a_mode == A_TMA && flat_eligible(p);
  t.TW = tw; t.TH = th;
  const bool patches = t.a_mode == A_TMA || t.a_mode == A_HALO;
  t.tiles_x = patches ? (t.flat ? cdiv(t.M, 128) : cdiv(p->Wo, tw)) : 0;
  t.tiles_y = patches ? (t.flat ? 1 : cdiv(p->Ho, th)) : 0;
  const int m_tiles = patches ? (t.flat ? t.tiles_x : p->N * t.tiles_x * t.tiles_y) : cdiv(t.M, TC_BM);
  t.m_tiles = m_tiles;
  {
    // CTA pairs: one tcgen05.mma.cta_group::2 spanning both SMs of a cluster of two, each CTA staging half of the weight
    // tile.  Chosen per layer by the autotuner (they win on the wide trunk layers); CTX_CONV_CLUSTER=2 makes them the
    // default where eligible.
    const char* e = getenv("CTX_CONV_CLUSTER");
    const bool want2 = tune_cluster ? tune_cluster == 2 : (e && e[0] == '2');
    t.cluster = (want2 && m_tiles >= 2 && t.bn % 32 == 0 && t.bn >= 128 && t.a_mode != A_HALO) ? 2 : 1;
    if (t.a_mode == A_HALO && tune_amode == 5) t.cluster = 2;          // CTA pairs with resident half-weights (checked below)
  }
  t.num_tiles = cdiv(m_tiles, t.cluster) * t.n_tiles_n;
  const int stage_bytes = TC_A_STAGE + (t.bn / t.cluster) * TC_BK * 2;
  pl->stages = stage_bytes <= 24 * 1024 ? 8 : (stage_bytes <= 32 * 1024 ? 6 : 4);
  {
    // commit group: enough MMA work per commit to hide its ~550 clk (4 MMAs of M128 x N x K16 take 2 N clk); must divide the ring
    int c = tune_commit > 0 ? tune_commit : (t.bn >= 192 ? 1 : (t.bn >= 96 ? 2 : 4));
    if (t.nk == 1) c = 1;                               // stem: one K-step per tile, latency matters more
    // gather mode signals K-step j only when it issues K-step j + S - 2 (cp.async look-ahead): the producers must be
    // able to run S - 2 slots ahead of the oldest unreleased group, i.e. look-ahead + group <= S
    if (t.a_mode == A_GATHER && c > 2) c = 2;
    while (c > 1 && (pl->stages % c || c > pl->stages / 2)) c >>= 1;
    t.clog = c >= 4 ? 2 : (c >= 2 ? 1 : 0);
  }
  const size_t bias_bytes = 4 * (((size_t)p->Cout + 31) / 32 * 32 + 32);
  if (t.a_mode == A_HALO) {
    // weight ring first (8 slots if they fit in ~half of shared memory, a multiple of the commit group), the rest — up to
    // 6 slots — is the patch ring: patch loads are the long-latency ones
    t.PW = 8 + 2 * p->dil; t.PH = 16 + 2 * p->dil;
    t.a_slot = (int)align_up((size_t)t.PW * t.PH * 128, 1024);
    // narrow tiles: a slot holds the three taps of a filter row and is released by its own commit — twelve MMAs per
    // barrier wait instead of four (the MMA thread's issue loop is the pace there); an explicit commit group keeps 1 tap
    const int clog_rule = t.clog;
    t.sa = 0;
    if (tune_amode == 5) {                  // CTA pairs: each CTA keeps half of the weight rows resident (conv_halo_pair_kernel)
      const size_t wbytes = (size_t)9 * t.cin_blocks * (t.bn / 2) * TC_BK * 2, budget = 232448 - 1024 - bias_bytes - 8 * (2 * 6 + 8) - 64;
      if (t.n_tiles_n == 1 && t.bn % 32 == 0 && t.bn >= 64 && t.bn <= 128 && t.fast_out && m_tiles >= 2 && wbytes + 3 * (size_t)t.a_slot <= budget) {
        t.resident = 1; t.occ = 1; t.acc_stride = 128; t.tps = 1; t.clog = 0;           // four accumulators of 128 TMEM columns
        t.sb = 9 * t.cin_blocks;
        t.sa = (int)std::min<size_t>(6, (budget - wbytes) / t.a_slot);
      } else {
        delete pl; set_error("conv_tc: HALO pair mode does not apply (tile %d, %d channel blocks)", t.bn, t.cin_blocks); return CTX_ERR_UNSUPPORTED;
      }
    }
    if (tune_amode == 4) {                  // weights resident: the whole layer is one N tile and fits beside >= 3 patches
      const size_t wbytes = (size_t)9 * t.cin_blocks * t.bn * TC_BK * 2, budget = 232448 - 1024 - bias_bytes - 8 * (2 * 6 + 2 * 18 + 4) - 64;
      if (t.n_tiles_n == 1 && t.cin_blocks <= 2 && wbytes + 3 * (size_t)t.a_slot <= budget) {
        t.resident = 1; t.occ = 1; t.acc_stride = 256; t.tps = 1; t.clog = 0;
        t.sb = 9 * t.cin_blocks;
        t.sa = (int)std::min<size_t>(6, (budget - wbytes) / t.a_slot);
      } else {
        delete pl; set_error("conv_tc: HALO mode with resident weights does not apply (tile %d, %d channel blocks)", t.bn, t.cin_blocks); return CTX_ERR_UNSUPPORTED;
      }
    }
    for (int occ = (tune_amode == 3 && t.bn <= 128) ? 2 : 1; occ >= 1 && !t.sa && !t.resident; --occ) {     // two CTAs per SM if they fit, else one
      t.occ = occ;
      t.acc_stride = 256 / occ;
      t.clog = clog_rule;
      t.tps = (t.bn <= 128 && tune_commit <= 0 && occ == 1) ? 3 : 1;
      if (t.tps == 3) t.clog = 0;
      if (occ == 2 && t.clog > 1) t.clog = 1;
      const size_t per_cta = occ == 2 ? (233472 - 2 * 1024) / 2 : 232448;
      const size_t b_stage = (size_t)t.tps * t.bn * TC_BK * 2, budget = per_cta - 1024 - bias_bytes - 8 * (2 * 6 + 2 * 8 + 4) - 64;
      const int c = 1 << t.clog;
      for (int sb = 8; sb >= 2 && !t.sa; --sb) {
        if (sb % c || sb < 2 * c || (size_t)sb * b_stage + 2 * (size_t)t.a_slot > budget) continue;
        const int sa = (int)std::min<size_t>(6, (budget - (size_t)sb * b_stage) / t.a_slot);
        if (sa >= 3 || sb <= 4) { t.sa = sa; t.sb = sb; }          // shrink the weight ring before going below 3 patches
      }
    }
    if (!t.sa) { delete pl; set_error("conv_tc: HALO mode does not fit shared memory (dilation %d, tile width %d)", p->dil, t.bn); return CTX_ERR_UNSUPPORTED; }
    if (t.occ == 1 && t.cluster == 1 && t.bn <= 128 && t.fast_out) t.acc_stride = 128;        // four accumulators in rotation (conv_halo_kernel<1, 4>)
    pl->stages = t.sb;
  }
  pl->smem = t.a_mode == A_HALO ? (size_t)t.sa * t.a_slot + (size_t)t.sb * t.tps * (t.bn / t.cluster) * TC_BK * 2 + 8 * (2 * t.sa + 2 * t.sb + 4) + 64 + bias_bytes + 1024 :
             (size_t)pl->stages * stage_bytes + 24 * pl->stages + 64 + 4 * (((size_t)p->Cout + 31) / 32 * 32 + 32) + 1024;
  // bulk-copy epilogue: pixel-linear tiles (stem / gather / flat), the whole pixel in one tile, dense output map, small rows
  if (t.fast_out && (t.a_mode == A_STEM || t.a_mode == A_GATHER || (t.a_mode == A_TMA && t.flat)) && t.n_tiles_n == 1 && t.cluster == 1 &&
      !p->pool2 && p->Cout <= 64 && s0.pix_stride == p->Cout && s0.img_stride == (long long)p->Ho * p->Wo * p->Cout &&
      pl->smem + 128 + 8 * 64 * (size_t)p->Cout <= 232448) {
    t.bulk_out = 1;
    pl->smem += 128 + 8 * 64 * (size_t)p->Cout;
  }
  pl->grid = std::min(t.num_tiles, num_sms() * (t.a_mode == A_HALO ? t.occ : 1) / t.cluster) * t.cluster;

  // weights: [Cout_pad][KH*KW*Cin_pad] 16-bit, K-major; box = 64 (K) x BN (Cout), SWIZZLE_128B, OOB rows read as zero
  const unsigned long long ktot = (unsigned long long)t.nk * TC_BK;
  const unsigned long long cout_pad = (unsigned long long)((p->Cout + 15) / 16 * 16);
  int rc = encode_2d_sw128(&pl->tmap_w, p->weight, t.is_bf16 != 0, cout_pad, ktot, (unsigned)(t.bn / t.cluster));
  if (!rc && t.a_mode == A_HALO)
    rc = encode_nhwc_sw128(&pl->tmap_a, p->in, t.is_bf16 != 0, p->N, p->H, p->W, p->in_cstride, p->in_coffset + p->Cin, (unsigned)t.PW, (unsigned)t.PH);
  if (!rc && t.a_mode == A_TMA)
    rc = t.flat ? encode_nhwc_sw128(&pl->tmap_a, p->in, t.is_bf16 != 0, 1, 1, t.M, p->in_cstride, p->in_coffset + p->Cin, 128u, 1u)
                : encode_nhwc_sw128(&pl->tmap_a, p->in, t.is_bf16 != 0, p->N, p->H, p->W, p->in_cstride, p->in_coffset + p->Cin, (unsigned)tw, (unsigned)th,
                                    (unsigned)p->stride);
  if (rc) { delete pl; return rc; }
  if (t.pool2 && !(((t.a_mode == A_TMA && t.TW == 16) || t.a_mode == A_HALO) && t.fast_out)) {
    delete pl;
    set_error("conv_tc: fused pooling needs the TMA patch mode and a single 16-bit output segment");
    return CTX_ERR_UNSUPPORTED;
  }
  *plan_out = pl;
  return CTX_OK;
}

extern "C" int ctx_conv2d_tc_plan_create(const CtxConvParams* p, void** plan_out) { return plan_create(p, 0, 0, -1, 0, plan_out); }

extern "C" int ctx_conv2d_tc_plan_create_tuned(const CtxConvParams* p, int n_tiles_n, int cluster, int a_mode, int commit_group, void** plan_out) {
  return plan_create(p, n_tiles_n, cluster, a_mode, commit_group, plan_out);
}

extern "C" int ctx_conv2d_tc_plan_info(void* plan, int* info8) {
  CTX_REQUIRE(plan && info8, "ctx_conv2d_tc_plan_info: null argument");
  const TcPlan* pl = (const TcPlan*)plan;
  int* info6 = info8;
  info8[6] = pl->p.a_mode == A_HALO && pl->p.tps == 3 ? 3 : 1 << pl->p.clog; info8[7] = pl->p.TW * 1000 + pl->p.TH;
  info6[0] = pl->p.bn; info6[1] = pl->p.n_tiles_n; info6[2] = pl->p.cluster; info6[3] = pl->p.a_mode == A_STEM2 ? 6 : pl->p.a_mode == A_HALO && pl->p.cluster == 2 ? 7 : pl->p.a_mode == A_HALO && pl->p.resident ? 5 : (pl->p.a_mode == A_HALO && pl->p.occ == 2 ? 4 : pl->p.a_mode); info6[4] = pl->stages; info6[5] = pl->grid;
  return CTX_OK;
}

extern "C" int ctx_conv2d_tc_plan_run(void* plan, void* stream) {
  CTX_REQUIRE(plan, "ctx_conv2d_tc_plan_run: null plan");
  const TcPlan* pl = (const TcPlan*)plan;
  cudaStream_t st = (cudaStream_t)stream;
  if (pl->p.a_mode == A_STEM2) return launch_stem2(pl, st);
  if (pl->p.a_mode == A_HALO && pl->p.cluster == 2) return launch_halo_pair(pl, st);
  if (pl->p.a_mode == A_HALO) return pl->p.occ == 2 ? launch_halo<2, 2>(pl, st) : (pl->p.acc_stride == 128 ? launch_halo<1, 4>(pl, st) : launch_halo<1, 2>(pl, st));
  if (pl->p.cluster == 2) {
    if (pl->stages == 8) return launch_tc<8, 2>(pl, st);
    if (pl->stages == 6) return launch_tc<6, 2>(pl, st);
    return launch_tc<4, 2>(pl, st);
  }
  if (pl->stages == 8) return launch_tc<8, 1>(pl, st);
  if (pl->stages == 6) return launch_tc<6, 1>(pl, st);
  return launch_tc<4, 1>(pl, st);
}

extern "C" void ctx_conv2d_tc_plan_destroy(void* plan) { delete (TcPlan*)plan; }
