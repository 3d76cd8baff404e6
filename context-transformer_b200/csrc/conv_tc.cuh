// conv_tc.cuh — pieces shared by the tensor-core conv kernels (conv_tc.cu: 16-bit modes; conv_x3.cu: fp32 emulated with
// fp16 hi/lo operand planes): tile constants, the kernel parameter block, tile -> pixel mapping, patch-shape selection.
#pragma once
#include "tc_common.cuh"

#include <algorithm>

namespace ctx {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 448;        // 4 producer + TMA + MMA + 2 x 4 epilogue warps
constexpr int TC_A_STAGE = TC_BM * TC_BK * 2;     // 16 KB
constexpr int A_GATHER = 0, A_TMA = 1, A_STEM = 2, A_HALO = 3, A_STEM2 = 4;      // A_STEM2: conv_stem2.cu (conv1_1 fused into conv1_2)
constexpr int HALO_THREADS = 352;      // patch TMA + weight TMA + MMA + 2 x 4 epilogue warps

struct TcParams {
  const void* in;
  const float* bias;
  const void* residual;
  int N, H, W, Cin, in_cstride, in_coffset;
  int Cout, KH, KW, stride, pad_h, pad_w, dil, Ho, Wo, relu, relu_cend;
  int M, cin_blocks, nk;
  int bn, n_tiles_n, num_tiles;          // num_tiles counts tile GROUPS: `cluster` M-adjacent tiles of one N tile
  int cluster_req;
  int m_tiles, cluster;                  // cluster = 2: CTA pairs, one 2-CTA MMA per K-step (opt-in)
  int a_mode, TW, TH, tiles_x, tiles_y;       // TMA mode: output patch TW x TH (<= 128 pixels), tiles per image
  int occ, acc_stride;                        // CTAs per SM the kernel is sized for (TMEM columns = 512 / occ), columns between the two accumulators
  int tps;                                    // HALO mode: filter taps per weight-ring slot (1 or 3)
  int resident;                               // HALO mode: the layer's whole weight tensor stays in shared memory (loaded once per CTA)
  int PW, PH, a_slot, sa, sb;                 // HALO mode: staged patch (TW + 2 dil) x (TH + 2 dil) pixels, slot bytes, A / B ring depths
  int flat;                                   // TMA mode, 1x1 convs: the whole batch is one pixel row, tiles are 128-pixel runs
  int res_cstride, res_coffset, res_dtype;
  int is_bf16;
  int pool2;                   // fused MaxPool2d(2,2): TMA mode with TW = 16 (2x2 windows live inside one warp)
  int clog;                    // log2 of the commit group: ring slots are handed back to the producers 1 << clog at a time
  int fast_out;                // single 16-bit segment, 8-channel aligned: 128-bit stores
  int bulk_out;                // fast_out + pixel-linear tiles of a dense map: each epilogue warp stages its 32 rows in shared memory
                               // and writes them with ONE bulk copy (the per-thread 16-B stores touch 32 lines per instruction)
  int vec_f32;                 // fp32 segments, 4-channel aligned: 128-bit stores
  SegTable segs;
  // split mode (conv_x3.cu): lo planes of input / residual / 16-bit output, per-channel accumulator scale
  const void* in_lo;
  const void* residual_lo;
  void* out_lo;
  const float* out_scale;
  float trunc_comp;            // expected relative loss of one chain to the tensor core's truncating adder (see conv_x3.cu)
  int epi_split;               // both epilogue sets drain every tile, half of its chunks each (else: alternate tiles) — conv_tc_epilogue.cuh
  const float* bias1;          // A_STEM2: bias of the fused first conv (conv1_1), 64 floats
  long long* dbg;              // development aid (ctx_debug_set_conv_timeline): clock64 stamps of CTA 0, [8 roles][64 tiles][6]; NULL = off
};
__device__ __forceinline__ void dbg_stamp(const TcParams& p, int role, uint32_t j, int k) {
  if (p.dbg && blockIdx.x == 0 && j < 64u) p.dbg[(role * 64 + (int)j) * 6 + k] = clock64();
}

// A planned tensor-core conv: the TMA descriptors (pointers baked in), the kernel parameters and the launch shape.
struct TcPlan {
  CUtensorMap tmap_w, tmap_a;
  CUtensorMap tmap_raw, tmap_w1;         // A_STEM2: raw fp32 NCHW input patches, conv1_1 weights
  TcParams p;
  float bias1_host[64], bias2_host[64];  // A_STEM2: the biases of conv1_1 / conv1_2 (passed to the kernel by value)
  int stages;
  int grid;
  size_t smem;
};
int launch_stem2(const TcPlan* pl, cudaStream_t st);      // conv_stem2.cu
long long* conv_timeline_ptr();                           // conv_stem2.cu: development aid (ctx_debug_set_conv_timeline)

// output pixel (image, linear pixel index, validity) of row r of M-tile mt
__device__ __forceinline__ bool tile_row_pixel(const TcParams& p, int mt, int r, int& n_img, int& pix) {
  if ((p.a_mode == A_TMA && !p.flat) || p.a_mode == A_HALO || p.a_mode == A_STEM2) {
    const int per_img = p.tiles_x * p.tiles_y;
    n_img = mt / per_img;
    const int t = mt - n_img * per_img;
    const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
    const int ry = r / p.TW;
    const int oy = ty * p.TH + ry, ox = tx * p.TW + (r - ry * p.TW);
    pix = oy * p.Wo + ox;
    return n_img < p.N && ry < p.TH && oy < p.Ho && ox < p.Wo;      // rows >= TW*TH of the M tile are never loaded
  }
  const int m = mt * TC_BM + r;
  const int HoWo = p.Ho * p.Wo;
  n_img = m / HoWo;
  pix = m - n_img * HoWo;
  return m < p.M;
}

static inline int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

// TMA activation mode: stride-1 conv whose output map is tiled by TW x TH pixel patches, TW * TH <= 128 (the rows of the
// M = 128 tile beyond TW * TH are never loaded nor stored).  The patch that needs the fewest tiles wins; `max_ratio_pct`
// bounds the tile count against the gather mode's ceil(M / 128) (gather packs pixels across rows and images).
static inline bool tma_eligible(const CtxConvParams* p, int max_stride = 1) {
  return p->stride >= 1 && p->stride <= max_stride && p->Cin % 8 == 0 && p->in_coffset % 8 == 0 && p->in_cstride % 8 == 0 && !p->in_nchw;
}
static inline bool flat_eligible(const CtxConvParams* p) {
  return tma_eligible(p) && p->KH == 1 && p->KW == 1 && p->pad_h == 0 && p->pad_w == 0 && !p->pool2;
}
// max_stride = 2 (conv_tc.cu only): the TMA unit walks the input with a traversal stride, so the TW x TH patch of OUTPUT pixels of a
// stride-2 conv is still one box per tap (the RFB blocks that halve the map: extras.1 / extras.2)
static inline bool choose_patch(const CtxConvParams* p, int* tw, int* th, int max_ratio_pct = 150, int max_stride = 1) {
  if (!tma_eligible(p, max_stride) || (p->stride > 1 && p->pool2)) return false;
  if (p->pool2) {                       // 2 x 2 windows must sit inside one warp of the epilogue: 16 x 8 patches
    if ((p->Ho | p->Wo) & 1) return false;
    *tw = 16; *th = 8;
    return (long long)cdiv(p->Wo, 16) * cdiv(p->Ho, 8) * p->N * 100 <= (long long)cdiv((long long)p->N * p->Ho * p->Wo, 128) * max_ratio_pct;
  }
  long long best = -1;
  int best_area = 0;
  for (int w = std::min(p->Wo, 128); w >= 1; --w) {
    const int h = std::min(128 / w, p->Ho);
    const long long tiles = (long long)cdiv(p->Wo, w) * cdiv(p->Ho, h);
    if (best < 0 || tiles < best || (tiles == best && w * h < best_area)) { best = tiles; best_area = w * h; *tw = w; *th = h; }
  }
  const long long gather_tiles = cdiv((long long)p->N * p->Ho * p->Wo, 128);
  return best * p->N * 100 <= gather_tiles * max_ratio_pct;
}

}  // namespace ctx
