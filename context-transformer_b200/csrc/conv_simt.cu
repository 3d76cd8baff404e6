// conv_simt.cu — CUDA-core (fp32-accumulate) implicit-GEMM convolution, NHWC max-pool, layout
// conversion and row softmax for sm_100a.
//
// Role: (1) the fp32 "parity" precision mode of the conv stack — every conv of
// models/RFB_Net_vgg.py (BasicConv :7-22, vgg() :323-343, heads :387-416) evaluated in fp32 so the
// 1e-4 parity bar against the fp32 reference is meaningful; (2) in the 16-bit throughput mode, the
// geometries the tcgen05 kernel does not take (Cin = 3 stem, strided convs).
//
// GEMM view: M = N*Ho*Wo output pixels (flattened over the whole batch, so small late-pyramid maps
// still fill tiles), N = Cout, K = KH*KW*Cin with k = tap*Cin + ci.  64x64x16 tiles, 256 threads,
// 4x4 outputs per thread, operands staged in shared memory; activations NHWC so the K-run of one
// tap is contiguous.  Epilogue: + bias (BatchNorm folded) [+ residual] [ReLU] -> up to three output
// segments (lets loc/conf/obj heads land directly in their concatenated [B,P,*] buffers, the
// permute(0,2,3,1).contiguous()+cat of RFB_Net_vgg.py:239-248 for free).
#include "common.cuh"
#include <algorithm>

namespace ctx {

constexpr int BM = 64, BN = 64, BK = 16, PADM = 68;

struct ConvGeom {
  int N, H, W, Cin, in_cstride, in_coffset, Cout, CoutP, KH, KW, stride, pad_h, pad_w, dil, Ho, Wo, relu, relu_cend;
  int K, M;
  int res_dtype, res_cstride, res_coffset;
};

template <typename TIn> struct Vec4Load;
template <> struct Vec4Load<float> {
  static __device__ __forceinline__ void load(const float* p, float* o) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct Vec4Load<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* o) {
    uint2 v = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&v.x), b = *reinterpret_cast<__nv_bfloat162*>(&v.y);
    o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
  }
};
template <> struct Vec4Load<__half> {
  static __device__ __forceinline__ void load(const __half* p, float* o) {
    uint2 v = *reinterpret_cast<const uint2*>(p);
    __half2 a = *reinterpret_cast<__half2*>(&v.x), b = *reinterpret_cast<__half2*>(&v.y);
    o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
  }
};

template <typename TIn, bool ALIGNED>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const TIn* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ bias,
                 const void* __restrict__ residual, ConvGeom g, SegTable segs) {
  __shared__ __align__(16) float As[BK][PADM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;

  // A-load role: pixel lm, k-sub [lk, lk+4)
  const int lm = tid >> 2, lk = (tid & 3) * 4;
  const int gm = m0 + lm;
  const bool m_ok = gm < g.M;
  int pn = 0, poy = 0, pox = 0;
  if (m_ok) { pn = gm / (g.Ho * g.Wo); int r = gm - pn * g.Ho * g.Wo; poy = r / g.Wo; pox = r - poy * g.Wo; }
  const int iy0 = poy * g.stride - g.pad_h, ix0 = pox * g.stride - g.pad_w;
  // B-load role: k row bk, 4 consecutive couts
  const int bk = tid >> 4, bn = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (ALIGNED) {
      // Cin % 16 == 0: the whole BK chunk lies inside one tap
      const int tap = k0 / g.Cin, ci = k0 - tap * g.Cin + lk;
      const int ky = tap / g.KW, kx = tap - ky * g.KW;
      const int iy = iy0 + ky * g.dil, ix = ix0 + kx * g.dil;
      if (m_ok && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
        Vec4Load<TIn>::load(in + ((size_t)(pn * g.H + iy) * g.W + ix) * g.in_cstride + g.in_coffset + ci, av);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + lk + e;
        if (m_ok && k < g.K) {
          const int tap = k / g.Cin, ci = k - tap * g.Cin;
          const int ky = tap / g.KW, kx = tap - ky * g.KW;
          const int iy = iy0 + ky * g.dil, ix = ix0 + kx * g.dil;
          if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
            av[e] = to_f32<TIn>(in[((size_t)(pn * g.H + iy) * g.W + ix) * g.in_cstride + g.in_coffset + ci]);
        }
      }
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 + bk < g.K && n0 + bn < g.CoutP)
      bv = *reinterpret_cast<const float4*>(wgt + (size_t)(k0 + bk) * g.CoutP + n0 + bn);
    __syncthreads();                 // previous tile fully consumed
#pragma unroll
    for (int e = 0; e < 4; ++e) As[lk + e][lm] = av[e];
    *reinterpret_cast<float4*>(&Bs[bk][bn]) = bv;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    const int n = m / (g.Ho * g.Wo);
    const int pix = m - n * g.Ho * g.Wo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= g.Cout) continue;
      float v = acc[i][j];
      if (bias) v += bias[c];
      if (residual) v += load_as(residual, (long long)m * g.res_cstride + g.res_coffset + c, g.res_dtype);
      if (g.relu && c < g.relu_cend) v = fmaxf(v, 0.f);
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if (s < segs.nseg && c >= segs.seg[s].c_begin && c < segs.seg[s].c_end) {
          const CtxOutSeg& sg = segs.seg[s];
          store_as(sg.ptr, (long long)n * sg.img_stride + (long long)pix * sg.pix_stride + sg.ch_offset + (c - sg.c_begin),
                   sg.dtype, v);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// nn.MaxPool2d on an NHWC view.  One thread per output element, channel fastest (coalesced).
// Window clipped to the input (padding never wins; ceil_mode is resolved by the host in Ho/Wo).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool_nhwc_kernel(CtxPoolParams p) {
  const long long total = (long long)p.N * p.Ho * p.Wo * p.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % p.C);
    long long r = i / p.C;
    const int ox = (int)(r % p.Wo); r /= p.Wo;
    const int oy = (int)(r % p.Ho);
    const int n = (int)(r / p.Ho);
    const int y0 = max(oy * p.stride - p.pad, 0), y1 = min(oy * p.stride - p.pad + p.k, p.H);
    const int x0 = max(ox * p.stride - p.pad, 0), x1 = min(ox * p.stride - p.pad + p.k, p.W);
    float m = -INFINITY;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x)
        m = fmaxf(m, load_as(p.in, (long long)n * p.in_img_stride + (long long)(y * p.W + x) * p.in_pix_stride + c, p.dtype));
    store_as(p.out, (long long)n * p.out_img_stride + (long long)(oy * p.Wo + ox) * p.out_pix_stride + c, p.dtype, m);
  }
}

// 16-bit fast path: 8 channels (16 bytes) per thread, fully coalesced 128-bit loads/stores.
template <bool BF16, int K = 0>      // K = 2, 3: window unrolled (all its loads in flight at once); 0: any window
__global__ void __launch_bounds__(256)
maxpool_nhwc_vec8_kernel(CtxPoolParams p) {
  // grid: x over (output column, 8-channel group), y over (image, output row) — no 64-bit division per element (four of them
  // per element made this kernel instruction-bound at half of the memory rate)
  const int C8 = p.C >> 3;
  const uint16_t* in = reinterpret_cast<const uint16_t*>(p.in);
  uint16_t* out = reinterpret_cast<uint16_t*>(p.out);
  const int n = (int)blockIdx.y / p.Ho, oy = (int)blockIdx.y - n * p.Ho;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.Wo * C8; i += gridDim.x * blockDim.x) {
    const int ox = i / C8, c = (i - ox * C8) * 8;
    const int y0 = max(oy * p.stride - p.pad, 0), y1 = min(oy * p.stride - p.pad + p.k, p.H);
    const int x0 = max(ox * p.stride - p.pad, 0), x1 = min(ox * p.stride - p.pad + p.k, p.W);
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
    auto take = [&](const uint4 v) {
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f;
        if (BF16) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
        else f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
        m[2 * e] = fmaxf(m[2 * e], f.x); m[2 * e + 1] = fmaxf(m[2 * e + 1], f.y);
      }
    };
    if (K > 0) {
      uint4 v[K > 0 ? K * K : 1];
      const int yb = oy * p.stride - p.pad, xb = ox * p.stride - p.pad;
#pragma unroll
      for (int dy = 0; dy < K; ++dy)
#pragma unroll
        for (int dx = 0; dx < K; ++dx) {
          const int y = yb + dy, x = xb + dx;
          // a clipped tap repeats a tap that is inside the window (the window always holds one): the maximum does not change
          const int yc = min(max(y, y0), y1 - 1), xc = min(max(x, x0), x1 - 1);
          v[dy * K + dx] = *reinterpret_cast<const uint4*>(in + (long long)n * p.in_img_stride + (long long)(yc * p.W + xc) * p.in_pix_stride + c);
        }
#pragma unroll
      for (int q = 0; q < K * K; ++q) take(v[q]);
    } else {
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x)
          take(*reinterpret_cast<const uint4*>(in + (long long)n * p.in_img_stride + (long long)(y * p.W + x) * p.in_pix_stride + c));
    }
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (BF16) { __nv_bfloat162 h = __floats2bfloat162_rn(m[2 * e], m[2 * e + 1]); o[e] = *reinterpret_cast<uint32_t*>(&h); }
      else { __half2 h = __floats2half2_rn(m[2 * e], m[2 * e + 1]); o[e] = *reinterpret_cast<uint32_t*>(&h); }
    }
    *reinterpret_cast<uint4*>(out + (long long)n * p.out_img_stride + (long long)(oy * p.Wo + ox) * p.out_pix_stride + c) =
        make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// fp32 fast path (the conf pools in front of the Context-Transformer): 4 channels (16 bytes) per thread, the same (image, row)
// grid and 32-bit index arithmetic as the 16-bit kernel above.  The generic kernel took 32 us for the 66 MB of the 38x38 level.
__global__ void __launch_bounds__(256)
maxpool_nhwc_f32x4_kernel(CtxPoolParams p) {
  const int C4 = p.C >> 2;
  const float* in = reinterpret_cast<const float*>(p.in);
  float* out = reinterpret_cast<float*>(p.out);
  const int n = (int)blockIdx.y / p.Ho, oy = (int)blockIdx.y - n * p.Ho;
  const int y0 = max(oy * p.stride - p.pad, 0), y1 = min(oy * p.stride - p.pad + p.k, p.H);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.Wo * C4; i += gridDim.x * blockDim.x) {
    const int ox = i / C4, c = (i - ox * C4) * 4;
    const int x0 = max(ox * p.stride - p.pad, 0), x1 = min(ox * p.stride - p.pad + p.k, p.W);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in + (long long)n * p.in_img_stride + (long long)(y * p.W + x) * p.in_pix_stride + c));
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    *reinterpret_cast<float4*>(out + (long long)n * p.out_img_stride + (long long)(oy * p.Wo + ox) * p.out_pix_stride + c) = m;
  }
}

// First-layer patch extraction for the tensor-core path: x[N,3,H,W] fp32 NCHW -> patches[N,H,W,64] 16-bit with
// channel = (ky*3 + kx)*3 + ci for the 3x3 / pad 1 neighbourhood (27 values) and 37 zero channels (one full
// 64-channel K-step = one 128-byte swizzle row, so the conv kernel's TMA activation path applies), so that
// conv1_1 (Cin = 3, reference base.0) becomes a K = 64 GEMM row per pixel instead of a CUDA-core convolution.
template <bool BF16>
__global__ void __launch_bounds__(256)
patch27_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, int N, int H, int W) {
  const long long hw = (long long)H * W, total = (long long)N * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / hw;
    const int r = (int)(i - n * hw), y = r / W, x = r - y * W;
    float v[32];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = y + ky - 1, ix = x + kx - 1;
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[(ky * 3 + kx) * 3 + c] = ok ? in[(n * 3 + c) * hw + (long long)iy * W + ix] : 0.f;
      }
#pragma unroll
    for (int e = 27; e < 32; ++e) v[e] = 0.f;
    uint32_t o[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      if (BF16) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]); o[e] = *reinterpret_cast<uint32_t*>(&h); }
      else { __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]); o[e] = *reinterpret_cast<uint32_t*>(&h); }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + i * 64);
#pragma unroll
    for (int e = 0; e < 4; ++e) dst[e] = make_uint4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
#pragma unroll
    for (int e = 4; e < 8; ++e) dst[e] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// x[N,C,H,W] fp32 -> NHWC (RFBNet.forward takes NCHW, RFB_Net_vgg.py:210)
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ in, void* __restrict__ out, int N, int C, int H, int W, int dtype) {
  const long long hw = (long long)H * W, total = (long long)N * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / hw, r = i - n * hw;
    for (int c = 0; c < C; ++c) store_as(out, i * C + c, dtype, in[(n * C + c) * hw + r]);
  }
}

// BaseTransform without the resize (data/data_augment.py:258-261 for an image that already has the network's size:
// cv2.resize to the same size is a copy): x[n, c, y, x] = float(img[n, y, x, c]) - means[c], uint8 HWC -> fp32 CHW.
// One thread per pixel: 3 bytes in (coalesced over the warp), three coalesced plane writes.
__global__ void __launch_bounds__(256)
u8hwc_to_f32chw_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int N, int H, int W, float m0, float m1, float m2) {
  const long long hw = (long long)H * W, total = (long long)N * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / hw, r = i - n * hw;
    const uint8_t* px = img + i * 3;
    float* o = out + n * 3 * hw + r;
    o[0] = (float)px[0] - m0;
    o[hw] = (float)px[1] - m1;
    o[2 * hw] = (float)px[2] - m2;
  }
}

// BaseTransform WITH the resize (data/data_augment.py:257-261): cv2.resize(img, (S, S), INTER_LINEAR) on the 8-bit image,
// then float - mean, HWC -> CHW.  OpenCV's 8-bit bilinear resize is fixed-point (imgproc/resize.cpp: resizeGeneric_ with
// HResizeLinear<uchar,int,short,2048> and VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>; the IPP variant is not
// used for 8-bit linear unless IPP "not exact" mode is switched on), and it is restated here operation for operation:
//   scale = 1 / (dst / src) in double;  f = (float)((d + 0.5) * scale - 0.5);  s = floor(f);  f -= s  (float)
//   columns: s < 0 -> f = 0, s = 0;  s >= src_w - 1 -> f = 0, s = src_w - 1;  rows: s and s + 1 clipped to the image
//   taps a1 = cvRound(f * 2048), a0 = cvRound((1.f - f) * 2048)  (round half to even), horizontal pass in int32,
//   vertical pass  (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
// No fused multiply-adds in the coordinate arithmetic (x86 builds of OpenCV do not contract).  One thread per output pixel.
__device__ __forceinline__ void cv_linear_tap(int d, double scale, int src, bool clamp_edges, int& s, int& a0, int& a1) {
  float f = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
  s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (clamp_edges) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

__global__ void __launch_bounds__(256)
resize_base_transform_kernel(const uint8_t* __restrict__ img, int sh, int sw, float* __restrict__ out, int S, float m0, float m1, float m2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * S) return;
  const int dy = i / S, dx = i - dy * S;
  const double scale_x = 1.0 / ((double)S / (double)sw), scale_y = 1.0 / ((double)S / (double)sh);
  int sx, a0, a1, sy, b0, b1;
  cv_linear_tap(dx, scale_x, sw, true, sx, a0, a1);
  cv_linear_tap(dy, scale_y, sh, false, sy, b0, b1);
  const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);
  const bool one_tap = sx + 1 >= sw;                          // dx >= xmax: D = S[sx] * 2048
  const uint8_t* p00 = img + ((long long)y0 * sw + sx) * 3;
  const uint8_t* p10 = img + ((long long)y1 * sw + sx) * 3;
  const float mean[3] = {m0, m1, m2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int r0 = one_tap ? (int)p00[c] * 2048 : (int)p00[c] * a0 + (int)p00[c + 3] * a1;
    const int r1 = one_tap ? (int)p10[c] * 2048 : (int)p10[c] * a0 + (int)p10[c + 3] * a1;
    const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    out[(long long)c * S * S + i] = (float)(v & 255) - mean[c];
  }
}

// softmax over the last dimension (output activation, RFB_Net_vgg.py:279-285); one thread per row
__global__ void __launch_bounds__(256)
softmax_lastdim_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int cols) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const float* x = in + r * cols;
    float m = -INFINITY;
    for (int c = 0; c < cols; ++c) m = fmaxf(m, x[c]);
    float s = 0.f;
    for (int c = 0; c < cols; ++c) s += expf(x[c] - m);
    const float inv = 1.0f / s;
    float* y = out + r * cols;
    for (int c = 0; c < cols; ++c) y[c] = expf(x[c] - m) * inv;
  }
}

// rows of at most 32 columns (the class dimension: 2, 20, 21): the row lives in registers, every exponential is evaluated once
// (same operations on the same values as the generic kernel: identical results); 128-bit accesses when COLS % 4 == 0
template <int COLS>
__global__ void __launch_bounds__(256)
softmax_lastdim_small_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float x[COLS];
    if (COLS % 4 == 0) {
      const float4* src = reinterpret_cast<const float4*>(in + r * COLS);
#pragma unroll
      for (int c = 0; c < COLS / 4; ++c) { const float4 v = src[c]; x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w; }
    } else {
#pragma unroll
      for (int c = 0; c < COLS; ++c) x[c] = in[r * COLS + c];
    }
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < COLS; ++c) m = fmaxf(m, x[c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < COLS; ++c) { x[c] = expf(x[c] - m); s += x[c]; }
    const float inv = 1.0f / s;
    if (COLS % 4 == 0) {
      float4* dst = reinterpret_cast<float4*>(out + r * COLS);
#pragma unroll
      for (int c = 0; c < COLS / 4; ++c) dst[c] = make_float4(x[4 * c] * inv, x[4 * c + 1] * inv, x[4 * c + 2] * inv, x[4 * c + 3] * inv);
    } else {
#pragma unroll
      for (int c = 0; c < COLS; ++c) out[r * COLS + c] = x[c] * inv;
    }
  }
}

static int validate_conv(const CtxConvParams* p) {
  CTX_REQUIRE(p, "conv: null params");
  CTX_REQUIRE(!p->pool2, "conv (CUDA-core path): fused pooling is a tensor-core-path feature");
  CTX_REQUIRE(p->in && p->weight, "conv: null tensor pointer");
  CTX_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->Cin > 0 && p->Cout > 0, "conv: bad dims");
  CTX_REQUIRE(p->KH > 0 && p->KW > 0 && p->stride > 0 && p->dil > 0 && p->pad_h >= 0 && p->pad_w >= 0, "conv: bad geometry");
  CTX_REQUIRE(p->in_cstride >= p->in_coffset + p->Cin, "conv: input channel slice out of range");
  const int ho = (p->H + 2 * p->pad_h - p->dil * (p->KH - 1) - 1) / p->stride + 1;
  const int wo = (p->W + 2 * p->pad_w - p->dil * (p->KW - 1) - 1) / p->stride + 1;
  CTX_REQUIRE(ho == p->Ho && wo == p->Wo, "conv: Ho/Wo (%d,%d) inconsistent with geometry (%d,%d)", p->Ho, p->Wo, ho, wo);
  CTX_REQUIRE(p->nseg >= 1 && p->nseg <= 3, "conv: nseg must be 1..3");
  for (int s = 0; s < p->nseg; ++s)
    CTX_REQUIRE(p->seg[s].ptr && p->seg[s].c_begin >= 0 && p->seg[s].c_end <= p->Cout && p->seg[s].c_begin < p->seg[s].c_end,
                "conv: bad output segment %d", s);
  return CTX_OK;
}

int conv_simt_launch(const CtxConvParams* p, cudaStream_t st) {
  int rc = validate_conv(p);
  if (rc) return rc;
  ConvGeom g;
  g.N = p->N; g.H = p->H; g.W = p->W; g.Cin = p->Cin; g.in_cstride = p->in_cstride; g.in_coffset = p->in_coffset;
  g.Cout = p->Cout; g.CoutP = (p->Cout + 3) & ~3; g.KH = p->KH; g.KW = p->KW; g.stride = p->stride;
  g.pad_h = p->pad_h; g.pad_w = p->pad_w; g.dil = p->dil; g.Ho = p->Ho; g.Wo = p->Wo; g.relu = p->relu; g.relu_cend = p->relu_channels > 0 ? p->relu_channels : p->Cout;
  g.K = p->KH * p->KW * p->Cin; g.M = p->N * p->Ho * p->Wo;
  g.res_dtype = p->res_dtype; g.res_cstride = p->res_cstride; g.res_coffset = p->res_coffset;
  SegTable segs; segs.nseg = p->nseg;
  for (int s = 0; s < 3; ++s) segs.seg[s] = p->seg[s < p->nseg ? s : 0];
  dim3 grid(cdiv(g.M, BM), cdiv(g.Cout, BN));
  const bool aligned = (p->Cin % 16 == 0) && (p->in_cstride % 4 == 0) && (p->in_coffset % 4 == 0);
  const float* w = (const float*)p->weight;
#define LAUNCH(T, A) conv_simt_kernel<T, A><<<grid, 256, 0, st>>>((const T*)p->in, w, p->bias, p->residual, g, segs)
  if (p->in_dtype == CTX_F32) { if (aligned) LAUNCH(float, true); else LAUNCH(float, false); }
  else if (p->in_dtype == CTX_BF16) { if (aligned) LAUNCH(__nv_bfloat16, true); else LAUNCH(__nv_bfloat16, false); }
  else if (p->in_dtype == CTX_F16) { if (aligned) LAUNCH(__half, true); else LAUNCH(__half, false); }
  else { set_error("conv: bad in_dtype %d", p->in_dtype); return CTX_ERR_INVALID; }
#undef LAUNCH
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

// Split mode (fp16 hi / lo planes, value = hi + lo): the maximum is taken on the fp32 sums and the winner's two halves are
// copied (re-splitting the fp32 sum would give the same pair: hi = fp16(hi + lo) by construction).  8 channels per thread.
__global__ void __launch_bounds__(256)
maxpool_nhwc_split_kernel(CtxPoolParams p) {
  const int C8 = p.C >> 3;
  const long long total = (long long)p.N * p.Ho * p.Wo * C8;
  const uint16_t* in_hi = reinterpret_cast<const uint16_t*>(p.in);
  const uint16_t* in_lo = reinterpret_cast<const uint16_t*>(p.in_lo);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    long long r = i / C8;
    const int ox = (int)(r % p.Wo); r /= p.Wo;
    const int oy = (int)(r % p.Ho);
    const int n = (int)(r / p.Ho);
    const int y0 = max(oy * p.stride - p.pad, 0), y1 = min(oy * p.stride - p.pad + p.k, p.H);
    const int x0 = max(ox * p.stride - p.pad, 0), x1 = min(ox * p.stride - p.pad + p.k, p.W);
    float m[8];
    uint16_t bh[8], bl[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { m[e] = -INFINITY; bh[e] = 0xFC00u; bl[e] = 0u; }
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        const long long off = (long long)n * p.in_img_stride + (long long)(y * p.W + x) * p.in_pix_stride + c;
        const uint4 vh = *reinterpret_cast<const uint4*>(in_hi + off), vl = *reinterpret_cast<const uint4*>(in_lo + off);
        const uint32_t wh[4] = {vh.x, vh.y, vh.z, vh.w}, wl[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint16_t hh = (uint16_t)(wh[e >> 1] >> ((e & 1) * 16)), ll = (uint16_t)(wl[e >> 1] >> ((e & 1) * 16));
          const float f = __half2float(__ushort_as_half(hh)) + __half2float(__ushort_as_half(ll));
          if (f > m[e]) { m[e] = f; bh[e] = hh; bl[e] = ll; }
        }
      }
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { oh[e] = (uint32_t)bh[2 * e] | ((uint32_t)bh[2 * e + 1] << 16); ol[e] = (uint32_t)bl[2 * e] | ((uint32_t)bl[2 * e + 1] << 16); }
    const long long o = (long long)n * p.out_img_stride + (long long)(oy * p.Wo + ox) * p.out_pix_stride + c;
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + o) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out_lo) + o) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  }
}

int maxpool_launch(const CtxPoolParams* p, cudaStream_t st) {
  CTX_REQUIRE(p && p->in && p->out, "maxpool: null pointer");
  if (p->in_lo || p->out_lo) {
    CTX_REQUIRE(p->in_lo && p->out_lo && p->dtype == CTX_F16 && p->C % 8 == 0 && p->in_pix_stride % 8 == 0 && p->out_pix_stride % 8 == 0 &&
                p->in_img_stride % 8 == 0 && p->out_img_stride % 8 == 0 && ((uintptr_t)p->in) % 16 == 0 && ((uintptr_t)p->out) % 16 == 0 &&
                ((uintptr_t)p->in_lo) % 16 == 0 && ((uintptr_t)p->out_lo) % 16 == 0,
                "maxpool (split mode): needs both fp16 planes, 8-channel aligned");
    CTX_REQUIRE(p->N > 0 && p->k > 0 && p->stride > 0 && p->Ho > 0 && p->Wo > 0 && (p->Ho - 1) * p->stride - p->pad < p->H && (p->Wo - 1) * p->stride - p->pad < p->W,
                "maxpool (split mode): bad dims");
    const long long total8 = (long long)p->N * p->Ho * p->Wo * (p->C / 8);
    maxpool_nhwc_split_kernel<<<(int)std::min<long long>((total8 + 255) / 256, 148LL * 32), 256, 0, st>>>(*p);
    CTX_LAUNCH_CHECK();
    return CTX_OK;
  }
  CTX_REQUIRE(p->N > 0 && p->C > 0 && p->k > 0 && p->stride > 0 && p->Ho > 0 && p->Wo > 0, "maxpool: bad dims");
  CTX_REQUIRE((p->Ho - 1) * p->stride - p->pad < p->H && (p->Wo - 1) * p->stride - p->pad < p->W,
              "maxpool: last window starts outside the input");
  long long total = (long long)p->N * p->Ho * p->Wo * p->C;
  const bool vec = p->dtype != CTX_F32 && p->C % 8 == 0 && p->in_pix_stride % 8 == 0 && p->out_pix_stride % 8 == 0 &&
                   p->in_img_stride % 8 == 0 && p->out_img_stride % 8 == 0 && ((uintptr_t)p->in) % 16 == 0 && ((uintptr_t)p->out) % 16 == 0;
  if (vec) {
    CTX_REQUIRE((long long)p->N * p->Ho <= 65535, "maxpool: too many output rows for the (image, row) grid dimension");
    const dim3 grid((unsigned)((p->Wo * (p->C / 8) + 255) / 256), (unsigned)(p->N * p->Ho));
    const bool bf = p->dtype == CTX_BF16;
    if (p->k == 2) { if (bf) maxpool_nhwc_vec8_kernel<true, 2><<<grid, 256, 0, st>>>(*p); else maxpool_nhwc_vec8_kernel<false, 2><<<grid, 256, 0, st>>>(*p); }
    else if (p->k == 3) { if (bf) maxpool_nhwc_vec8_kernel<true, 3><<<grid, 256, 0, st>>>(*p); else maxpool_nhwc_vec8_kernel<false, 3><<<grid, 256, 0, st>>>(*p); }
    else if (bf) maxpool_nhwc_vec8_kernel<true><<<grid, 256, 0, st>>>(*p);
    else maxpool_nhwc_vec8_kernel<false><<<grid, 256, 0, st>>>(*p);
    CTX_LAUNCH_CHECK();
    return CTX_OK;
  }
  const bool vec4 = p->dtype == CTX_F32 && p->C % 4 == 0 && p->in_pix_stride % 4 == 0 && p->out_pix_stride % 4 == 0 && p->in_img_stride % 4 == 0 &&
                    p->out_img_stride % 4 == 0 && ((uintptr_t)p->in) % 16 == 0 && ((uintptr_t)p->out) % 16 == 0 && (long long)p->N * p->Ho <= 65535;
  if (vec4) {
    const dim3 grid((unsigned)((p->Wo * (p->C / 4) + 255) / 256), (unsigned)(p->N * p->Ho));
    maxpool_nhwc_f32x4_kernel<<<grid, 256, 0, st>>>(*p);
    CTX_LAUNCH_CHECK();
    return CTX_OK;
  }
  int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
  maxpool_nhwc_kernel<<<blocks, 256, 0, st>>>(*p);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

int nchw_to_nhwc_launch(const float* in, void* out, int N, int C, int H, int W, int dtype, cudaStream_t st) {
  CTX_REQUIRE(in && out && N > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad arguments");
  long long total = (long long)N * H * W;
  int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
  nchw_to_nhwc_kernel<<<blocks, 256, 0, st>>>(in, out, N, C, H, W, dtype);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

int base_transform_launch(const unsigned char* img, float* out, int N, int H, int W, const float* means3, cudaStream_t st) {
  CTX_REQUIRE(img && out && means3 && N > 0 && H > 0 && W > 0, "base_transform: bad arguments");
  long long total = (long long)N * H * W;
  int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
  u8hwc_to_f32chw_kernel<<<blocks, 256, 0, st>>>(img, out, N, H, W, means3[0], means3[1], means3[2]);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

int resize_base_transform_launch(const unsigned char* img, int sh, int sw, float* out, int S, const float* means3, cudaStream_t st) {
  CTX_REQUIRE(img && out && means3 && sh > 0 && sw > 0 && S > 0, "base_transform_resize: bad arguments");
  resize_base_transform_kernel<<<cdiv((long long)S * S, 256), 256, 0, st>>>(img, sh, sw, out, S, means3[0], means3[1], means3[2]);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

int patch27_launch(const float* in, void* out, int N, int H, int W, int dtype, cudaStream_t st) {
  CTX_REQUIRE(in && out && N > 0 && H > 0 && W > 0, "patch27: bad arguments");
  CTX_REQUIRE(dtype == CTX_BF16 || dtype == CTX_F16, "patch27: 16-bit output only");
  long long total = (long long)N * H * W;
  int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
  if (dtype == CTX_BF16) patch27_kernel<true><<<blocks, 256, 0, st>>>(in, (uint16_t*)out, N, H, W);
  else patch27_kernel<false><<<blocks, 256, 0, st>>>(in, (uint16_t*)out, N, H, W);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

int softmax_launch(const float* in, float* out, long long rows, int cols, cudaStream_t st) {
  CTX_REQUIRE(in && out && rows >= 0 && cols > 0, "softmax: bad arguments");
  if (rows == 0) return CTX_OK;
  int blocks = (int)std::min<long long>((rows + 255) / 256, 148LL * 32);
  const bool al = ((uintptr_t)in) % 16 == 0 && ((uintptr_t)out) % 16 == 0;
  if (cols == 20 && al) softmax_lastdim_small_kernel<20><<<blocks, 256, 0, st>>>(in, out, rows);
  else if (cols == 21) softmax_lastdim_small_kernel<21><<<blocks, 256, 0, st>>>(in, out, rows);
  else if (cols == 2) softmax_lastdim_small_kernel<2><<<blocks, 256, 0, st>>>(in, out, rows);
  else softmax_lastdim_kernel<<<blocks, 256, 0, st>>>(in, out, rows, cols);
  CTX_LAUNCH_CHECK();
  return CTX_OK;
}

}  // namespace ctx

extern "C" int ctx_conv2d_simt(const CtxConvParams* p, void* stream) { return ctx::conv_simt_launch(p, (cudaStream_t)stream); }
extern "C" int ctx_maxpool2d_nhwc(const CtxPoolParams* p, void* stream) { return ctx::maxpool_launch(p, (cudaStream_t)stream); }
extern "C" int ctx_nchw_to_nhwc(const float* in, void* out, int N, int C, int H, int W, int out_dtype, void* stream) {
  return ctx::nchw_to_nhwc_launch(in, out, N, C, H, W, out_dtype, (cudaStream_t)stream);
}
extern "C" int ctx_base_transform_resize(const unsigned char* img_hwc, int src_h, int src_w, float* out_chw, int size, const float* means3, void* stream) {
  return ctx::resize_base_transform_launch(img_hwc, src_h, src_w, out_chw, size, means3, (cudaStream_t)stream);
}
extern "C" int ctx_base_transform(const unsigned char* img_hwc, float* out_chw, int N, int H, int W, const float* means3, void* stream) {
  return ctx::base_transform_launch(img_hwc, out_chw, N, H, W, means3, (cudaStream_t)stream);
}
extern "C" int ctx_nchw_to_patch27(const float* in, void* out, int N, int H, int W, int out_dtype, void* stream) {
  return ctx::patch27_launch(in, out, N, H, W, out_dtype, (cudaStream_t)stream);
}
extern "C" int ctx_softmax_lastdim(const float* in, float* out, long long rows, int cols, void* stream) {
  return ctx::softmax_launch(in, out, rows, cols, (cudaStream_t)stream);
}
