// SPADEGenerator4 inference path (reference models/SPADE_related.py:70-85,128-149,1404-1605; eval mode, batch-sharded).
//
// Activations are NHWC fp32 so that every convolution is an implicit GEMM  out[pixel, cout] = sum_k A(pixel, k) W[cout, k]
// with k = (ky, kx, cin) and cin contiguous: the A operand is the `Im2col` functor below (reflection padding is index
// arithmetic in its row token — no padded copy), the B operand the re-packed weight matrix, and the
// contraction runs on the tcgen05 3xTF32 engine of tc_gemm.cuh (fp32-level accuracy: SURVEY App. F shows single-pass
// TF32/BF16 miss the 1e-4 contract by 40-300x).  The gamma and beta convolutions of a SPADE4 are ONE contraction over
// interleaved weight rows whose epilogue (TcEpiSpade) computes lrelu(x_hat * (1 + gamma) + beta) directly: gamma and beta
// never reach HBM (they are 61 % of the generator's FLOPs and would be 2 x C x H x W floats per SPADE4 otherwise).
#include <type_traits>

#include "../../include/sln_b200.h"
// 16 producer warps, like the VAE engine: the long-K convolutions accumulate over several TMEM segments into a per-thread
// running sum, and with the accumulator columns split over PROD_WARPS/4 column groups (tc_gemm.cuh) that sum is 32 registers
// at 16 warps, which fits the 96-register cap (with the 64-register sum of an 8-warp-style split the 16-warp build spilled
// and lost: 346 -> 332 images/s).  Measured, same box: 8 warps 349 images/s, 16 warps 392.
// Two producer groups of 8 warps alternate chunks, one chunk prefetched per group (tc_gemm.cuh defaults).
#include "gemm.cuh"
#include "tc_gemm.cuh"

namespace sln {
namespace {

__device__ __forceinline__ int reflect(int v, int n) { return v < 0 ? -v : (v >= n ? 2 * n - 2 - v : v); }

// Virtual [B*H*W, ks*ks*C] matrix over an NHWC tensor: row = output pixel, column k = (ky*ks + kx)*C + c, reading the
// input at the REFLECTED position (y + ky - ks/2, x + kx - ks/2)   (nn.ReflectionPad2d(1) + Conv2d(k=3, padding=0)), or the
// pixel itself for ks == 1.  Optional lazy ReLU on load.
template <bool ZP>   // ZP: zero padding (compile-time: the reflect variant carries no validity bits and no extra branch in its hot loop)
struct Im2colT {
  static constexpr bool kTwoLoads = false;
  const float* p;
  int H, W, C, ks, rows, cols, relu, vec;
  int cshift;   // log2(C) when C is a power of two (the usual case), else -1: k -> (tap, c) without an integer division
  static constexpr int zpad = ZP ? 1 : 0;   // 0: ReflectionPad2d(1) (SPADE4 / SPADEResnetBlock4); 1: zero padding = nn.Conv2d(padding=1) of the plain SPADE blocks
  __device__ __forceinline__ float elem(int r, int k) const {
    const int tap = ks == 3 ? k / C : 0, c = k - tap * C;
    const int ky = ks == 3 ? tap / 3 : 1, kx = ks == 3 ? tap - 3 * (tap / 3) : 1;
    const int hw = H * W, b = r / hw, rem = r - b * hw, y = rem / W, x = rem - y * W;
    if (ks == 3 && zpad && ((unsigned)(y + ky - 1) >= (unsigned)H || (unsigned)(x + kx - 1) >= (unsigned)W)) return 0.f;
    const int yy = ks == 3 ? reflect(y + ky - 1, H) : y, xx = ks == 3 ? reflect(x + kx - 1, W) : x;
    float v = __ldg(p + ((size_t)b * hw + (size_t)yy * W + xx) * C + c);
    return relu ? fmaxf(v, 0.f) : v;
  }
  __device__ __forceinline__ float at_t(int r, int k) const { return (r < rows && k < cols) ? elem(r, k) : 0.f; }
  __device__ __forceinline__ float4 ld4(int r, int k) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= rows || k >= cols) return v;
    if (vec) {   // C % 4 == 0: the quad stays inside one tap
      const int tap = ks == 3 ? k / C : 0, c = k - tap * C;
      const int ky = ks == 3 ? tap / 3 : 1, kx = ks == 3 ? tap - 3 * (tap / 3) : 1;
      const int hw = H * W, b = r / hw, rem = r - b * hw, y = rem / W, x = rem - y * W;
      if (ks == 3 && zpad && ((unsigned)(y + ky - 1) >= (unsigned)H || (unsigned)(x + kx - 1) >= (unsigned)W)) return v;
      const int yy = ks == 3 ? reflect(y + ky - 1, H) : y, xx = ks == 3 ? reflect(x + kx - 1, W) : x;
      v = ldg4(p + ((size_t)b * hw + (size_t)yy * W + xx) * C + c);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    } else {
      v.x = elem(r, k);
      if (k + 1 < cols) v.y = elem(r, k + 1);
      if (k + 2 < cols) v.z = elem(r, k + 2);
      if (k + 3 < cols) v.w = elem(r, k + 3);
    }
    return v;
  }
  // two-phase API of the tensor-core loader: the token holds the image base and the three reflected row / column offsets
  struct TokR { const float* base; int y0, y1, y2, x0, x1, x2; };
  struct TokZ { const float* base; int y0, y1, y2, x0, x1, x2; int valid; };   // valid: bit ky = row tap inside the image, bit 3 + kx = column tap
  using Tok = typename std::conditional<ZP, TokZ, TokR>::type;
  __device__ __forceinline__ Tok token(int r) const {
    r = min(r, rows - 1);
    const int hw = H * W, b = r / hw, rem = r - b * hw, y = rem / W, x = rem - y * W;
    Tok t;
    t.base = p + (size_t)b * hw * C;
    if (ks == 3) {
      t.y0 = reflect(y - 1, H) * W; t.y1 = y * W; t.y2 = reflect(y + 1, H) * W;
      t.x0 = reflect(x - 1, W); t.x1 = x; t.x2 = reflect(x + 1, W);
      if constexpr (ZP) t.valid = (y >= 1 ? 1 : 0) | 2 | (y + 1 < H ? 4 : 0) | (x >= 1 ? 8 : 0) | 16 | (x + 1 < W ? 32 : 0);
    } else {
      t.y0 = t.y1 = t.y2 = y * W; t.x0 = t.x1 = t.x2 = x;
      if constexpr (ZP) t.valid = 63;
    }
    return t;
  }
  __device__ __forceinline__ int clampc(int c) const { return min(c, cols - 4); }
  __device__ __forceinline__ void fetch4(const Tok& t, int k, float4& a, float4&) const {
    const int tap = ks == 3 ? (cshift >= 0 ? (k >> cshift) : k / C) : 0, c = k - tap * C;
    const int ky = ks == 3 ? (tap * 11) >> 5 : 1, kx = ks == 3 ? tap - 3 * ky : 1;      // tap / 3 for tap < 9
    const int yo = ky == 0 ? t.y0 : (ky == 1 ? t.y1 : t.y2), xo = kx == 0 ? t.x0 : (kx == 1 ? t.x1 : t.x2);
    if constexpr (ZP) {
      if (!(((t.valid >> ky) & 1) && ((t.valid >> (3 + kx)) & 1))) { a = make_float4(0.f, 0.f, 0.f, 0.f); return; }   // outside: zero
    }
    a = ldg4(t.base + (size_t)(unsigned)((yo + xo) * C + c));
  }
  __device__ __forceinline__ float4 finish4(const Tok&, int, float4 v, float4) const {
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    return v;
  }
  bool vec_ok() const { return vec != 0; }
};

using Im2col = Im2colT<false>;
using Im2colZ = Im2colT<true>;
template <bool ZP>
Im2colT<ZP> make_im2col(const float* p, int B, int H, int W, int C, int ks, int relu) {
  Im2colT<ZP> a;
  a.p = p; a.H = H; a.W = W; a.C = C; a.ks = ks; a.rows = B * H * W; a.cols = ks * ks * C; a.relu = relu;
  a.vec = (C % 4 == 0 && (uintptr_t)p % 16 == 0) ? 1 : 0;
  a.cshift = -1;
  for (int sh = 0; sh < 30; ++sh) if ((1 << sh) == C) a.cshift = sh;
  return a;
}

// SPADE4 modulation epilogue (reference SPADE_related.py:139-144,1451) fused with the block's leaky_relu (:1490-1491,1504):
// out[pix, c] = act( (x[pix, c] - mean[b]) * inv[b] * (1 + gamma) + beta ),  gamma/beta = the two halves of the tile + bias.
struct TcEpiSpade {
  float* out; const float* x; int C;
  const float* bias_g; const float* bias_b;
  const float* mean; const float* inv; int HW; float slope;
  int sb, sc;   // statistics layout: value of (sample b, channel c) at [b * sb + c * sc] — (1, 0): one per sample (LayerNorm2D);
                // (C, 1): one per sample and channel (InstanceNorm2d); (0, 1): one per channel (eval-mode BatchNorm2d)
  static constexpr bool kStats = false;
  static constexpr bool kPaired = true;
  __device__ __forceinline__ bool wants_stats() const { return false; }
  __device__ __forceinline__ void apply4(int, int, int, float4, float (&)[4], float (&)[4]) const {}
  __device__ __forceinline__ float4 load_pair(int i, int c) const { return ldg4(x + (size_t)i * C + c); }   // issued before the accumulator read-out
  __device__ __forceinline__ void apply_pair(int i, int c, float4 g, float4 b, float4 xv) const {
    const int bi = i / HW;
    float4 m, iv;
    if (sc) { m = ldg4(mean + (size_t)bi * sb + c); iv = ldg4(inv + (size_t)bi * sb + c); }
    else { const float m1 = __ldg(mean + bi * sb), i1 = __ldg(inv + bi * sb); m = make_float4(m1, m1, m1, m1); iv = make_float4(i1, i1, i1, i1); }
    const float4 bg = ldg4(bias_g + c), bb = ldg4(bias_b + c);
    float4 o;
    o.x = fmaf((xv.x - m.x) * iv.x, 1.f + (g.x + bg.x), b.x + bb.x);
    o.y = fmaf((xv.y - m.y) * iv.y, 1.f + (g.y + bg.y), b.y + bb.y);
    o.z = fmaf((xv.z - m.z) * iv.z, 1.f + (g.z + bg.z), b.z + bb.z);
    o.w = fmaf((xv.w - m.w) * iv.w, 1.f + (g.w + bg.w), b.w + bb.w);
    o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
    o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
    *reinterpret_cast<float4*>(out + (size_t)i * C + c) = o;
  }
  __device__ __forceinline__ void finalize(int, double, double, double, double, int) const {}
  __device__ __forceinline__ const sln_bn_sync* sync() const { return nullptr; }
  __device__ __forceinline__ int slot0() const { return 0; }
  __device__ __forceinline__ int rows() const { return 0; }
  __device__ __forceinline__ float* partial() const { return nullptr; }
  __device__ __forceinline__ unsigned* counter() const { return nullptr; }
};

// direct fallback of the modulation for channel counts the paired tiles cannot hold (2C < 32: reduced test models only)
template <class AT>
__global__ void k_modulate_direct(const AT A, const float* __restrict__ Wg, const float* __restrict__ Wb, const float* __restrict__ bg,
                                  const float* __restrict__ bb, int C, const float* __restrict__ x, const float* __restrict__ mean,
                                  const float* __restrict__ inv, int HW, float slope, float* out, int sb, int sc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= A.rows * C) return;
  const int i = idx / C, c = idx - i * C;
  float g = __ldg(bg + c), b = __ldg(bb + c);
  for (int k = 0; k < A.cols; ++k) {
    const float a = A.at_t(i, k);
    g = fmaf(a, __ldg(Wg + (size_t)c * A.cols + k), g);
    b = fmaf(a, __ldg(Wb + (size_t)c * A.cols + k), b);
  }
  const int bi = i / HW;
  float o = fmaf((__ldg(x + idx) - __ldg(mean + (size_t)bi * sb + c * sc)) * __ldg(inv + (size_t)bi * sb + c * sc), 1.f + g, b);
  out[idx] = o > 0.f ? o : o * slope;
}

// ------------------------------------------------------------------------------------------------ InstanceNorm2d (plain SPADE / Conv2dBlock)
// per (sample, channel): mean and 1 / sqrt(biased var + eps) over H*W of an NHWC tensor (nn.InstanceNorm2d(affine=False,
// track_running_stats=False), reference SPADE_related.py:34,311).  CTA = (32 channels, sample); 8 pixel lanes per channel, fp64,
// fixed-order shared-memory fold: deterministic.
__global__ void __launch_bounds__(256) k_in_stats(const float* __restrict__ x, int HW, int C, float eps, float* mean, float* inv) {
  const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), pl = threadIdx.x >> 5;
  double s = 0.0, q = 0.0;
  if (c < C) {
    const float* xb = x + (size_t)b * HW * C + c;
    for (int p = pl; p < HW; p += 8) { const double v = (double)__ldg(xb + (size_t)p * C); s += v; q += v * v; }
  }
  __shared__ double sh[2][8][32];
  sh[0][pl][threadIdx.x & 31] = s; sh[1][pl][threadIdx.x & 31] = q;
  __syncthreads();
  if (pl == 0 && c < C) {
    double S = 0.0, Q = 0.0;
    for (int k = 0; k < 8; ++k) { S += sh[0][k][threadIdx.x]; Q += sh[1][k][threadIdx.x]; }
    const double m = S / HW;
    double var = Q / HW - m * m;
    if (var < 0.0) var = 0.0;
    mean[(size_t)b * C + c] = (float)m;
    inv[(size_t)b * C + c] = (float)(1.0 / sqrt(var + (double)eps));
  }
}
// out = act((x - mean[b,c]) * inv[b,c]);  act: 0 none, 1 ReLU   (Conv2dBlock.norm + activation, SPADE_related.py:58-63)
__global__ void k_norm_act(const float* __restrict__ x, long long n4, int HWC, int C, const float* __restrict__ mean, const float* __restrict__ inv,
                           int sb, int sc, int act, float* out) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= n4) return;
  const long long i = i4 * 4;
  const int b = (int)(i / HWC), c = (int)(i % C);
  const float4 v = ldg4(x + i);
  float4 m, iv;
  if (sc) { m = ldg4(mean + (size_t)b * sb + c); iv = ldg4(inv + (size_t)b * sb + c); }
  else { const float m1 = __ldg(mean + b * sb), i1 = __ldg(inv + b * sb); m = make_float4(m1, m1, m1, m1); iv = make_float4(i1, i1, i1, i1); }
  float4 o = make_float4((v.x - m.x) * iv.x, (v.y - m.y) * iv.y, (v.z - m.z) * iv.z, (v.w - m.w) * iv.w);
  if (act == 1) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
  *reinterpret_cast<float4*>(out + i) = o;
}
// NCHW label map [B, nc, S, S] -> NHWC [B, h, w, cpad] (channels >= nc zero): bilinear with align_corners=False (F.interpolate(...,
// mode='bilinear'), SPADE_related.py:330) or nearest (F.interpolate default, :226,228).  A thread per output pixel and channel quad.
__global__ void k_seg_resize(const float* __restrict__ seg, int B, int nc, int S, int h, int w, int cpad, int nearest, float* out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = cpad / 4;
  if (idx >= (long long)B * h * w * q) return;
  const int cq = (int)(idx % q) * 4;
  const long long pix = idx / q;
  const int x = (int)(pix % w), y = (int)((pix / w) % h), b = (int)(pix / ((long long)w * h));
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  if (nearest) {
    const int sy = min((int)floorf(y * ((float)S / h)), S - 1), sx = min((int)floorf(x * ((float)S / w)), S - 1);
    for (int e = 0; e < 4; ++e) if (cq + e < nc) o[e] = __ldg(seg + (((size_t)b * nc + cq + e) * S + sy) * S + sx);
  } else {
    float fy = ((float)y + 0.5f) * ((float)S / h) - 0.5f, fx = ((float)x + 0.5f) * ((float)S / w) - 0.5f;
    fy = fmaxf(fy, 0.f); fx = fmaxf(fx, 0.f);
    const int y0 = min((int)fy, S - 1), x0 = min((int)fx, S - 1), y1 = min(y0 + 1, S - 1), x1 = min(x0 + 1, S - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    for (int e = 0; e < 4; ++e) {
      if (cq + e >= nc) continue;
      const float* pl = seg + ((size_t)b * nc + cq + e) * S * S;
      o[e] = hy * (hx * __ldg(pl + (size_t)y0 * S + x0) + lx * __ldg(pl + (size_t)y0 * S + x1)) +
             ly * (hx * __ldg(pl + (size_t)y1 * S + x0) + lx * __ldg(pl + (size_t)y1 * S + x1));
    }
  }
  *reinterpret_cast<float4*>(out + idx * 4) = make_float4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------------ LayerNorm2D statistics
// per sample: mean and 1 / (unbiased std + eps) over C*H*W (reference :139-144).  fp64 accumulation.
__global__ void __launch_bounds__(256) k_ln_partial(const float* __restrict__ x, long long n, double* acc) {
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * n;
  double s = 0.0, q = 0.0;
  const long long n4 = n / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {           // four loads in flight per thread
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(xb) + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s += (double)v[u].x + (double)v[u].y + (double)v[u].z + (double)v[u].w;
      q += (double)v[u].x * v[u].x + (double)v[u].y * v[u].y + (double)v[u].z * v[u].z + (double)v[u].w * v[u].w;
    }
  }
  for (; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xb) + i);
    s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
    q += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) { s += xb[i]; q += (double)xb[i] * xb[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  __shared__ double ss[8], sq[8];
  if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0.0, Q = 0.0;
    for (int w = 0; w < 8; ++w) { S += ss[w]; Q += sq[w]; }
    atomicAdd(acc + 2 * b, S);
    atomicAdd(acc + 2 * b + 1, Q);
  }
}
__global__ void k_ln_final(const double* acc, int B, long long n, float eps, float* mean, float* inv) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double m = acc[2 * b] / (double)n;
  double var = (acc[2 * b + 1] - (double)n * m * m) / (double)(n > 1 ? n - 1 : 1);
  if (var < 0.0) var = 0.0;
  mean[b] = (float)m;
  inv[b] = (float)(1.0 / (sqrt(var) + (double)eps));
}

// ------------------------------------------------------------------------------------------------ segmentation features
// SPADE4 part 2 up to the concat (reference :1444-1447): resize the map to (h, w) [bilinear, align_corners=False; or the
// nearest-neighbour map seg_1 that head_0 receives, :1579], depth channel -> reflect-pad 3x3 conv (1 -> nd) + LeakyReLU(0.01),
// concatenated with the nc-1 label channels -> NHWC [B, h, w, nd + nc - 1].
__device__ __forceinline__ float seg_sample(const float* __restrict__ ch, int S, int h, int w, int mode, int y, int x) {
  if (mode == 1) {   // nearest (torch: min(floor(dst * scale), S - 1), scale = S / size)
    const int sy = min((int)floorf((float)y * ((float)S / (float)h)), S - 1), sx = min((int)floorf((float)x * ((float)S / (float)w)), S - 1);
    return __ldg(ch + (size_t)sy * S + sx);
  }
  if (h == S && w == S) return __ldg(ch + (size_t)y * S + x);
  const float fy = fmaxf(((float)S / (float)h) * ((float)y + 0.5f) - 0.5f, 0.f), fx = fmaxf(((float)S / (float)w) * ((float)x + 0.5f) - 0.5f, 0.f);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < S - 1 ? 1 : 0), x1 = x0 + (x0 < S - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const float v00 = __ldg(ch + (size_t)y0 * S + x0), v01 = __ldg(ch + (size_t)y0 * S + x1);
  const float v10 = __ldg(ch + (size_t)y1 * S + x0), v11 = __ldg(ch + (size_t)y1 * S + x1);
  return (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}
// A CTA of 256 threads owns 64 consecutive pixels: thread (pixel p = tid % 64, group g = tid / 64) computes the features
// j = g, g + 4, ... of pixel p (a thread-per-pixel loop is a chain of nc dependent-latency loads: 33 us even for an 8 x 8 map).
// The 64 rows are contiguous in `out`: they are staged in shared memory (odd row stride: conflict-free) and written out as one
// coalesced block (thread-per-row stores touch 32 sectors per instruction).
constexpr int kSegPix = 64, kSegGroups = 4;
__global__ void __launch_bounds__(kSegPix * kSegGroups) k_seg_features(const float* __restrict__ seg, int B, int nc, int S, int mode, int h, int w,
                                                                       const float* __restrict__ dw, const float* __restrict__ db, int nd,
                                                                       float* __restrict__ out) {
  extern __shared__ float s_row[];   // [kSegPix][C | 1]
  const int C = nd + nc - 1, LD = C | 1;
  const int p = threadIdx.x % kSegPix, g = threadIdx.x / kSegPix;
  const int pix0 = blockIdx.x * kSegPix, pix = pix0 + p, P = B * h * w;
  if (pix < P) {
    const int b = pix / (h * w), rem = pix - b * h * w, y = rem / w, x = rem - y * w;
    const float* sb = seg + (size_t)b * nc * S * S;
    float* o = s_row + p * LD;
    if (g < nd) {                     // warp-uniform (64 pixels per group = 2 warps)
      float d[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) d[t] = seg_sample(sb, S, h, w, mode, reflect(y + t / 3 - 1, h), reflect(x + t % 3 - 1, w));
      for (int j = g; j < nd; j += kSegGroups) {
        float acc = __ldg(db + j);
#pragma unroll
        for (int t = 0; t < 9; ++t) acc = fmaf(d[t], __ldg(dw + j * 9 + t), acc);
        o[j] = acc > 0.f ? acc : acc * 0.01f;
      }
    }
    int c = 1 + g;
    for (; c + 3 * kSegGroups < nc; c += 4 * kSegGroups) {      // four independent samples in flight
      const float v0 = seg_sample(sb + (size_t)c * S * S, S, h, w, mode, y, x);
      const float v1 = seg_sample(sb + (size_t)(c + kSegGroups) * S * S, S, h, w, mode, y, x);
      const float v2 = seg_sample(sb + (size_t)(c + 2 * kSegGroups) * S * S, S, h, w, mode, y, x);
      const float v3 = seg_sample(sb + (size_t)(c + 3 * kSegGroups) * S * S, S, h, w, mode, y, x);
      o[nd + c - 1] = v0; o[nd + c + kSegGroups - 1] = v1; o[nd + c + 2 * kSegGroups - 1] = v2; o[nd + c + 3 * kSegGroups - 1] = v3;
    }
    for (; c < nc; c += kSegGroups) o[nd + c - 1] = seg_sample(sb + (size_t)c * S * S, S, h, w, mode, y, x);
  }
  __syncthreads();
  const int n = min(kSegPix, P - pix0) * C;
  float* ob = out + (size_t)pix0 * C;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int r = e / C;
    ob[e] = s_row[r * LD + (e - r * C)];
  }
}

// ------------------------------------------------------------------------------------------------ 2x upsampling (NHWC)
// nn.Upsample(scale_factor=2, mode='nearest' | 'bilinear' [align_corners=False])   (reference :1544-1545)
// blockIdx.x = output row (b, oy), blockIdx.y = 256-float4 slice of the row: 32-bit index arithmetic only (a flat 64-bit index
// costs four 64-bit divisions per float4: 293 us for the 128 -> 256 step, 1.1 TB/s).
__global__ void __launch_bounds__(256) k_upsample2x(const float* __restrict__ x, int B, int H, int W, int C, int bilinear, float* __restrict__ out) {
  const int c4n = C / 4;
  const int e = blockIdx.y * blockDim.x + threadIdx.x;     // (ox, c4) within the row
  if (e >= 2 * W * c4n) return;
  const int ox = e / c4n, c4 = e - ox * c4n;
  const int row = blockIdx.x, b = row / (2 * H), oy = row - b * 2 * H;
  const float* xb = x + (size_t)b * H * W * C + c4 * 4;
  float4 v;
  if (!bilinear) {
    v = ldg4(xb + ((size_t)(oy >> 1) * W + (ox >> 1)) * C);
  } else {
    const float fy = fmaxf(0.5f * ((float)oy + 0.5f) - 0.5f, 0.f), fx = fmaxf(0.5f * ((float)ox + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)fy, x0 = (int)fx, y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float4 a = ldg4(xb + ((size_t)y0 * W + x0) * C), bq = ldg4(xb + ((size_t)y0 * W + x1) * C);
    const float4 c = ldg4(xb + ((size_t)y1 * W + x0) * C), dq = ldg4(xb + ((size_t)y1 * W + x1) * C);
    v.x = (1.f - ly) * ((1.f - lx) * a.x + lx * bq.x) + ly * ((1.f - lx) * c.x + lx * dq.x);
    v.y = (1.f - ly) * ((1.f - lx) * a.y + lx * bq.y) + ly * ((1.f - lx) * c.y + lx * dq.y);
    v.z = (1.f - ly) * ((1.f - lx) * a.z + lx * bq.z) + ly * ((1.f - lx) * c.z + lx * dq.z);
    v.w = (1.f - ly) * ((1.f - lx) * a.w + lx * bq.w) + ly * ((1.f - lx) * c.w + lx * dq.w);
  }
  *reinterpret_cast<float4*>(out + ((size_t)row * 2 * W + ox) * C + c4 * 4) = v;
}

// ------------------------------------------------------------------------------------------------ squeeze-excite + residual
// SEBlock2 (reference :81-85, reduction 8) and the block's residual add (:1493).
// pooling partials: a CTA sums rows [r0, r1) of image b.  With C / 4 < 256 column lanes the threads also split the rows (RL row lanes
// per column, combined through shared memory) — at C = 64 a column-only mapping left 240 of the 256 threads idle (248 us at 256 x 256).
__global__ void __launch_bounds__(256) k_se_pool(const float* __restrict__ dx, int HW, int C, int rows_per, float* __restrict__ partial) {
  __shared__ float4 s_acc[256];
  const int b = blockIdx.y, chunk = blockIdx.x, nch = gridDim.x;
  const int r0 = chunk * rows_per, r1 = min(HW, r0 + rows_per);
  const int c4n = C / 4;
  if (c4n >= 256 || 256 % c4n != 0) {
    for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = r0; r < r1; ++r) {
        const float4 v = ldg4(dx + ((size_t)b * HW + r) * C + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(partial + ((size_t)b * nch + chunk) * C + c) = acc;
    }
    return;
  }
  const int RL = 256 / c4n, cl = threadIdx.x % c4n, rl = threadIdx.x / c4n;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* base = dx + (size_t)b * HW * C + cl * 4;
  int r = r0 + rl;
  for (; r + 3 * RL < r1; r += 4 * RL) {
    const float4 v0 = ldg4(base + (size_t)r * C), v1 = ldg4(base + (size_t)(r + RL) * C);
    const float4 v2 = ldg4(base + (size_t)(r + 2 * RL) * C), v3 = ldg4(base + (size_t)(r + 3 * RL) * C);
    acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
    acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
  }
  for (; r < r1; r += RL) {
    const float4 v = ldg4(base + (size_t)r * C);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  if (rl == 0) {
    for (int k = 1; k < RL; ++k) {             // fixed order: reproducible
      const float4 v = s_acc[k * c4n + cl];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(partial + ((size_t)b * nch + chunk) * C + cl * 4) = acc;
  }
}
// 1024 threads per image: fc1 and fc2 are warp-per-output-row dot products with lanes along the (contiguous) reduction axis
__global__ void __launch_bounds__(1024) k_se_fc(const float* __restrict__ partial, int nch, int HW, int C, const float* __restrict__ W1,
                                                const float* __restrict__ W2, int Ch, float* __restrict__ svec) {
  extern __shared__ float sm[];   // pooled [C] | hidden [Ch]
  float* pooled = sm;
  float* hidden = sm + C;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < nch; ++k) s += __ldg(partial + ((size_t)b * nch + k) * C + c);
    pooled[c] = s / (float)HW;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < Ch; j += nw) {
    float s = 0.f;
#pragma unroll 4
    for (int c = lane; c < C; c += 32) s = fmaf(__ldg(W1 + (size_t)j * C + c), pooled[c], s);
    s = warp_sum(s);
    if (lane == 0) hidden[j] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int c = warp; c < C; c += nw) {
    float s = 0.f;
    for (int j = lane; j < Ch; j += 32) s = fmaf(__ldg(W2 + (size_t)c * Ch + j), hidden[j], s);
    s = warp_sum(s);
    if (lane == 0) svec[(size_t)b * C + c] = 1.f / (1.f + expf(-s));
  }
}
__global__ void k_se_apply(const float* __restrict__ dx, const float* __restrict__ xs, const float* __restrict__ svec, long long n4, int HWC4, int C4,
                           float* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int b = (int)(i / HWC4), c4 = (int)(i % C4);
  const float4 d = __ldg(reinterpret_cast<const float4*>(dx) + i), s = __ldg(reinterpret_cast<const float4*>(svec) + (size_t)b * C4 + c4);
  const float4 r = __ldg(reinterpret_cast<const float4*>(xs) + i);
  reinterpret_cast<float4*>(out)[i] = make_float4(fmaf(d.x, s.x, r.x), fmaf(d.y, s.y, r.y), fmaf(d.z, s.z, r.z), fmaf(d.w, s.w, r.w));
}

// ------------------------------------------------------------------------------------------------ conv_img + tanh
// leaky_relu(x, 0.2) -> Conv2d(nf, 3, 5, padding=2 [zeros]) -> tanh (reference :1602-1603); NHWC in, NCHW out.  N = 3 is far
// too narrow for tensor-core tiles: one thread per pixel, weights in shared memory (broadcast reads).
template <int MAXC>
__global__ void __launch_bounds__(128) k_to_rgb(const float* __restrict__ x, int B, int H, int W, int Cin, const float* __restrict__ Wt,
                                                const float* __restrict__ bias, int Cout, int ks, float slope, float* pre, float* out) {
  extern __shared__ float s_w[];   // [ks*ks*Cin][MAXC]
  const int K = ks * ks * Cin;
  for (int e = threadIdx.x; e < K * MAXC; e += blockDim.x) {
    const int k = e / MAXC, co = e - k * MAXC;
    s_w[e] = co < Cout ? __ldg(Wt + (size_t)co * K + k) : 0.f;
  }
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= B * H * W) return;
  const int b = pix / (H * W), rem = pix - b * H * W, y = rem / W, xq = rem - y * W;
  float acc[MAXC];
#pragma unroll
  for (int co = 0; co < MAXC; ++co) acc[co] = co < Cout ? __ldg(bias + co) : 0.f;
  const int pad = ks / 2;
  for (int ky = 0; ky < ks; ++ky) {
    const int yy = y + ky - pad;
    if (yy < 0 || yy >= H) continue;
    for (int kx = 0; kx < ks; ++kx) {
      const int xx = xq + kx - pad;
      if (xx < 0 || xx >= W) continue;
      const float* src = x + ((size_t)b * H * W + (size_t)yy * W + xx) * Cin;
      const float* wk = s_w + (size_t)((ky * ks + kx) * Cin) * MAXC;
      for (int c = 0; c < Cin; c += 4) {
        float4 v = ldg4(src + c);
        v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
        v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
#pragma unroll
        for (int co = 0; co < MAXC; ++co)
          acc[co] = fmaf(v.x, wk[(c + 0) * MAXC + co], fmaf(v.y, wk[(c + 1) * MAXC + co], fmaf(v.z, wk[(c + 2) * MAXC + co], fmaf(v.w, wk[(c + 3) * MAXC + co], acc[co]))));
      }
    }
  }
  for (int co = 0; co < Cout; ++co) {
    const size_t o = ((size_t)b * Cout + co) * H * W + (size_t)y * W + xq;
    if (pre) pre[o] = acc[co];
    out[o] = tanhf(acc[co]);
  }
}

// The same layer for the shapes the generators use (KS x KS, Cin % 16 == 0): a CTA of 128 threads owns a 32 x 16 pixel tile, stages
// the activated input (tile + halo, 16 channels at a time, pixel stride 20 floats: conflict-free LDS.128) in shared memory with
// coalesced loads, and every thread computes 4 vertically adjacent pixels so that one weight read feeds 4 pixels and one input
// read up to KS taps (the thread-per-pixel kernel above re-reads the 25 x 256 B neighbourhood of every pixel with 32 cache lines
// per load instruction: 1.9 ms at 16 x 256 x 256 x 64).
constexpr int kRgbTX = 32, kRgbTY = 16, kRgbCC = 16, kRgbLD = kRgbCC + 4;
template <int KS>
__global__ void __launch_bounds__(128, 2) k_to_rgb_tiled(const float* __restrict__ x, int B, int H, int W, int Cin, const float* __restrict__ Wt,
                                                         const float* __restrict__ bias, int Cout, float slope, float* pre, float* out) {
  constexpr int PAD = KS / 2, SW = kRgbTX + KS - 1, SH = kRgbTY + KS - 1, NR = 4 + KS - 1;
  extern __shared__ __align__(16) float s_rgb[];
  float* s_w = s_rgb;                          // [KS*KS*Cin][4]
  float* s_x = s_rgb + KS * KS * Cin * 4;      // [SH][SW][kRgbLD]
  const int K = KS * KS * Cin;
  for (int e = threadIdx.x; e < K * 4; e += blockDim.x) {
    const int k = e >> 2, co = e & 3;
    s_w[e] = co < Cout ? __ldg(Wt + (size_t)co * K + k) : 0.f;
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x0 = blockIdx.x * kRgbTX, y0 = blockIdx.y * kRgbTY, b = blockIdx.z;
  const float* xb = x + (size_t)b * H * W * Cin;
  float acc[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int co = 0; co < 3; ++co) acc[i][co] = co < Cout ? __ldg(bias + co) : 0.f;
  for (int cc = 0; cc < Cin; cc += kRgbCC) {
    __syncthreads();                           // the previous chunk has been consumed (first pass: nothing to wait for but s_w's writers)
    for (int e = threadIdx.x; e < SH * SW * (kRgbCC / 4); e += blockDim.x) {
      const int p = e / (kRgbCC / 4), q = e - p * (kRgbCC / 4);
      const int py = p / SW, px = p - py * SW;
      const int gy = y0 + py - PAD, gx = x0 + px - PAD;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        v = ldg4(xb + ((size_t)gy * W + gx) * Cin + cc + q * 4);
        v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
        v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
      }
      *reinterpret_cast<float4*>(s_x + (size_t)p * kRgbLD + q * 4) = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int kx = 0; kx < KS; ++kx) {
#pragma unroll 1
      for (int c4 = 0; c4 < kRgbCC / 4; ++c4) {
        float4 xv[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) xv[r] = *reinterpret_cast<const float4*>(s_x + (size_t)((ty * 4 + r) * SW + tx + kx) * kRgbLD + c4 * 4);
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          const float4* wk = reinterpret_cast<const float4*>(s_w) + (size_t)((ky * KS + kx) * Cin + cc + c4 * 4);
          const float4 w0 = wk[0], w1 = wk[1], w2 = wk[2], w3 = wk[3];     // 4 input channels x (up to) 4 output channels, broadcast reads
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 v = xv[i + ky];
            acc[i][0] = fmaf(v.x, w0.x, fmaf(v.y, w1.x, fmaf(v.z, w2.x, fmaf(v.w, w3.x, acc[i][0]))));
            acc[i][1] = fmaf(v.x, w0.y, fmaf(v.y, w1.y, fmaf(v.z, w2.y, fmaf(v.w, w3.y, acc[i][1]))));
            acc[i][2] = fmaf(v.x, w0.z, fmaf(v.y, w1.z, fmaf(v.z, w2.z, fmaf(v.w, w3.z, acc[i][2]))));
          }
        }
      }
    }
  }
  const int gx = x0 + tx;
  if (gx >= W) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gy = y0 + ty * 4 + i;
    if (gy >= H) break;
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      if (co >= Cout) break;
      const size_t o = ((size_t)b * Cout + co) * H * W + (size_t)gy * W + gx;
      if (pre) pre[o] = acc[i][co];
      out[o] = tanhf(acc[i][co]);
    }
  }
}

// W [N][K] row-major -> the pre-split, pre-tiled image described at tc::PackedB (one CTA per 32 x 32 unit)
__global__ void __launch_bounds__(256) k_pack_weights(const float* __restrict__ Wm, int N, int K, int kchunks, float* out) {
  const int unit = blockIdx.x, n32 = unit / kchunks, kc = unit - n32 * kchunks;
  const int r = threadIdx.x >> 3, q = threadIdx.x & 7;
  const int n = n32 * 32 + r, k = kc * 32 + q * 4;
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = (n < N && k + e < K) ? __ldg(Wm + (size_t)n * K + k + e) : 0.f;
  float h[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = tc::tf32_hi(v[e]);
  float* u = out + (size_t)unit * tc::PACK_UNIT_FLOATS;
  const uint32_t off = tc::sw128(r, q * 4) / 4;
  *reinterpret_cast<float4*>(u + off) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(u + 1024 + off) = make_float4(v[0] - h[0], v[1] - h[1], v[2] - h[2], v[3] - h[3]);
}

int check_img(int64_t B, int64_t H, int64_t W, int64_t C) {
  SLN_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && C >= 1 && B * H * W < (1ll << 31) && B * H * W * C < (1ll << 40), "image extents out of range");
  return SLN_OK;
}

}  // namespace
}  // namespace sln

using namespace sln;

extern "C" {

size_t sln_packed_weights_bytes(int64_t N, int64_t K) { return (N < 1 || K < 1) ? 0 : tc::packed_weight_floats(N, K) * sizeof(float); }

int sln_pack_weights(const float* Wm, int64_t N, int64_t K, float* out, void* stream) {
  SLN_CHECK_ARG(Wm && out && N >= 1 && K >= 1 && N < (1ll << 24) && K < (1ll << 24) && (uintptr_t)out % 16 == 0, "bad argument");
  const int kchunks = (int)ceil_div64(K, 32);
  const long long units = ceil_div64(N, 128) * 4 * kchunks;
  SLN_CHECK_ARG(units < (1ll << 31), "weight matrix too large");
  ProfScope prof((cudaStream_t)stream, PROF_SPADE_MISC, 12.0 * (double)N * K);
  k_pack_weights<<<(unsigned)units, 256, 0, (cudaStream_t)stream>>>(Wm, (int)N, (int)K, kchunks, out);
  return check_launch("pack_weights");
}

int sln_spade_conv(const float* x, int64_t B, int64_t H, int64_t W, int64_t Cin, int32_t ks, int32_t relu_in, const float* Wp, const float* Wpacked,
                   const float* bias, int64_t Cout, float* out, void* stream) {
  return sln_conv2d_nhwc(x, B, H, W, Cin, ks, relu_in, 0, Wp, Wpacked, bias, Cout, out, stream);
}

}  // extern "C" (templates cannot have C linkage)
template <bool ZP>
static int conv2d_impl(const float* x, int64_t B, int64_t H, int64_t W, int64_t Cin, int32_t ks, int32_t relu_in, const float* Wp,
                       const float* Wpacked, const float* bias, int64_t Cout, float* out, cudaStream_t st) {
  const int M = (int)(B * H * W), N = (int)Cout, K = (int)(ks * ks * Cin);
  Im2colT<ZP> A = make_im2col<ZP>(x, (int)B, (int)H, (int)W, (int)Cin, ks, relu_in);
  MatView Wv = make_view(Wp, K, N, K);
  EpiStore epi; memset(&epi, 0, sizeof(epi));
  epi.C = out; epi.ldc = N; epi.bias = bias;
  if (tc::tc_eligible(M, N, K) && A.vec_ok() && Wv.vec_ok()) {
    tc::TcEpiStore te{out, N, bias, epi.fin};
    if (Wpacked) {   // weights pre-split and pre-tiled once per weight version: their tiles arrive by cp.async.bulk
      tc::PackedB pb{Wpacked, ceil_div(K, tc::BK)};
      SLN_CHECK_ARG(pb.vec_ok(), "packed weights must be 16-byte aligned");
      return tc::launch_tc<true, true>(st, A, pb, te, M, N, K, false, "spade_conv_tc_packed", PROF_SPADE_CONV);
    }
    return tc::launch_tc<true, true>(st, A, Wv, te, M, N, K, false, "spade_conv_tc", PROF_SPADE_CONV);
  }
  return launch_gemm<true, true>(st, A, Wv, epi, M, N, K, false, "spade_conv", PROF_SPADE_CONV);
}

extern "C" {
int sln_conv2d_nhwc(const float* x, int64_t B, int64_t H, int64_t W, int64_t Cin, int32_t ks, int32_t relu_in, int32_t pad_mode, const float* Wp,
                    const float* Wpacked, const float* bias, int64_t Cout, float* out, void* stream) {
  SLN_TRY(check_img(B, H, W, Cin));
  SLN_CHECK_ARG(pad_mode == 0 || pad_mode == 1, "pad_mode: 0 reflection, 1 zeros");
  SLN_CHECK_ARG(x && Wp && out && Cout >= 1, "null pointer");
  SLN_CHECK_ARG(ks == 1 || ks == 3, "kernel size must be 1 or 3");
  SLN_CHECK_ARG(ks == 1 || pad_mode == 1 || (H >= 2 && W >= 2), "reflection padding needs at least 2 rows and columns");
  SLN_CHECK_ARG(H * W * Cin < (1ll << 31), "one image must have fewer than 2^31 elements");
  cudaStream_t st = (cudaStream_t)stream;
  if (pad_mode == 1 && ks == 3) return conv2d_impl<true>(x, B, H, W, Cin, ks, relu_in, Wp, Wpacked, bias, Cout, out, st);
  return conv2d_impl<false>(x, B, H, W, Cin, ks, relu_in, Wp, Wpacked, bias, Cout, out, st);
}

int sln_spade_modulate(const float* actv, int64_t B, int64_t H, int64_t W, int64_t Ca, const float* Wgb, const float* Wgb_packed, const float* bias_g,
                       const float* bias_b, int64_t C, int32_t pair, const float* x, const float* mean, const float* inv, float slope, float* out,
                       void* stream) {
  return sln_spade_modulate_ex(actv, B, H, W, Ca, Wgb, Wgb_packed, bias_g, bias_b, C, pair, x, mean, inv, 1, 0, 0, slope, out, stream);
}

}  // extern "C"
template <bool ZP>
static int modulate_impl(const float* actv, int64_t B, int64_t H, int64_t W, int64_t Ca, const float* Wgb, const float* Wgb_packed,
                         const float* bias_g, const float* bias_b, int64_t C, int32_t pair, const float* x, const float* mean, const float* inv,
                         int32_t stat_stride_b, int32_t stat_stride_c, float slope, float* out, cudaStream_t st) {
  const int M = (int)(B * H * W), N = (int)(2 * C), K = (int)(9 * Ca);
  Im2colT<ZP> A = make_im2col<ZP>(actv, (int)B, (int)H, (int)W, (int)Ca, 3, 1);   // ReLU of mlp_shared applied on load
  MatView Wv = make_view(Wgb, K, N, K);
  const bool tc_ok = (pair == 32 || pair == 64 || pair == 128) && C % 4 == 0 && A.vec_ok() && Wv.vec_ok() && ((uintptr_t)x % 16 == 0) &&
                     ((uintptr_t)out % 16 == 0) && ((uintptr_t)bias_g % 16 == 0) && ((uintptr_t)bias_b % 16 == 0);
  ProfScope prof(st, PROF_SPADE_CONV, 2.0 * (double)M * N * K);
  if (tc_ok) {
    TcEpiSpade te{out, x, (int)C, bias_g, bias_b, mean, inv, (int)(H * W), slope, stat_stride_b, stat_stride_c};
    tc::TcChoice ch{pair, 1, ceil_div(K, tc::BK) * tc::BK};
    int rc;
    if (Wgb_packed) {
      tc::PackedB pb{Wgb_packed, ceil_div(K, tc::BK)};
      SLN_CHECK_ARG(pb.vec_ok(), "packed weights must be 16-byte aligned");
      if (pair == 128) rc = tc::launch_tc_bn<128, true, true>(st, A, pb, te, M, N, K, ch);
      else if (pair == 64) rc = tc::launch_tc_bn<64, true, true>(st, A, pb, te, M, N, K, ch);
      else rc = tc::launch_tc_bn<32, true, true>(st, A, pb, te, M, N, K, ch);
    } else if (pair == 128) rc = tc::launch_tc_bn<128, true, true>(st, A, Wv, te, M, N, K, ch);
    else if (pair == 64) rc = tc::launch_tc_bn<64, true, true>(st, A, Wv, te, M, N, K, ch);
    else rc = tc::launch_tc_bn<32, true, true>(st, A, Wv, te, M, N, K, ch);
    if (rc != SLN_OK) return rc;
    return check_launch("spade_modulate_tc");
  }
  // interleaved weight rows: tile t = [gamma of channels t*half .. | beta of the same]; the direct kernel needs them per channel
  SLN_CHECK_ARG(pair == 2 * C, "the direct modulation fallback expects a single pair tile (pair == 2C)");
  const long long total = (long long)M * C;
  k_modulate_direct<<<(int)ceil_div64(total, 256), 256, 0, st>>>(A, Wgb, Wgb + (size_t)C * K, bias_g, bias_b, (int)C, x, mean, inv, (int)(H * W), slope, out,
                                                                    stat_stride_b, stat_stride_c);
  return check_launch("spade_modulate_direct");
}

extern "C" {
int sln_spade_modulate_ex(const float* actv, int64_t B, int64_t H, int64_t W, int64_t Ca, const float* Wgb, const float* Wgb_packed,
                          const float* bias_g, const float* bias_b, int64_t C, int32_t pair, const float* x, const float* mean, const float* inv,
                          int32_t stat_stride_b, int32_t stat_stride_c, int32_t pad_mode, float slope, float* out, void* stream) {
  SLN_TRY(check_img(B, H, W, Ca));
  SLN_CHECK_ARG((pad_mode == 0 || pad_mode == 1) && (stat_stride_c == 0 || stat_stride_c == 1) && stat_stride_b >= 0, "bad statistics layout / pad mode");
  SLN_CHECK_ARG(stat_stride_c == 0 || (((uintptr_t)mean | (uintptr_t)inv) % 16 == 0 && stat_stride_b % 4 == 0), "per-channel statistics must be 16-byte aligned rows");
  SLN_CHECK_ARG(actv && Wgb && bias_g && bias_b && x && mean && inv && out, "null pointer");
  SLN_CHECK_ARG((pad_mode == 1 || (H >= 2 && W >= 2)) && C >= 1 && pair >= 2 && (2 * C) % pair == 0, "bad modulation shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (pad_mode == 1) return modulate_impl<true>(actv, B, H, W, Ca, Wgb, Wgb_packed, bias_g, bias_b, C, pair, x, mean, inv, stat_stride_b, stat_stride_c, slope, out, st);
  return modulate_impl<false>(actv, B, H, W, Ca, Wgb, Wgb_packed, bias_g, bias_b, C, pair, x, mean, inv, stat_stride_b, stat_stride_c, slope, out, st);
}

int sln_spade_ln_stats(const float* x, int64_t B, int64_t n_per_sample, float eps, void* scratch, float* mean, float* inv, void* stream) {
  SLN_CHECK_ARG(x && scratch && mean && inv && B >= 1 && B <= 65535 && n_per_sample >= 1, "bad argument");
  SLN_CHECK_ARG((uintptr_t)x % 16 == 0 && (uintptr_t)scratch % 8 == 0 && n_per_sample % 4 == 0, "LayerNorm statistics need 16-byte aligned samples");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(st, PROF_SPADE_MISC, 4.0 * (double)B * n_per_sample);
  SLN_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * B, st));
  int chunks = (int)max((int64_t)1, min(ceil_div64(n_per_sample / 4, 256 * 8), (int64_t)(4 * kNumSMs) / B + 1));
  k_ln_partial<<<dim3(chunks, (unsigned)B), 256, 0, st>>>(x, n_per_sample, (double*)scratch);
  SLN_TRY(check_launch("ln_partial"));
  k_ln_final<<<ceil_div((int)B, 128), 128, 0, st>>>((const double*)scratch, (int)B, n_per_sample, eps, mean, inv);
  return check_launch("ln_final");
}

int sln_instnorm_stats(const float* x, int64_t B, int64_t HW, int64_t C, float eps, float* mean, float* inv, void* stream) {
  SLN_CHECK_ARG(x && mean && inv && B >= 1 && B <= 65535 && HW >= 1 && C >= 1 && HW < (1ll << 31) && C < (1ll << 24), "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(st, PROF_SPADE_MISC, 4.0 * (double)B * HW * C);
  k_in_stats<<<dim3((unsigned)ceil_div64(C, 32), (unsigned)B), 256, 0, st>>>(x, (int)HW, (int)C, eps, mean, inv);
  return check_launch("instnorm_stats");
}

int sln_norm_act(const float* x, int64_t B, int64_t HW, int64_t C, const float* mean, const float* inv, int32_t stat_stride_b, int32_t stat_stride_c,
                 int32_t act, float* out, void* stream) {
  SLN_CHECK_ARG(x && mean && inv && out && B >= 1 && HW >= 1 && C >= 4 && C % 4 == 0 && (act == 0 || act == 1), "bad argument (C must be a multiple of 4)");
  SLN_CHECK_ARG(((uintptr_t)x | (uintptr_t)out) % 16 == 0 && HW * C < (1ll << 31), "tensors must be 16-byte aligned");
  SLN_CHECK_ARG(stat_stride_c == 0 || (stat_stride_c == 1 && ((uintptr_t)mean | (uintptr_t)inv) % 16 == 0 && stat_stride_b % 4 == 0), "bad statistics layout");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n4 = (long long)B * HW * C / 4;
  ProfScope prof(st, PROF_SPADE_MISC, 8.0 * (double)n4 * 4);
  k_norm_act<<<(unsigned)ceil_div64(n4, 256), 256, 0, st>>>(x, n4, (int)(HW * C), (int)C, mean, inv, stat_stride_b, stat_stride_c, act, out);
  return check_launch("norm_act");
}

int sln_seg_resize_nhwc(const float* seg, int64_t B, int32_t nc, int32_t S, int32_t nearest, int64_t h, int64_t w, int32_t cpad, float* out, void* stream) {
  SLN_CHECK_ARG(seg && out && B >= 1 && nc >= 1 && S >= 1 && h >= 1 && w >= 1 && cpad >= nc && cpad % 4 == 0 && (nearest == 0 || nearest == 1), "bad argument");
  SLN_CHECK_ARG((uintptr_t)out % 16 == 0, "output must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * h * w * (cpad / 4);
  ProfScope prof(st, PROF_SPADE_MISC, 4.0 * (double)n * 4);
  k_seg_resize<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(seg, (int)B, nc, S, (int)h, (int)w, cpad, nearest, out);
  return check_launch("seg_resize");
}

int sln_spade_seg_features(const float* seg, int64_t B, int32_t nc, int32_t S, int32_t mode, int64_t h, int64_t w, const float* dw, const float* db,
                           int32_t nd, float* out, void* stream) {
  SLN_TRY(check_img(B, h, w, nd + nc - 1));
  SLN_CHECK_ARG(seg && dw && db && out && nc >= 2 && S >= 1 && (mode == 0 || mode == 1) && h >= 2 && w >= 2 && nd >= 1, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int P = (int)(B * h * w);
  ProfScope prof(st, PROF_SPADE_MISC, 4.0 * P * (double)(nd + nc - 1));
  k_seg_features<<<ceil_div(P, kSegPix), kSegPix * kSegGroups, (size_t)kSegPix * ((nd + nc - 1) | 1) * sizeof(float), st>>>(seg, (int)B, nc, S, mode, (int)h, (int)w, dw, db, nd, out);
  return check_launch("seg_features");
}

int sln_spade_upsample2x(const float* x, int64_t B, int64_t H, int64_t W, int64_t C, int32_t bilinear, float* out, void* stream) {
  SLN_TRY(check_img(B, 2 * H, 2 * W, C));
  SLN_CHECK_ARG(x && out && C % 4 == 0, "upsample2x needs C % 4 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)B * 4 * H * W * (C / 4);
  ProfScope prof(st, PROF_SPADE_MISC, 20.0 * (double)total);
  const long long rowlen = 2 * W * (C / 4);
  SLN_CHECK_ARG(rowlen <= 65535ll * 256, "upsample2x row too long");
  k_upsample2x<<<dim3((unsigned)(B * 2 * H), (unsigned)ceil_div64(rowlen, 256)), 256, 0, st>>>(x, (int)B, (int)H, (int)W, (int)C, bilinear, out);
  return check_launch("upsample2x");
}

int sln_spade_se_residual(const float* dx, const float* xs, int64_t B, int64_t H, int64_t W, int64_t C, const float* W1, const float* W2, int32_t Ch,
                          void* scratch, size_t scratch_bytes, float* out, void* stream) {
  SLN_TRY(check_img(B, H, W, C));
  SLN_CHECK_ARG(dx && xs && W1 && W2 && scratch && out && C % 4 == 0 && Ch >= 1, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = (int)(H * W);
  int nch = min(HW, 64);
  int rows_per = ceil_div(HW, nch);
  nch = ceil_div(HW, rows_per);
  const size_t need = ((size_t)B * nch * C + (size_t)B * C) * sizeof(float);
  SLN_CHECK_ARG(scratch_bytes >= need && (uintptr_t)scratch % 16 == 0, "SE scratch too small (%zu < %zu bytes) or misaligned", scratch_bytes, need);
  float* partial = (float*)scratch;
  float* svec = partial + (size_t)B * nch * C;
  ProfScope prof(st, PROF_SPADE_MISC, 16.0 * (double)B * HW * C);
  k_se_pool<<<dim3(nch, (unsigned)B), 256, 0, st>>>(dx, HW, (int)C, rows_per, partial);
  SLN_TRY(check_launch("se_pool"));
  k_se_fc<<<(unsigned)B, 1024, (size_t)(C + Ch) * sizeof(float), st>>>(partial, nch, HW, (int)C, W1, W2, Ch, svec);
  SLN_TRY(check_launch("se_fc"));
  const long long n4 = (long long)B * HW * (C / 4);
  k_se_apply<<<(unsigned)ceil_div64(n4, 256), 256, 0, st>>>(dx, xs, svec, n4, HW * (int)(C / 4), (int)(C / 4), out);
  return check_launch("se_apply");
}

int sln_spade_to_rgb(const float* x, int64_t B, int64_t H, int64_t W, int64_t Cin, const float* Wt, const float* bias, int32_t Cout, int32_t ks,
                     float slope, float* pre, float* out, void* stream) {
  SLN_TRY(check_img(B, H, W, Cin));
  SLN_CHECK_ARG(x && Wt && bias && out && Cout >= 1 && Cout <= 4 && (ks & 1) && Cin % 4 == 0, "to_rgb supports <= 4 output channels, odd kernels, Cin % 4 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)ks * ks * Cin * 4 * sizeof(float);
  SLN_CHECK_ARG(smem <= 200 * 1024, "to_rgb weights do not fit shared memory");
  const int P = (int)(B * H * W);
  ProfScope prof(st, PROF_SPADE_CONV, 2.0 * (double)P * ks * ks * Cin * Cout);
  const size_t smem_t = smem + (size_t)(kRgbTX + 4) * (kRgbTY + 4) * kRgbLD * sizeof(float);
  if (ks == 5 && Cin % kRgbCC == 0 && Cout <= 3 && smem_t <= 110 * 1024 && B <= 65535) {
    static unsigned long long configured_t = 0ull;
    if (first_use_on_device(configured_t)) {
      SLN_CUDA_TRY(cudaFuncSetAttribute(k_to_rgb_tiled<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    }
    k_to_rgb_tiled<5><<<dim3(ceil_div((int)W, kRgbTX), ceil_div((int)H, kRgbTY), (unsigned)B), 128, smem_t, st>>>(x, (int)B, (int)H, (int)W, (int)Cin, Wt, bias,
                                                                                                            Cout, slope, pre, out);
    return check_launch("to_rgb");
  }
  static unsigned long long configured = 0ull;   // devices configured (per call site / instantiation)
  if (first_use_on_device(configured)) {
    SLN_CUDA_TRY(cudaFuncSetAttribute(k_to_rgb<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  k_to_rgb<4><<<ceil_div(P, 128), 128, smem, st>>>(x, (int)B, (int)H, (int)W, (int)Cin, Wt, bias, Cout, ks, slope, pre, out);
  return check_launch("to_rgb");
}

#ifdef SLN_TC_TRACE
// tuning build only (tools/build_trace.sh spade): the phase stamps of CTA (0,0,0) of this translation unit's last contraction launch
int sln_debug_tc_trace_spade(long long* out48) {
  SLN_CUDA_TRY(cudaDeviceSynchronize());
  SLN_CUDA_TRY(cudaMemcpyFromSymbol(out48, tc::g_tc_trace, sizeof(long long) * 48));
  return SLN_OK;
}
#endif

}  // extern "C"
