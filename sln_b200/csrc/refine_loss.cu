// Fused multi-scale refinement loss of the layout-refinement loop.
// Reference: testing/test_render_refine.py:332-352 (null-fill of the last depth plane, PSP pyramids of the 29 depth planes and
// the 40 class planes, L1 * 0.5 * 100 + sum_scales CrossEntropy(ignore -100) / 800 * 100) with PSP_pool_new (:192-215):
// every plane is resized 256 -> s (bilinear, align_corners=True) -> 96 (bilinear, align_corners=False) for s in sizes.
// The reference (and the torch restatement in models/refine.py) spends ~150 small kernels per iteration on this; here the
// forward AND the gradient w.r.t. the rendered image are seven launches, all gather-style (no atomics: deterministic):
//   k_null_fill    null mask + filled last plane                                   (:333)
//   k_psp_down     256x256 -> s x s for every (plane, scale)                        (nn.Upsample(size=s, align_corners=True))
//   k_psp_loss     s x s -> 96x96 on the fly; |.| and log-softmax/NLL per pixel; per-CTA loss partials; d(loss)/d(upsampled)
//   k_psp_tables   per source coordinate: which destinations tap it, with which weight (built from the forward taps)
//   k_psp_up_bwd   transpose of the upsample: d(upsampled) -> d(s x s)
//   k_psp_down_bwd transpose of the downsample over all scales -> d(image), null-filled pixels of the last plane get 0
// plus k_refine_loss_final (fixed-order sum of the partials -> the scalar loss).
#include "../../include/sln_b200.h"
#include "common.cuh"

namespace sln {
namespace {

constexpr int kScales = 4;
constexpr int kMaxSem = 64;

struct PspGeom {
  int S;                 // input image size (256)
  int top;               // output size of every pyramid level (sizes[last] = 96)
  int size[kScales];     // 32, 48, 64, 96
  int off[kScales + 1];  // offsets of the levels inside one plane of the intermediate buffer (prefix sums of size^2)
  int n_sem, n_dep;      // 40 class planes (image channels 1..n_sem), 29 depth planes (channels 1+n_sem ..)
};

struct Tap { int i0, i1; float w0, w1; };
// torch's upsample_bilinear2d source index + lambdas (aten/native/UpSample.h: area_pixel_compute_scale / _source_index)
__device__ __forceinline__ Tap tap_corners(int o, int in, int out) {            // align_corners = True
  const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  const float src = scale * (float)o;
  Tap t; t.i0 = (int)src; t.i1 = t.i0 + (t.i0 < in - 1 ? 1 : 0); t.w1 = src - (float)t.i0; t.w0 = 1.f - t.w1;
  return t;
}
__device__ __forceinline__ Tap tap_centers(int o, int in, int out) {            // align_corners = False
  const float scale = (float)in / (float)out;
  float src = scale * ((float)o + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  Tap t; t.i0 = (int)src; t.i1 = t.i0 + (t.i0 < in - 1 ? 1 : 0); t.w1 = src - (float)t.i0; t.w0 = 1.f - t.w1;
  return t;
}
__device__ __forceinline__ float bilerp(const float* __restrict__ p, int ld, const Tap& ty, const Tap& tx) {
  return ty.w0 * (tx.w0 * __ldg(p + ty.i0 * ld + tx.i0) + tx.w1 * __ldg(p + ty.i0 * ld + tx.i1)) +
         ty.w1 * (tx.w0 * __ldg(p + ty.i1 * ld + tx.i0) + tx.w1 * __ldg(p + ty.i1 * ld + tx.i1));
}

// image [1 + n_sem + n_dep, S, S].  filled[S*S] = last plane with 1.0 where the depth planes sum to < 0.5; mask = that predicate.
__global__ void __launch_bounds__(256) k_null_fill(const float* __restrict__ image, PspGeom g, float* __restrict__ filled,
                                                   unsigned char* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int P = g.S * g.S;
  if (i >= P) return;
  const float* d = image + (size_t)(1 + g.n_sem) * P + i;
  float s = 0.f;
  for (int c = 0; c < g.n_dep; ++c) s += __ldg(d + (size_t)c * P);
  const bool null = s < 0.5f;
  filled[i] = null ? 1.f : __ldg(d + (size_t)(g.n_dep - 1) * P);
  mask[i] = null ? 1 : 0;
}

// plane p in [0, n_sem + n_dep) = image channel 1 + p (the last one read from `filled`); inter [planes][off[4]]
__global__ void __launch_bounds__(256) k_psp_down(const float* __restrict__ image, const float* __restrict__ filled, PspGeom g,
                                                  float* __restrict__ inter) {
  const int planes = g.n_sem + g.n_dep;
  const int p = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.off[kScales]) return;
  int s = 0;
  while (i >= g.off[s + 1]) ++s;
  const int n = g.size[s], r = i - g.off[s];
  const float* src = p == planes - 1 ? filled : image + (size_t)(1 + p) * g.S * g.S;
  inter[(size_t)p * g.off[kScales] + i] = bilerp(src, g.S, tap_corners(r / n, g.S, n), tap_corners(r % n, g.S, n));
}

// One thread per output pixel of one pyramid level.  grid = (ceil(top^2 / 256), kScales).
// partial[(s * gridDim.x + blockIdx.x) * 2 + {0,1}] = sum |depth - target|, sum NLL of this CTA.
// d_up [kScales][planes][top^2] = d(loss)/d(upsampled plane) (only when d_up != nullptr).
__global__ void __launch_bounds__(256) k_psp_loss(const float* __restrict__ inter, PspGeom g, const float* __restrict__ t_depth,
                                                  const int64_t* __restrict__ t_labels, float g_depth, float4 g_sem4,
                                                  float* __restrict__ partial, float* __restrict__ d_up) {
  const int s = blockIdx.y, n = g.size[s], T2 = g.top * g.top;
  const int planes = g.n_sem + g.n_dep, stride = g.off[kScales];
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const float g_sem = s == 0 ? g_sem4.x : s == 1 ? g_sem4.y : s == 2 ? g_sem4.z : g_sem4.w;
  float l1 = 0.f, nll = 0.f;
  if (o < T2) {
    const Tap ty = tap_centers(o / g.top, n, g.top), tx = tap_centers(o % g.top, n, g.top);
    const float* base = inter + g.off[s];
    float* du = d_up ? d_up + (size_t)s * planes * T2 + o : nullptr;
    // ---- class planes: log-softmax over n_sem logits, NLL against the target label (ignore_index -100)
    float logit[kMaxSem];
    float m = -INFINITY;
#pragma unroll 8
    for (int c = 0; c < g.n_sem; ++c) { logit[c] = bilerp(base + (size_t)c * stride, n, ty, tx); m = fmaxf(m, logit[c]); }
    float se = 0.f;
    for (int c = 0; c < g.n_sem; ++c) se += expf(logit[c] - m);
    const float lse = m + logf(se);
    const int64_t lab = __ldg(t_labels + (size_t)s * T2 + o);
    const bool live = lab >= 0 && lab < g.n_sem;
    if (live) nll = lse - logit[(int)lab];
    if (du)
      for (int c = 0; c < g.n_sem; ++c) du[(size_t)c * T2] = live ? g_sem * (expf(logit[c] - lse) - (c == (int)lab ? 1.f : 0.f)) : 0.f;
    // ---- depth planes: L1 against the pooled target (channel s * n_dep + c of the concatenated pyramid)
    for (int c = 0; c < g.n_dep; ++c) {
      const float v = bilerp(base + (size_t)(g.n_sem + c) * stride, n, ty, tx);
      const float diff = v - __ldg(t_depth + ((size_t)s * g.n_dep + c) * T2 + o);
      l1 += fabsf(diff);
      if (du) du[(size_t)(g.n_sem + c) * T2] = diff > 0.f ? g_depth : diff < 0.f ? -g_depth : 0.f;
    }
  }
  // deterministic CTA reduction: warp shuffles, then warp 0 over the 8 warp sums
  __shared__ float sh[2][8];
  l1 = warp_sum(l1); nll = warp_sum(nll);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = l1; sh[1][threadIdx.x >> 5] = nll; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
    partial[(s * gridDim.x + blockIdx.x) * 2] = a;
    partial[(s * gridDim.x + blockIdx.x) * 2 + 1] = b;
  }
}

// loss = 100 * 0.5 * mean|depth diff| + 100 * sum_s (sum NLL_s / count_s) / 800   (test_render_refine.py:347-349)
__global__ void k_refine_loss_final(const float* __restrict__ partial, int ctas_per_scale, float inv_n_depth, float4 inv_count,
                                    float* __restrict__ loss3) {
  // one warp per level; lanes stride the level's CTA partials, fixed shuffle tree: deterministic
  __shared__ float sh[kScales][2];
  const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float d = 0.f, n = 0.f;
  for (int i = lane; i < ctas_per_scale; i += 32) { d += partial[(s * ctas_per_scale + i) * 2]; n += partial[(s * ctas_per_scale + i) * 2 + 1]; }
  d = warp_sum(d); n = warp_sum(n);
  if (lane == 0) { sh[s][0] = d; sh[s][1] = n; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  float dsum = 0.f, sem = 0.f;
  for (int k = 0; k < kScales; ++k) {
    dsum += sh[k][0];
    sem += sh[k][1] * (k == 0 ? inv_count.x : k == 1 ? inv_count.y : k == 2 ? inv_count.z : inv_count.w) / 800.f;
  }
  const float depth = dsum * inv_n_depth * 0.5f;
  loss3[0] = depth * 100.f + sem * 100.f;
  loss3[1] = depth;
  loss3[2] = sem;
}

// ---- transposed resampling without atomics: per (direction, level, source coordinate) the list of destination coordinates
// that tap it, with the (merged) weight, built by scanning the forward taps — so the transposes use exactly the forward
// lambdas.  dir 0: upsample n -> top (source coordinate < n, destinations < top); dir 1: downsample S -> n.
constexpr int kMaxTaps = 8;
constexpr int kMaxCoord = 4096;
struct TapList { int n; int o[kMaxTaps]; float w[kMaxTaps]; };

// grid = (kScales, 2), one WARP per source coordinate: the lanes scan the destinations 32 at a time, an ordered ballot compaction
// appends the ones that tap it
__global__ void __launch_bounds__(256) k_psp_tables(PspGeom g, TapList* __restrict__ tables, int stride, int* __restrict__ overflow) {
  const int s = blockIdx.x, dir = blockIdx.y, n = g.size[s];
  const int n_src = dir == 0 ? n : g.S, n_dst = dir == 0 ? g.top : n;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x >> 5; i < n_src; i += blockDim.x >> 5) {
    TapList* t = tables + (size_t)(dir * kScales + s) * stride + i;
    int cnt = 0;
    for (int o0 = 0; o0 < n_dst; o0 += 32) {
      const int o = o0 + lane;
      bool hit = false; float w = 0.f;
      if (o < n_dst) {
        const Tap tp = dir == 0 ? tap_centers(o, n, g.top) : tap_corners(o, g.S, n);
        hit = tp.i0 == i || tp.i1 == i;
        w = (tp.i0 == i ? tp.w0 : 0.f) + (tp.i1 == i ? tp.w1 : 0.f);
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      const int pos = cnt + __popc(m & ((1u << lane) - 1u));
      if (hit) { if (pos < kMaxTaps) { t->o[pos] = o; t->w[pos] = w; } else atomicExch(overflow, 1); }
      cnt += __popc(m);
    }
    if (lane == 0) t->n = min(cnt, kMaxTaps);
  }
}

// d_inter[p][off[s] + iy*n + ix] = sum over the output pixels that tap (iy, ix) of weight * d_up.  grid = (ceil(off[4]/256), planes)
__global__ void __launch_bounds__(256) k_psp_up_bwd(const float* __restrict__ d_up, PspGeom g, const TapList* __restrict__ tables, int stride,
                                                    float* __restrict__ d_inter) {
  const int planes = g.n_sem + g.n_dep, T2 = g.top * g.top;
  const int p = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.off[kScales]) return;
  int s = 0;
  while (i >= g.off[s + 1]) ++s;
  const int n = g.size[s], r = i - g.off[s];
  const TapList* ty = tables + (size_t)s * stride + r / n;
  const TapList* tx = tables + (size_t)s * stride + r % n;
  const float* du = d_up + ((size_t)s * planes + p) * T2;
  const int ny = ty->n, nx = tx->n;
  float acc = 0.f;
  for (int a = 0; a < ny; ++a) {
    const float* row_p = du + ty->o[a] * g.top;
    float row = 0.f;
    for (int b = 0; b < nx; ++b) row += tx->w[b] * __ldg(row_p + tx->o[b]);
    acc += ty->w[a] * row;
  }
  d_inter[(size_t)p * g.off[kScales] + i] = acc;
}

// d_image[ch][y][x]: transpose of the four downsamples.  One thread per PIXEL: the (level, row, column) taps of the pixel are the
// same for every plane, so they are gathered once (<= 3 x 3 per level in practice, kMaxDownTaps) and reused for all planes.
constexpr int kMaxDownTaps = 36;
__global__ void __launch_bounds__(256) k_psp_down_bwd(const float* __restrict__ d_inter, PspGeom g, const TapList* __restrict__ tables, int stride,
                                                      const unsigned char* __restrict__ mask, float* __restrict__ d_image) {
  const int planes = g.n_sem + g.n_dep, P = g.S * g.S;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int y = i / g.S, x = i % g.S;
  int t_off[kMaxDownTaps]; float t_w[kMaxDownTaps];
  int nt = 0;
  for (int s = 0; s < kScales; ++s) {
    const int n = g.size[s];
    const TapList* ty = tables + (size_t)(kScales + s) * stride + y;
    const TapList* tx = tables + (size_t)(kScales + s) * stride + x;
    const int ny = ty->n, nx = tx->n;
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b)
        if (nt < kMaxDownTaps) { t_off[nt] = g.off[s] + ty->o[a] * n + tx->o[b]; t_w[nt] = ty->w[a] * tx->w[b]; ++nt; }
  }
  const bool null = mask[i] != 0;
  if (blockIdx.y == 0) d_image[i] = 0.f;                              // channel 0 (the plain depth render) is not part of the loss
  const int per = (planes + gridDim.y - 1) / gridDim.y;               // the planes are split over grid.y for parallelism
  const int p_lo = blockIdx.y * per, p_hi = min(planes, p_lo + per);
  for (int p = p_lo; p < p_hi; ++p) {
    const float* di = d_inter + (size_t)p * g.off[kScales];
    float acc = 0.f;
    for (int k = 0; k < nt; ++k) acc += t_w[k] * __ldg(di + t_off[k]);
    if (p == planes - 1 && null) acc = 0.f;                           // null-filled pixels of the last plane are constants (:333)
    d_image[(size_t)(1 + p) * P + i] = acc;
  }
}

struct Ws { float* filled; unsigned char* mask; float* inter; float* d_up; float* d_inter; float* partial; TapList* tables; int stride; int* overflow; size_t total; };
bool make_geom(int S, const int32_t* sizes, int n_sem, int n_dep, PspGeom* g) {
  g->S = S; g->n_sem = n_sem; g->n_dep = n_dep; g->off[0] = 0;
  for (int s = 0; s < kScales; ++s) {
    if (sizes[s] < 1 || sizes[s] > 4096) return false;
    g->size[s] = sizes[s]; g->off[s + 1] = g->off[s] + sizes[s] * sizes[s];
  }
  g->top = sizes[kScales - 1];
  return true;
}
Ws carve(void* ws, const PspGeom& g) {
  Ws w; size_t at = 0;
  auto take = [&](size_t bytes) { size_t r = at; at += (bytes + 255) / 256 * 256; return r; };
  const size_t P = (size_t)g.S * g.S, planes = g.n_sem + g.n_dep, T2 = (size_t)g.top * g.top;
  char* b = (char*)ws;
  w.filled = (float*)(b + take(P * 4));
  w.mask = (unsigned char*)(b + take(P));
  w.inter = (float*)(b + take(planes * g.off[kScales] * 4));
  w.d_up = (float*)(b + take(kScales * planes * T2 * 4));
  w.d_inter = (float*)(b + take(planes * g.off[kScales] * 4));
  w.partial = (float*)(b + take(kScales * ((T2 + 255) / 256) * 2 * 4));
  w.stride = g.S > g.top ? g.S : g.top;
  w.tables = (TapList*)(b + take((size_t)2 * kScales * w.stride * sizeof(TapList)));
  w.overflow = (int*)(b + take(256));
  w.total = at;
  return w;
}

}  // namespace
}  // namespace sln

using namespace sln;

extern "C" {

size_t sln_refine_loss_workspace_bytes(int32_t image_size, const int32_t* sizes4, int32_t n_sem, int32_t n_dep) {
  PspGeom g;
  if (!sizes4 || image_size < 2 || n_sem < 1 || n_dep < 1 || !make_geom(image_size, sizes4, n_sem, n_dep, &g)) return 0;
  return carve(nullptr, g).total;
}

int sln_refine_loss(const float* image, int32_t image_size, const int32_t* sizes4, int32_t n_sem, int32_t n_dep, const float* t_depth,
                    const int64_t* t_labels, const float* label_counts4, float* loss3, float* d_image, void* ws, size_t ws_bytes,
                    void* stream) {
  SLN_CHECK_ARG(image && sizes4 && t_depth && t_labels && label_counts4 && loss3 && ws, "refine_loss: null pointer");
  SLN_CHECK_ARG(image_size >= 2 && n_sem >= 1 && n_sem <= kMaxSem && n_dep >= 1, "refine_loss: bad extents (n_sem <= %d)", kMaxSem);
  PspGeom g;
  SLN_CHECK_ARG(make_geom(image_size, sizes4, n_sem, n_dep, &g), "refine_loss: bad pyramid sizes");
  for (int i = 0; i < kScales; ++i) {
    // a source coordinate is tapped by <= 2*ceil(ratio)+1 destinations; the lists hold kMaxTaps
    const int up = (g.top + g.size[i] - 1) / g.size[i], down = (g.size[i] + g.S - 1) / g.S;
    SLN_CHECK_ARG(g.size[i] <= g.S, "refine_loss: pyramid level %d (%d) is larger than the image", i, g.size[i]);   // <= 3 x 3 taps per pixel and level
    SLN_CHECK_ARG(2 * up + 1 <= kMaxTaps && 2 * down + 1 <= kMaxTaps && g.size[i] <= g.top && g.S <= kMaxCoord,
                  "refine_loss: pyramid level %d (%d) needs more than %d taps per coordinate", i, g.size[i], kMaxTaps);
  }
  Ws w = carve(ws, g);
  if (ws_bytes < w.total) { set_error("refine_loss: workspace too small (%zu < %zu)", ws_bytes, w.total); return SLN_EWORKSPACE; }
  SLN_CHECK_ARG((uintptr_t)ws % 16 == 0, "refine_loss: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int P = g.S * g.S, planes = n_sem + n_dep, T2 = g.top * g.top, ctas = (T2 + 255) / 256;
  ProfScope prof(st, PROF_MISC, 4.0 * (double)(1 + planes) * P * (d_image ? 2 : 1));
  k_null_fill<<<(P + 255) / 256, 256, 0, st>>>(image, g, w.filled, w.mask);
  SLN_TRY(check_launch("null_fill"));
  k_psp_down<<<dim3((g.off[kScales] + 255) / 256, planes), 256, 0, st>>>(image, w.filled, g, w.inter);
  SLN_TRY(check_launch("psp_down"));
  // d(loss)/d(pooled depth) = 100 * 0.5 / numel ; d(loss)/d(NLL_s) = 100 / 800 / count_s  (count 0 -> the reference's 0/0 = NaN)
  const float inv_nd = 1.f / ((float)kScales * (float)n_dep * (float)T2);
  float4 inv_cnt = make_float4(1.f / label_counts4[0], 1.f / label_counts4[1], 1.f / label_counts4[2], 1.f / label_counts4[3]);
  float4 g_sem = make_float4(inv_cnt.x * 0.125f, inv_cnt.y * 0.125f, inv_cnt.z * 0.125f, inv_cnt.w * 0.125f);
  k_psp_loss<<<dim3(ctas, kScales), 256, 0, st>>>(w.inter, g, t_depth, t_labels, 50.f * inv_nd, g_sem, w.partial, d_image ? w.d_up : nullptr);
  SLN_TRY(check_launch("psp_loss"));
  k_refine_loss_final<<<1, 32 * kScales, 0, st>>>(w.partial, ctas, inv_nd, inv_cnt, loss3);
  SLN_TRY(check_launch("refine_loss_final"));
  if (d_image) {
    k_psp_tables<<<dim3(kScales, 2), 256, 0, st>>>(g, w.tables, w.stride, w.overflow);
    SLN_TRY(check_launch("psp_tables"));
    k_psp_up_bwd<<<dim3((g.off[kScales] + 255) / 256, planes), 256, 0, st>>>(w.d_up, g, w.tables, w.stride, w.d_inter);
    SLN_TRY(check_launch("psp_up_bwd"));
    k_psp_down_bwd<<<dim3((P + 255) / 256, 8), 256, 0, st>>>(w.d_inter, g, w.tables, w.stride, w.mask, d_image);
    SLN_TRY(check_launch("psp_down_bwd"));
  }
  return SLN_OK;
}

}  // extern "C"
