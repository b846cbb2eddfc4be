// FP32 SIMT tiled contraction with fused operand transforms and fused epilogues.
//
// The graph-conv MLPs of 3D_SLN (reference models/graph.py:10-27,84,109) are small dense contractions
// (M = #triples or #nodes, N,K <= 640) separated by training-mode BatchNorm.  Everything that the reference
// does as separate aten ops around each Linear is folded into the operand loaders and epilogues here:
//   * loaders apply the previous layer's BatchNorm+ReLU lazily (x = relu(y*scale+shift)) and the
//     [obj[s] | pred | obj[o]] gather+concat of graph.py:78-83, so neither is ever materialised;
//   * the forward epilogue adds the bias, writes the pre-BN activation and reduces the BN batch statistics
//     (deterministic per-row-tile partials; the last CTA finalises mean/rstd/scale/shift + running stats);
//   * the backward-data epilogue applies the ReLU mask and reduces the BN-backward column sums;
//   * the backward-weight contraction is split along the sample dimension and reduced with RED.ADD.F32.
// All arithmetic is fp32 FMA (the parity contract is 1e-4 against an fp32 reference).
#pragma once
#include "common.cuh"

namespace sln {

// ---------------------------------------------------------------- operand functors
// Row-major matrix with an optional lazy per-column affine + ReLU: v(r,c) = act(p[r*ld+c]*scale[c]+shift[c]).
struct MatView {
  static constexpr bool kTwoLoads = false;   // fetch4 fills only `a` (tc_gemm.cuh: prefetch depth)
  const float* p;
  int ld, rows, cols;
  const float* scale;  // null -> identity
  const float* shift;
  int relu;
  int vec;  // rows/base 16B aligned, cols % 4 == 0, scale/shift 16B aligned

  __device__ __forceinline__ float at(int r, int c) const {
    float v = __ldg(p + (size_t)r * ld + c);
    if (scale) v = fmaf(v, __ldg(scale + c), __ldg(shift + c));
    if (relu) v = fmaxf(v, 0.f);
    return v;
  }
  // bounds-checked single element in storage coordinates
  __device__ __forceinline__ float at_t(int r, int c) const { return (r < rows && c < cols) ? at(r, c) : 0.f; }
  // Two-phase access used by the tensor-core loader (requires vec): fetch4 issues the raw 16-byte load, finish4 applies
  // the lazy transform.  Neither checks bounds: token() clamps the row and the caller clamps the column with clampc(), so
  // padded tile rows/columns read valid (finite) memory whose products are never stored; the caller zeroes K padding.
  struct Tok { const float* rp; };
  __device__ __forceinline__ Tok token(int r) const { return Tok{p + (size_t)min(r, rows - 1) * ld}; }
  __device__ __forceinline__ int clampc(int c) const { return min(c, cols - 4); }
  __device__ __forceinline__ void fetch4(const Tok& t, int c, float4& a, float4& b) const { a = ldg4(t.rp + c); }
  __device__ __forceinline__ float4 finish4(const Tok& t, int c, float4 v, float4) const {
    if (scale) {
      float4 s = ldg4(scale + c), h = ldg4(shift + c);
      v.x = fmaf(v.x, s.x, h.x); v.y = fmaf(v.y, s.y, h.y); v.z = fmaf(v.z, s.z, h.z); v.w = fmaf(v.w, s.w, h.w);
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    return v;
  }
  bool vec_ok() const { return vec != 0; }
  // 4 consecutive columns starting at c (c % 4 == 0); zero outside the matrix.
  __device__ __forceinline__ float4 ld4(int r, int c) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= rows || c >= cols) return v;
    if (vec) {
      v = ldg4(p + (size_t)r * ld + c);
      if (scale) {
        float4 s = ldg4(scale + c), t = ldg4(shift + c);
        v.x = fmaf(v.x, s.x, t.x); v.y = fmaf(v.y, s.y, t.y); v.z = fmaf(v.z, s.z, t.z); v.w = fmaf(v.w, s.w, t.w);
      }
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    } else {
      v.x = at(r, c);
      if (c + 1 < cols) v.y = at(r, c + 1);
      if (c + 2 < cols) v.z = at(r, c + 2);
      if (c + 3 < cols) v.w = at(r, c + 3);
    }
    return v;
  }
};

inline MatView make_view(const float* p, int ld, int rows, int cols, const float* scale = nullptr,
                         const float* shift = nullptr, int relu = 0) {
  MatView v;
  v.p = p; v.ld = ld; v.rows = rows; v.cols = cols; v.scale = scale; v.shift = shift; v.relu = relu;
  bool ok = ((uintptr_t)p % 16 == 0) && (ld % 4 == 0) && (cols % 4 == 0);
  if (scale) ok = ok && ((uintptr_t)scale % 16 == 0) && ((uintptr_t)shift % 16 == 0);
  v.vec = ok ? 1 : 0;
  return v;
}

// Virtual [T, 3D] matrix  [ obj[s_t] | pred[t] | obj[o_t] ]   (reference graph.py:78-83), D % 4 == 0.
struct GatherCat {
  static constexpr bool kTwoLoads = false;
  MatView obj, pred;
  const int* s_idx;
  const int* o_idx;
  int D, rows, cols;
  __device__ __forceinline__ float4 ld4(int t, int c) const {
    if (t >= rows || c >= cols) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) return obj.ld4(__ldg(s_idx + t), c);
    if (c < 2 * D) return pred.ld4(t, c - D);
    return obj.ld4(__ldg(o_idx + t), c - 2 * D);
  }
  __device__ __forceinline__ float at_t(int t, int c) const {
    if (t >= rows || c >= cols) return 0.f;
    if (c < D) return obj.at(__ldg(s_idx + t), c);
    if (c < 2 * D) return pred.at(t, c - D);
    return obj.at(__ldg(o_idx + t), c - 2 * D);
  }
  struct Tok { const float* sp; const float* pp; const float* op; };   // the three source rows of a triple, resolved once per row
  __device__ __forceinline__ Tok token(int t) const {
    t = min(t, rows - 1);
    return Tok{obj.p + (size_t)__ldg(s_idx + t) * obj.ld, pred.p + (size_t)t * pred.ld, obj.p + (size_t)__ldg(o_idx + t) * obj.ld};
  }
  __device__ __forceinline__ int clampc(int c) const { return min(c, cols - 4); }
  __device__ __forceinline__ void fetch4(const Tok& k, int c, float4& a, float4& b) const {
    a = ldg4(c < D ? k.sp + c : (c < 2 * D ? k.pp + (c - D) : k.op + (c - 2 * D)));
  }
  __device__ __forceinline__ float4 finish4(const Tok& k, int c, float4 v, float4) const {
    const MatView& m = (c >= D && c < 2 * D) ? pred : obj;
    const int cc = c < D ? c : (c < 2 * D ? c - D : c - 2 * D);
    if (m.scale) {
      float4 s = ldg4(m.scale + cc), h = ldg4(m.shift + cc);
      v.x = fmaf(v.x, s.x, h.x); v.y = fmaf(v.y, s.y, h.y); v.z = fmaf(v.z, s.z, h.z); v.w = fmaf(v.w, s.w, h.w);
    }
    if (m.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    return v;
  }
  bool vec_ok() const { return obj.vec && pred.vec && D % 4 == 0; }
};

// Virtual [rows, a.cols + b.cols] matrix [ a | b ]  (decoder box_net input: cat([obj_vecs, attr_vecs]),
// reference Sg2ScVAE_model.py:166-167).  a.cols % 4 == 0.
struct Concat2 {
  static constexpr bool kTwoLoads = false;
  MatView a, b;
  int rows, cols;
  __device__ __forceinline__ float4 ld4(int r, int c) const {
    if (r >= rows || c >= cols) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < a.cols) return a.ld4(r, c);
    return b.ld4(r, c - a.cols);
  }
  __device__ __forceinline__ float at_t(int r, int c) const {
    if (r >= rows || c >= cols) return 0.f;
    return c < a.cols ? a.at(r, c) : b.at(r, c - a.cols);
  }
  struct Tok { const float* ap; const float* bp; };
  __device__ __forceinline__ Tok token(int r) const {
    r = min(r, rows - 1);
    return Tok{a.p + (size_t)r * a.ld, b.p + (size_t)r * b.ld};
  }
  __device__ __forceinline__ int clampc(int c) const { return min(c, cols - 4); }
  __device__ __forceinline__ void fetch4(const Tok& t, int c, float4& va, float4& vb) const {
    va = ldg4(c < a.cols ? t.ap + c : t.bp + (c - a.cols));
  }
  __device__ __forceinline__ float4 finish4(const Tok& t, int c, float4 v, float4) const {
    const MatView& m = c < a.cols ? a : b;
    const int cc = c < a.cols ? c : c - a.cols;
    if (m.scale) {
      float4 s = ldg4(m.scale + cc), h = ldg4(m.shift + cc);
      v.x = fmaf(v.x, s.x, h.x); v.y = fmaf(v.y, s.y, h.y); v.z = fmaf(v.z, s.z, h.z); v.w = fmaf(v.w, s.w, h.w);
    }
    if (m.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    return v;
  }
  bool vec_ok() const { return a.vec && b.vec && a.cols % 4 == 0; }
};

// Gradient w.r.t. a pre-BN activation, formed on load:  dy = g*p + y*q + r  (per-column p,q,r).
//   p == null        : dy = g                         (no normalisation)
//   q == null        : dy = g*p                       (eval-mode BN: p = gamma*rstd_running)
//   otherwise        : training-mode BN backward, see bn_bwd_finalize()
struct DyView {
  static constexpr bool kTwoLoads = true;    // g and y: two 16-byte loads per quad
  const float* g;
  int ldg;
  const float* y;
  int ldy;
  const float* p;
  const float* q;
  const float* r;
  int rows, cols;
  int vec;
  __device__ __forceinline__ float at(int i, int c) const {
    float v = __ldg(g + (size_t)i * ldg + c);
    if (p) {
      v *= __ldg(p + c);
      if (q) v += fmaf(__ldg(y + (size_t)i * ldy + c), __ldg(q + c), __ldg(r + c));
    }
    return v;
  }
  __device__ __forceinline__ float at_t(int i, int c) const { return (i < rows && c < cols) ? at(i, c) : 0.f; }
  struct Tok { const float* gp; const float* yp; };
  __device__ __forceinline__ Tok token(int r) const {
    r = min(r, rows - 1);
    return Tok{g + (size_t)r * ldg, q ? y + (size_t)r * ldy : nullptr};
  }
  __device__ __forceinline__ int clampc(int c) const { return min(c, cols - 4); }
  __device__ __forceinline__ void fetch4(const Tok& t, int c, float4& a, float4& b) const {
    a = ldg4(t.gp + c);
    if (q) b = ldg4(t.yp + c);
  }
  __device__ __forceinline__ float4 finish4(const Tok& t, int c, float4 v, float4 yy) const {
    if (p) {
      float4 pp = ldg4(p + c);
      v.x *= pp.x; v.y *= pp.y; v.z *= pp.z; v.w *= pp.w;
      if (q) {
        float4 qq = ldg4(q + c), rr = ldg4(r + c);
        v.x += fmaf(yy.x, qq.x, rr.x); v.y += fmaf(yy.y, qq.y, rr.y);
        v.z += fmaf(yy.z, qq.z, rr.z); v.w += fmaf(yy.w, qq.w, rr.w);
      }
    }
    return v;
  }
  bool vec_ok() const { return vec != 0; }
  __device__ __forceinline__ float4 ld4(int i, int c) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i >= rows || c >= cols) return v;
    if (vec) {
      v = ldg4(g + (size_t)i * ldg + c);
      if (p) {
        float4 pp = ldg4(p + c);
        v.x *= pp.x; v.y *= pp.y; v.z *= pp.z; v.w *= pp.w;
        if (q) {
          float4 yy = ldg4(y + (size_t)i * ldy + c), qq = ldg4(q + c), rr = ldg4(r + c);
          v.x += fmaf(yy.x, qq.x, rr.x); v.y += fmaf(yy.y, qq.y, rr.y);
          v.z += fmaf(yy.z, qq.z, rr.z); v.w += fmaf(yy.w, qq.w, rr.w);
        }
      }
    } else {
      v.x = at(i, c);
      if (c + 1 < cols) v.y = at(i, c + 1);
      if (c + 2 < cols) v.z = at(i, c + 2);
      if (c + 3 < cols) v.w = at(i, c + 3);
    }
    return v;
  }
};

inline DyView make_dy(const float* g, int ldg, int rows, int cols, const float* y = nullptr, int ldy = 0,
                      const float* p = nullptr, const float* q = nullptr, const float* r = nullptr) {
  DyView v;
  v.g = g; v.ldg = ldg; v.y = y; v.ldy = ldy; v.p = p; v.q = q; v.r = r; v.rows = rows; v.cols = cols;
  bool ok = ((uintptr_t)g % 16 == 0) && (ldg % 4 == 0) && (cols % 4 == 0);
  if (p) ok = ok && ((uintptr_t)p % 16 == 0);
  if (q) ok = ok && ((uintptr_t)q % 16 == 0) && ((uintptr_t)r % 16 == 0) && ((uintptr_t)y % 16 == 0) && (ldy % 4 == 0);
  v.vec = ok ? 1 : 0;
  return v;
}

// ---------------------------------------------------------------- BatchNorm finalisation (device)
enum NormMode { NORM_NONE = 0, NORM_BN_TRAIN = 1, NORM_BN_EVAL = 2 };

struct BnFwdFin {  // forward: batch statistics -> scale/shift (+ running stats), reference graph.py:14-15
  const sln_bn_sync* sync;   // null: per-rank statistics; else SyncBatchNorm over all ranks (include/sln_b200.h)
  int slot0;                 // sync slot of this layer's column block 0
  int enabled;     // 0: no statistics wanted
  float* partial;  // [row_tiles][2][N]
  unsigned* counter;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* nbt;
  float* mean;   // saved for backward
  float* rstd;
  float* scale;  // gamma*rstd
  float* shift;  // beta - mean*scale
  float eps, momentum;
  int M;
};

// (s, q) and M are the GLOBAL sums / row count when statistics are synchronised across ranks
__device__ __forceinline__ void bn_fwd_apply(const BnFwdFin& f0, int col, double s, double q, int M) {
  BnFwdFin f = f0; f.M = M;
  double mean = s / f.M;
  double var = q / f.M - mean * mean;
  if (var < 0.0) var = 0.0;
  float rstd = (float)(1.0 / sqrt(var + (double)f.eps));
  float sc = f.gamma[col] * rstd;
  f.mean[col] = (float)mean;
  f.rstd[col] = rstd;
  f.scale[col] = sc;
  f.shift[col] = f.beta[col] - (float)mean * sc;
  if (f.running_mean) {
    double unb = f.M > 1 ? var * ((double)f.M / (double)(f.M - 1)) : var;
    f.running_mean[col] = (1.f - f.momentum) * f.running_mean[col] + f.momentum * (float)mean;
    f.running_var[col] = (1.f - f.momentum) * f.running_var[col] + f.momentum * (float)unb;
    if (col == 0 && f.nbt) f.nbt[0] += 1;
  }
}

struct BnBwdFin {  // backward: column sums of g and g*yhat -> (p,q,r) of DyView + parameter gradients
  const sln_bn_sync* sync;   // null: per-rank statistics (set only for training-mode BatchNorm layers)
  int slot0;
  int mode;        // NormMode of the layer whose pre-activation gradient is being formed
  float* partial;  // [row_tiles][2][N]
  unsigned* counter;
  const float* gamma;
  const float* mean;
  const float* rstd;
  const float* scale;  // forward scale (gamma*rstd)
  float* p;
  float* q;
  float* r;
  float* dgamma;  // accumulated (+=)
  float* dbeta;
  float* dbias;  // Linear bias gradient = sum_rows dy
  int M;
};

// (sg, sgy): this rank's sums -> parameter gradients (the gradient all-reduce sums them over ranks);
// (sg_all, sgy_all, M): sums / rows over all ranks -> the dy formula of training-mode BatchNorm (equal to the local ones without sync)
__device__ __forceinline__ void bn_bwd_apply(const BnBwdFin& f0, int col, double sg, double sgy, double sg_all, double sgy_all, int M) {
  BnBwdFin f = f0; f.M = M;
  if (f.mode == NORM_NONE) {
    if (f.dbias) f.dbias[col] += (float)sg;
    return;
  }
  if (f.dgamma) f.dgamma[col] += (float)sgy;
  if (f.dbeta) f.dbeta[col] += (float)sg;
  float s = f.scale[col];
  if (f.mode == NORM_BN_EVAL) {
    f.p[col] = s;
    if (f.dbias) f.dbias[col] += s * (float)sg;
    return;
  }
  // training-mode BN:  dy = s*(g - c1 - yhat*c2),  yhat = (y-mean)*rstd,  c1 = sum(g)/M, c2 = sum(g*yhat)/M
  double c1 = sg_all / f.M, c2 = sgy_all / f.M;
  double rs = f.rstd[col], mu = f.mean[col];
  f.p[col] = s;
  f.q[col] = (float)(-(double)s * rs * c2);
  f.r[col] = (float)((double)s * (mu * rs * c2 - c1));
  // dbias is exactly zero under training-mode BN (the mean subtraction removes it); leave the zeroed buffer.
}

// ---------------------------------------------------------------- column-statistics finalisation
// Protocol shared by every kernel that produces per-row-tile column partials partial[tile][2][N]:
// after a CTA has written its partials for the column block blockIdx.x it takes a ticket on counter[blockIdx.x]; the
// last of the gridDim.y row tiles reduces that block's columns over all tiles (fp64, fixed order -> deterministic) with
// every thread of the CTA (threads = column x tile-group, independent loads in flight) and applies `fin`.
// One counter per column block keeps the finalisation parallel across column blocks and off the critical path of the
// other CTAs (a single "last CTA reduces everything" tail cost ~60 us per Linear at config 2).
constexpr int kCounterStride = 16;   // counters reserved per BatchNorm layer (>= number of column blocks)

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// fin(col, S_local, Q_local, S_all, Q_all, rows_all)
template <int NT, class FinFn>
__device__ __forceinline__ void finalize_column_block(const float* partial, unsigned* counter, int n0, int ncols, int N, int tid,
                                                      double* sred, int* s_last, FinFn fin, const sln_bn_sync* sync = nullptr,
                                                      int slot0 = 0, int Mlocal = 0) {
  const int tiles = gridDim.y;
  const int cb = n0 / ncols;            // column block (== blockIdx.x unless the launch covers a subset of the column tiles)
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned ticket = atomicAdd(counter + cb, 1u);
    *s_last = (ticket == (unsigned)tiles - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  // thread = (column c, tile group g); every launch configuration has NT >= W >= ncols
  const int W = ncols <= 32 ? 32 : (ncols <= 64 ? 64 : (ncols <= 128 ? 128 : 256));
  const int G = NT / W;
  const int c = tid % W, g = tid / W;
  const bool live = g < G && c < ncols && n0 + c < N;
  double s = 0.0, q = 0.0;
  if (live) {
    const float* p0 = partial + n0 + c;
    int b = g;
    for (; b + 3 * G < tiles; b += 4 * G) {   // 8 independent L2 loads in flight per thread
      float a0 = __ldcg(p0 + ((size_t)b * 2) * N), a1 = __ldcg(p0 + ((size_t)(b + G) * 2) * N);
      float a2 = __ldcg(p0 + ((size_t)(b + 2 * G) * 2) * N), a3 = __ldcg(p0 + ((size_t)(b + 3 * G) * 2) * N);
      float c0 = __ldcg(p0 + ((size_t)b * 2 + 1) * N), c1 = __ldcg(p0 + ((size_t)(b + G) * 2 + 1) * N);
      float c2 = __ldcg(p0 + ((size_t)(b + 2 * G) * 2 + 1) * N), c3 = __ldcg(p0 + ((size_t)(b + 3 * G) * 2 + 1) * N);
      s += (double)a0; s += (double)a1; s += (double)a2; s += (double)a3;
      q += (double)c0; q += (double)c1; q += (double)c2; q += (double)c3;
    }
    for (; b < tiles; b += G) {
      s += (double)__ldcg(p0 + ((size_t)b * 2) * N);
      q += (double)__ldcg(p0 + ((size_t)b * 2 + 1) * N);
    }
  }
  sred[2 * tid] = s;
  sred[2 * tid + 1] = q;
  __syncthreads();
  double S = 0.0, Q = 0.0;
  if (live && g == 0) {
    for (int gg = 0; gg < G; ++gg) { S += sred[2 * (gg * W + c)]; Q += sred[2 * (gg * W + c) + 1]; }
  }
  if (sync == nullptr) {
    if (live && g == 0) fin(n0 + c, S, Q, S, Q, Mlocal);
  } else {
    // SyncBatchNorm: one-shot all-gather of this column block's sums over peer memory (NVLink), reduced in rank order
    const int world = sync->world, rank = sync->rank;
    const size_t slot = (size_t)slot0 + cb;
    if (live && g == 0) {
      for (int r = 0; r < world; ++r) {
        double* dst = sync->recv[r] + ((slot * world + rank) * SLN_BN_SYNC_COLS + c) * 3;
        dst[0] = S; dst[1] = Q; dst[2] = (double)Mlocal;
      }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const unsigned expect = (sync->use[slot] += 1u) * (unsigned)world;
      for (int r = 0; r < world; ++r) atomicAdd_system(sync->flag[r] + slot, 1u);
      while (ld_acquire_sys(sync->flag[rank] + slot) < expect) { __nanosleep(64); }
    }
    __syncthreads();
    if (live && g == 0) {
      double Sa = 0.0, Qa = 0.0, Ma = 0.0;
      for (int r = 0; r < world; ++r) {
        const volatile double* src = sync->recv[rank] + ((slot * world + r) * SLN_BN_SYNC_COLS + c) * 3;
        Sa += src[0]; Qa += src[1]; Ma += src[2];
      }
      fin(n0 + c, S, Q, Sa, Qa, (int)Ma);
    }
  }
  if (tid == 0) counter[cb] = 0u;
}

// ---------------------------------------------------------------- epilogues
// Every epilogue sees the 8x8 register micro-tile of one thread: rows i0 + {0..3} and i0 + BM/2 + {0..3},
// columns j0 + {0..3} and j0 + BN/2 + {0..3}.
template <int BM, int BN>
struct TileCoord {
  int m0, n0, ty, tx, tid;
  __device__ __forceinline__ int row(int a) const { return m0 + (a < 4 ? ty * 4 + a : BM / 2 + ty * 4 + (a - 4)); }
  __device__ __forceinline__ int col(int b) const { return n0 + (b < 4 ? tx * 4 + b : BN / 2 + tx * 4 + (b - 4)); }
};

// block-level: reduce per-thread column partials (s1,s2 over the thread's 8 rows) across the BM/8 thread rows,
// write them to partial[blockIdx.y][{0,1}][N]; returns true in the last CTA of the grid (all partials visible).
template <int BM, int BN>
__device__ __forceinline__ void column_partials(const TileCoord<BM, BN>& tc, float (&s1)[8], float (&s2)[8],
                                                float* smem, float* partial, int N) {
  constexpr int TY = BM / 8;
  float* r1 = smem;            // [TY][BN]
  float* r2 = smem + TY * BN;  // [TY][BN]
  __syncthreads();             // the main loop's tiles are dead
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    int jl = (b < 4 ? tc.tx * 4 + b : BN / 2 + tc.tx * 4 + (b - 4));
    r1[tc.ty * BN + jl] = s1[b];
    r2[tc.ty * BN + jl] = s2[b];
  }
  __syncthreads();
  if (tc.tid < BN) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int t = 0; t < TY; ++t) { a += r1[t * BN + tc.tid]; c += r2[t * BN + tc.tid]; }
    int j = tc.n0 + tc.tid;
    if (j < N) {
      partial[((size_t)blockIdx.y * 2 + 0) * N + j] = a;
      partial[((size_t)blockIdx.y * 2 + 1) * N + j] = c;
    }
  }
}

// C = acc + bias (+ BN batch statistics).
struct EpiStore {
  float* C;
  int ldc;
  const float* bias;  // per output column, may be null
  BnFwdFin fin;
  template <int BM, int BN>
  __device__ __forceinline__ void run(float (&acc)[8][8], const TileCoord<BM, BN>& tc, float* smem, int M, int N) const {
    constexpr int NT = (BM / 8) * (BN / 8);
    float s1[8], s2[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) { s1[b] = 0.f; s2[b] = 0.f; }
    const bool vec = ((uintptr_t)C % 16 == 0) && (ldc % 4 == 0);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      int i = tc.row(a);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int j = tc.col(h * 4);
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float bb = (bias && j + e < N) ? __ldg(bias + j + e) : 0.f;
          v[e] = acc[a][h * 4 + e] + bb;
        }
        if (i < M) {
          if (vec && j + 3 < N) {
            *reinterpret_cast<float4*>(C + (size_t)i * ldc + j) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (j + e < N) C[(size_t)i * ldc + j + e] = v[e];
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) { s1[h * 4 + e] += v[e]; s2[h * 4 + e] = fmaf(v[e], v[e], s2[h * 4 + e]); }
        }
      }
    }
    if (fin.enabled) {
      column_partials<BM, BN>(tc, s1, s2, smem, fin.partial, N);
      __shared__ int s_last;
      const BnFwdFin& f = fin;
      finalize_column_block<NT>(fin.partial, fin.counter, tc.n0, BN, N, tc.tid, reinterpret_cast<double*>(smem), &s_last,
                                [&](int col, double, double, double S, double Q, int Mt) { bn_fwd_apply(f, col, S, Q, Mt); }, f.sync, f.slot0, f.M);
    }
  }
};

// G = relu_mask(yprev) ? (acc + add) : 0, plus BN-backward column sums of the layer behind the activation.
struct EpiMaskReduce {
  float* G;
  int ldg;
  const float* add;  // optional second gradient contribution, same shape
  int ldadd;
  const float* yprev;  // pre-activation of the producing layer
  int ldy;
  const float* scale;  // forward affine of that layer (null: identity)
  const float* shift;
  const float* mean;  // BN statistics of that layer (null when NORM_NONE)
  const float* rstd;
  BnBwdFin fin;
  template <int BM, int BN>
  __device__ __forceinline__ void run(float (&acc)[8][8], const TileCoord<BM, BN>& tc, float* smem, int M, int N) const {
    constexpr int NT = (BM / 8) * (BN / 8);
    float s1[8], s2[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) { s1[b] = 0.f; s2[b] = 0.f; }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      int i = tc.row(a);
      if (i >= M) continue;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        int j = tc.col(b);
        if (j >= N) continue;
        float y = __ldg(yprev + (size_t)i * ldy + j);
        float pre = scale ? fmaf(y, __ldg(scale + j), __ldg(shift + j)) : y;
        float d = acc[a][b];
        if (add) d += __ldg(add + (size_t)i * ldadd + j);
        float g = pre > 0.f ? d : 0.f;
        G[(size_t)i * ldg + j] = g;
        s1[b] += g;
        float yh = mean ? (y - __ldg(mean + j)) * __ldg(rstd + j) : 0.f;
        s2[b] = fmaf(g, yh, s2[b]);
      }
    }
    column_partials<BM, BN>(tc, s1, s2, smem, fin.partial, N);
    __shared__ int s_last;
    const BnBwdFin& f = fin;
    finalize_column_block<NT>(fin.partial, fin.counter, tc.n0, BN, N, tc.tid, reinterpret_cast<double*>(smem), &s_last,
                              [&](int col, double S, double Q, double Sa, double Qa, int Mt) { bn_bwd_apply(f, col, S, Q, Sa, Qa, Mt); }, f.sync, f.slot0, f.M);
  }
};

// C += acc  (split-K weight gradients; C pre-zeroed by the caller).
struct EpiAtomic {
  float* C;
  int ldc;
  template <int BM, int BN>
  __device__ __forceinline__ void run(float (&acc)[8][8], const TileCoord<BM, BN>& tc, float* smem, int M, int N) const {
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      int i = tc.row(a);
      if (i >= M) continue;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        int j = tc.col(b);
        if (j < N) red_add(C + (size_t)i * ldc + j, acc[a][b]);
      }
    }
  }
};

// ---------------------------------------------------------------- the kernel
// C[i,j] = sum_k A(i,k) * B(j,k).  X_RC == true : operand stored [out][k] (k contiguous), read as ld4(out, k)
//                                  X_RC == false: operand stored [k][out] (out contiguous), read as ld4(k, out)
// grid = (ceil(N/BN), ceil(M/BM), splits); each z-slice reduces k in [z*kchunk, (z+1)*kchunk).
template <int BM, int BN, bool A_RC, bool B_RC, class AOp, class BOp, class Epi>
__global__ void __launch_bounds__((BM / 8) * (BN / 8), 512 / ((BM / 8) * (BN / 8)))
gemm_kernel(const AOp A, const BOp B, const Epi epi, int M, int N, int K, int kchunk) {
  constexpr int BK = 16;
  constexpr int NT = (BM / 8) * (BN / 8);
  constexpr int LDA = BM + 4, LDB = BN + 4;
  constexpr int A_F4 = BM * BK / 4 / NT, B_F4 = BN * BK / 4 / NT;
  __shared__ __align__(16) float smem[2 * BK * (LDA + LDB)];
  float* As = smem;
  float* Bs = smem + 2 * BK * LDA;

  const int tid = threadIdx.x;
  TileCoord<BM, BN> tc;
  tc.tid = tid; tc.tx = tid % (BN / 8); tc.ty = tid / (BN / 8);
  tc.m0 = blockIdx.y * BM; tc.n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);

  float4 ra[A_F4], rb[B_F4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      int f = tid + i * NT;
      if (A_RC) { int o = f % BM, kq = f / BM; ra[i] = (k0 + kq * 4 < kend) ? A.ld4(tc.m0 + o, k0 + kq * 4) : make_float4(0.f, 0.f, 0.f, 0.f); }
      else { int o4 = f % (BM / 4), k = f / (BM / 4); ra[i] = (k0 + k < kend) ? A.ld4(k0 + k, tc.m0 + o4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f); }
    }
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      int f = tid + i * NT;
      if (B_RC) { int o = f % BN, kq = f / BN; rb[i] = (k0 + kq * 4 < kend) ? B.ld4(tc.n0 + o, k0 + kq * 4) : make_float4(0.f, 0.f, 0.f, 0.f); }
      else { int o4 = f % (BN / 4), k = f / (BN / 4); rb[i] = (k0 + k < kend) ? B.ld4(k0 + k, tc.n0 + o4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f); }
    }
  };
  auto sstore = [&](int buf) {
    float* a = As + buf * BK * LDA;
    float* b = Bs + buf * BK * LDB;
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      int f = tid + i * NT;
      if (A_RC) {
        int o = f % BM, kq = f / BM;
        a[(kq * 4 + 0) * LDA + o] = ra[i].x; a[(kq * 4 + 1) * LDA + o] = ra[i].y;
        a[(kq * 4 + 2) * LDA + o] = ra[i].z; a[(kq * 4 + 3) * LDA + o] = ra[i].w;
      } else {
        int o4 = f % (BM / 4), k = f / (BM / 4);
        *reinterpret_cast<float4*>(a + k * LDA + o4 * 4) = ra[i];
      }
    }
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      int f = tid + i * NT;
      if (B_RC) {
        int o = f % BN, kq = f / BN;
        b[(kq * 4 + 0) * LDB + o] = rb[i].x; b[(kq * 4 + 1) * LDB + o] = rb[i].y;
        b[(kq * 4 + 2) * LDB + o] = rb[i].z; b[(kq * 4 + 3) * LDB + o] = rb[i].w;
      } else {
        int o4 = f % (BN / 4), k = f / (BN / 4);
        *reinterpret_cast<float4*>(b + k * LDB + o4 * 4) = rb[i];
      }
    }
  };

  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

  const int nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;
  if (nk > 0) {
    gload(kbeg);
    sstore(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) gload(kbeg + (kt + 1) * BK);
    const float* a = As + cur * BK * LDA + tc.ty * 4;
    const float* b = Bs + cur * BK * LDB + tc.tx * 4;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(a + k * LDA);
      float4 a1 = *reinterpret_cast<const float4*>(a + k * LDA + BM / 2);
      float4 b0 = *reinterpret_cast<const float4*>(b + k * LDB);
      float4 b1 = *reinterpret_cast<const float4*>(b + k * LDB + BN / 2);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) sstore(cur ^ 1);
    __syncthreads();
  }
  epi.template run<BM, BN>(acc, tc, smem, M, N);
}

// ---------------------------------------------------------------- host-side launch
struct TileChoice { int bm, bn; int splits; int kchunk; };

// Pick the tile shape that minimises (max CTAs per SM) x (tile area / efficiency) on 148 SMs.
inline TileChoice pick_tile(int M, int N, int K, bool allow_split) {
  const int bms[3] = {128, 128, 64}, bns[3] = {128, 64, 64};
  const double eff[3] = {1.0, 0.92, 0.80};
  const int occ[3] = {2, 4, 8};  // resident CTAs per SM (register bound)
  double best = 1e300;
  TileChoice c{64, 64, 1, K};
  for (int t = 0; t < 3; ++t) {
    int tiles = ceil_div(M, bms[t]) * ceil_div(N, bns[t]);
    int splits = 1;
    if (allow_split) {
      int want = 2 * kNumSMs;
      splits = tiles >= want ? 1 : ceil_div(want, tiles);
      int max_splits = ceil_div(K, 256);  // keep >= 256 reduction rows per CTA so RED traffic stays minor
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
    }
    int kchunk = ceil_div(ceil_div(K, splits), 16) * 16;
    splits = ceil_div(K, kchunk);
    int ctas = tiles * splits;
    int per_sm = ceil_div(ctas, kNumSMs);
    // CTAs beyond the resident set serialise; resident ones share the SM's FMA pipe
    (void)occ;
    double cost = (double)per_sm * bms[t] * bns[t] * (double)kchunk / eff[t];
    if (cost < best) { best = cost; c = TileChoice{bms[t], bns[t], splits, kchunk}; }
  }
  return c;
}

inline int max_row_tiles(int M) { return ceil_div(M, 16); }

template <bool A_RC, bool B_RC, class AOp, class BOp, class Epi>
int launch_gemm(cudaStream_t st, const AOp& A, const BOp& B, const Epi& epi, int M, int N, int K, bool allow_split,
                const char* what, int prof_cls) {
  if (M <= 0 || N <= 0) return SLN_OK;
  ProfScope prof(st, prof_cls, 2.0 * (double)M * (double)N * (double)K);
  TileChoice c = pick_tile(M, N, K, allow_split);
  dim3 grid(ceil_div(N, c.bn), ceil_div(M, c.bm), c.splits);
  if (c.bm == 128 && c.bn == 128)
    gemm_kernel<128, 128, A_RC, B_RC, AOp, BOp, Epi><<<grid, 256, 0, st>>>(A, B, epi, M, N, K, c.kchunk);
  else if (c.bm == 128 && c.bn == 64)
    gemm_kernel<128, 64, A_RC, B_RC, AOp, BOp, Epi><<<grid, 128, 0, st>>>(A, B, epi, M, N, K, c.kchunk);
  else
    gemm_kernel<64, 64, A_RC, B_RC, AOp, BOp, Epi><<<grid, 64, 0, st>>>(A, B, epi, M, N, K, c.kchunk);
  return check_launch(what);
}

}  // namespace sln
