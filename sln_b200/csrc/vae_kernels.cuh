// Non-contraction kernels of the VAE-graph hot path: graph preparation (CSR by node), embeddings,
// avg-pool scatter expressed as a deterministic gather, BN-backward "prep" passes, losses, Adam.
// All of them are HBM/L2-bound: one pass over their operands, coalesced along the feature dimension.
#pragma once
#include "gemm.cuh"

namespace sln {

// ================================================================ graph preparation
// reference: Sg2ScVAE_model.py:117-119 (split triples), graph.py:74-75,92-107 (s/o indices, degree counts)
struct Graph {
  int O, T;
  int* s_idx;    // [T]
  int* p_idx;    // [T]
  int* o_idx;    // [T]
  int* row_ptr;  // [O+1]  CSR over nodes; entries = triples touching the node as subject (side 0) or object (side 1)
  int* ent;      // [2T]   (side << 30) | t, sorted ascending inside each row (= reference scatter_add order)
  float* cnt;  // [O]  max(deg,1) as float (reference graph.py:102-108 divides by the clamped count)
  int* cursor;     // [O]  scratch
};

// Index validation: out-of-range ids are remapped to row 0 (so no kernel ever reads or writes out of bounds) and recorded as a bit
// in the workspace's index flag (include/sln_b200.h SLN_IDX_*), which the caller reads back — the reference raises IndexError.
__global__ void k_split_triples(const long long* __restrict__ triples, int T, int O, int num_preds, int* s_idx, int* p_idx, int* o_idx,
                                int* deg, int* err) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  long long s = triples[(size_t)t * 3 + 0], p = triples[(size_t)t * 3 + 1], o = triples[(size_t)t * 3 + 2];
  if (s < 0 || s >= O || o < 0 || o >= O) { atomicOr(err, 8); s = 0; o = 0; }
  if (p < 0 || p >= num_preds) { atomicOr(err, 16); p = 0; }
  s_idx[t] = (int)s; p_idx[t] = (int)p; o_idx[t] = (int)o;
  atomicAdd(deg + s, 1);
  atomicAdd(deg + o, 1);
}

// single-CTA exclusive scan (O is at most a few 10^5; runs once per batch)
__global__ void k_scan_deg(const int* __restrict__ deg, int O, int* row_ptr, float* cnt) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int base = 0; base < O; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < O ? deg[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
      int tot = lane < nw ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, tot, o); if (lane >= o) tot += y; }
      if (lane < nw) warp_tot[lane] = tot;  // inclusive over warps
    }
    __syncthreads();
    int warp_off = w > 0 ? warp_tot[w - 1] : 0;
    int carry = carry_s;
    if (i < O) {
      row_ptr[i] = carry + warp_off + x - v;
      cnt[i] = (float)max(v, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + warp_tot[nw - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) row_ptr[O] = carry_s;
}

__global__ void k_fill_csr(const int* __restrict__ s_idx, const int* __restrict__ o_idx, int T, const int* __restrict__ row_ptr,
                           int* cursor, int* ent) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int s = s_idx[t], o = o_idx[t];
  ent[row_ptr[s] + atomicAdd(cursor + s, 1)] = t;
  ent[row_ptr[o] + atomicAdd(cursor + o, 1)] = (1 << 30) | t;
}

// One thread per node sorts short rows by insertion; rows above 16 entries (the room node of a scene: one entry per object) are
// rank-sorted by the whole warp — every lane counts, for its keys, the smaller keys of the row (keys are unique: entry = side bit |
// triple index), then writes each key to its rank: no dependent chain (a 62-entry insertion sort in global memory took 17 us).
constexpr int kSortShort = 16, kSortLaneKeys = 8;        // warp path: rows up to 32 * kSortLaneKeys entries
__global__ void __launch_bounds__(128) k_sort_rows(const int* __restrict__ row_ptr, int O, int* ent) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  const int b = o < O ? row_ptr[o] : 0, n = o < O ? row_ptr[o + 1] - b : 0;
  const bool warp_row = n > kSortShort && n <= 32 * kSortLaneKeys;
  if (!warp_row) {
    for (int i = b + 1; i < b + n; ++i) {  // insertion sort: rows are short (degree of a scene-graph node)
      int key = ent[i], j = i - 1;
      while (j >= b && ent[j] > key) { ent[j + 1] = ent[j]; --j; }
      ent[j + 1] = key;
    }
  }
  unsigned todo = __ballot_sync(0xffffffffu, warp_row);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const int bb = __shfl_sync(0xffffffffu, b, src), nn = __shfl_sync(0xffffffffu, n, src);
    int key[kSortLaneKeys], rank[kSortLaneKeys];
#pragma unroll
    for (int q = 0; q < kSortLaneKeys; ++q) {
      const int i = q * 32 + lane;
      key[q] = i < nn ? ent[bb + i] : 0;
      rank[q] = 0;
    }
    for (int m = 0; m < nn; ++m) {
      const int other = ent[bb + m];                     // same address in every lane: one broadcast load
#pragma unroll
      for (int q = 0; q < kSortLaneKeys; ++q) rank[q] += other < key[q] ? 1 : 0;
    }
    __syncwarp();                                        // every lane has read the row before anyone overwrites it
#pragma unroll
    for (int q = 0; q < kSortLaneKeys; ++q)
      if (q * 32 + lane < nn) ent[bb + rank[q]] = key[q];
    __syncwarp();
  }
}

__global__ void k_i64_to_i32(const long long* __restrict__ src, int n, int* dst, int limit, int* err, int bit) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long v = src[i];
  if (v < 0 || v >= limit) { atomicOr(err, bit); v = 0; }
  dst[i] = (int)v;
}

// ================================================================ embeddings
// out[i, col_off + c] = table[idx ? idx[i] : i][c]           (reference Sg2ScVAE_model.py:121-129,150-154)
__global__ void k_gather_rows(const float* __restrict__ table, int ldt, const int* __restrict__ idx, int n, int width,
                              float* out, int ldo, int col_off) {
  int i = blockIdx.x * blockDim.y + threadIdx.y;
  if (i >= n) return;
  int r = idx ? __ldg(idx + i) : i;
  for (int c = threadIdx.x; c < width; c += blockDim.x) out[(size_t)i * ldo + col_off + c] = __ldg(table + (size_t)r * ldt + c);
}

// out[i, col_off + c] = sum_k x[i,k] * W[c,k] + b[c]   for tiny k (box_embeddings: Linear(6 or 4, 3E/4))
__global__ void k_small_linear(const float* __restrict__ x, int n, int kin, const float* __restrict__ W, const float* __restrict__ b,
                               int width, float* out, int ldo, int col_off) {
  int i = blockIdx.x * blockDim.y + threadIdx.y;
  if (i >= n) return;
  for (int c = threadIdx.x; c < width; c += blockDim.x) {
    float acc = b ? __ldg(b + c) : 0.f;
    for (int k = 0; k < kin; ++k) acc = fmaf(__ldg(x + (size_t)i * kin + k), __ldg(W + (size_t)c * kin + k), acc);
    out[(size_t)i * ldo + col_off + c] = acc;
  }
}

// table_grad[r, c] += sum_{i: idx[i]==r} ( a[i, c] (+ b[i, c]) ); idx == null -> single row 0 (column sum: Linear bias grads).
// grid (table rows, index splits, column blocks of 128); the 8 warps of a CTA walk interleaved entries of their index
// range (the match test is warp-uniform, the row read a coalesced 128..512-byte segment), reduce through shared memory in
// a fixed order and add the CTA's partial with one RED.ADD per column.
__global__ void __launch_bounds__(256) k_embed_bwd(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                                   const int* __restrict__ idx, int n, int width, float* table_grad, int ldt, int per_cta) {
  const int r = blockIdx.x;
  const int i0 = blockIdx.y * per_cta, i1 = min(n, i0 + per_cta);
  const int c0 = blockIdx.z * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = i0 + warp; i < i1; i += 8) {
    int id = idx ? __ldg(idx + i) : r;
    if (id != r) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = c0 + lane + 32 * j;
      if (c < width) {
        float v = __ldg(a + (size_t)i * lda + c);
        if (b) v += __ldg(b + (size_t)i * ldb + c);
        acc[j] += v;
      }
    }
  }
  __shared__ float red[8][128];
#pragma unroll
  for (int j = 0; j < 4; ++j) red[warp][lane + 32 * j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 128 && c0 + threadIdx.x < width) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    if (s != 0.f) red_add(table_grad + (size_t)r * ldt + c0 + threadIdx.x, s);
  }
}
// Small tables (the four embedding tables: <= 33 rows, <= 128 columns): every warp accumulates its entries into a private copy of the
// table in shared memory (fixed order inside a warp), the copies are summed in warp order and the CTA adds one partial per element.
__global__ void __launch_bounds__(256) k_embed_bwd_tab(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                                       const int* __restrict__ idx, int n, int width, int rows, float* table_grad, int per_warp) {
  extern __shared__ float tab[];   // [8][rows * width]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sz = rows * width;
  for (int e = threadIdx.x; e < 8 * sz; e += 256) tab[e] = 0.f;
  __syncthreads();
  float* mine = tab + warp * sz;
  const int i0 = (blockIdx.x * 8 + warp) * per_warp, i1 = min(n, i0 + per_warp);
  for (int i = i0; i < i1; i += 4) {             // four rows per trip: their index and value loads are independent (adds stay in row order)
    int id[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) id[u] = i + u < i1 ? __ldg(idx + i + u) : 0;
    for (int c = lane; c < width; c += 32) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = i + u < i1 ? __ldg(a + (size_t)(i + u) * lda + c) : 0.f;
        if (b && i + u < i1) v[u] += __ldg(b + (size_t)(i + u) * ldb + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u < i1) mine[id[u] * width + c] += v[u];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < sz; e += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += tab[w * sz + e];
    if (s != 0.f) red_add(table_grad + e, s);
  }
}
inline int launch_embed_bwd(cudaStream_t st, const float* a, int lda, const float* b, int ldb, const int* idx, int n, int width,
                            float* table_grad, int rows) {
  if (!table_grad || n <= 0 || width <= 0) return SLN_OK;
  if (idx && (size_t)rows * width * 8 * sizeof(float) <= 64 * 1024) {
    const int per_warp = 16;
    const size_t smem = (size_t)rows * width * 8 * sizeof(float);
    if (smem > 48 * 1024) {
      static unsigned long long configured = 0ull;   // devices configured (per call site / instantiation)
      if (first_use_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(k_embed_bwd_tab, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(k_embed_bwd_tab) failed: %s", cudaGetErrorString(e)); return SLN_ECUDA; }
      }
    }
    k_embed_bwd_tab<<<ceil_div(n, 8 * per_warp), 256, smem, st>>>(a, lda, b, ldb, idx, n, width, rows, table_grad, per_warp);
    return check_launch("embed_bwd_tab");
  }
  int splits = max(1, min(ceil_div(n, 64), (2 * kNumSMs) / max(rows, 1)));
  int per_cta = ceil_div(n, splits);
  splits = ceil_div(n, per_cta);
  dim3 grid(rows, splits, ceil_div(width, 128));
  k_embed_bwd<<<grid, 256, 0, st>>>(a, lda, b, ldb, idx, n, width, table_grad, width, per_cta);
  return check_launch("embed_bwd");
}

// ================================================================ avg pooling  (reference graph.py:92-108)
// pooled[o, c] = ( sum_{t: s_t = o} a2[t, c] + sum_{t: o_t = o} a2[t, H + D + c] ) / max(deg[o], 1),  a2 = relu(bn(y2)) lazily.
// Streaming segmented sum over the CSR.  A group of `lanes` threads (V columns each; whole warps) owns `npg` <= 32 CONSECUTIVE
// nodes and walks their concatenated entry list as one stream, so the index chain (row_ptr -> entries -> rows) is paid once
// per group, not once per node:
//   * lane j of every warp holds row_ptr[o0 + j + 1] — node boundaries are warp shuffles, not dependent loads;
//   * entries are read 32 at a time, one per lane (coalesced), the next 32 prefetched while rows are in flight;
//   * BATCH independent row reads per thread (contiguous V*lanes*4-byte segments per group) are issued before the first add.
// Sums are taken in CSR order per column (= the reference's two scatter_add calls: all subject-side rows in triple order,
// then all object-side rows), then truly divided: bit-identical to scatter_add + clamp + divide.  No atomics; the only
// barrier is the one after staging the lazy-BatchNorm vectors in shared memory.
template <int V> __device__ __forceinline__ void pool_ld(const float* p, float (&o)[V]);
template <> __device__ __forceinline__ void pool_ld<1>(const float* p, float (&o)[1]) { o[0] = __ldg(p); }
template <> __device__ __forceinline__ void pool_ld<2>(const float* p, float (&o)[2]) { float2 t = __ldg(reinterpret_cast<const float2*>(p)); o[0] = t.x; o[1] = t.y; }
template <> __device__ __forceinline__ void pool_ld<4>(const float* p, float (&o)[4]) { float4 t = ldg4(p); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }

template <int V, int BATCH, int MINB>
__global__ void __launch_bounds__(256, MINB) k_pool_fwd(const MatView a2, const int* __restrict__ row_ptr, const int* __restrict__ ent,
                                                        int O, int H, int D, float* pooled, int lanes, int npg) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next contraction may start its prologue (it waits for us before reading)
  extern __shared__ float s_bn[];                                    // [4][lanes*V]: scale_s, shift_s, scale_o, shift_o of this column block
  const int W = lanes * V;
  const int cl = (threadIdx.x % lanes) * V;                          // column within the block
  const int c = blockIdx.y * W + cl;                                 // first of this thread's V columns
  const bool bn = a2.scale != nullptr;
  if (bn) {
    for (int i = threadIdx.x; i < W; i += 256) {
      const int cc = blockIdx.y * W + i;
      const bool ok = cc < H;
      s_bn[i] = ok ? __ldg(a2.scale + cc) : 0.f;                 s_bn[W + i] = ok ? __ldg(a2.shift + cc) : 0.f;
      s_bn[2 * W + i] = ok ? __ldg(a2.scale + H + D + cc) : 0.f; s_bn[3 * W + i] = ok ? __ldg(a2.shift + H + D + cc) : 0.f;
    }
    __syncthreads();
  }
  const int per_cta = 256 / lanes;
  const int g = threadIdx.x / lanes;
  const int o0 = (blockIdx.x * per_cta + g) * npg;
  if (g >= per_cta || o0 >= O) return;                               // warp-uniform: lanes is a multiple of 32
  const int lane = threadIdx.x & 31;
  const bool col_ok = c < H;                                         // idle lanes stay for the shuffles
  const int n_nodes = min(O, o0 + npg) - o0;
  const int rp = __ldg(row_ptr + min(o0 + lane + 1, O));             // lane j: end of node o0 + j
  int k = __ldg(row_ptr + o0);
  const int k_end = __shfl_sync(0xffffffffu, rp, n_nodes - 1);
  int j_node = 0, b = k, e = __shfl_sync(0xffffffffu, rp, 0);        // current node o0 + j_node and its entry range
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  auto flush = [&]() {
    const float ic = (float)max(e - b, 1);                           // = clamp(count, min=1) of graph.py:102-107
    if (col_ok) {
      float* dst = pooled + (size_t)(o0 + j_node) * H + c;
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = __fdiv_rn(acc[j], ic);
      if (V == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      else if (V == 2) *reinterpret_cast<float2*>(dst) = make_float2(acc[0], acc[1]);
      else dst[0] = acc[0];
    }
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
  };
  int en = k + lane < k_end ? __ldg(ent + k + lane) : -1;
  while (k < k_end) {
    int nx = -1;
#pragma unroll
    for (int h = 0; h < 32 / BATCH; ++h) {
      float v[BATCH][V];
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const int x = __shfl_sync(0xffffffffu, en, h * BATCH + u);
        if (x >= 0 && col_ok) pool_ld<V>(a2.p + (size_t)(x & ((1 << 30) - 1)) * a2.ld + ((x >> 30) ? (H + D) : 0) + c, v[u]);
      }
      if (h == 0 && k + 32 + lane < k_end) nx = __ldg(ent + k + 32 + lane);
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const int x = __shfl_sync(0xffffffffu, en, h * BATCH + u);
        if (x < 0) break;                                            // past the group's last entry (warp-uniform)
        const int kk = k + h * BATCH + u;
        while (kk == e) { flush(); ++j_node; b = e; e = __shfl_sync(0xffffffffu, rp, j_node); }   // node boundary (steps over empty nodes too)
        if (col_ok) {
          const float* sb = s_bn + ((x >> 30) ? 2 * W : 0) + cl;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            float t = v[u][j];
            if (bn) t = fmaf(t, sb[j], sb[W + j]);
            if (a2.relu) t = fmaxf(t, 0.f);
            acc[j] += t;
          }
        }
      }
    }
    k += 32;
    en = nx;
  }
  for (;;) {                                                         // the last node and any trailing empty ones
    flush();
    if (++j_node >= n_nodes) break;
    b = e; e = __shfl_sync(0xffffffffu, rp, j_node);
  }
}

// Small graphs (BASELINE configs[1]: 2048 nodes, 10 MB, L2-resident): ONE WARP PER NODE, lanes along the columns (QPL float4 per lane),
// 8 rows in flight per lane, sums in CSR order (bit-identical to k_pool_fwd and to scatter_add + clamp + divide).  The streaming
// kernel above is unrolled for bandwidth (32 entries x 16 rows per iteration: ~7000 SASS instructions that every warp steps through
// even for a 3-entry node — ncu at 64 scenes: 14.5 us, issue slots 47 % busy, DRAM idle, profiles/r2_prof_pool64_*); this one is a
// short loop whose makespan is the 62-entry room node of a scene (8 dependent batches).
template <int QPL>
__global__ void __launch_bounds__(256) k_pool_fwd_node(const MatView a2, const int* __restrict__ row_ptr, const int* __restrict__ ent,
                                                       int O, int H, int D, float* pooled) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int node = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (node >= O) return;
  const int b = __ldg(row_ptr + node), e = __ldg(row_ptr + node + 1);
  const bool bn = a2.scale != nullptr;
  float4 acc[QPL], ss[QPL], hs[QPL], so[QPL], ho[QPL];
#pragma unroll
  for (int q = 0; q < QPL; ++q) {
    acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = (lane + 32 * q) * 4;
    if (bn) { ss[q] = ldg4(a2.scale + c); hs[q] = ldg4(a2.shift + c); so[q] = ldg4(a2.scale + H + D + c); ho[q] = ldg4(a2.shift + H + D + c); }
  }
  int mine = (b + lane < e) ? __ldg(ent + b + lane) : 0;             // 32 entries per outer step, one per lane
  for (int k0 = b; k0 < e; k0 += 32) {
    const int next = (k0 + 32 + lane < e) ? __ldg(ent + k0 + 32 + lane) : 0;   // in flight while this step's rows are summed
    for (int k = k0; k < min(e, k0 + 32); k += 8) {
      const int n = min(8, e - k), base = k - k0;
      float4 v[8][QPL];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int x = __shfl_sync(0xffffffffu, mine, (base + u) & 31);
        if (u < n) {
          const float* row = a2.p + (size_t)(x & ((1 << 30) - 1)) * a2.ld + ((x >> 30) ? (H + D) : 0);
#pragma unroll
          for (int q = 0; q < QPL; ++q) v[u][q] = ldg4(row + (lane + 32 * q) * 4);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int x = __shfl_sync(0xffffffffu, mine, (base + u) & 31);
        if (u < n) {
          const bool oside = (x >> 30) != 0;
#pragma unroll
          for (int q = 0; q < QPL; ++q) {
            float4 t = v[u][q];
            if (bn) {
              const float4 sc = oside ? so[q] : ss[q], sh = oside ? ho[q] : hs[q];
              t.x = fmaf(t.x, sc.x, sh.x); t.y = fmaf(t.y, sc.y, sh.y); t.z = fmaf(t.z, sc.z, sh.z); t.w = fmaf(t.w, sc.w, sh.w);
            }
            if (a2.relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
            acc[q].x += t.x; acc[q].y += t.y; acc[q].z += t.z; acc[q].w += t.w;
          }
        }
      }
    }
    mine = next;
  }
  const float ic = (float)max(e - b, 1);                              // = clamp(count, min=1) of graph.py:102-107
#pragma unroll
  for (int q = 0; q < QPL; ++q) {
    float4 o = make_float4(__fdiv_rn(acc[q].x, ic), __fdiv_rn(acc[q].y, ic), __fdiv_rn(acc[q].z, ic), __fdiv_rn(acc[q].w, ic));
    *reinterpret_cast<float4*>(pooled + (size_t)node * H + (lane + 32 * q) * 4) = o;
  }
}

// Launch plan of k_pool_fwd.  V = 4 needs 16-byte aligned rows and halves; anything else takes the scalar-column variant.
// Tuning override (sweeps only): env SLN_POOL="V,npg,batch".
struct PoolPlan { int V, lanes, npg, batch; dim3 grid; size_t smem; };
inline PoolPlan pool_plan(const MatView& a2, int O, int H, int D) {
  static int env_v = -1, env_npg = 0, env_batch = 0;
  if (env_v < 0) {
    env_v = 0;
    if (const char* s = getenv("SLN_POOL")) sscanf(s, "%d,%d,%d", &env_v, &env_npg, &env_batch);
  }
  PoolPlan p;
  const bool aligned = a2.vec && H % 4 == 0 && D % 4 == 0;
  p.V = aligned ? 4 : 1;
  if (aligned && (env_v == 1 || env_v == 2)) p.V = env_v;
  p.batch = env_batch;
  p.lanes = min(256, ceil_div(ceil_div(H, p.V), 32) * 32);
  while (256 % p.lanes) p.lanes += 32;                               // whole groups per CTA
  const int col_blocks = ceil_div(H, p.lanes * p.V);
  // Measured on B200 (tools/bench_pool.py, profiles/): a 32-node stream per group with 4 rows in flight per thread and 4 CTAs
  // per SM is the bandwidth optimum (72-75 % of the HBM peak at O = 262144); small graphs are latency-bound and want one
  // node per group with 16 rows in flight (a 62-entry hub row = 4 round trips).
  const int per_cta = 256 / p.lanes;
  p.npg = env_npg > 0 ? min(env_npg, 32) : max(1, min(32, O / (per_cta * kNumSMs * 2)));
  if (p.npg > 16 && env_npg <= 0) p.npg = 32;
  if (p.batch == 0) p.batch = p.npg == 32 ? 4 : 16;
  p.grid = dim3(ceil_div(ceil_div(O, p.npg), per_cta), col_blocks);
  p.smem = a2.scale ? (size_t)4 * p.lanes * p.V * sizeof(float) : 0;
  return p;
}
inline void launch_pool_fwd(cudaStream_t st, const MatView& a2, const int* row_ptr, const int* ent, int O, int H, int D, float* pooled) {
  static int small_ok = -1;                                            // env SLN_POOL_NODE=0 disables the warp-per-node kernel (A/B measurements)
  if (small_ok < 0) { const char* e = getenv("SLN_POOL_NODE"); small_ok = (e && e[0] == '0') ? 0 : 1; }
  const bool aligned = a2.vec && H % 128 == 0 && D % 4 == 0 && ((uintptr_t)pooled % 16 == 0) &&
                       (!a2.scale || (((uintptr_t)a2.scale | (uintptr_t)a2.shift) % 16 == 0));
  if (small_ok && aligned && H <= 512 && O <= 16 * kNumSMs * 2 && !getenv("SLN_POOL")) {   // every node's warp resident at once (<= 2 CTAs of 8 warps per SM)
    const int grid = ceil_div(O, 8);
    if (H == 128) k_pool_fwd_node<1><<<grid, 256, 0, st>>>(a2, row_ptr, ent, O, H, D, pooled);
    else if (H == 256) k_pool_fwd_node<2><<<grid, 256, 0, st>>>(a2, row_ptr, ent, O, H, D, pooled);
    else if (H == 384) k_pool_fwd_node<3><<<grid, 256, 0, st>>>(a2, row_ptr, ent, O, H, D, pooled);
    else k_pool_fwd_node<4><<<grid, 256, 0, st>>>(a2, row_ptr, ent, O, H, D, pooled);
    return;
  }
  PoolPlan p = pool_plan(a2, O, H, D);
  if (p.V == 4 && p.batch == 16) k_pool_fwd<4, 16, 1><<<p.grid, 256, p.smem, st>>>(a2, row_ptr, ent, O, H, D, pooled, p.lanes, p.npg);
  else if (p.V == 4 && p.batch == 4) k_pool_fwd<4, 4, 4><<<p.grid, 256, p.smem, st>>>(a2, row_ptr, ent, O, H, D, pooled, p.lanes, p.npg);
  else if (p.V == 4) k_pool_fwd<4, 8, 4><<<p.grid, 256, p.smem, st>>>(a2, row_ptr, ent, O, H, D, pooled, p.lanes, p.npg);   // <= 64 registers: 4 CTAs / SM
  else if (p.V == 2) k_pool_fwd<2, 16, 3><<<p.grid, 256, p.smem, st>>>(a2, row_ptr, ent, O, H, D, pooled, p.lanes, p.npg);
  else k_pool_fwd<1, 32, 4><<<p.grid, 256, p.smem, st>>>(a2, row_ptr, ent, O, H, D, pooled, p.lanes, p.npg);
}

// ================================================================ backward "prep" passes
// G[i,j] = mask(y[i,j]) ? src(i,j) : 0 and the BN-backward column sums, for gradients that are assembled by a gather
// rather than produced by a contraction epilogue.
struct PoolBwdSrc {  // gradient w.r.t. a2 = relu(bn(y2)) [T, 2H+D] from d pooled [O,H] and d new_pred [T,D]
  const float* dpooled;  // [O,H]
  const float* cnt;
  const int* s_idx;
  const int* o_idx;
  const float* dpred;  // [T, ldd] slice or null
  int ldd, H, D;
  __device__ __forceinline__ float at(int t, int c) const {
    if (c < H) { int s = __ldg(s_idx + t); return __fdiv_rn(__ldg(dpooled + (size_t)s * H + c), __ldg(cnt + s)); }
    if (c < H + D) return dpred ? __ldg(dpred + (size_t)t * ldd + (c - H)) : 0.f;
    int o = __ldg(o_idx + t);
    return __fdiv_rn(__ldg(dpooled + (size_t)o * H + (c - H - D)), __ldg(cnt + o));
  }
};
struct NodeGatherSrc {  // gradient w.r.t. obj_vecs [O,D] from d cat [T,3D]: transpose of the s/o gathers (graph.py:78-79)
  const float* dcat;
  int ld, D;
  const int* row_ptr;
  const int* ent;
  __device__ __forceinline__ float at(int o, int c) const {
    int b = __ldg(row_ptr + o), e = __ldg(row_ptr + o + 1);
    float acc = 0.f;
    int k = b;
    for (; k + 8 <= e; k += 8) {   // 8 independent (entry -> row) load chains in flight, summed in CSR order
      int en[8];
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) en[u] = __ldg(ent + k + u);
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(dcat + (size_t)(en[u] & ((1 << 30) - 1)) * ld + ((en[u] >> 30) ? 2 * D + c : c));
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; k < e; ++k) {
      int en = __ldg(ent + k);
      int t = en & ((1 << 30) - 1);
      acc += __ldg(dcat + (size_t)t * ld + ((en >> 30) ? 2 * D + c : c));
    }
    return acc;
  }
};
struct PlainSrc {
  const float* p;
  int ld;
  __device__ __forceinline__ float at(int i, int c) const { return __ldg(p + (size_t)i * ld + c); }
};
// the four rows i, i+4, i+8, i+12 (< r1) of column c that one k_prep thread handles per trip
template <class Src>
__device__ __forceinline__ void prep_at4(const Src& src, int i, int r1, int c, float (&d)[4]) {
#pragma unroll
  for (int u = 0; u < 4; ++u) d[u] = i + 4 * u < r1 ? src.at(i + 4 * u, c) : 0.f;
}
// NodeGatherSrc: at() is a loop of dependent (entry -> row) loads, and four calls in a row serialise their chains (measured: 15.6 us
// per launch for 2048 x 128 outputs).  Here the first four entries of all four rows are loaded as 16 independent chains; nodes of
// higher degree (the room node: 62) continue in batches of 16.  Sums run in CSR order exactly as in at() (bit-identical).
__device__ __forceinline__ void prep_at4(const NodeGatherSrc& src, int i, int r1, int c, float (&d)[4]) {
  const int mask30 = (1 << 30) - 1;
  int b[4], n[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const bool ok = i + 4 * u < r1;
    b[u] = ok ? __ldg(src.row_ptr + i + 4 * u) : 0;
    n[u] = ok ? __ldg(src.row_ptr + i + 4 * u + 1) - b[u] : 0;
  }
  int en[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int k = 0; k < 4; ++k) en[u][k] = k < n[u] ? __ldg(src.ent + b[u] + k) : -1;
  float v[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int k = 0; k < 4; ++k)
      v[u][k] = en[u][k] >= 0 ? __ldg(src.dcat + (size_t)(en[u][k] & mask30) * src.ld + ((en[u][k] >> 30) ? 2 * src.D + c : c)) : 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (k < n[u]) acc += v[u][k];
    int k = b[u] + 4;
    const int e = b[u] + n[u];
    for (; k + 16 <= e; k += 16) {
      int e16[16];
      float v16[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) e16[q] = __ldg(src.ent + k + q);
#pragma unroll
      for (int q = 0; q < 16; ++q) v16[q] = __ldg(src.dcat + (size_t)(e16[q] & mask30) * src.ld + ((e16[q] >> 30) ? 2 * src.D + c : c));
#pragma unroll
      for (int q = 0; q < 16; ++q) acc += v16[q];
    }
    for (; k < e; ++k) {
      const int e1 = __ldg(src.ent + k);
      acc += __ldg(src.dcat + (size_t)(e1 & mask30) * src.ld + ((e1 >> 30) ? 2 * src.D + c : c));
    }
    d[u] = acc;
  }
}

struct ActInfo {  // the activation whose input gradient is being formed: a = relu(y*scale+shift)
  int has_act;    // 0: plain copy (no mask, no statistics)
  const float* y;
  int ldy;
  const float* scale;
  const float* shift;
  const float* mean;
  const float* rstd;
};

// grid (ceil(N/128), row_tiles), block (128, 4); each CTA handles `rows_per_tile` rows.
template <class Src>
__global__ void __launch_bounds__(512) k_prep(const Src src, const ActInfo act, float* G, int ldg, BnBwdFin fin, int M, int N,
                                              int rows_per_tile) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_tile;
  const int r1 = min(M, r0 + rows_per_tile);
  float s1 = 0.f, s2 = 0.f;
  if (j < N) {
    float sc = 1.f, sh = 0.f, mu = 0.f, rs = 0.f;
    if (act.has_act && act.scale) { sc = __ldg(act.scale + j); sh = __ldg(act.shift + j); }
    if (act.has_act && act.mean) { mu = __ldg(act.mean + j); rs = __ldg(act.rstd + j); }
    for (int i = r0 + threadIdx.y; i < r1; i += 16) {   // 4 rows per trip: their (dependent) load chains overlap
      float d[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ii = i + 4 * u;
        y[u] = (ii < r1 && act.has_act) ? __ldg(act.y + (size_t)ii * act.ldy + j) : 0.f;
      }
      prep_at4(src, i, r1, j, d);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ii = i + 4 * u;
        if (ii >= r1) break;
        if (act.has_act) {
          float g = fmaf(y[u], sc, sh) > 0.f ? d[u] : 0.f;
          G[(size_t)ii * ldg + j] = g;
          s1 += g;
          s2 = fmaf(g, (y[u] - mu) * rs, s2);
        } else {
          G[(size_t)ii * ldg + j] = d[u];
        }
      }
    }
  }
  if (!act.has_act) return;
  __shared__ float red[2][4][128];
  __shared__ int s_last;
  red[0][threadIdx.y][threadIdx.x] = s1;
  red[1][threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && j < N) {
    float a = red[0][0][threadIdx.x] + red[0][1][threadIdx.x] + red[0][2][threadIdx.x] + red[0][3][threadIdx.x];
    float c = red[1][0][threadIdx.x] + red[1][1][threadIdx.x] + red[1][2][threadIdx.x] + red[1][3][threadIdx.x];
    fin.partial[((size_t)blockIdx.y * 2 + 0) * N + j] = a;
    fin.partial[((size_t)blockIdx.y * 2 + 1) * N + j] = c;
  }
  const int tid = threadIdx.y * 128 + threadIdx.x;
  __shared__ double sred[2 * 512];
  finalize_column_block<512>(fin.partial, fin.counter, blockIdx.x * 128, 128, N, tid, sred, &s_last,
                             [&](int col, double S, double Q, double Sa, double Qa, int Mt) { bn_bwd_apply(fin, col, S, Q, Sa, Qa, Mt); },
                             fin.sync, fin.slot0, fin.M);
}

template <class Src>
int launch_prep(cudaStream_t st, const Src& src, const ActInfo& act, float* G, int ldg, const BnBwdFin& fin, int M, int N,
                const char* what) {
  if (M <= 0 || N <= 0) return SLN_OK;
  // aim for ~2 waves of CTAs; partial buffer is sized for max_row_tiles(M) = ceil(M/16) tiles
  int col_blocks = ceil_div(N, 128);
  static int waves = -1;                                   // env SLN_PREP_WAVES (tuning experiments)
  if (waves < 0) { const char* e = getenv("SLN_PREP_WAVES"); waves = e ? atoi(e) : 2; if (waves < 1) waves = 2; }
  int want_tiles = max(1, (waves * kNumSMs) / col_blocks);
  int rows = max(16, ceil_div(M, want_tiles));
  rows = ceil_div(rows, 4) * 4;
  dim3 grid(col_blocks, ceil_div(M, rows));
  ProfScope prof(st, PROF_PREP, 12.0 * (double)M * (double)N);  // read src + y, write G
  k_prep<Src><<<grid, dim3(128, 4), 0, st>>>(src, act, G, ldg, fin, M, N, rows);
  return check_launch(what);
}

// ================================================================ skinny Linear layers (heads: N or K <= 32)
// The output / input heads of the model (box_net.1: 256 -> 6, angle_net.1: 256 -> 24, angle_mean/var: 128 -> 16,
// box_embeddings: 6 -> 48; reference Sg2ScVAE_model.py:61-104) are far too narrow for 128-wide contraction tiles: they
// are latency problems.  Three small kernels cover them.
//
// forward: out[i, j] = bias[j] + sum_k A(i,k) W[j,k], N <= 32.  W^T is staged in shared memory ([k][32], conflict-free),
// one warp per row (lane = output column), the row of A is read as broadcast float4s through the lazy-BN/ReLU view.
template <class AOp>
__global__ void __launch_bounds__(256) k_skinny_fwd(const AOp A, const float* __restrict__ W, const float* __restrict__ bias, int M, int N, int K,
                                                    float* out, int ldo) {
  extern __shared__ float s_wt[];   // [K][33]: coalesced reads of W (k fastest), transposed writes with an odd row stride (conflict-free)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#pragma unroll 8                                          // 8 weight loads in flight per thread (24 dependent round trips otherwise)
  for (int e = threadIdx.x; e < N * K; e += blockDim.x) {
    int j = e / K, k = e - j * K;
    s_wt[k * 33 + j] = __ldg(W + e);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float b = (bias && lane < N) ? __ldg(bias + lane) : 0.f;
  for (int i = blockIdx.x * 8 + warp; i < M; i += gridDim.x * 8) {
    float acc0 = b, acc1 = 0.f;
    int k = 0;
#pragma unroll 4                                        // 8 row loads in flight per warp (one iteration = one exposed L2 round trip otherwise)
    for (; k + 8 <= K; k += 8) {
      float4 a = A.ld4(i, k), c = A.ld4(i, k + 4);
      acc0 = fmaf(a.x, s_wt[(k + 0) * 33 + lane], acc0); acc1 = fmaf(a.y, s_wt[(k + 1) * 33 + lane], acc1);
      acc0 = fmaf(a.z, s_wt[(k + 2) * 33 + lane], acc0); acc1 = fmaf(a.w, s_wt[(k + 3) * 33 + lane], acc1);
      acc0 = fmaf(c.x, s_wt[(k + 4) * 33 + lane], acc0); acc1 = fmaf(c.y, s_wt[(k + 5) * 33 + lane], acc1);
      acc0 = fmaf(c.z, s_wt[(k + 6) * 33 + lane], acc0); acc1 = fmaf(c.w, s_wt[(k + 7) * 33 + lane], acc1);
    }
    for (; k < K; ++k) acc0 = fmaf(A.at_t(i, k), s_wt[k * 33 + lane], acc0);
    if (lane < N) out[(size_t)i * ldo + lane] = acc0 + acc1;
  }
}
template <class AOp>
int launch_skinny_fwd(cudaStream_t st, const AOp& A, const float* W, const float* bias, int M, int N, int K, float* out, int ldo) {
  if (M <= 0) return SLN_OK;
  size_t smem = (size_t)K * 33 * sizeof(float);
  if (smem > 48 * 1024) {
    static unsigned long long configured = 0ull;   // devices configured (per call site / instantiation)
    if (first_use_on_device(configured)) {
      cudaError_t e = cudaFuncSetAttribute(k_skinny_fwd<AOp>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(k_skinny_fwd) failed: %s", cudaGetErrorString(e)); return SLN_ECUDA; }
    }
  }
  ProfScope prof(st, PROF_GEMM_FWD, 2.0 * (double)M * N * K);
  k_skinny_fwd<AOp><<<min(ceil_div(M, 8), 2 * kNumSMs), 256, smem, st>>>(A, W, bias, M, N, K, out, ldo);
  return check_launch("skinny_fwd");
}

// backward-data with a tiny reduction: src(i, j) = sum_{k < Kt} dy[i, k] W[k, j] (+ add[i, j]), used as a k_prep source so
// that the ReLU mask and the BatchNorm-backward column sums of the producing layer are fused exactly as in the contraction
// epilogue.  dy [M, Kt] plain (a loss gradient), W [Kt, N] = the Linear weight [out, in].
struct SmallKSrc {
  const float* dy; int lddy; int Kt;
  const float* W; int ldw;
  const float* add; int ldadd;
  __device__ __forceinline__ float at(int i, int j) const {
    float acc = add ? __ldg(add + (size_t)i * ldadd + j) : 0.f;
#pragma unroll 8
    for (int k = 0; k < Kt; ++k) acc = fmaf(__ldg(dy + (size_t)i * lddy + k), __ldg(W + (size_t)k * ldw + j), acc);
    return acc;
  }
};

// backward-weight with a tiny output dimension:  C[p, q] += sum_i P[i, p] * Q(i, q),  p < KP <= 32 (P plain, e.g. a loss
// gradient), q < NQ (Q through a lazy view).  Written to Cout[p * ldc + q], or transposed (Cout[q * ldc + p]) for the
// box-embedding weight whose SMALL side is the input.  db (optional) += column sums of P (the Linear bias gradient).
// grid (ceil(NQ/128), row splits), 128 threads = 128 columns q; P rows are staged in shared memory per 32-row slab.
template <int KP, class QOp>
__global__ void __launch_bounds__(128) k_skinny_bwd_w(const float* __restrict__ P, int ldp, int kp, const QOp Q, int M, int NQ, float* C, int ldc,
                                                      int transposed, float* db, int rows_per_cta) {
  __shared__ float s_p[32][KP];
  const int q = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float acc[KP];
#pragma unroll
  for (int e = 0; e < KP; ++e) acc[e] = 0.f;
  float bsum = 0.f;   // thread t < kp of column block 0 accumulates the bias gradient of output t
  for (int base = r0; base < r1; base += 32) {
    const int nr = min(32, r1 - base);
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * KP; e += 128) {
      int r = e / KP, c = e - r * KP;
      s_p[r][c] = (r < nr && c < kp) ? __ldg(P + (size_t)(base + r) * ldp + c) : 0.f;
    }
    __syncthreads();
    if (q < NQ) {
      for (int r = 0; r < nr; r += 4) {        // 4 independent loads in flight
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = r + u < nr ? Q.at_t(base + r + u, q) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int e = 0; e < KP; ++e) acc[e] = fmaf(s_p[(r + u) & 31][e], x[u], acc[e]);
      }
    }
    if (db && blockIdx.x == 0 && threadIdx.x < kp)
      for (int r = 0; r < nr; ++r) bsum += s_p[r][threadIdx.x];
  }
  if (q < NQ) {
#pragma unroll
    for (int e = 0; e < KP; ++e)
      if (e < kp) red_add(transposed ? C + (size_t)q * ldc + e : C + (size_t)e * ldc + q, acc[e]);
  }
  if (db && blockIdx.x == 0 && threadIdx.x < kp) red_add(db + threadIdx.x, bsum);
}
template <class QOp>
int launch_skinny_bwd_w(cudaStream_t st, const float* P, int ldp, int kp, const QOp& Q, int M, int NQ, float* C, int ldc, bool transposed,
                        float* db) {
  if (M <= 0 || (!C && !db)) return SLN_OK;
  if (kp > 32) { set_error("internal: skinny backward-weight needs <= 32 small-side columns (got %d)", kp); return SLN_EINVAL; }
  const int col_blocks = ceil_div(NQ, 128);
  int splits = max(1, min(ceil_div(M, 32), (2 * kNumSMs) / col_blocks));
  int rows = ceil_div(ceil_div(M, splits), 32) * 32;
  dim3 grid(col_blocks, ceil_div(M, rows));
  ProfScope prof(st, PROF_GEMM_BWD_W, 2.0 * (double)M * kp * NQ);
  if (kp <= 8) k_skinny_bwd_w<8, QOp><<<grid, 128, 0, st>>>(P, ldp, kp, Q, M, NQ, C, ldc, transposed ? 1 : 0, db, rows);
  else if (kp <= 16) k_skinny_bwd_w<16, QOp><<<grid, 128, 0, st>>>(P, ldp, kp, Q, M, NQ, C, ldc, transposed ? 1 : 0, db, rows);
  else k_skinny_bwd_w<32, QOp><<<grid, 128, 0, st>>>(P, ldp, kp, Q, M, NQ, C, ldc, transposed ? 1 : 0, db, rows);
  return check_launch("skinny_bwd_w");
}

// scale/shift (+mean/rstd) of an eval-mode BatchNorm from its running statistics
__global__ void k_bn_eval_prep(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rm,
                               const float* __restrict__ rv, float eps, int C, float* mean, float* rstd, float* scale, float* shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float rs = 1.f / sqrtf(rv[c] + eps);
  float sc = gamma[c] * rs;
  mean[c] = rm[c]; rstd[c] = rs; scale[c] = sc; shift[c] = beta[c] - rm[c] * sc;
}

// out[i, c] = view(i, c)   (materialise a lazy BN+ReLU view)
__global__ void k_materialize(const MatView v, float* out, int ldo) {
  int i = blockIdx.x * blockDim.y + threadIdx.y;
  if (i >= v.rows) return;
  for (int c = threadIdx.x * 4; c < v.cols; c += blockDim.x * 4) {
    float4 x = v.ld4(i, c);
    float* d = out + (size_t)i * ldo + c;
    d[0] = x.x; if (c + 1 < v.cols) d[1] = x.y; if (c + 2 < v.cols) d[2] = x.z; if (c + 3 < v.cols) d[3] = x.w;
  }
}

// ================================================================ heads: log-softmax, reparameterisation, losses
// reference Sg2ScVAE_model.py:171 — one warp per row (n_angle = 24 <= 32 columns fast path; generic loop otherwise)
__global__ void k_log_softmax_fwd(const float* __restrict__ logits, int n, int C, float* out) {
  int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (i >= n) return;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, logits[(size_t)i * C + c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(logits[(size_t)i * C + c] - m);
  s = warp_sum(s);
  float lse = m + logf(s);
  for (int c = lane; c < C; c += 32) out[(size_t)i * C + c] = logits[(size_t)i * C + c] - lse;
}
// d logits = d out - exp(out) * sum_c d out
__global__ void k_log_softmax_bwd(const float* __restrict__ dout, const float* __restrict__ out, int n, int C, float* dlogits) {
  int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (i >= n) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += dout[(size_t)i * C + c];
  s = warp_sum(s);
  for (int c = lane; c < C; c += 32) dlogits[(size_t)i * C + c] = dout[(size_t)i * C + c] - expf(out[(size_t)i * C + c]) * s;
}

// z = eps * exp(0.5*logvar) + mu      (reference Sg2ScVAE_model.py:180-183)
__global__ void k_reparam_fwd(const float* __restrict__ mu, const float* __restrict__ logvar, const float* __restrict__ eps, int n, float* z) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] = fmaf(eps[i], expf(0.5f * logvar[i]), mu[i]);
}
// dmu += dz ; dlogvar += dz * eps * 0.5 * exp(0.5*logvar)
__global__ void k_reparam_bwd(const float* __restrict__ dz, const float* __restrict__ logvar, const float* __restrict__ eps, int n,
                              float* dmu, float* dlogvar) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float d = dz[i];
    dmu[i] += d;
    dlogvar[i] += d * eps[i] * 0.5f * expf(0.5f * logvar[i]);
  }
}

// Fused loss forward + gradient seeds (reference utils.py:12-33,139-146):
//   bbox = mean|pred-gt| ; angle = -mean logp[i, gt_i] ; kld = -0.5*sum(1+lv-mu^2-e^lv)/O ; total = bbox + angle + w*kld
// losses[0..3] = {bbox, angle, w*kld, total}.  Deterministic: per-CTA partials, last CTA sums in order.
struct LossArgs {
  const float* boxes_pred; const float* boxes_gt; int BD;
  const float* logp; const long long* angles_gt; int NA;
  const float* mu; const float* logvar; int Z;  // mu == null -> AE mode, no KL
  float kl_weight;
  const float* kl_weight_dev;  // non-null: read the KL weight from device memory (schedules change it without re-capturing a CUDA graph)
  int O;
  int logits_grad;  // 1: d_logits = d total / d logits (log-softmax backward fused); 0: gradient w.r.t. the log-probs
  float* d_boxes; float* d_logits; float* d_mu; float* d_logvar;  // gradient outputs (may be null: loss only)
  float* partial;   // [gridDim.x][3]
  unsigned* counter;
  float* losses;    // [4]
};
__global__ void __launch_bounds__(256) k_vae_loss(const LossArgs a) {
  const int rows_per = 64;
  int r0 = blockIdx.x * rows_per, r1 = min(a.O, r0 + rows_per);
  float l1 = 0.f, nll = 0.f, kl = 0.f;
  const float klw = a.kl_weight_dev ? __ldg(a.kl_weight_dev) : a.kl_weight;
  const float inv_bb = 1.f / ((float)a.O * (float)a.BD), inv_o = 1.f / (float)a.O;
  // inputs and gradient seeds never alias: restrict-qualified views let the unrolled loops keep several loads in flight
  const float* __restrict__ bp = a.boxes_pred; const float* __restrict__ bg = a.boxes_gt; float* __restrict__ dbx = a.d_boxes;
  const float* __restrict__ lpv = a.logp; const long long* __restrict__ ag = a.angles_gt; float* __restrict__ dlg = a.d_logits;
  const float* __restrict__ muv = a.mu; const float* __restrict__ lvv = a.logvar; float* __restrict__ dmu = a.d_mu; float* __restrict__ dlv = a.d_logvar;
#pragma unroll 2
  for (int e = threadIdx.x; e < (r1 - r0) * a.BD; e += blockDim.x) {
    size_t k = (size_t)r0 * a.BD + e;
    float d = bp[k] - bg[k];
    l1 += fabsf(d);
    if (dbx) dbx[k] = (d > 0.f ? inv_bb : (d < 0.f ? -inv_bb : 0.f));
  }
#pragma unroll 4
  for (int e = threadIdx.x; e < (r1 - r0) * a.NA; e += blockDim.x) {
    int i = r0 + e / a.NA, c = e % a.NA;
    float lp = lpv[(size_t)i * a.NA + c];
    bool hit = ((long long)c == ag[i]);
    if (hit) nll -= lp;
    if (dlg) dlg[(size_t)i * a.NA + c] = a.logits_grad ? (expf(lp) - (hit ? 1.f : 0.f)) * inv_o : (hit ? -inv_o : 0.f);
  }
  if (muv) {
#pragma unroll 4
    for (int e = threadIdx.x; e < (r1 - r0) * a.Z; e += blockDim.x) {
      size_t k = (size_t)r0 * a.Z + e;
      float m = muv[k], lv = lvv[k], ex = expf(lv);
      kl += 1.f + lv - m * m - ex;
      if (dmu) { dmu[k] = klw * m * inv_o; dlv[k] = klw * 0.5f * (ex - 1.f) * inv_o; }
    }
  }
  __shared__ float red[3][8];
  __shared__ int s_last;
  l1 = warp_sum(l1); nll = warp_sum(nll); kl = warp_sum(kl);
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = l1; red[1][w] = nll; red[2][w] = kl; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { s0 += red[0][k]; s1 += red[1][k]; s2 += red[2][k]; }
    a.partial[blockIdx.x * 3 + 0] = s0; a.partial[blockIdx.x * 3 + 1] = s1; a.partial[blockIdx.x * 3 + 2] = s2;
    __threadfence();
    unsigned ticket = atomicAdd(a.counter, 1u);
    s_last = (ticket == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double t0 = 0, t1 = 0, t2 = 0;
    for (unsigned b = 0; b < gridDim.x; ++b) {
      t0 += (double)__ldcg(a.partial + b * 3 + 0); t1 += (double)__ldcg(a.partial + b * 3 + 1); t2 += (double)__ldcg(a.partial + b * 3 + 2);
    }
    float bbox = (float)(t0 / ((double)a.O * a.BD));
    float ang = (float)(t1 / (double)a.O);
    float kld = a.mu ? klw * (float)(-0.5 * t2 / (double)a.O) : 0.f;
    a.losses[0] = bbox; a.losses[1] = ang; a.losses[2] = kld; a.losses[3] = bbox + ang + kld;
    *a.counter = 0u;
  }
}

// ================================================================ Adam over one flat parameter arena
// torch.optim.Adam semantics (train.py:15,82-84): m = b1*m+(1-b1)g ; v = b2*v+(1-b2)g^2 ;
// p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps), bc_i = 1 - b_i^step.  `step` lives on the device so the launch is graph-safe.
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                       float lr, float b1, float b2, float eps, float weight_decay, float grad_scale, const long long* step_ptr,
                       const float* lr_dev, const float* guard) {
  // guard: the reference skips backward + step when the loss is not finite (train.py:78-80 "not backpropping"); here the whole
  // update (and k_inc_step) is predicated on the device-side loss so that one NaN step cannot poison parameters or moments
  if (guard != nullptr && !isfinite(__ldg(guard))) return;
  if (lr_dev != nullptr) lr = __ldg(lr_dev);
  long long step = *step_ptr;
  float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 pp = *reinterpret_cast<float4*>(p + i), gg = *reinterpret_cast<const float4*>(g + i);
    float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
    float* P = &pp.x; float* G = &gg.x; float* Mm = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float gr = G[e] * grad_scale + weight_decay * P[e];
      Mm[e] = b1 * Mm[e] + (1.f - b1) * gr;
      V[e] = b2 * V[e] + (1.f - b2) * gr * gr;
      P[e] -= step_size * Mm[e] / (sqrtf(V[e]) * inv_sqrt_bc2 + eps);
    }
    *reinterpret_cast<float4*>(p + i) = pp; *reinterpret_cast<float4*>(m + i) = mm; *reinterpret_cast<float4*>(v + i) = vv;
  } else {
    for (long long k = i; k < n; ++k) {
      float gr = g[k] * grad_scale + weight_decay * p[k];
      m[k] = b1 * m[k] + (1.f - b1) * gr;
      v[k] = b2 * v[k] + (1.f - b2) * gr * gr;
      p[k] -= step_size * m[k] / (sqrtf(v[k]) * inv_sqrt_bc2 + eps);
    }
  }
}
__global__ void k_inc_step(long long* step_ptr, const float* guard) {
  if (guard != nullptr && !isfinite(*guard)) return;
  *step_ptr += 1;
}

}  // namespace sln
