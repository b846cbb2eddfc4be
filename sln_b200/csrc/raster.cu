// Differentiable mesh rasterizer for the layout-refinement path (reference models/diff_render.py:359-366,398 through the
// un-vendored `neural_renderer` package; semantics restated in oracle/raster_oracle.c).
//
// B200 design (not upstream's one-thread-per-pixel loop over ALL faces / one-thread-per-face serial backward):
//   forward   k_project -> k_face_setup (per face: back-face test, pixel-space inverse, conservative pixel bounding box)
//             -> k_raster_tiles: one CTA per 16x16 pixel tile; 4096 faces at a time are culled against the tile (warps ballot
//             32 bounding boxes per step into a shared bit mask, barrier-free, each CTA starting at a different offset of the
//             box array; a block scan turns the mask into the ORDERED id list), then the survivors' records are staged in shared memory 256 at a time and
//             every pixel thread walks them in ascending face order (strict '<' z-test, so ties keep the lower face index
//             exactly as upstream).  Work drops from P*F to P*(faces touching the tile).
//   backward  k_backward_rgb_warp / _cta: a WARP per (face, edge, axis) job, long edges deferred to 32-warp CTAs; lanes stride each d1 sweep,
//             warp-shuffle reduction, one RED.ADD per touched gradient slot; k_backward_depth: per covered pixel.
//   The arithmetic that decides coverage and depth order uses explicit round-to-nearest intrinsics (no FMA contraction)
//   in the oracle's operation order, so face_index maps are bit-identical to the CPU oracle.
#include "../../include/sln_b200.h"
#include "common.cuh"

namespace sln {
namespace {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }

// ------------------------------------------------------------------------------------------------ projection
// nr.projection with the README.md:13-18 patch (no lens distortion): see oracle ro_project.
__global__ void k_project(const float* __restrict__ verts, int V, const float* __restrict__ K, const float* __restrict__ R,
                          const float* __restrict__ t, float orig_size, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const float eps = 1e-9f;
  const float half = dvd(orig_size, 2.0f);
  const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
  const float xc = add(add(add(mul(x, R[0]), mul(y, R[1])), mul(z, R[2])), t[0]);
  const float yc = add(add(add(mul(x, R[3]), mul(y, R[4])), mul(z, R[5])), t[1]);
  const float zc = add(add(add(mul(x, R[6]), mul(y, R[7])), mul(z, R[8])), t[2]);
  const float x_ = dvd(xc, add(zc, eps)), y_ = dvd(yc, add(zc, eps));
  float u = add(add(mul(x_, K[0]), mul(y_, K[1])), K[2]);
  float v = add(add(mul(x_, K[3]), mul(y_, K[4])), K[5]);
  v = sub(orig_size, v);
  u = dvd(mul(2.0f, sub(u, half)), orig_size);
  v = dvd(mul(2.0f, sub(v, half)), orig_size);
  out[3 * i] = u; out[3 * i + 1] = v; out[3 * i + 2] = zc;
}

__device__ __forceinline__ float fix_to_float(long long v);
__global__ void k_project_bwd(const float* __restrict__ verts, int V, const float* __restrict__ K, const float* __restrict__ R,
                              const float* __restrict__ t, float orig_size, const long long* __restrict__ grad_fix, float* __restrict__ grad_verts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const float eps = 1e-9f;
  const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
  const float xc = x * R[0] + y * R[1] + z * R[2] + t[0];
  const float yc = x * R[3] + y * R[4] + z * R[5] + t[1];
  const float zc = x * R[6] + y * R[7] + z * R[8] + t[2];
  const float zi = 1.0f / (zc + eps);
  const float grad_out[3] = {fix_to_float(grad_fix[3 * i]), fix_to_float(grad_fix[3 * i + 1]), fix_to_float(grad_fix[3 * i + 2])};
  const float du = grad_out[0] * (2.0f / orig_size), dv = -grad_out[1] * (2.0f / orig_size);
  const float dx_ = K[0] * du + K[3] * dv, dy_ = K[1] * du + K[4] * dv;
  const float dxc = dx_ * zi, dyc = dy_ * zi;
  const float dzc = grad_out[2] - (dx_ * xc + dy_ * yc) * zi * zi;
  grad_verts[3 * i] = R[0] * dxc + R[3] * dyc + R[6] * dzc;
  grad_verts[3 * i + 1] = R[1] * dxc + R[4] * dyc + R[7] * dzc;
  grad_verts[3 * i + 2] = R[2] * dxc + R[5] * dyc + R[8] * dzc;
}

// ------------------------------------------------------------------------------------------------ per-face setup
__device__ __forceinline__ bool backside(const float* f) {
  return mul(sub(f[7], f[1]), sub(f[3], f[0])) < mul(sub(f[4], f[1]), sub(f[6], f[0]));
}

// fv [F2,9] (vertices_to_faces with fill_back: face F+f = face f reversed), finv [F2,9], fbox [F2] = pixel bounding box
// (x0,x1,y0,y1 inclusive, expanded by one pixel so that rounding in the edge functions can never place a covered pixel
// outside it; x0 > x1 marks a face that can never be drawn: back-facing or off-screen).
__global__ void k_face_setup(const float* __restrict__ pv, const int* __restrict__ faces, int F, int fill_back, int is,
                             float* __restrict__ fv, float* __restrict__ finv, int4* __restrict__ fbox) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int F2 = fill_back ? 2 * F : F;
  if (f >= F2) return;
  const int src = f < F ? f : f - F;
  float face[9];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int vi = faces[3 * src + (f < F ? k : 2 - k)];
    face[3 * k] = pv[3 * vi]; face[3 * k + 1] = pv[3 * vi + 1]; face[3 * k + 2] = pv[3 * vi + 2];
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) fv[9 * (size_t)f + k] = face[k];
  float inv[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) inv[k] = 0.f;
  int4 box = make_int4(1, 0, 1, 0);
  if (!backside(face)) {
    const float fis = (float)is;
    float p[3][2];
#pragma unroll
    for (int n = 0; n < 3; ++n)
#pragma unroll
      for (int d = 0; d < 2; ++d) p[n][d] = mul(0.5f, sub(add(mul(face[3 * n + d], fis), fis), 1.0f));
    inv[0] = sub(p[1][1], p[2][1]); inv[1] = sub(p[2][0], p[1][0]); inv[2] = sub(mul(p[1][0], p[2][1]), mul(p[2][0], p[1][1]));
    inv[3] = sub(p[2][1], p[0][1]); inv[4] = sub(p[0][0], p[2][0]); inv[5] = sub(mul(p[2][0], p[0][1]), mul(p[0][0], p[2][1]));
    inv[6] = sub(p[0][1], p[1][1]); inv[7] = sub(p[1][0], p[0][0]); inv[8] = sub(mul(p[0][0], p[1][1]), mul(p[1][0], p[0][1]));
    const float den = add(add(mul(p[2][0], sub(p[0][1], p[1][1])), mul(p[0][0], sub(p[1][1], p[2][1]))), mul(p[1][0], sub(p[2][1], p[0][1])));
#pragma unroll
    for (int k = 0; k < 9; ++k) inv[k] = dvd(inv[k], den);
    // conservative bounding box in pixel indices: pixel centre xi <-> ndc (2 xi + 1 - is)/is  <=>  xi = p-space coordinate
    float xmin = fminf(fminf(p[0][0], p[1][0]), p[2][0]), xmax = fmaxf(fmaxf(p[0][0], p[1][0]), p[2][0]);
    float ymin = fminf(fminf(p[0][1], p[1][1]), p[2][1]), ymax = fmaxf(fmaxf(p[0][1], p[1][1]), p[2][1]);
    if (xmin == xmin && xmax == xmax && ymin == ymin && ymax == ymax) {   // NaN coordinates never pass the inside test
      float lim = (float)is + 4.f;
      int x0 = (int)floorf(fmaxf(xmin, -4.f)) - 1, x1 = (int)ceilf(fminf(xmax, lim)) + 1;
      int y0 = (int)floorf(fmaxf(ymin, -4.f)) - 1, y1 = (int)ceilf(fminf(ymax, lim)) + 1;
      x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, is - 1); y1 = min(y1, is - 1);
      if (x0 <= x1 && y0 <= y1) box = make_int4(x0, x1, y0, y1);
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) finv[9 * (size_t)f + k] = inv[k];
  fbox[f] = box;
}

// ------------------------------------------------------------------------------------------------ tiled z-buffer
constexpr int TILE = 16;
constexpr int CHUNK = 256;   // faces examined per round = threads per CTA
constexpr int BIG_AREA = 32;     // pixels of the tile a face's box may cover and still be drawn by a single thread
constexpr int LIST_CAP = 4096;   // faces culled per macro round = capacity of the tile's survivor list (LIST_CAP / 32 <= 256 threads)


// TWO: a second z-buffer with its own near plane is kept in the same pass (the reference's depth render clips at the rasterizer
// default near = 0.1, its class renders at the constructor's near = 0.001; both see identical geometry).
template <bool TWO>
__global__ void __launch_bounds__(256, 2) k_raster_tiles(const float* __restrict__ fv, const float* __restrict__ finv, const int4* __restrict__ fbox,
                                                      int F2, int is, float near, float far, int* __restrict__ face_index_map,
                                                      float* __restrict__ weight_map, float* __restrict__ depth_map, float near2,
                                                      int* __restrict__ face_index_map2, float* __restrict__ weight_map2,
                                                      float* __restrict__ depth_map2) {
  __shared__ int s_ids[LIST_CAP];
  __shared__ unsigned s_mask[LIST_CAP / 32];
  __shared__ int s_cnt[8];
  __shared__ unsigned long long s_key[TWO ? 2 : 1][TILE * TILE];   // per pixel: (depth bits << 32) | face id; min = nearest, ties -> lower id
  __shared__ float s_xp[TILE], s_yp[TILE];
  __shared__ int s_big[LIST_CAP];   // faces covering more than BIG_AREA pixels of the tile: drawn pixel-parallel after the small ones
  __shared__ int s_nbig;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE;
  const float fis = (float)is;
  const int tx1 = min(tx0 + TILE - 1, is - 1), ty1 = min(ty0 + TILE - 1, is - 1);
  const int4 none = make_int4(1, 0, 1, 0);
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  const unsigned long long empty = ((unsigned long long)__float_as_uint(far) << 32) | 0xffffffffull;
  s_key[0][tid] = empty;
  if (tid == 0) s_nbig = 0;
  if (TWO) s_key[1][tid] = empty;
  if (tid < TILE) s_xp[tid] = dvd(sub(add(mul(2.0f, (float)(tx0 + tid)), 1.0f), fis), fis);
  else if (tid < 2 * TILE) s_yp[tid - TILE] = dvd(sub(add(mul(2.0f, (float)(ty0 + tid - TILE)), 1.0f), fis), fis);

  for (int base = 0; base < F2; base += LIST_CAP) {
    // ---- cull: which of the next LIST_CAP faces touch this tile?  Every warp tests 32-face groups on its own (one ballot word per
    // group into s_mask, no barrier, loads pipelined), and every CTA starts at a different group of the shared box array.
    const int ng = min(LIST_CAP / 32, (F2 - base + 31) / 32);
    const int rot = (cta * 37) % ng;
#pragma unroll 4
    for (int i = warp; i < ng; i += 8) {
      int g = i + rot; if (g >= ng) g -= ng;
      const int f = base + g * 32 + lane;
      const int4 b = f < F2 ? __ldg(fbox + f) : none;
      const bool hit = b.x <= b.y && b.x <= tx1 && b.y >= tx0 && b.z <= ty1 && b.w >= ty0;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) s_mask[g] = m;
    }
    __syncthreads();
    // compact id list from the masks: exclusive scan of the group counts, then each thread expands its group's bits
    const unsigned mine = tid < ng ? s_mask[tid] : 0u;
    const int c = __popc(mine);
    int x = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_cnt[warp] = x;
    __syncthreads();
    int off = x - c, n = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const int cw = s_cnt[w]; off += (w < warp) ? cw : 0; n += cw; }
    for (unsigned m = mine; m; m &= m - 1) s_ids[off++] = base + tid * 32 + (__ffs(m) - 1);
    __syncthreads();
    // ---- draw, FACE-parallel: a thread owns a surviving face and visits only the pixels of its bounding box inside the tile.
    // Scene meshes project to triangles of a few pixels, hundreds to thousands of them per tile: testing every face at every pixel
    // (the pixel-parallel walk this replaces) spent > 99 % of its tests on misses and made the dense tiles the kernel's tail.
    // The shared 64-bit min picks the nearest face and, on equal depth, the lower face index = upstream's strict '<' in ascending
    // face order; coverage and depth use the same expressions as before, so the maps stay bit-identical to the oracle.
    for (int q = tid; q < n; q += CHUNK) {
      const int f = s_ids[q];
      float face[9], inv[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) { face[k] = __ldg(fv + 9 * (size_t)f + k); inv[k] = __ldg(finv + 9 * (size_t)f + k); }
      const int4 b = __ldg(fbox + f);
      const int x0 = max(b.x, tx0), x1 = min(b.y, tx1), y0 = max(b.z, ty0), y1 = min(b.w, ty1);
      if ((x1 - x0 + 1) * (y1 - y0 + 1) > BIG_AREA) { s_big[atomicAdd(&s_nbig, 1)] = f; continue; }   // order is irrelevant: min-keys
      for (int yi = y0; yi <= y1; ++yi) {
        const float yp = s_yp[yi - ty0], fyi = (float)yi;
        for (int xi = x0; xi <= x1; ++xi) {
          const float xp = s_xp[xi - tx0], fxi = (float)xi;
          if (mul(sub(yp, face[1]), sub(face[3], face[0])) < mul(sub(xp, face[0]), sub(face[4], face[1])) ||
              mul(sub(yp, face[4]), sub(face[6], face[3])) < mul(sub(xp, face[3]), sub(face[7], face[4])) ||
              mul(sub(yp, face[7]), sub(face[0], face[6])) < mul(sub(xp, face[6]), sub(face[1], face[7])))
            continue;
          float w[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) w[k] = add(add(mul(inv[3 * k], fxi), mul(inv[3 * k + 1], fyi)), inv[3 * k + 2]);
          float wsum = 0.f;
#pragma unroll
          for (int k = 0; k < 3; ++k) { w[k] = fminf(fmaxf(w[k], 0.f), 1.f); wsum = add(wsum, w[k]); }
#pragma unroll
          for (int k = 0; k < 3; ++k) w[k] = dvd(w[k], wsum);
          const float zp = dvd(1.0f, add(add(dvd(w[0], face[2]), dvd(w[1], face[5])), dvd(w[2], face[8])));
          if (!(zp < far)) continue;                       // far <= zp, or NaN (never selected by the '<' test either)
          const unsigned long long key = ((unsigned long long)__float_as_uint(zp) << 32) | (unsigned)f;
          const int px = (yi - ty0) * TILE + (xi - tx0);
          if (TWO && !(zp <= near2)) atomicMin(&s_key[1][px], key);
          if (zp <= near) continue;
          atomicMin(&s_key[0][px], key);
        }
      }
    }
    __syncthreads();
    // ---- the large faces of this round (room shell, close-ups), PIXEL-parallel: one thread per pixel walks the short list
    const int nbig = s_nbig;
    if (nbig > 0) {
      const int xi = tx0 + (tid % TILE), yi = ty0 + (tid / TILE);
      const float xp = s_xp[tid % TILE], yp = s_yp[tid / TILE], fxi = (float)xi, fyi = (float)yi;
      unsigned long long best = s_key[0][tid], best2 = TWO ? s_key[1][tid] : 0ull;
      if (xi < is && yi < is) {
        for (int k = 0; k < nbig; ++k) {
          const int f = s_big[k];
          const float* face = fv + 9 * (size_t)f;
          const float f0 = __ldg(face), f1 = __ldg(face + 1), f3 = __ldg(face + 3), f4 = __ldg(face + 4), f6 = __ldg(face + 6), f7 = __ldg(face + 7);
          if (mul(sub(yp, f1), sub(f3, f0)) < mul(sub(xp, f0), sub(f4, f1)) ||
              mul(sub(yp, f4), sub(f6, f3)) < mul(sub(xp, f3), sub(f7, f4)) ||
              mul(sub(yp, f7), sub(f0, f6)) < mul(sub(xp, f6), sub(f1, f7)))
            continue;
          const float* inv = finv + 9 * (size_t)f;
          float w[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) w[j] = add(add(mul(__ldg(inv + 3 * j), fxi), mul(__ldg(inv + 3 * j + 1), fyi)), __ldg(inv + 3 * j + 2));
          float wsum = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) { w[j] = fminf(fmaxf(w[j], 0.f), 1.f); wsum = add(wsum, w[j]); }
#pragma unroll
          for (int j = 0; j < 3; ++j) w[j] = dvd(w[j], wsum);
          const float zp = dvd(1.0f, add(add(dvd(w[0], __ldg(face + 2)), dvd(w[1], __ldg(face + 5))), dvd(w[2], __ldg(face + 8))));
          if (!(zp < far)) continue;
          const unsigned long long key = ((unsigned long long)__float_as_uint(zp) << 32) | (unsigned)f;
          if (TWO && !(zp <= near2) && key < best2) best2 = key;
          if (zp <= near) continue;
          if (key < best) best = key;
        }
      }
      s_key[0][tid] = best;                               // this phase: a pixel is touched by its own thread only
      if (TWO) s_key[1][tid] = best2;
      __syncthreads();
      if (tid == 0) s_nbig = 0;                           // ordered before the next round's draw by that round's cull barriers
    }
  }
  // ---- resolve: the winning face's barycentric weights are recomputed for the pixel (same expressions as in the draw loop)
  const int xi = tx0 + (tid % TILE), yi = ty0 + (tid / TILE);
  if (xi < is && yi < is) {
    const int pn = yi * is + xi;
    const float fxi = (float)xi, fyi = (float)yi;
#pragma unroll
    for (int z = 0; z < (TWO ? 2 : 1); ++z) {
      const unsigned long long key = s_key[z][tid];
      const int f = (int)(unsigned)(key & 0xffffffffull);
      float w[3] = {0.f, 0.f, 0.f};
      float depth = far;
      if (f >= 0) {
        depth = __uint_as_float((unsigned)(key >> 32));
        float wsum = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          w[k] = add(add(mul(__ldg(finv + 9 * (size_t)f + 3 * k), fxi), mul(__ldg(finv + 9 * (size_t)f + 3 * k + 1), fyi)), __ldg(finv + 9 * (size_t)f + 3 * k + 2));
          w[k] = fminf(fmaxf(w[k], 0.f), 1.f); wsum = add(wsum, w[k]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = dvd(w[k], wsum);
      }
      int* fim = z == 0 ? face_index_map : face_index_map2;
      float* wm = z == 0 ? weight_map : weight_map2;
      float* dm = z == 0 ? depth_map : depth_map2;
      fim[pn] = f; dm[pn] = depth;
      wm[3 * pn] = w[0]; wm[3 * pn + 1] = w[1]; wm[3 * pn + 2] = w[2];
    }
  }
}

// ------------------------------------------------------------------------------------------------ texture sampling
__global__ void k_texture_sample(const float* __restrict__ fv, const float* __restrict__ textures, const int* __restrict__ face_index_map,
                                 const float* __restrict__ weight_map, const float* __restrict__ depth_map, int is, int ts, int F_tex,
                                 float eps, float* __restrict__ rgb_map) {
  int pn = blockIdx.x * blockDim.x + threadIdx.x;
  if (pn >= is * is) return;
  float acc[3] = {0.f, 0.f, 0.f};
  const int fi = face_index_map[pn];
  if (fi >= 0) {
    const float* face = fv + 9 * (size_t)fi;
    // fill_back: the back copy (fi >= F_tex) samples the front face's texture with the first and last texel axes swapped
    // (upstream renderer.py: textures.permute(0,1,4,3,2,5))
    const bool back = fi >= F_tex;
    const float* tex = textures + (size_t)(back ? fi - F_tex : fi) * ts * ts * ts * 3;
    const float depth = depth_map[pn];
    float tif[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = mul(mul(weight_map[3 * pn + k], (float)(ts - 1)), dvd(depth, face[3 * k + 2]));
      v = fmaxf(v, 0.f);
      v = fminf(v, sub((float)(ts - 1), eps));
      tif[k] = v;
    }
    for (int pnn = 0; pnn < 8; ++pnn) {
      float w = 1.f;
      int ti[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int base = (int)tif[k];
        if (((pnn >> k) % 2) == 0) { w = mul(w, sub(1.f, sub(tif[k], (float)base))); ti[k] = base; }
        else { w = mul(w, sub(tif[k], (float)base)); ti[k] = base + 1; }
      }
      const int isc = back ? (ti[2] * ts + ti[1]) * ts + ti[0] : (ti[0] * ts + ti[1]) * ts + ti[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[k] = add(acc[k], mul(w, tex[isc * 3 + k]));
    }
  }
  rgb_map[3 * pn] = acc[0]; rgb_map[3 * pn + 1] = acc[1]; rgb_map[3 * pn + 2] = acc[2];
}

// ------------------------------------------------------------------------------------------------ backward: rgb (Kato)
// One warp per (face, edge, axis).  `C` channels per pixel in `img`/`gimg` (3 for an rgb render).  For the fused scene
// path (class != null) the image is implicit: channel value of class c at a pixel = sval[pixel] if cls[face_index] == c
// else 0, and the per-class clamp `diff_grad <= 0 -> skip` is applied per class exactly as 32 separate renders would.
struct RgbBwdArgs {
  const float* fv; const int* face_index_map; int F2, is; float eps;
  const int* fvis;   // [F2] 1 = the face shows at >= 1 pixel (a face that shows nowhere contributes to neither sweep)
  // explicit image mode
  const float* img; const float* gimg; int C;
  // fused class mode
  const int* face_cls; const float* sval; const float* gcls; int n_cls;   // gcls [n_cls, is, is] in internal orientation
  float* grad_faces;
  float* jobg;       // [6 F2][2]: every (face, edge, axis) job's two partial sums, written exactly once; k_rgb_combine adds them
                     // into grad_faces in a fixed order (no floating-point atomics: bit-reproducible gradients)
};

// The reference pixel of a sweep is fixed (the in-pixel for the out sweep, the out-pixel for the in sweep): its class and sample
// are read once per column instead of once per visited pixel.
struct RefPix { int idx, cls; float val; };
__device__ __forceinline__ RefPix ref_pixel(const RgbBwdArgs& a, int idx_ref) {
  RefPix r; r.idx = idx_ref; r.cls = -1; r.val = 0.f;
  if (a.face_cls != nullptr) {
    const int fb = a.face_index_map[idx_ref];
    if (fb >= 0) { r.cls = a.face_cls[fb]; r.val = a.sval[idx_ref]; }
  }
  return r;
}
__device__ __forceinline__ float pix_diff_grad(const RgbBwdArgs& a, int idx, const RefPix& ref) {
  float d = 0.f;
  if (a.face_cls == nullptr) {
    for (int k = 0; k < a.C; ++k) d += (a.img[(size_t)a.C * idx + k] - a.img[(size_t)a.C * ref.idx + k]) * a.gimg[(size_t)a.C * idx + k];
    return d > 0.f ? d : 0.f;
  }
  // fused: only the classes of the two pixels' faces have a non-zero image difference; each class is a separate render
  // (3 identical channels whose gradient is g/3 each), clamped separately
  const int fa = a.face_index_map[idx];
  const float sv = a.sval[idx];
  const int ca = fa >= 0 ? a.face_cls[fa] : -1, cb = ref.cls;
  const float va = fa >= 0 ? sv : 0.f, vb = ref.val;
  const size_t P = (size_t)a.is * a.is;
  const float gb = (cb >= 0 && cb != ca) ? a.gcls[(size_t)cb * P + idx] : 0.f;     // independent of the fa -> ca chain
  float tot = 0.f;
  if (ca >= 0) {
    const float g3 = a.gcls[(size_t)ca * P + idx] * (1.0f / 3.0f);
    const float diff = va - (cb == ca ? vb : 0.f);
    float dg = 0.f;
    for (int k = 0; k < 3; ++k) dg += diff * g3;
    if (dg > 0.f) tot += dg;
  }
  if (cb >= 0 && cb != ca) {
    const float g3 = gb * (1.0f / 3.0f);
    const float diff = 0.f - vb;
    float dg = 0.f;
    for (int k = 0; k < 3; ++k) dg += diff * g3;
    if (dg > 0.f) tot += dg;
  }
  return tot;
}

// Work split of the Kato backward.  A job = (face, edge, axis).  Launch 1: a WARP per job; edges spanning more than kSplitCols pixel
// columns are not processed but appended to a job list.  Launch 2: 1024-thread CTAs walk that list, 32 warps striding the edge's
// columns (room-shell triangles span hundreds of columns, each with a sweep to the image border: as single-warp jobs they were a
// 400 us tail behind 25 us of work).  Per job the warps are summed in a fixed order before the one RED.ADD per gradient slot, so
// the result does not depend on the order of the list.
constexpr int kSplitCols = 4;
constexpr int kBigWarps = 32;
// returns 0 = nothing to add, 1 = g0/g1 hold this warp's partial sums for slots (pi[0], 1-axis), (pi[1], 1-axis), 2 = deferred
template <int WPJ>
__device__ __forceinline__ int rgb_job(const RgbBwdArgs& a, int job, int w_idx, int lane, float& g0, float& g1, int (&pi)[3], int& axis_out) {
  const int fn = job / 6, edge = (job % 6) >> 1, axis = job & 1;
  if (!a.fvis[fn]) return 0;   // both sweeps only add where face_index_map == fn
  const float* face = a.fv + 9 * (size_t)fn;
  if (backside(face)) return 0;
  const int is = a.is;
  const float fis = (float)is;
  float p[3][2];
#pragma unroll
  for (int n = 0; n < 3; ++n) pi[n] = (edge + n) % 3;
#pragma unroll
  for (int n = 0; n < 3; ++n)
#pragma unroll
    for (int d = 0; d < 2; ++d) p[n][d] = mul(0.5f, sub(add(mul(face[3 * pi[n] + ((d + axis) % 2)], fis), fis), 1.0f));
  int direction;
  if (axis == 0) direction = (p[0][0] < p[1][0]) ? -1 : 1;
  else direction = (p[0][0] < p[1][0]) ? 1 : -1;
  const int d0_from = (int)fmaxf(ceilf(fminf(p[0][0], p[1][0])), 0.f);
  const int d0_to = (int)fminf(fmaxf(p[0][0], p[1][0]), fis - 1.f);
  if (WPJ == 1 && d0_to - d0_from + 1 > kSplitCols) return 2;      // deferred to the CTA-per-job launch
  g0 = 0.f; g1 = 0.f;   // gradient of vertex pi[0] / pi[1], component (1 - axis)
  // Triangles are a few pixels wide but the out sweep runs to the image border: the d0 columns of the edge are walked
  // serially (warp-uniform set-up) and the LANES stride each sweep along d1.
  for (int d0 = d0_from + w_idx; d0 <= d0_to; d0 += WPJ) {
    const float fd0 = (float)d0;
    const float d1_cross = add(mul(dvd(sub(p[1][1], p[0][1]), sub(p[1][0], p[0][0])), sub(fd0, p[0][0])), p[0][1]);
    const int d1_in = (0 < direction) ? (int)floorf(d1_cross) : (int)ceilf(d1_cross);
    const int d1_out = d1_in + direction;
    if (d1_in < 0 || is <= d1_in) continue;
    if (d1_out < 0 || is <= d1_out) continue;
    const int idx_in = axis == 0 ? d1_in * is + d0 : d0 * is + d1_in;
    const int idx_out = axis == 0 ? d1_out * is + d0 : d0 * is + d1_out;
    const bool use0 = p[1][0] != fd0, use1 = p[0][0] != fd0;
    const float c0 = use0 ? dvd(sub(p[1][0], p[0][0]), sub(p[1][0], fd0)) : 0.f;
    const float c1 = use1 ? dvd(sub(p[1][0], p[0][0]), sub(fd0, p[0][0])) : 0.f;
    if (a.face_index_map[idx_in] == fn) {   // out sweep: from the out-pixel to the image border
      const int d1_limit = (0 < direction) ? is - 1 : 0;
      const int d1_from = max(min(d1_out, d1_limit), 0), d1_to = min(max(d1_out, d1_limit), is - 1);
      const RefPix ref = ref_pixel(a, idx_in);
      for (int d1 = d1_from + lane; d1 <= d1_to; d1 += 32) {
        const int idx = axis == 0 ? d1 * is + d0 : d0 * is + d1;
        const float dg = pix_diff_grad(a, idx, ref);
        if (dg <= 0.f) continue;
        const float t = sub((float)d1, d1_cross);
        if (use0) { float dist = dvd(mul(mul(c0, t), 2.0f), fis); dist = (0.f < dist) ? dist + a.eps : dist - a.eps; g0 -= dg / dist; }
        if (use1) { float dist = dvd(mul(mul(c1, t), 2.0f), fis); dist = (0.f < dist) ? dist + a.eps : dist - a.eps; g1 -= dg / dist; }
      }
    }
    {   // in sweep: from the in-pixel to the opposite edge crossing, pixels showing this face
      float d0_cross2;
      if (mul(sub(fd0, p[0][0]), sub(fd0, p[2][0])) < 0.f)
        d0_cross2 = add(mul(dvd(sub(p[2][1], p[0][1]), sub(p[2][0], p[0][0])), sub(fd0, p[0][0])), p[0][1]);
      else
        d0_cross2 = add(mul(dvd(sub(p[1][1], p[2][1]), sub(p[1][0], p[2][0])), sub(fd0, p[2][0])), p[2][1]);
      const int d1_limit = (0 < direction) ? (int)ceilf(d0_cross2) : (int)floorf(d0_cross2);
      const int d1_from = max(min(d1_in, d1_limit), 0), d1_to = min(max(d1_in, d1_limit), is - 1);
      const RefPix ref = ref_pixel(a, idx_out);
      for (int d1 = d1_from + lane; d1 <= d1_to; d1 += 32) {
        const int idx = axis == 0 ? d1 * is + d0 : d0 * is + d1;
        if (a.face_index_map[idx] != fn) continue;
        const float dg = pix_diff_grad(a, idx, ref);
        if (dg <= 0.f) continue;
        const float t = sub((float)d1, d1_cross);
        if (use0) { float dist = dvd(mul(mul(c0, t), 2.0f), fis); dist = (0.f < dist) ? dist + a.eps : dist - a.eps; g0 -= dg / dist; }
        if (use1) { float dist = dvd(mul(mul(c1, t), 2.0f), fis); dist = (0.f < dist) ? dist + a.eps : dist - a.eps; g1 -= dg / dist; }
      }
    }
  }
  axis_out = axis;
  return 1;
}

__global__ void __launch_bounds__(256) k_backward_rgb_warp(const RgbBwdArgs a, int* __restrict__ big_jobs) {
  const int job = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (job >= a.F2 * 6) return;
  float g0, g1; int pi[3], axis;
  const int r = rgb_job<1>(a, job, 0, lane, g0, g1, pi, axis);
  if (r == 2) { if (lane == 0) big_jobs[1 + atomicAdd(big_jobs, 1)] = job; return; }   // (integer atomic: the ORDER of the list does not matter)
  if (r == 0) { g0 = 0.f; g1 = 0.f; }
  g0 = warp_sum(g0); g1 = warp_sum(g1);
  if (lane == 0) { a.jobg[2 * (size_t)job] = g0; a.jobg[2 * (size_t)job + 1] = g1; }
}

__global__ void __launch_bounds__(32 * kBigWarps) k_backward_rgb_cta(const RgbBwdArgs a, const int* __restrict__ big_jobs) {
  __shared__ float sh[2][kBigWarps];
  const int lane = threadIdx.x & 31, w_idx = threadIdx.x >> 5;
  const int n = big_jobs[0];
  for (int li = blockIdx.x; li < n; li += gridDim.x) {
    const int job = big_jobs[1 + li];
    float g0, g1; int pi[3], axis;
    const int r = rgb_job<kBigWarps>(a, job, w_idx, lane, g0, g1, pi, axis);   // CTA-uniform return value
    if (r != 1) { g0 = 0.f; g1 = 0.f; }
    g0 = warp_sum(g0); g1 = warp_sum(g1);
    if (lane == 0) { sh[0][w_idx] = g0; sh[1][w_idx] = g1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float t0 = 0.f, t1 = 0.f;
      for (int w = 0; w < kBigWarps; ++w) { t0 += sh[0][w]; t1 += sh[1][w]; }
      a.jobg[2 * (size_t)job] = t0; a.jobg[2 * (size_t)job + 1] = t1;
    }
    __syncthreads();
  }
}
// grad_faces[fn, k, c] += partial of job (fn, edge k, axis 1-c) for vertex pi[0] = k  +  partial of job (fn, edge k+2, axis 1-c) for
// vertex pi[1] = k: the only two jobs that touch the slot, added in this order.
__global__ void k_rgb_combine(const float* __restrict__ jobg, int F2, float* __restrict__ grad_faces) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;     // (fn, k, c)
  if (i >= F2 * 6) return;
  const int fn = i / 6, k = (i % 6) >> 1, c = i & 1, axis = 1 - c;
  const float a0 = jobg[2 * ((size_t)fn * 6 + k * 2 + axis)], a1 = jobg[2 * ((size_t)fn * 6 + ((k + 2) % 3) * 2 + axis) + 1];
  const float t = a0 + a1;
  if (t != 0.f) grad_faces[9 * (size_t)fn + 3 * k + c] += t;
}
__global__ void k_mark_visible(const int* __restrict__ face_index_map, int P, int F2, int* __restrict__ fvis) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int f = face_index_map[p];
  if (f >= 0 && f < F2) fvis[f] = 1;   // benign race: every writer stores the same value
}
inline int launch_backward_rgb(RgbBwdArgs a, int* fvis, int* big_jobs, float* jobg, cudaStream_t st, const char* what) {
  a.jobg = jobg;
  const int P = a.is * a.is;
  // fvis [F2] and the job counter big_jobs[0] are adjacent in the workspace plan: one memset clears both
  cudaError_t e = cudaMemsetAsync(fvis, 0, ((size_t)a.F2 + 1) * sizeof(int), st);
  if (e != cudaSuccess) { set_error("%s: cudaMemsetAsync failed: %s", what, cudaGetErrorString(e)); return SLN_ECUDA; }
  k_mark_visible<<<ceil_div(P, 256), 256, 0, st>>>(a.face_index_map, P, a.F2, fvis);
  SLN_TRY(check_launch(what));
  a.fvis = fvis;
  k_backward_rgb_warp<<<ceil_div(a.F2 * 6, 8), 256, 0, st>>>(a, big_jobs);
  SLN_TRY(check_launch(what));
  k_backward_rgb_cta<<<min(a.F2 * 6, 4 * kNumSMs), 32 * kBigWarps, 0, st>>>(a, big_jobs);
  SLN_TRY(check_launch(what));
  k_rgb_combine<<<ceil_div(a.F2 * 6, 256), 256, 0, st>>>(jobg, a.F2, a.grad_faces);
  return check_launch(what);
}

// ------------------------------------------------------------------------------------------------ backward: depth
// One warp per face gathers the pixels that show the face inside its pixel box (lanes stride the box row-major, fixed-order
// shuffle tree): a single writer per face, no floating-point atomics.  Per pixel the arithmetic is the per-pixel scatter of the
// published kernel (d depth / d z_k through the barycentric weights, d depth / d xy through the pixel-space inverse).
constexpr int kDepthBigPixels = 256;      // pixel boxes above this are gathered by a whole CTA (room-shell faces can span the image)
struct DepthBwdArgs {
  const float* fv; const float* finv; const int4* fbox; const float* depth_map; const int* face_index_map; const float* weight_map;
  const float* grad_depth_map; int is, F2; float* grad_faces;
};
// partial sums of face fn over the pixels i = first, first + stride, ... of its box (row-major): 9 accumulators per thread
__device__ __forceinline__ bool depth_face_partial(const DepthBwdArgs& a, int fn, const int4 bx, int first, int stride, float (&acc)[9]) {
  const float* face = a.fv + 9 * (size_t)fn;
  const float* fi = a.finv + 9 * (size_t)fn;
  float tmp[2] = {0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int l = 0; l < 3; ++l) tmp[k] += -fi[3 * l + k] / face[3 * l + 2];
  const float fis = (float)a.is;
#pragma unroll
  for (int j = 0; j < 9; ++j) acc[j] = 0.f;
  const int w = bx.y - bx.x + 1, n = w * (bx.w - bx.z + 1);
  bool any = false;
  for (int i = first; i < n; i += stride) {
    const int pn = (bx.z + i / w) * a.is + bx.x + i % w;
    if (a.face_index_map[pn] != fn) continue;
    const float g = a.grad_depth_map[pn];
    if (g == 0.f) continue;
    any = true;
    const float depth = a.depth_map[pn], depth2 = depth * depth;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float wk = a.weight_map[3 * pn + k], zk = face[3 * k + 2];
      acc[3 * k + 2] += g * wk * depth2 / (zk * zk);
#pragma unroll
      for (int l = 0; l < 2; ++l) acc[3 * k + l] += -g * tmp[l] * wk * depth2 * fis / 2.f;
    }
  }
  return any;
}
__global__ void __launch_bounds__(256) k_backward_depth(const DepthBwdArgs a, int* __restrict__ big_list) {
  const int fn = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (fn >= a.F2) return;
  const int4 bx = a.fbox[fn];                                // x0, x1, y0, y1 inclusive; x0 > x1: never drawn
  if (bx.x > bx.y || bx.z > bx.w) return;
  if ((bx.y - bx.x + 1) * (bx.w - bx.z + 1) > kDepthBigPixels) {   // deferred to the CTA-per-face launch (integer atomic: list order is irrelevant)
    if (lane == 0) big_list[1 + atomicAdd(big_list, 1)] = fn;
    return;
  }
  float acc[9];
  const bool any = depth_face_partial(a, fn, bx, lane, 32, acc);
  if (!__any_sync(0xffffffffu, any)) return;
#pragma unroll
  for (int j = 0; j < 9; ++j) acc[j] = warp_sum(acc[j]);
  if (lane == 0) {
    float* gf = a.grad_faces + 9 * (size_t)fn;
#pragma unroll
    for (int j = 0; j < 9; ++j) if (acc[j] != 0.f) gf[j] += acc[j];
  }
}
__global__ void __launch_bounds__(1024) k_backward_depth_cta(const DepthBwdArgs a, const int* __restrict__ big_list) {
  __shared__ float sh[32][9];
  const int lane = threadIdx.x & 31, w_idx = threadIdx.x >> 5;
  const int n = big_list[0];
  for (int li = blockIdx.x; li < n; li += gridDim.x) {
    const int fn = big_list[1 + li];
    float acc[9];
    depth_face_partial(a, fn, a.fbox[fn], threadIdx.x, 1024, acc);
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 9; ++j) sh[w_idx][j] = acc[j];
    }
    __syncthreads();
    if (threadIdx.x < 9) {
      float t = 0.f;
      for (int w = 0; w < 32; ++w) t += sh[w][threadIdx.x];
      if (t != 0.f) a.grad_faces[9 * (size_t)fn + threadIdx.x] += t;
    }
    __syncthreads();
  }
}

// grad_fv [F2,9] -> grad_pv [V,3]  (transpose of vertices_to_faces incl. the fill_back copies).  A vertex is shared by a handful of
// faces whose ids are arbitrary, so the scatter stays — but into 64-bit FIXED-POINT accumulators (2^-32 resolution, +-2.1e9 range):
// integer addition is associative, the result does not depend on the order in which the atomics land.
constexpr double kFixScale = 4294967296.0;   // 2^32
__device__ __forceinline__ float fix_to_float(long long v) { return (float)((double)v * (1.0 / kFixScale)); }
__global__ void k_faces_to_vertices_bwd(const float* __restrict__ grad_fv, const int* __restrict__ faces, int F, int fill_back, long long* grad_pv) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;   // one (face, corner) per thread
  const int F2 = fill_back ? 2 * F : F;
  if (e >= F2 * 3) return;
  const int f = e / 3, k = e % 3;
  const int src = f < F ? f : f - F;
  const int vi = faces[3 * src + (f < F ? k : 2 - k)];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float g = grad_fv[9 * (size_t)f + 3 * k + d];
    if (g != 0.f) {
      g = fminf(fmaxf(g, -2.0e9f), 2.0e9f);
      atomicAdd(reinterpret_cast<unsigned long long*>(grad_pv + 3 * (size_t)vi + d), (unsigned long long)__double2ll_rn((double)g * kFixScale));
    }
  }
}

// ------------------------------------------------------------------------------------------------ fused scene passes
// Per covered pixel: s = what a constant-1 texture samples to (the sum of the 8 trilinear weights, evaluated exactly as
// k_texture_sample does), so class image c = s where cls[face] == c, else 0 — bit-identical to 32 separate renders.
__global__ void k_scene_sval(const float* __restrict__ fv, const int* __restrict__ face_index_map, const float* __restrict__ weight_map,
                             const float* __restrict__ depth_map, int is, int ts, float eps, float* __restrict__ sval) {
  int pn = blockIdx.x * blockDim.x + threadIdx.x;
  if (pn >= is * is) return;
  const int fi = face_index_map[pn];
  float acc = 0.f;
  if (fi >= 0) {
    const float* face = fv + 9 * (size_t)fi;
    const float depth = depth_map[pn];
    float tif[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = mul(mul(weight_map[3 * pn + k], (float)(ts - 1)), dvd(depth, face[3 * k + 2]));
      v = fmaxf(v, 0.f);
      v = fminf(v, sub((float)(ts - 1), eps));
      tif[k] = v;
    }
    for (int pnn = 0; pnn < 8; ++pnn) {
      float w = 1.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int base = (int)tif[k];
        if (((pnn >> k) % 2) == 0) w = mul(w, sub(1.f, sub(tif[k], (float)base)));
        else w = mul(w, sub(tif[k], (float)base));
      }
      acc = add(acc, mul(w, 1.0f));
    }
  }
  sval[pn] = acc;
}

// class images [n_cls, is, is] in OUTPUT orientation (vertically flipped): value = (s+s+s)*(1/3) as torch.sum(images,1)/3.0 on CUDA
__global__ void k_scene_class_images(const int* __restrict__ face_index_map, const int* __restrict__ face_cls, const float* __restrict__ sval,
                                     int is, int n_cls, float* __restrict__ images) {
  int pn = blockIdx.x * blockDim.x + threadIdx.x;
  const int P = is * is;
  if (pn >= P) return;
  const int yi = pn / is, xi = pn % is;
  const int fi = face_index_map[pn];
  const int c = fi >= 0 ? face_cls[fi] : -1;
  const float s = sval[pn];
  // torch evaluates `tensor / 3.0` on CUDA as a multiplication by the fp32 reciprocal (ATen div_true_kernel_cuda with a
  // CPU-scalar divisor), and the reference's renderer only exists on CUDA, so that is the arithmetic to reproduce
  const float v = mul(add(add(s, s), s), 1.0f / 3.0f);
  const size_t o = (size_t)(is - 1 - yi) * is + xi;
  for (int k = 0; k < n_cls; ++k) images[(size_t)k * P + o] = (k == c) ? v : 0.f;
}

}  // namespace
}  // namespace sln

using namespace sln;

// workspace layout (all 256-byte aligned): pv [V,3] | fv [F2,9] | finv [F2,9] | fbox [F2] int4
namespace {
struct RasterPlan { float* pv; float* fv; float* finv; int4* fbox; int* fvis; float* jobg; size_t bytes; };
RasterPlan plan_raster(void* ws, int64_t V, int64_t F2) {
  Arena ar(ws, (size_t)-1);
  RasterPlan p;
  p.pv = ar.take<float>(3 * (size_t)V);
  p.fv = ar.take<float>(9 * (size_t)F2);
  p.finv = ar.take<float>(9 * (size_t)F2);
  p.fbox = ar.take<int4>((size_t)F2);
  p.fvis = ar.take<int>((size_t)F2 + 1 + 6 * (size_t)F2);   // backward scratch: fvis [F2] (does the face own a pixel?) | big_jobs [1 + 6 F2] (count, job ids)
  p.jobg = ar.take<float>(12 * (size_t)F2);                 // backward scratch: the two partial sums of every (face, edge, axis) job
  p.bytes = ar.off;
  return p;
}
int check_raster_args(int64_t V, int64_t F, int32_t is) {
  SLN_CHECK_ARG(V >= 1 && V < (1ll << 28) && F >= 0 && F < (1ll << 27), "vertex/face count out of range");
  SLN_CHECK_ARG(is >= 1 && is <= 8192, "image_size out of range");
  return SLN_OK;
}
}  // namespace

extern "C" {

size_t sln_raster_workspace_bytes(int64_t V, int64_t F, int32_t fill_back) {
  return plan_raster(nullptr, V, fill_back ? 2 * F : F).bytes;
}

int sln_raster_setup(const float* vertices, int64_t V, const int32_t* faces, int64_t F, int32_t fill_back, const float* K, const float* R,
                     const float* t, float orig_size, int32_t image_size, void* ws, size_t ws_bytes, void* stream) {
  SLN_TRY(check_raster_args(V, F, image_size));
  SLN_CHECK_ARG(vertices && (faces || F == 0) && K && R && t && ws, "null pointer");
  const int64_t F2 = fill_back ? 2 * F : F;
  RasterPlan p = plan_raster(ws, V, F2);
  SLN_CHECK_ARG((uintptr_t)ws % 256 == 0 && ws_bytes >= p.bytes, "raster workspace misaligned or too small");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(st, PROF_RASTER_FWD, 12.0 * V + 12.0 * F + 88.0 * F2);
  k_project<<<ceil_div((int)V, 256), 256, 0, st>>>(vertices, (int)V, K, R, t, orig_size, p.pv);
  SLN_TRY(check_launch("project"));
  if (F2 > 0) {
    k_face_setup<<<ceil_div((int)F2, 256), 256, 0, st>>>(p.pv, faces, (int)F, fill_back, image_size, p.fv, p.finv, p.fbox);
    SLN_TRY(check_launch("face_setup"));
  }
  return SLN_OK;
}

int sln_raster_face_arrays(void* ws, int64_t V, int64_t F, int32_t fill_back, const float** proj_vertices, const float** face_vertices,
                           const float** face_inv) {
  RasterPlan p = plan_raster(ws, V, fill_back ? 2 * F : F);
  if (proj_vertices) *proj_vertices = p.pv;
  if (face_vertices) *face_vertices = p.fv;
  if (face_inv) *face_inv = p.finv;
  return SLN_OK;
}

int sln_raster_forward(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, float near, float far,
                       int32_t* face_index_map, float* weight_map, float* depth_map, void* stream) {
  SLN_TRY(check_raster_args(V, F, image_size));
  SLN_CHECK_ARG(ws && face_index_map && weight_map && depth_map, "null pointer");
  const int64_t F2 = fill_back ? 2 * F : F;
  RasterPlan p = plan_raster((void*)ws, V, F2);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = ceil_div(image_size, TILE);
  ProfScope prof(st, PROF_RASTER_FWD, 88.0 * F2 + 20.0 * image_size * image_size);
  k_raster_tiles<false><<<dim3(tiles, tiles), 256, 0, st>>>(p.fv, p.finv, p.fbox, (int)F2, image_size, near, far, face_index_map, weight_map, depth_map,
                                                             0.f, nullptr, nullptr, nullptr);
  return check_launch("raster_tiles");
}

int sln_raster_forward2(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, float near_a, float near_b, float far,
                        int32_t* face_index_map_a, float* weight_map_a, float* depth_map_a, int32_t* face_index_map_b, float* weight_map_b,
                        float* depth_map_b, void* stream) {
  SLN_TRY(check_raster_args(V, F, image_size));
  SLN_CHECK_ARG(ws && face_index_map_a && weight_map_a && depth_map_a && face_index_map_b && weight_map_b && depth_map_b, "null pointer");
  const int64_t F2 = fill_back ? 2 * F : F;
  RasterPlan p = plan_raster((void*)ws, V, F2);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = ceil_div(image_size, TILE);
  ProfScope prof(st, PROF_RASTER_FWD, 88.0 * F2 + 40.0 * image_size * image_size);
  k_raster_tiles<true><<<dim3(tiles, tiles), 256, 0, st>>>(p.fv, p.finv, p.fbox, (int)F2, image_size, near_a, far, face_index_map_a, weight_map_a,
                                                            depth_map_a, near_b, face_index_map_b, weight_map_b, depth_map_b);
  return check_launch("raster_tiles2");
}

int sln_raster_texture_sample(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, const float* textures,
                              int32_t texture_size, float eps, const int32_t* face_index_map, const float* weight_map,
                              const float* depth_map, float* rgb_map, void* stream) {
  SLN_TRY(check_raster_args(V, F, image_size));
  SLN_CHECK_ARG(ws && textures && face_index_map && weight_map && depth_map && rgb_map && texture_size >= 2, "bad argument");
  RasterPlan p = plan_raster((void*)ws, V, fill_back ? 2 * F : F);
  const int P = image_size * image_size;
  ProfScope prof((cudaStream_t)stream, PROF_RASTER_FWD, 32.0 * P);
  k_texture_sample<<<ceil_div(P, 256), 256, 0, (cudaStream_t)stream>>>(p.fv, textures, face_index_map, weight_map, depth_map, image_size,
                                                                      texture_size, (int)F, eps, rgb_map);
  return check_launch("texture_sample");
}

int sln_raster_backward_rgb(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, float eps,
                            const int32_t* face_index_map, const float* rgb_map, const float* grad_rgb_map, float* grad_faces, void* stream) {
  SLN_TRY(check_raster_args(V, F, image_size));
  SLN_CHECK_ARG(ws && face_index_map && rgb_map && grad_rgb_map && grad_faces, "null pointer");
  const int64_t F2 = fill_back ? 2 * F : F;
  if (F2 == 0) return SLN_OK;
  RasterPlan p = plan_raster((void*)ws, V, F2);
  RgbBwdArgs a; memset(&a, 0, sizeof(a));
  a.fv = p.fv; a.face_index_map = face_index_map; a.F2 = (int)F2; a.is = image_size; a.eps = eps;
  a.img = rgb_map; a.gimg = grad_rgb_map; a.C = 3; a.grad_faces = grad_faces;
  ProfScope prof((cudaStream_t)stream, PROF_RASTER_BWD, 36.0 * F2 + 28.0 * image_size * image_size);
  return launch_backward_rgb(a, p.fvis, p.fvis + F2, p.jobg, (cudaStream_t)stream, "backward_rgb");
}

int sln_raster_backward_depth(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, const int32_t* face_index_map,
                              const float* weight_map, const float* depth_map, const float* grad_depth_map, float* grad_faces, void* stream) {
  SLN_TRY(check_raster_args(V, F, image_size));
  SLN_CHECK_ARG(ws && face_index_map && weight_map && depth_map && grad_depth_map && grad_faces, "null pointer");
  RasterPlan p = plan_raster((void*)ws, V, fill_back ? 2 * F : F);
  const int P = image_size * image_size;
  ProfScope prof((cudaStream_t)stream, PROF_RASTER_BWD, 24.0 * P);
  const int F2 = (int)(fill_back ? 2 * F : F);
  (void)P;
  if (F2 == 0) return SLN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int* big_list = p.fvis + F2;                               // the deferred-job list of the workspace plan (count, ids): shared with the rgb pass, used in turn
  SLN_CUDA_TRY(cudaMemsetAsync(big_list, 0, sizeof(int), st));
  DepthBwdArgs a{p.fv, p.finv, p.fbox, depth_map, face_index_map, weight_map, grad_depth_map, image_size, F2, grad_faces};
  k_backward_depth<<<ceil_div(F2, 8), 256, 0, st>>>(a, big_list);
  SLN_TRY(check_launch("backward_depth"));
  k_backward_depth_cta<<<min(F2, 2 * kNumSMs), 1024, 0, st>>>(a, big_list);
  return check_launch("backward_depth_cta");
}

int sln_raster_vertex_grad(const void* ws, const float* vertices, int64_t V, const int32_t* faces, int64_t F, int32_t fill_back,
                           const float* K, const float* R, const float* t, float orig_size, const float* grad_faces, float* grad_proj_scratch,
                           float* grad_vertices, void* stream) {
  SLN_CHECK_ARG(ws && vertices && (faces || F == 0) && K && R && t && grad_faces && grad_proj_scratch && grad_vertices, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t F2 = fill_back ? 2 * F : F;
  ProfScope prof(st, PROF_RASTER_BWD, 36.0 * F2 + 36.0 * V);
  SLN_CHECK_ARG((uintptr_t)grad_proj_scratch % 8 == 0, "grad_proj_scratch must be 8-byte aligned (24 V bytes of fixed-point accumulators)");
  long long* acc = reinterpret_cast<long long*>(grad_proj_scratch);
  SLN_CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(long long) * 3 * (size_t)V, st));
  if (F2 > 0) {
    k_faces_to_vertices_bwd<<<ceil_div((int)F2 * 3, 256), 256, 0, st>>>(grad_faces, faces, (int)F, fill_back, acc);
    SLN_TRY(check_launch("faces_to_vertices_bwd"));
  }
  k_project_bwd<<<ceil_div((int)V, 256), 256, 0, st>>>(vertices, (int)V, K, R, t, orig_size, acc, grad_vertices);
  return check_launch("project_bwd");
}

int sln_scene_classes_fwd(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, int32_t texture_size, float eps,
                          const int32_t* face_index_map, const float* weight_map, const float* depth_map, const int32_t* face_cls,
                          int32_t n_cls, float* sval, float* class_images, void* stream) {
  SLN_TRY(check_raster_args(V, F, image_size));
  SLN_CHECK_ARG(ws && face_index_map && weight_map && depth_map && face_cls && sval && class_images && n_cls >= 1, "bad argument");
  RasterPlan p = plan_raster((void*)ws, V, fill_back ? 2 * F : F);
  cudaStream_t st = (cudaStream_t)stream;
  const int P = image_size * image_size;
  ProfScope prof(st, PROF_RASTER_FWD, (28.0 + 4.0 * n_cls) * P);
  k_scene_sval<<<ceil_div(P, 256), 256, 0, st>>>(p.fv, face_index_map, weight_map, depth_map, image_size, texture_size, eps, sval);
  SLN_TRY(check_launch("scene_sval"));
  k_scene_class_images<<<ceil_div(P, 256), 256, 0, st>>>(face_index_map, face_cls, sval, image_size, n_cls, class_images);
  return check_launch("scene_class_images");
}

int sln_scene_classes_bwd(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, float eps,
                          const int32_t* face_index_map, const int32_t* face_cls, int32_t n_cls, const float* sval,
                          const float* grad_class_images_internal, float* grad_faces, void* stream) {
  SLN_TRY(check_raster_args(V, F, image_size));
  SLN_CHECK_ARG(ws && face_index_map && face_cls && sval && grad_class_images_internal && grad_faces, "null pointer");
  const int64_t F2 = fill_back ? 2 * F : F;
  if (F2 == 0) return SLN_OK;
  RasterPlan p = plan_raster((void*)ws, V, F2);
  RgbBwdArgs a; memset(&a, 0, sizeof(a));
  a.fv = p.fv; a.face_index_map = face_index_map; a.F2 = (int)F2; a.is = image_size; a.eps = eps;
  a.face_cls = face_cls; a.sval = sval; a.gcls = grad_class_images_internal; a.n_cls = n_cls; a.grad_faces = grad_faces;
  ProfScope prof((cudaStream_t)stream, PROF_RASTER_BWD, 36.0 * F2 + (12.0 + 4.0 * n_cls) * image_size * image_size);
  return launch_backward_rgb(a, p.fvis, p.fvis + F2, p.jobg, (cudaStream_t)stream, "scene_backward_rgb");
}

}  // extern "C"
