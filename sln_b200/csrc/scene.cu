// Scene assembly of the differentiable-render path (SURVEY §8 a10 stages i-iii, §8f N3).
// Reference: models/diff_render.py:76-159 (per object: scale = min(box size / model size), Ry(theta) with theta = -angle*2pi/24,
// trans = box centre - scale * R * model centre, vertices = scale * R * v + trans — a Python loop over the objects with a dozen
// small torch ops each) and :344-356 (faces with a vertex closer than 0.06 to the camera plane are dropped).
// Here the meshes stay resident on the device (one flat vertex array, objects contiguous) and a layout update is
//   forward   k_obj_params (per object)  ->  k_assemble_vertices (per vertex)  ->  k_cull_faces (per face; a culled face keeps
//             its slot as the zero-area triangle (0,0,0), so shapes are static and the iteration is CUDA-graph capturable)
//   backward  k_assemble_bwd: one CTA per layout row, fixed-order block reduction over the object's vertices (no atomics),
//             then the chain rule to the box corners and the angle (the min() routes to the limiting axis, as torch.min does).
#include "../../include/sln_b200.h"
#include "common.cuh"

namespace sln {
namespace {

struct ObjParams { float s, c, sn, cx, cy, cz; int kmin; int pad; };
constexpr float kAngleStep = 6.283185307179586f / 24.0f;   // 2*pi/24 (diff_render.py:84)

__global__ void k_obj_params(const float* __restrict__ boxes, const float* __restrict__ angles, const int* __restrict__ kept, int n_kept,
                             float rx, float ry, float rz, const float* __restrict__ msize, ObjParams* __restrict__ prm,
                             float* __restrict__ size_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_kept) return;
  const int r = kept[j];
  const float room[3] = {rx, ry, rz};
  float size[3], cen[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float lo = boxes[6 * r + k] * room[k], hi = boxes[6 * r + 3 + k] * room[k];
    cen[k] = (hi + lo) / 2.f; size[k] = hi - lo;
    size_out[3 * j + k] = size[k];
  }
  float s = size[0] / msize[3 * j]; int kmin = 0;
#pragma unroll
  for (int k = 1; k < 3; ++k) { const float q = size[k] / msize[3 * j + k]; if (q < s) { s = q; kmin = k; } }
  const float theta = -angles[r] * kAngleStep;
  ObjParams p; p.s = s; p.c = cosf(theta); p.sn = sinf(theta); p.cx = cen[0]; p.cy = cen[1]; p.cz = cen[2]; p.kmin = kmin; p.pad = 0;
  prm[j] = p;
}

// R = [[c,0,s],[0,1,0],[-s,0,c]]  (diff_render.py:85-90)
__device__ __forceinline__ void rot_y(const ObjParams& p, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = p.c * x + p.sn * z; oy = y; oz = -p.sn * x + p.c * z;
}

__global__ void __launch_bounds__(256) k_assemble_vertices(const float* __restrict__ mv, const int* __restrict__ vobj, int n_obj_verts,
                                                           const float* __restrict__ shell_v, int n_shell, const float* __restrict__ mcent,
                                                           const ObjParams* __restrict__ prm, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_obj_verts + n_shell) return;
  if (i >= n_obj_verts) {
    const int k = i - n_obj_verts;
    out[3 * i] = shell_v[3 * k]; out[3 * i + 1] = shell_v[3 * k + 1]; out[3 * i + 2] = shell_v[3 * k + 2];
    return;
  }
  const int j = vobj[i];
  const ObjParams p = prm[j];
  float wx, wy, wz, mx, my, mz;
  rot_y(p, mv[3 * i], mv[3 * i + 1], mv[3 * i + 2], wx, wy, wz);
  rot_y(p, mcent[3 * j], mcent[3 * j + 1], mcent[3 * j + 2], mx, my, mz);
  out[3 * i] = p.s * wx + (p.cx - p.s * mx);
  out[3 * i + 1] = p.s * wy + (p.cy - p.s * my);
  out[3 * i + 2] = p.s * wz + (p.cz - p.s * mz);
}

__global__ void __launch_bounds__(256) k_cull_faces(const float* __restrict__ verts, const int* __restrict__ faces, int F,
                                                    const float* __restrict__ R, const float* __restrict__ t, float eps, int* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  bool keep = true;
  int v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v[k] = faces[3 * f + k];
    const float z = verts[3 * v[k]] * R[6] + verts[3 * v[k] + 1] * R[7] + verts[3 * v[k] + 2] * R[8] + t[2];
    if (z < eps) keep = false;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) out[3 * f + k] = keep ? v[k] : 0;
}

// grid = layout rows (n + 1).  row_to_kept[r] = j or -1.  vstart[j] .. vstart[j+1] = the vertices of kept object j.
__global__ void __launch_bounds__(256) k_assemble_bwd(const float* __restrict__ g_verts, const float* __restrict__ g_size,
                                                      const float* __restrict__ mv, const int* __restrict__ vstart,
                                                      const int* __restrict__ row_to_kept, const float* __restrict__ mcent,
                                                      const float* __restrict__ msize, const ObjParams* __restrict__ prm, float rx, float ry,
                                                      float rz, int fix_corners, float angle_scale, float* __restrict__ d_boxes,
                                                      float* __restrict__ d_angles) {
  const int r = blockIdx.x;
  const int j = row_to_kept[r];
  if (j < 0) {
    if (threadIdx.x < 6) d_boxes[6 * r + threadIdx.x] = 0.f;
    if (threadIdx.x == 6) d_angles[r] = 0.f;
    return;
  }
  const ObjParams p = prm[j];
  // acc: G (3) | sum g . (R m) | sum g . (dR/dtheta m)
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = vstart[j] + threadIdx.x; i < vstart[j + 1]; i += blockDim.x) {
    const float gx = g_verts[3 * i], gy = g_verts[3 * i + 1], gz = g_verts[3 * i + 2];
    const float x = mv[3 * i], y = mv[3 * i + 1], z = mv[3 * i + 2];
    acc[0] += gx; acc[1] += gy; acc[2] += gz;
    acc[3] += gx * (p.c * x + p.sn * z) + gy * y + gz * (-p.sn * x + p.c * z);
    acc[4] += gx * (-p.sn * x + p.c * z) + gz * (-p.c * x - p.sn * z);
  }
  __shared__ float sh[5][8];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  float t[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) { t[k] = 0.f; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t[k] += sh[k][w]; }
  const float mx = mcent[3 * j], my = mcent[3 * j + 1], mz = mcent[3 * j + 2];
  // v = s R m + (centre - s R mc)
  const float d_s = t[3] - (t[0] * (p.c * mx + p.sn * mz) + t[1] * my + t[2] * (-p.sn * mx + p.c * mz));
  const float d_theta = p.s * (t[4] - (t[0] * (-p.sn * mx + p.c * mz) + t[2] * (-p.c * mx - p.sn * mz)));
  const float room[3] = {rx, ry, rz};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float d_size = g_size ? g_size[3 * j + k] : 0.f;
    if (k == p.kmin) d_size += d_s / msize[3 * j + k];
    float lo = (t[k] * 0.5f - d_size) * room[k], hi = (t[k] * 0.5f + d_size) * room[k];
    if (fix_corners) { const float avg = hi / 2.0f + lo / 2.0f; lo = hi = avg; }   // fix_grad (test_render_refine.py:220-225): boxes translate, sizes stay
    d_boxes[6 * r + k] = lo;
    d_boxes[6 * r + 3 + k] = hi;
  }
  d_angles[r] = -d_theta * kAngleStep * angle_scale;                               // quad_grad (:227-230) when angle_scale = 4
}


// ------------------------------------------------------------------------------------------------ compositing (diff_render.py:366-434)
// depth [P] (one z-buffer render), images [C, P] (class masks) -> out [1 + (n_onehot - 1) + n_keep, P]:
//   out[0] = d = depth > 15 ? -1 : depth (:367);  out[ch] = images[c] for the class c with index[c] == ch, else 0 (:429-431);
//   out[n_onehot + k] = images[keep[k]] > 0.1 ? d / wall_max : mean_c / wall_max  (:401-421), mean_c = mean of d over the hard mask
//   of class c (wall_max if the mask is empty), wall_max = max of d over the wall mask (10 if empty; detached).
// Reductions are two-level with fixed order (deterministic).  stats [2C + 1] = cnt[C] | fill[C] = mean_c / wall_max | wall_max.
constexpr int kCompBlocks = 16;
constexpr float kHard = 0.1f;

__global__ void __launch_bounds__(256) k_comp_reduce(const float* __restrict__ depth, const float* __restrict__ images, int P, int wall,
                                                     float* __restrict__ partial) {
  const int c = blockIdx.x, b = blockIdx.y;
  const int per = (P + kCompBlocks - 1) / kCompBlocks, lo = b * per, hi = min(P, lo + per);
  float cnt = 0.f, sum = 0.f, mx = -INFINITY;
  for (int p = lo + threadIdx.x; p < hi; p += blockDim.x) {
    if (__ldg(images + (size_t)c * P + p) > kHard) {
      float d = __ldg(depth + p); d = d > 15.f ? -1.f : d;
      cnt += 1.f; sum += d; mx = fmaxf(mx, d);
    }
  }
  __shared__ float sh[3][8];
  cnt = warp_sum(cnt); sum = warp_sum(sum); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = cnt; sh[1][threadIdx.x >> 5] = sum; sh[2][threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, s2 = 0.f, m = -INFINITY;
    for (int w = 0; w < 8; ++w) { a += sh[0][w]; s2 += sh[1][w]; m = fmaxf(m, sh[2][w]); }
    float* o = partial + ((size_t)c * kCompBlocks + b) * 3;
    o[0] = a; o[1] = s2; o[2] = (c == wall) ? m : 0.f;
  }
}

__global__ void k_comp_stats(const float* __restrict__ partial, int C, int wall, float* __restrict__ stats) {
  __shared__ float s_wall;
  const int c = threadIdx.x;
  float cnt = 0.f, sum = 0.f, mx = -INFINITY;
  if (c < C)
    for (int b = 0; b < kCompBlocks; ++b) {
      const float* o = partial + ((size_t)c * kCompBlocks + b) * 3;
      cnt += o[0]; sum += o[1]; if (c == wall) mx = fmaxf(mx, o[2]);
    }
  if (c == wall) s_wall = cnt > 0.f ? mx : 10.0f;                       // :408-410
  __syncthreads();
  if (c >= C) return;
  const float wm = s_wall;
  const float mean = cnt > 0.f ? sum / fmaxf(cnt, 1.f) : wm;             // :411-419
  stats[c] = cnt; stats[C + c] = mean / wm;
  if (c == 0) stats[2 * C] = wm;
}

// inv_index [n_onehot]: class whose image goes to one-hot channel ch, or -1
__global__ void __launch_bounds__(256) k_comp_assemble(const float* __restrict__ depth, const float* __restrict__ images, int C, int P,
                                                       const int* __restrict__ inv_index, int n_onehot, const int* __restrict__ keep, int n_keep,
                                                       const float* __restrict__ stats, float* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float d = __ldg(depth + p); d = d > 15.f ? -1.f : d;
  const float wm = stats[2 * C];
  out[p] = d;
  for (int ch = 1; ch < n_onehot; ++ch) {
    const int c = inv_index[ch];
    out[(size_t)ch * P + p] = c >= 0 ? __ldg(images + (size_t)c * P + p) : 0.f;
  }
  const float dn = d / wm;
  for (int k = 0; k < n_keep; ++k) {
    const int c = keep[k];
    out[(size_t)(n_onehot + k) * P + p] = __ldg(images + (size_t)c * P + p) > kHard ? dn : stats[C + c];
  }
}

// S_k = sum over the pixels OUTSIDE the hard mask of class keep[k] of g_out[n_onehot + k]  (the gradient of the fill value)
__global__ void __launch_bounds__(256) k_comp_bwd_reduce(const float* __restrict__ images, const float* __restrict__ g_out, int P, const int* __restrict__ keep,
                                                         int n_onehot, float* __restrict__ partial) {
  const int k = blockIdx.x, b = blockIdx.y, c = keep[k];
  const int per = (P + kCompBlocks - 1) / kCompBlocks, lo = b * per, hi = min(P, lo + per);
  float s = 0.f;
  for (int p = lo + threadIdx.x; p < hi; p += blockDim.x)
    if (!(__ldg(images + (size_t)c * P + p) > kHard)) s += __ldg(g_out + (size_t)(n_onehot + k) * P + p);
  __shared__ float sh[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float a = 0.f; for (int w = 0; w < 8; ++w) a += sh[w]; partial[k * kCompBlocks + b] = a; }
}

__global__ void __launch_bounds__(256) k_comp_bwd(const float* __restrict__ depth, const float* __restrict__ images, const float* __restrict__ g_out, int C,
                                                  int P, const int* __restrict__ index, const int* __restrict__ keep, int n_keep, int n_onehot,
                                                  const float* __restrict__ stats, const float* __restrict__ partial, float* __restrict__ g_depth,
                                                  float* __restrict__ g_images) {
  __shared__ float s_fill[64];     // per kept class: d(loss)/d(mean_c) / wall_max / cnt_c, 0 when the mask is empty
  if (threadIdx.x < n_keep) {
    const int k = threadIdx.x, c = keep[k];
    float a = 0.f;
    for (int b = 0; b < kCompBlocks; ++b) a += partial[k * kCompBlocks + b];
    const float cnt = stats[c];
    s_fill[k] = cnt > 0.f ? a / stats[2 * C] / fmaxf(cnt, 1.f) : 0.f;
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  for (int c = 0; c < C; ++c) {
    const int ch = index[c];
    g_images[(size_t)c * P + p] = (ch >= 1 && ch < n_onehot) ? __ldg(g_out + (size_t)ch * P + p) : 0.f;   // the mask threshold is not differentiable (.detach(), :401)
  }
  float g = __ldg(g_out + p);
  const float wm = stats[2 * C];
  for (int k = 0; k < n_keep; ++k)
    if (__ldg(images + (size_t)keep[k] * P + p) > kHard) g += __ldg(g_out + (size_t)(n_onehot + k) * P + p) / wm + s_fill[k];
  g_depth[p] = __ldg(depth + p) > 15.f ? 0.f : g;
}

}  // namespace
}  // namespace sln

using namespace sln;

extern "C" {

size_t sln_scene_assemble_workspace_bytes(int64_t n_kept) { return (size_t)(n_kept > 0 ? n_kept : 1) * sizeof(ObjParams); }

int sln_scene_assemble_fwd(const float* boxes, const float* angles, int64_t n_rows, const int32_t* kept, int64_t n_kept, const float* room3_host,
                           const float* model_verts, const int32_t* vert_obj, int64_t n_obj_verts, const float* shell_verts, int64_t n_shell,
                           const float* model_size, const float* model_center, const int32_t* faces, int64_t F, const float* R, const float* t,
                           float cull_eps, float* vertices, float* sizes, int32_t* faces_out, void* ws, size_t ws_bytes, void* stream) {
  SLN_CHECK_ARG(boxes && angles && room3_host && vertices && ws, "scene_assemble_fwd: null pointer");
  SLN_CHECK_ARG(n_rows >= 1 && n_kept >= 0 && n_kept <= n_rows && n_obj_verts >= 0 && n_shell >= 0 && F >= 0, "scene_assemble_fwd: bad extents");
  SLN_CHECK_ARG(n_kept == 0 || (kept && model_verts && vert_obj && model_size && model_center && sizes), "scene_assemble_fwd: null object arrays");
  SLN_CHECK_ARG(F == 0 || (faces && faces_out && R && t), "scene_assemble_fwd: null face arrays");
  if (ws_bytes < sln_scene_assemble_workspace_bytes(n_kept)) { set_error("scene_assemble_fwd: workspace too small"); return SLN_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  ObjParams* prm = (ObjParams*)ws;
  if (n_kept > 0) {
    k_obj_params<<<(unsigned)((n_kept + 127) / 128), 128, 0, st>>>(boxes, angles, kept, (int)n_kept, room3_host[0], room3_host[1], room3_host[2],
                                                                 model_size, prm, sizes);
    SLN_TRY(check_launch("obj_params"));
  }
  const int64_t V = n_obj_verts + n_shell;
  if (V > 0) {
    k_assemble_vertices<<<(unsigned)((V + 255) / 256), 256, 0, st>>>(model_verts, vert_obj, (int)n_obj_verts, shell_verts, (int)n_shell, model_center,
                                                                     prm, vertices);
    SLN_TRY(check_launch("assemble_vertices"));
  }
  if (F > 0) {
    k_cull_faces<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(vertices, faces, (int)F, R, t, cull_eps, faces_out);
    SLN_TRY(check_launch("cull_faces"));
  }
  return SLN_OK;
}

int sln_scene_assemble_bwd(const float* grad_vertices, const float* grad_sizes, int64_t n_rows, const int32_t* row_to_kept, int64_t n_kept,
                           const float* room3_host, const float* model_verts, const int32_t* vert_start, const float* model_size,
                           const float* model_center, const void* ws, size_t ws_bytes, int32_t fix_corners, float angle_grad_scale,
                           float* d_boxes, float* d_angles, void* stream) {
  SLN_CHECK_ARG(grad_vertices && row_to_kept && room3_host && ws && d_boxes && d_angles, "scene_assemble_bwd: null pointer");
  SLN_CHECK_ARG(n_rows >= 1 && n_kept >= 0, "scene_assemble_bwd: bad extents");
  SLN_CHECK_ARG(n_kept == 0 || (model_verts && vert_start && model_size && model_center), "scene_assemble_bwd: null object arrays");
  if (ws_bytes < sln_scene_assemble_workspace_bytes(n_kept)) { set_error("scene_assemble_bwd: workspace too small"); return SLN_EWORKSPACE; }
  k_assemble_bwd<<<(unsigned)n_rows, 256, 0, (cudaStream_t)stream>>>(grad_vertices, grad_sizes, model_verts, vert_start, row_to_kept, model_center,
                                                                    model_size, (const ObjParams*)ws, room3_host[0], room3_host[1], room3_host[2],
                                                                    fix_corners, angle_grad_scale, d_boxes, d_angles);
  return check_launch("assemble_bwd");
}


size_t sln_composite_workspace_bytes(int32_t C) { return (size_t)(C > 0 ? C : 1) * kCompBlocks * 3 * sizeof(float); }

int sln_composite_fwd(const float* depth, const float* images, int32_t C, int64_t P, int32_t wall, const int32_t* inv_index, int32_t n_onehot,
                      const int32_t* keep, int32_t n_keep, float* out, float* stats, void* ws, size_t ws_bytes, void* stream) {
  SLN_CHECK_ARG(depth && images && inv_index && out && stats && ws && (n_keep == 0 || keep), "composite_fwd: null pointer");
  SLN_CHECK_ARG(C >= 1 && C <= 1024 && P >= 1 && P < (1ll << 30) && wall >= 0 && wall < C && n_onehot >= 1 && n_keep >= 0 && n_keep <= 64,
                "composite_fwd: bad extents (C <= 1024, n_keep <= 64)");
  if (ws_bytes < sln_composite_workspace_bytes(C)) { set_error("composite_fwd: workspace too small"); return SLN_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  k_comp_reduce<<<dim3(C, kCompBlocks), 256, 0, st>>>(depth, images, (int)P, wall, (float*)ws);
  SLN_TRY(check_launch("comp_reduce"));
  k_comp_stats<<<1, ((C + 31) / 32) * 32, 0, st>>>((const float*)ws, C, wall, stats);
  SLN_TRY(check_launch("comp_stats"));
  k_comp_assemble<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(depth, images, C, (int)P, inv_index, n_onehot, keep, n_keep, stats, out);
  return check_launch("comp_assemble");
}

int sln_composite_bwd(const float* depth, const float* images, const float* grad_out, int32_t C, int64_t P, const int32_t* index, int32_t n_onehot,
                      const int32_t* keep, int32_t n_keep, const float* stats, float* grad_depth, float* grad_images, void* ws, size_t ws_bytes,
                      void* stream) {
  SLN_CHECK_ARG(depth && images && grad_out && index && stats && grad_depth && grad_images && ws && (n_keep == 0 || keep), "composite_bwd: null pointer");
  SLN_CHECK_ARG(C >= 1 && C <= 1024 && P >= 1 && P < (1ll << 30) && n_onehot >= 1 && n_keep >= 0 && n_keep <= 64, "composite_bwd: bad extents");
  if (ws_bytes < sln_composite_workspace_bytes(C > n_keep ? C : n_keep)) { set_error("composite_bwd: workspace too small"); return SLN_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  if (n_keep > 0) {
    k_comp_bwd_reduce<<<dim3(n_keep, kCompBlocks), 256, 0, st>>>(images, grad_out, (int)P, keep, n_onehot, (float*)ws);
    SLN_TRY(check_launch("comp_bwd_reduce"));
  }
  k_comp_bwd<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(depth, images, grad_out, C, (int)P, index, keep, n_keep, n_onehot, stats, (const float*)ws,
                                                          grad_depth, grad_images);
  return check_launch("comp_bwd");
}

}  // extern "C"
