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
                                                      float rz, float* __restrict__ d_boxes, float* __restrict__ d_angles) {
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
    d_boxes[6 * r + k] = (t[k] * 0.5f - d_size) * room[k];
    d_boxes[6 * r + 3 + k] = (t[k] * 0.5f + d_size) * room[k];
  }
  d_angles[r] = -d_theta * kAngleStep;
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
                           const float* model_center, const void* ws, size_t ws_bytes, float* d_boxes, float* d_angles, void* stream) {
  SLN_CHECK_ARG(grad_vertices && row_to_kept && room3_host && ws && d_boxes && d_angles, "scene_assemble_bwd: null pointer");
  SLN_CHECK_ARG(n_rows >= 1 && n_kept >= 0, "scene_assemble_bwd: bad extents");
  SLN_CHECK_ARG(n_kept == 0 || (model_verts && vert_start && model_size && model_center), "scene_assemble_bwd: null object arrays");
  if (ws_bytes < sln_scene_assemble_workspace_bytes(n_kept)) { set_error("scene_assemble_bwd: workspace too small"); return SLN_EWORKSPACE; }
  k_assemble_bwd<<<(unsigned)n_rows, 256, 0, (cudaStream_t)stream>>>(grad_vertices, grad_sizes, model_verts, vert_start, row_to_kept, model_center,
                                                                    model_size, (const ObjParams*)ws, room3_host[0], room3_host[1], room3_host[2],
                                                                    d_boxes, d_angles);
  return check_launch("assemble_bwd");
}

}  // extern "C"
