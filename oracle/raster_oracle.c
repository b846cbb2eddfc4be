/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of the Neural 3D Mesh Renderer operators that the reference reaches
 * through `nr.Renderer` (reference models/misc.py:7 import, models/diff_render.py:359-361 ctor, :366 mode='depth',
 * :398 mode='rgb').  Nothing under 3d_sln_b200/ links or loads this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED.  The algorithm lives in a third-party dependency that is NOT in /root/reference:
 *   neural_renderer, "daniilidis version" (reference README.md:12), no version or commit pinned anywhere in the reference
 *   (no requirements.txt / setup.py), locally patched per README.md:13-18 (lens distortion removed from projection.py).
 * Upstream is a CUDA-only extension, is not installed here, and there is no network, so nothing in this file could be
 * checked against upstream outputs.  It restates upstream's published algorithm (Kato et al., CVPR 2018; the package's
 * rasterize_cuda_kernel.cu / projection.py / rasterize.py) from knowledge of that code; the reference holds no golden
 * vectors or tests for this path.  The restatement is validated by self-consistency only
 * (tests/test_oracle_raster.py): float64 brute-force rasterization away from ties, finite differences of the depth
 * gradient, and closed-form cases.
 *
 * Arithmetic contract: plain IEEE fp32, no fused multiply-add (compile with -ffp-contract=off), operations in the order
 * written here.  The CUDA kernels follow the same order with __fmul_rn/__fadd_rn so that index buffers are bit-identical.
 * (Upstream mixes double literals into float expressions — 0.5 *, 2. *, 1. / — every such case is a correctly-rounded
 * single operation whose double-then-float rounding equals the direct float rounding (53 >= 2*24+2), so fp32 is exact
 * to upstream's intent.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RO_API __attribute__((visibility("default")))

/* ---- nr.projection (upstream projection.py, patched per README.md:13-18: x__ = x_, y__ = y_):
 * v_cam = R v + t ; x_ = x/(z+eps), y_ = y/(z+eps), eps = 1e-9 ; [u v] = K [x_ y_ 1] ; v = orig - v ;
 * u,v -> 2*(. - orig/2)/orig ; output [u, v, z_cam]. */
RO_API void ro_project(const float* verts, int V, const float* K, const float* R, const float* t, float orig_size, float* out) {
  const float eps = 1e-9f;
  const float half = orig_size / 2.0f;
  for (int i = 0; i < V; ++i) {
    const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    const float xc = ((x * R[0] + y * R[1]) + z * R[2]) + t[0];
    const float yc = ((x * R[3] + y * R[4]) + z * R[5]) + t[1];
    const float zc = ((x * R[6] + y * R[7]) + z * R[8]) + t[2];
    const float x_ = xc / (zc + eps), y_ = yc / (zc + eps);
    float u = (x_ * K[0] + y_ * K[1]) + K[2];
    float v = (x_ * K[3] + y_ * K[4]) + K[5];
    v = orig_size - v;
    u = 2.0f * (u - half) / orig_size;
    v = 2.0f * (v - half) / orig_size;
    out[3 * i] = u; out[3 * i + 1] = v; out[3 * i + 2] = zc;
  }
}

/* gradient of ro_project w.r.t. the world vertices (what torch autograd computes upstream) */
RO_API void ro_project_bwd(const float* verts, int V, const float* K, const float* R, const float* t, float orig_size,
                           const float* grad_out, float* grad_verts) {
  const float eps = 1e-9f;
  for (int i = 0; i < V; ++i) {
    const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    const float xc = ((x * R[0] + y * R[1]) + z * R[2]) + t[0];
    const float yc = ((x * R[3] + y * R[4]) + z * R[5]) + t[1];
    const float zc = ((x * R[6] + y * R[7]) + z * R[8]) + t[2];
    const float zi = 1.0f / (zc + eps);
    const float du = grad_out[3 * i] * (2.0f / orig_size), dv = -grad_out[3 * i + 1] * (2.0f / orig_size);
    const float dx_ = K[0] * du + K[3] * dv, dy_ = K[1] * du + K[4] * dv;
    const float dxc = dx_ * zi, dyc = dy_ * zi;
    const float dzc = grad_out[3 * i + 2] - (dx_ * xc + dy_ * yc) * zi * zi;
    grad_verts[3 * i] = R[0] * dxc + R[3] * dyc + R[6] * dzc;
    grad_verts[3 * i + 1] = R[1] * dxc + R[4] * dyc + R[7] * dzc;
    grad_verts[3 * i + 2] = R[2] * dxc + R[5] * dyc + R[8] * dzc;
  }
}

/* ---- fill_back + nr.vertices_to_faces: faces [F,3] -> fv [F2,3,3], F2 = fill_back ? 2F : F; face F+f = face f with the
 * vertex order reversed (upstream renderer.py: torch.cat((faces, faces[:, :, ::-1]))). */
RO_API void ro_gather_faces(const float* pv, const int32_t* faces, int F, int fill_back, float* fv) {
  const int F2 = fill_back ? 2 * F : F;
  for (int f = 0; f < F2; ++f) {
    const int src = f < F ? f : f - F;
    for (int k = 0; k < 3; ++k) {
      const int vi = faces[3 * src + (f < F ? k : 2 - k)];
      fv[9 * f + 3 * k] = pv[3 * vi]; fv[9 * f + 3 * k + 1] = pv[3 * vi + 1]; fv[9 * f + 3 * k + 2] = pv[3 * vi + 2];
    }
  }
}
/* its transpose: grad_fv [F2,9] -> grad_pv [V,3] (index_put accumulate upstream; summed here in face order) */
RO_API void ro_gather_faces_bwd(const float* grad_fv, const int32_t* faces, int F, int fill_back, int V, float* grad_pv) {
  const int F2 = fill_back ? 2 * F : F;
  memset(grad_pv, 0, sizeof(float) * 3 * (size_t)V);
  for (int f = 0; f < F2; ++f) {
    const int src = f < F ? f : f - F;
    for (int k = 0; k < 3; ++k) {
      const int vi = faces[3 * src + (f < F ? k : 2 - k)];
      for (int d = 0; d < 3; ++d) grad_pv[3 * vi + d] += grad_fv[9 * f + 3 * k + d];
    }
  }
}

static int ro_backside(const float* face) {
  return (face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]);
}

/* ---- forward_face_index_map kernel 1: per-face inverse of [[x0,x1,x2],[y0,y1,y2],[1,1,1]] in pixel space.
 * Back faces keep zeros (upstream returns before writing). */
RO_API void ro_face_inv(const float* fv, int F2, int is, float* face_inv_all) {
  const float fis = (float)is;
  memset(face_inv_all, 0, sizeof(float) * 9 * (size_t)F2);
  for (int f = 0; f < F2; ++f) {
    const float* face = fv + 9 * f;
    if (ro_backside(face)) continue;
    float p[3][2];
    for (int n = 0; n < 3; ++n)
      for (int d = 0; d < 2; ++d) p[n][d] = 0.5f * ((face[3 * n + d] * fis + fis) - 1.0f);
    float inv[9] = {
        p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
        p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
        p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
    const float den = (p[2][0] * (p[0][1] - p[1][1]) + p[0][0] * (p[1][1] - p[2][1])) + p[1][0] * (p[2][1] - p[0][1]);
    for (int k = 0; k < 9; ++k) face_inv_all[9 * f + k] = inv[k] / den;
  }
}

/* ---- forward_face_index_map kernel 2: per pixel, brute force over the faces in index order; strict '<' on the z-buffer,
 * so ties keep the lower face index.  Maps are in the renderer's internal (un-flipped) orientation: pixel (yi, xi). */
RO_API void ro_face_index_map(const float* fv, const float* face_inv_all, int F2, int is, float near, float far,
                              int32_t* face_index_map, float* weight_map, float* depth_map, float* face_inv_map) {
  const float fis = (float)is;
  for (int pn = 0; pn < is * is; ++pn) {
    const int yi = pn / is, xi = pn % is;
    const float yp = ((2.0f * (float)yi + 1.0f) - fis) / fis;
    const float xp = ((2.0f * (float)xi + 1.0f) - fis) / fis;
    float depth_min = far;
    int face_min = -1;
    float wmin[3] = {0.f, 0.f, 0.f};
    for (int fn = 0; fn < F2; ++fn) {
      const float* face = fv + 9 * fn;
      const float* finv = face_inv_all + 9 * fn;
      if (ro_backside(face)) continue;
      if (((yp - face[1]) * (face[3] - face[0]) < (xp - face[0]) * (face[4] - face[1])) ||
          ((yp - face[4]) * (face[6] - face[3]) < (xp - face[3]) * (face[7] - face[4])) ||
          ((yp - face[7]) * (face[0] - face[6]) < (xp - face[6]) * (face[1] - face[7])))
        continue;
      float w[3];
      for (int k = 0; k < 3; ++k) w[k] = (finv[3 * k] * (float)xi + finv[3 * k + 1] * (float)yi) + finv[3 * k + 2];
      float wsum = 0.f;
      for (int k = 0; k < 3; ++k) {
        w[k] = fminf(fmaxf(w[k], 0.f), 1.f);
        wsum += w[k];
      }
      for (int k = 0; k < 3; ++k) w[k] /= wsum;
      const float zp = 1.0f / ((w[0] / face[2] + w[1] / face[5]) + w[2] / face[8]);
      if (zp <= near || far <= zp) continue;
      if (zp < depth_min) {
        depth_min = zp;
        face_min = fn;
        wmin[0] = w[0]; wmin[1] = w[1]; wmin[2] = w[2];
      }
    }
    face_index_map[pn] = face_min;
    depth_map[pn] = depth_min;   /* stays `far` on empty pixels */
    for (int k = 0; k < 3; ++k) weight_map[3 * pn + k] = face_min >= 0 ? wmin[k] : 0.f;
    if (face_inv_map)
      for (int k = 0; k < 9; ++k) face_inv_map[9 * pn + k] = face_min >= 0 ? face_inv_all[9 * face_min + k] : 0.f;
  }
}

/* ---- forward_texture_sampling: trilinear blend of the face's [ts,ts,ts,3] texture (eps = 1e-3), background 0. */
RO_API void ro_texture_sampling(const float* fv, const float* textures, const int32_t* face_index_map, const float* weight_map,
                                const float* depth_map, int is, int ts, float eps, float* rgb_map) {
  for (int pn = 0; pn < is * is; ++pn) {
    float* pixel = rgb_map + 3 * pn;
    pixel[0] = pixel[1] = pixel[2] = 0.f;
    const int fi = face_index_map[pn];
    if (fi < 0) continue;
    const float* face = fv + 9 * fi;
    const float* tex = textures + (size_t)fi * ts * ts * ts * 3;
    const float* weight = weight_map + 3 * pn;
    const float depth = depth_map[pn];
    float tif[3];
    for (int k = 0; k < 3; ++k) {
      float v = (weight[k] * (float)(ts - 1)) * (depth / face[3 * k + 2]);
      v = fmaxf(v, 0.f);
      v = fminf(v, ((float)(ts - 1)) - eps);
      tif[k] = v;
    }
    float acc[3] = {0.f, 0.f, 0.f};
    for (int pnn = 0; pnn < 8; ++pnn) {
      float w = 1.f;
      int ti[3];
      for (int k = 0; k < 3; ++k) {
        const int base = (int)tif[k];
        if (((pnn >> k) % 2) == 0) { w *= 1.f - (tif[k] - (float)base); ti[k] = base; }
        else { w *= tif[k] - (float)base; ti[k] = base + 1; }
      }
      const int isc = (ti[0] * ts + ti[1]) * ts + ti[2];
      for (int k = 0; k < 3; ++k) acc[k] += w * tex[isc * 3 + k];
    }
    pixel[0] = acc[0]; pixel[1] = acc[1]; pixel[2] = acc[2];
  }
}

/* ---- backward_pixel_map (rgb channels only: the reference never asks for alpha): Kato's hand-crafted gradient of the
 * image w.r.t. the x,y of each front face's vertices.  grad_faces [F2,9] is OVERWRITTEN for front faces, untouched for
 * back faces (caller zeroes), as upstream. */
RO_API void ro_backward_pixel_map(const float* fv, const int32_t* face_index_map, const float* rgb_map, const float* grad_rgb_map,
                                  int F2, int is, float eps, float* grad_faces) {
  const float fis = (float)is;
  for (int fn = 0; fn < F2; ++fn) {
    const float* face = fv + 9 * fn;
    float grad_face[9] = {0};
    if (ro_backside(face)) continue;
    for (int edge = 0; edge < 3; ++edge) {
      int pi[3];
      float pp[3][2];
      for (int n = 0; n < 3; ++n) pi[n] = (edge + n) % 3;
      for (int n = 0; n < 3; ++n)
        for (int d = 0; d < 2; ++d) pp[n][d] = 0.5f * ((face[3 * pi[n] + d] * fis + fis) - 1.0f);
      for (int axis = 0; axis < 2; ++axis) {
        float p[3][2];
        for (int n = 0; n < 3; ++n)
          for (int d = 0; d < 2; ++d) p[n][d] = pp[n][(d + axis) % 2];
        int direction;
        if (axis == 0) direction = (p[0][0] < p[1][0]) ? -1 : 1;
        else direction = (p[0][0] < p[1][0]) ? 1 : -1;
        const int d0_from = (int)fmaxf(ceilf(fminf(p[0][0], p[1][0])), 0.f);
        const int d0_to = (int)fminf(fmaxf(p[0][0], p[1][0]), fis - 1.f);
        for (int d0 = d0_from; d0 <= d0_to; ++d0) {
          const float fd0 = (float)d0;
          const float d1_cross = ((p[1][1] - p[0][1]) / (p[1][0] - p[0][0])) * (fd0 - p[0][0]) + p[0][1];
          int d1_in = (0 < direction) ? (int)floorf(d1_cross) : (int)ceilf(d1_cross);
          const int d1_out = d1_in + direction;
          if (d1_in < 0 || is <= d1_in) continue;
          if (d1_out < 0 || is <= d1_out) continue;
          int idx_in, idx_out;
          if (axis == 0) { idx_in = d1_in * is + d0; idx_out = d1_out * is + d0; }
          else { idx_in = d0 * is + d1_in; idx_out = d0 * is + d1_out; }
          const float* rgb_in = rgb_map + 3 * idx_in;
          const float* rgb_out = rgb_map + 3 * idx_out;
          /* out: from the out-pixel to the image border, only if the in-pixel shows this face */
          if (face_index_map[idx_in] == fn) {
            const int d1_limit = (0 < direction) ? is - 1 : 0;
            int d1_from = d1_out < d1_limit ? d1_out : d1_limit; if (d1_from < 0) d1_from = 0;
            int d1_to = d1_out > d1_limit ? d1_out : d1_limit; if (d1_to > is - 1) d1_to = is - 1;
            for (int d1 = d1_from; d1 <= d1_to; ++d1) {
              const int idx = axis == 0 ? d1 * is + d0 : d0 * is + d1;
              float diff_grad = 0.f;
              for (int k = 0; k < 3; ++k) diff_grad += (rgb_map[3 * idx + k] - rgb_in[k]) * grad_rgb_map[3 * idx + k];
              if (diff_grad <= 0.f) continue;
              if (p[1][0] != fd0) {
                float dist = (((p[1][0] - p[0][0]) / (p[1][0] - fd0)) * ((float)d1 - d1_cross)) * 2.0f / fis;
                dist = (0.f < dist) ? dist + eps : dist - eps;
                grad_face[pi[0] * 3 + (1 - axis)] -= diff_grad / dist;
              }
              if (p[0][0] != fd0) {
                float dist = (((p[1][0] - p[0][0]) / (fd0 - p[0][0])) * ((float)d1 - d1_cross)) * 2.0f / fis;
                dist = (0.f < dist) ? dist + eps : dist - eps;
                grad_face[pi[1] * 3 + (1 - axis)] -= diff_grad / dist;
              }
            }
          }
          /* in: from the in-pixel to the opposite edge crossing, pixels showing this face */
          {
            float d0_cross2;
            if ((fd0 - p[0][0]) * (fd0 - p[2][0]) < 0.f)
              d0_cross2 = ((p[2][1] - p[0][1]) / (p[2][0] - p[0][0])) * (fd0 - p[0][0]) + p[0][1];
            else
              d0_cross2 = ((p[1][1] - p[2][1]) / (p[1][0] - p[2][0])) * (fd0 - p[2][0]) + p[2][1];
            const int d1_limit = (0 < direction) ? (int)ceilf(d0_cross2) : (int)floorf(d0_cross2);
            int d1_from = d1_in < d1_limit ? d1_in : d1_limit; if (d1_from < 0) d1_from = 0;
            int d1_to = d1_in > d1_limit ? d1_in : d1_limit; if (d1_to > is - 1) d1_to = is - 1;
            for (int d1 = d1_from; d1 <= d1_to; ++d1) {
              const int idx = axis == 0 ? d1 * is + d0 : d0 * is + d1;
              if (face_index_map[idx] != fn) continue;
              float diff_grad = 0.f;
              for (int k = 0; k < 3; ++k) diff_grad += (rgb_map[3 * idx + k] - rgb_out[k]) * grad_rgb_map[3 * idx + k];
              if (diff_grad <= 0.f) continue;
              if (p[1][0] != fd0) {
                float dist = (((p[1][0] - p[0][0]) / (p[1][0] - fd0)) * ((float)d1 - d1_cross)) * 2.0f / fis;
                dist = (0.f < dist) ? dist + eps : dist - eps;
                grad_face[pi[0] * 3 + (1 - axis)] -= diff_grad / dist;
              }
              if (p[0][0] != fd0) {
                float dist = (((p[1][0] - p[0][0]) / (fd0 - p[0][0])) * ((float)d1 - d1_cross)) * 2.0f / fis;
                dist = (0.f < dist) ? dist + eps : dist - eps;
                grad_face[pi[1] * 3 + (1 - axis)] -= diff_grad / dist;
              }
            }
          }
        }
      }
    }
    for (int k = 0; k < 9; ++k) grad_faces[9 * fn + k] = grad_face[k];
  }
}

/* ---- backward_depth_map: per covered pixel, d depth / d (vertex z and x,y); ACCUMULATES into grad_faces (atomicAdd
 * upstream; pixel order here). */
RO_API void ro_backward_depth_map(const float* fv, const float* depth_map, const int32_t* face_index_map, const float* face_inv_map,
                                  const float* weight_map, const float* grad_depth_map, int is, float* grad_faces) {
  const float fis = (float)is;
  for (int pn = 0; pn < is * is; ++pn) {
    const int fn = face_index_map[pn];
    if (fn < 0) continue;
    const float* face = fv + 9 * fn;
    const float depth = depth_map[pn], depth2 = depth * depth;
    const float* finv = face_inv_map + 9 * pn;
    const float* weight = weight_map + 3 * pn;
    const float g = grad_depth_map[pn];
    float* gf = grad_faces + 9 * fn;
    for (int k = 0; k < 3; ++k) {
      const float zk = face[3 * k + 2];
      gf[3 * k + 2] += g * weight[k] * depth2 / (zk * zk);
    }
    float tmp[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < 3; ++k)
      for (int l = 0; l < 3; ++l) tmp[k] += -finv[3 * l + k] / face[3 * l + 2];
    for (int k = 0; k < 3; ++k)
      for (int l = 0; l < 2; ++l) gf[3 * k + l] += -g * tmp[l] * weight[k] * depth2 * fis / 2.f;
  }
}
