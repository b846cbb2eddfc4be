"""TEST INFRASTRUCTURE ONLY — Python face of the rasterizer oracle (oracle/raster_oracle.c, built by oracle/build_oracle.py)
plus a torch-CPU restatement of the reference's compositing (models/diff_render.py:344-434) and camera (:13-46).

PARITY UNPINNED for the rasterizer core itself: see the header of raster_oracle.c (third-party `neural_renderer`, un-pinned,
not installed, CUDA-only; restated from its published algorithm).  The compositing / camera parts restate reference code
that IS in /root/reference (models/diff_render.py) but cannot be imported there (its `models/misc.py` needs pywavefront,
pymesh, SUNCG metadata and `np.float`); they ARE pinned by execution: oracle/gen_golden_render.py runs the reference's own
get_cam_mat / mesh_render_func from the file's AST (renderer stubbed by this C oracle) into tests/golden/render_*.npz, and
tests/test_oracle_render.py checks `get_cam_mat` (bit-exact) and `composite` (1e-6) below against those goldens.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import build_oracle

_lib = None
_F = ctypes.c_float
_I = ctypes.c_int
_P = ctypes.c_void_p


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.lib_path()
        if not os.path.exists(path):
            build_oracle.build()
        L = ctypes.CDLL(path)
        L.ro_project.argtypes = [_P, _I, _P, _P, _P, _F, _P]
        L.ro_project_bwd.argtypes = [_P, _I, _P, _P, _P, _F, _P, _P]
        L.ro_gather_faces.argtypes = [_P, _P, _I, _I, _P]
        L.ro_gather_faces_bwd.argtypes = [_P, _P, _I, _I, _I, _P]
        L.ro_face_inv.argtypes = [_P, _I, _I, _P]
        L.ro_face_index_map.argtypes = [_P, _P, _I, _I, _F, _F, _P, _P, _P, _P]
        L.ro_texture_sampling.argtypes = [_P, _P, _P, _P, _P, _I, _I, _F, _P]
        L.ro_backward_pixel_map.argtypes = [_P, _P, _P, _P, _I, _I, _F, _P]
        L.ro_backward_depth_map.argtypes = [_P, _P, _P, _P, _P, _P, _I, _P]
        for f in ("ro_project", "ro_project_bwd", "ro_gather_faces", "ro_gather_faces_bwd", "ro_face_inv", "ro_face_index_map",
                  "ro_texture_sampling", "ro_backward_pixel_map", "ro_backward_depth_map"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(_P)


def project(verts, K, R, t, orig_size):
    verts, K, R, t = _f32(verts).reshape(-1, 3), _f32(K).reshape(9), _f32(R).reshape(9), _f32(t).reshape(3)
    out = np.empty_like(verts)
    lib().ro_project(_p(verts), len(verts), _p(K), _p(R), _p(t), float(orig_size), _p(out))
    return out


def project_bwd(verts, K, R, t, orig_size, grad_out):
    verts, K, R, t, grad_out = _f32(verts).reshape(-1, 3), _f32(K).reshape(9), _f32(R).reshape(9), _f32(t).reshape(3), _f32(grad_out)
    g = np.empty_like(verts)
    lib().ro_project_bwd(_p(verts), len(verts), _p(K), _p(R), _p(t), float(orig_size), _p(grad_out), _p(g))
    return g


def gather_faces(pv, faces, fill_back=True):
    pv = _f32(pv)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    F = len(faces)
    fv = np.empty(((2 * F if fill_back else F), 9), dtype=np.float32)
    lib().ro_gather_faces(_p(pv), _p(faces), F, int(fill_back), _p(fv))
    return fv


def gather_faces_bwd(grad_fv, faces, V, fill_back=True):
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    grad_fv = _f32(grad_fv)
    g = np.empty((V, 3), dtype=np.float32)
    lib().ro_gather_faces_bwd(_p(grad_fv), _p(faces), len(faces), int(fill_back), V, _p(g))
    return g


def face_index_map(fv, image_size, near, far):
    """-> dict(face_inv [F2,9], face_index [is,is] i32 (-1 = background), weight [is,is,3], depth [is,is] (far on empty),
    face_inv_map [is,is,9]) in the renderer's internal (un-flipped) orientation."""
    fv = _f32(fv)
    F2, n = len(fv), image_size
    finv = np.empty((F2, 9), dtype=np.float32)
    lib().ro_face_inv(_p(fv), F2, n, _p(finv))
    fi = np.empty((n, n), dtype=np.int32)
    w = np.empty((n, n, 3), dtype=np.float32)
    d = np.empty((n, n), dtype=np.float32)
    fim = np.empty((n, n, 9), dtype=np.float32)
    lib().ro_face_index_map(_p(fv), _p(finv), F2, n, float(near), float(far), _p(fi), _p(w), _p(d), _p(fim))
    return dict(face_inv=finv, face_index=fi, weight=w, depth=d, face_inv_map=fim)


def texture_sampling(fv, textures, maps, texture_size, eps=1e-3):
    fv, textures = _f32(fv), _f32(textures)
    n = maps["face_index"].shape[0]
    rgb = np.empty((n, n, 3), dtype=np.float32)
    lib().ro_texture_sampling(_p(fv), _p(textures), _p(maps["face_index"]), _p(maps["weight"]), _p(maps["depth"]), n, texture_size,
                              float(eps), _p(rgb))
    return rgb


def backward_pixel_map(fv, maps, rgb, grad_rgb, eps=1e-3):
    fv, rgb, grad_rgb = _f32(fv), _f32(rgb), _f32(grad_rgb)
    n = maps["face_index"].shape[0]
    g = np.zeros((len(fv), 9), dtype=np.float32)
    lib().ro_backward_pixel_map(_p(fv), _p(maps["face_index"]), _p(rgb), _p(grad_rgb), len(fv), n, float(eps), _p(g))
    return g


def backward_depth_map(fv, maps, grad_depth, grad_faces=None):
    fv, grad_depth = _f32(fv), _f32(grad_depth)
    n = maps["face_index"].shape[0]
    g = np.zeros((len(fv), 9), dtype=np.float32) if grad_faces is None else grad_faces
    lib().ro_backward_depth_map(_p(fv), _p(maps["depth"]), _p(maps["face_index"]), _p(maps["face_inv_map"]), _p(maps["weight"]),
                                _p(grad_depth), n, _p(g))
    return g


# ----------------------------------------------------------------------------------------------------------------------
class RendererOracle(object):
    """What `nr.Renderer(camera_mode='projection', image_size, K, R, t, anti_aliasing=False, orig_size, near, ...)` computes
    for mode='depth' and mode='rgb' (upstream renderer.py render_depth / render_rgb with fill_back=True, ambient 1.0,
    directional 0.0 => textures pass through lighting unchanged), plus the gradient w.r.t. the world-space vertices.

    near: upstream's render_depth calls rasterize_depth WITHOUT near/far, so the depth pass uses the rasterizer defaults
    (near 0.1, far 100) while render_rgb uses the constructor's near (0.001 in the reference) — SURVEY.md App. C item 8.
    Both are parameters here so that either reading can be selected."""

    def __init__(self, image_size, K, R, t, orig_size, near_rgb=0.001, near_depth=0.1, far=100.0, fill_back=True, eps=1e-3):
        self.n, self.K, self.R, self.t, self.orig = image_size, _f32(K).reshape(9), _f32(R).reshape(9), _f32(t).reshape(3), orig_size
        self.near_rgb, self.near_depth, self.far, self.fill_back, self.eps = near_rgb, near_depth, far, fill_back, eps

    def _faces(self, vertices, faces):
        pv = project(vertices, self.K, self.R, self.t, self.orig)
        return pv, gather_faces(pv, faces, self.fill_back)

    def _fill_tex(self, textures):
        if not self.fill_back:
            return _f32(textures)
        t = _f32(textures)          # [F, ts, ts, ts, 3]; back faces: permute(0,1,4,3,2,5) upstream = swap first/last texel axes
        return np.concatenate([t, np.ascontiguousarray(t.transpose(0, 3, 2, 1, 4))], axis=0)

    def depth(self, vertices, faces):
        pv, fv = self._faces(vertices, faces)
        maps = face_index_map(fv, self.n, self.near_depth, self.far)
        return maps["depth"][::-1].copy(), dict(fv=fv, maps=maps)          # vertical flip on output (rasterize.py)

    def depth_bwd(self, vertices, faces, ctx, grad_depth):
        g = backward_depth_map(ctx["fv"], ctx["maps"], np.ascontiguousarray(_f32(grad_depth)[::-1]))
        gpv = gather_faces_bwd(g, faces, len(vertices), self.fill_back)
        return project_bwd(vertices, self.K, self.R, self.t, self.orig, gpv)

    def rgb(self, vertices, faces, textures):
        pv, fv = self._faces(vertices, faces)
        maps = face_index_map(fv, self.n, self.near_rgb, self.far)
        ts = textures.shape[1]
        rgb = texture_sampling(fv, self._fill_tex(textures), maps, ts, self.eps)
        out = np.ascontiguousarray(rgb[::-1].transpose(2, 0, 1))            # [3, is, is], flipped
        return out, dict(fv=fv, maps=maps, rgb=rgb)

    def rgb_bwd(self, vertices, faces, ctx, grad_rgb):
        g_int = np.ascontiguousarray(_f32(grad_rgb).transpose(1, 2, 0)[::-1])
        g = backward_pixel_map(ctx["fv"], ctx["maps"], ctx["rgb"], g_int, self.eps)
        gpv = gather_faces_bwd(g, faces, len(vertices), self.fill_back)
        return project_bwd(vertices, self.K, self.R, self.t, self.orig, gpv)


def get_cam_mat(room_box):
    """reference models/diff_render.py:13-46 — K [3,3], R [3,3], t [3] from the room box (last row of `boxes`)."""
    theta = -0.4
    fl = 400.0
    inter_out = 512
    K = np.array([[fl * inter_out / 1024, 0, inter_out / 2.0], [0, fl * inter_out / 1024, inter_out / 2.0], [0, 0, 1.0]], dtype="float32")
    Rw = torch.from_numpy(np.array([[1, 0, 0], [0, np.cos(theta), np.sin(theta)], [0, -np.sin(theta), np.cos(theta)]], dtype="float32"))
    cam = torch.zeros(3, 1)
    cam[0, 0] = float(room_box[3]) / 2.0
    cam[1, 0] = float(room_box[4]) / 2.0 + min(0.1, abs(float(room_box[4]) / 2.0))
    cam[2, 0] = float(room_box[5])
    t_w2c = torch.matmul(Rw, -cam)
    cv = torch.tensor([[1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=torch.float)
    return K, torch.matmul(cv, Rw).numpy(), torch.matmul(cv, t_w2c).numpy().reshape(3)


NYU_CLASS = ['wall', 'floor', 'cabinet', 'bed', 'chair', 'sofa', 'table', 'door', 'window', 'bookshelf', 'picture', 'counter', 'blinds',
             'desk', 'shelves', 'curtain', 'dresser', 'pillow', 'mirror', 'floor mat', 'clothes', 'ceiling', 'books', 'refridgerator',
             'television', 'paper', 'towel', 'shower curtain', 'box', 'whiteboard', 'person', 'night stand', 'toilet', 'sink', 'lamp',
             'bathtub', 'bag', 'otherstructure', 'otherfurniture', 'otherprop']      # reference models/diff_render.py:3


def composite(depth_data, class_images, class_names):
    """reference models/diff_render.py:366-434 as torch ops (autograd supplies the gradient oracle).

    depth_data   [1,H,W] (output of mode='depth', differentiable)
    class_images list of [1,H,W] tensors = sum(rgb, dim=1)/3 of the per-class mode='rgb' renders, in `class_names` order
                 (wall first, :372-374)
    -> final [1, 1+40+(len-3), H, W]"""
    depth_data = depth_data.clone()
    depth_data[depth_data > 15] = -1
    H, W = depth_data.shape[1:]
    one_hot = torch.zeros(41, H, W)
    planes = []
    wall_max = None
    for name, image in zip(class_names, class_images):
        hard_mask = image.detach() > 0.1
        depth_masked = depth_data[hard_mask]
        class_depth = torch.zeros_like(depth_data)
        mean = torch.mean(depth_masked)
        if name == "wall":
            wall_max = torch.max(depth_data[hard_mask]).detach() if hard_mask.any() else torch.tensor(float("nan"))
            if torch.isnan(wall_max):
                wall_max = 10.0
        if torch.isnan(mean):
            mean = wall_max
        class_depth[~hard_mask] = mean / wall_max
        class_depth[hard_mask] = depth_data[hard_mask] / wall_max
        if name not in ("wall", "floor", "ceiling"):
            planes.append(class_depth)
        one_hot[NYU_CLASS.index(name.replace("_", " ")) + 1] = image[0]
    return torch.cat([depth_data, one_hot[1:]] + planes, dim=0)[None]
