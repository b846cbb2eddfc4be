"""Drop-in for the slice of the third-party ``neural_renderer`` package that the reference uses
(``import neural_renderer as nr`` at reference models/misc.py:7; ``nr.Renderer(camera_mode='projection', image_size=256,
K=..., R=..., t=..., anti_aliasing=False, orig_size=512, near=0.001, light_intensity_ambient=1.0,
light_intensity_directional=0.0)`` at models/diff_render.py:359-361; ``renderer(vertices, faces, textures, mode='depth')``
at :366 and ``mode='rgb'`` at :398).

Everything runs in libsln_b200.so (csrc/raster.cu) — tiled z-buffer forward with exact face-index output and the Neural 3D
Mesh Renderer's hand-crafted backward.  There is no CPU fallback.  Beyond the upstream-shaped ``Renderer`` this module
has ``render_scene_classes``: depth + ALL per-class mask images of a scene from one rasterization (what
mesh_render_func obtains with 33 separate renderer calls).
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib

DEFAULT_NEAR = 0.1      # upstream rasterize.py defaults, used by mode='depth' (render_depth passes no near/far)
DEFAULT_FAR = 100.0
RASTERIZER_EPS = 1e-3


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sln_b200.neural_renderer runs on CUDA (sm_100a) only; got a %s tensor. No CPU fallback." % t.device)


class _Raster(object):
    """One projected + set-up mesh living in a caller-owned workspace tensor."""

    def __init__(self, vertices, faces, K, R, t, orig_size, image_size, fill_back):
        lib = _lib.load()
        self.lib = lib
        self.dev = vertices.device
        self.V, self.F = vertices.size(0), faces.size(0)
        self.fill_back, self.n, self.orig = int(fill_back), image_size, float(orig_size)
        self.vertices, self.faces = vertices, faces
        self.K, self.R, self.t = K, R, t
        nbytes = lib.sln_raster_workspace_bytes(self.V, self.F, self.fill_back)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        self.st = _lib.cur_stream(self.dev)
        _lib.check(lib.sln_raster_setup(vertices.data_ptr(), self.V, faces.data_ptr(), self.F, self.fill_back, K.data_ptr(), R.data_ptr(),
                                        t.data_ptr(), self.orig, image_size, self.ws.data_ptr(), nbytes, self.st), "raster_setup")

    @property
    def F2(self):
        return self.F * (2 if self.fill_back else 1)

    def forward(self, near, far):
        n = self.n
        fi = torch.empty(n, n, dtype=torch.int32, device=self.dev)
        w = torch.empty(n, n, 3, dtype=torch.float32, device=self.dev)
        d = torch.empty(n, n, dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.sln_raster_forward(self.ws.data_ptr(), self.V, self.F, self.fill_back, n, float(near), float(far), fi.data_ptr(),
                                               w.data_ptr(), d.data_ptr(), self.st), "raster_forward")
        return fi, w, d

    def forward2(self, near_a, near_b, far):
        """Two z-buffers (near planes near_a / near_b) from one pass over the faces."""
        n = self.n
        out = []
        for _ in range(2):
            out.append((torch.empty(n, n, dtype=torch.int32, device=self.dev), torch.empty(n, n, 3, dtype=torch.float32, device=self.dev),
                        torch.empty(n, n, dtype=torch.float32, device=self.dev)))
        (fa, wa, da), (fb, wb, db) = out
        _lib.check(self.lib.sln_raster_forward2(self.ws.data_ptr(), self.V, self.F, self.fill_back, n, float(near_a), float(near_b), float(far),
                                                fa.data_ptr(), wa.data_ptr(), da.data_ptr(), fb.data_ptr(), wb.data_ptr(), db.data_ptr(), self.st),
                   "raster_forward2")
        return out[0], out[1]

    def face_arrays(self):
        pv, fv, finv = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(self.lib.sln_raster_face_arrays(self.ws.data_ptr(), self.V, self.F, self.fill_back, ctypes.byref(pv), ctypes.byref(fv),
                                                   ctypes.byref(finv)), "face_arrays")
        base = self.ws.data_ptr()

        def view(p, count):
            return self.ws[p.value - base: p.value - base + 4 * count].view(torch.float32)
        return view(pv, 3 * self.V).view(self.V, 3), view(fv, 9 * self.F2).view(self.F2, 9), view(finv, 9 * self.F2).view(self.F2, 9)

    def vertex_grad(self, grad_faces):
        scratch = torch.empty(self.V, 3, dtype=torch.int64, device=self.dev)     # 64-bit fixed-point accumulators (deterministic scatter)
        gv = torch.empty(self.V, 3, dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.sln_raster_vertex_grad(self.ws.data_ptr(), self.vertices.data_ptr(), self.V, self.faces.data_ptr(), self.F,
                                                   self.fill_back, self.K.data_ptr(), self.R.data_ptr(), self.t.data_ptr(), self.orig,
                                                   grad_faces.data_ptr(), scratch.data_ptr(), gv.data_ptr(), self.st), "vertex_grad")
        return gv


def _prep(vertices, faces, K, R, t):
    _req_cuda(vertices, faces, K, R, t)
    if vertices.dim() != 3 or vertices.size(0) != 1 or faces.dim() != 3 or faces.size(0) != 1:
        raise NotImplementedError("sln_b200.neural_renderer renders one mesh per call (batch size 1), as the reference does")
    v = vertices[0].contiguous().float()
    f = faces[0].contiguous().to(torch.int32)
    return v, f, K.reshape(-1).contiguous().float(), R.reshape(-1).contiguous().float(), t.reshape(-1).contiguous().float()


class _DepthFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertices, faces, K, R, t, cfg):
        v, f, Kf, Rf, tf = _prep(vertices, faces, K, R, t)
        r = _Raster(v, f, Kf, Rf, tf, cfg["orig_size"], cfg["image_size"], cfg["fill_back"])
        fi, w, d = r.forward(cfg["near_depth"], cfg["far"])
        ctx.r, ctx.maps = r, (fi, w, d)
        return d.flip(0)[None]                      # rows flipped on output (upstream rasterize.py)

    @staticmethod
    def backward(ctx, grad_depth):
        r, (fi, w, d) = ctx.r, ctx.maps
        g = grad_depth[0].flip(0).contiguous().float()
        gf = torch.zeros(r.F2, 9, dtype=torch.float32, device=r.dev)
        _lib.check(r.lib.sln_raster_backward_depth(r.ws.data_ptr(), r.V, r.F, r.fill_back, r.n, fi.data_ptr(), w.data_ptr(), d.data_ptr(),
                                                   g.data_ptr(), gf.data_ptr(), _lib.cur_stream(r.dev)), "backward_depth")
        return r.vertex_grad(gf)[None], None, None, None, None, None


class _RgbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertices, faces, textures, K, R, t, cfg):
        v, f, Kf, Rf, tf = _prep(vertices, faces, K, R, t)
        _req_cuda(textures)
        if textures.dim() != 6 or textures.size(0) != 1 or textures.size(1) != f.size(0):
            raise ValueError("textures must be [1, F, ts, ts, ts, 3]")
        tex = textures[0].contiguous().float()
        ts = tex.size(1)
        r = _Raster(v, f, Kf, Rf, tf, cfg["orig_size"], cfg["image_size"], cfg["fill_back"])
        fi, w, d = r.forward(cfg["near"], cfg["far"])
        rgb = torch.empty(r.n, r.n, 3, dtype=torch.float32, device=r.dev)
        _lib.check(r.lib.sln_raster_texture_sample(r.ws.data_ptr(), r.V, r.F, r.fill_back, r.n, tex.data_ptr(), ts, RASTERIZER_EPS,
                                                   fi.data_ptr(), w.data_ptr(), d.data_ptr(), rgb.data_ptr(), r.st), "texture_sample")
        ctx.r, ctx.maps, ctx.rgb = r, (fi, w, d), rgb
        return rgb.flip(0).permute(2, 0, 1)[None].contiguous()

    @staticmethod
    def backward(ctx, grad_rgb):
        r, (fi, w, d), rgb = ctx.r, ctx.maps, ctx.rgb
        g = grad_rgb[0].permute(1, 2, 0).flip(0).contiguous().float()
        gf = torch.zeros(r.F2, 9, dtype=torch.float32, device=r.dev)
        _lib.check(r.lib.sln_raster_backward_rgb(r.ws.data_ptr(), r.V, r.F, r.fill_back, r.n, RASTERIZER_EPS, fi.data_ptr(), rgb.data_ptr(),
                                                 g.data_ptr(), gf.data_ptr(), _lib.cur_stream(r.dev)), "backward_rgb")
        return r.vertex_grad(gf)[None], None, None, None, None, None, None


class _SceneFn(torch.autograd.Function):
    """depth [1,is,is] + class images [n_cls,is,is] of one scene from a single set-up mesh."""

    @staticmethod
    def forward(ctx, vertices, faces, face_cls, n_cls, K, R, t, cfg):
        v, f, Kf, Rf, tf = _prep(vertices, faces, K, R, t)
        r = _Raster(v, f, Kf, Rf, tf, cfg["orig_size"], cfg["image_size"], cfg["fill_back"])
        cls = face_cls.contiguous().to(torch.int32)
        if cls.numel() != r.F:
            raise ValueError("face_cls must have one entry per face")
        cls2 = torch.cat([cls, cls]) if r.fill_back else cls
        if cfg["near"] == cfg["near_depth"]:
            dmaps = cmaps = r.forward(cfg["near_depth"], cfg["far"])
        else:
            dmaps, cmaps = r.forward2(cfg["near_depth"], cfg["near"], cfg["far"])
        sval = torch.empty(r.n, r.n, dtype=torch.float32, device=r.dev)
        images = torch.empty(n_cls, r.n, r.n, dtype=torch.float32, device=r.dev)
        _lib.check(r.lib.sln_scene_classes_fwd(r.ws.data_ptr(), r.V, r.F, r.fill_back, r.n, cfg["texture_size"], RASTERIZER_EPS,
                                               cmaps[0].data_ptr(), cmaps[1].data_ptr(), cmaps[2].data_ptr(), cls2.data_ptr(), n_cls,
                                               sval.data_ptr(), images.data_ptr(), r.st), "scene_classes_fwd")
        ctx.r, ctx.dmaps, ctx.cmaps, ctx.cls2, ctx.sval, ctx.n_cls = r, dmaps, cmaps, cls2, sval, n_cls
        ctx.face_index = cmaps[0]
        return dmaps[2].flip(0)[None], images

    @staticmethod
    def backward(ctx, grad_depth, grad_images):
        r = ctx.r
        st = _lib.cur_stream(r.dev)
        gf = torch.zeros(r.F2, 9, dtype=torch.float32, device=r.dev)
        if grad_images is not None:
            gi = grad_images.flip(1).contiguous().float()
            _lib.check(r.lib.sln_scene_classes_bwd(r.ws.data_ptr(), r.V, r.F, r.fill_back, r.n, RASTERIZER_EPS, ctx.cmaps[0].data_ptr(),
                                                   ctx.cls2.data_ptr(), ctx.n_cls, ctx.sval.data_ptr(), gi.data_ptr(), gf.data_ptr(), st),
                       "scene_classes_bwd")
        if grad_depth is not None:
            fi, w, d = ctx.dmaps
            g = grad_depth[0].flip(0).contiguous().float()
            _lib.check(r.lib.sln_raster_backward_depth(r.ws.data_ptr(), r.V, r.F, r.fill_back, r.n, fi.data_ptr(), w.data_ptr(), d.data_ptr(),
                                                       g.data_ptr(), gf.data_ptr(), st), "backward_depth")
        return r.vertex_grad(gf)[None], None, None, None, None, None, None, None


class Renderer(nn.Module):
    """Same constructor keywords and call convention as upstream ``nr.Renderer`` for the configuration the reference uses.
    Anything else (look/look_at cameras, anti-aliasing, directional light, silhouettes, batches) raises loudly."""

    def __init__(self, image_size=256, anti_aliasing=True, background_color=[0, 0, 0], fill_back=True, camera_mode='projection',
                 K=None, R=None, t=None, dist_coeffs=None, orig_size=1024, perspective=True, viewing_angle=30,
                 camera_direction=[0, 0, 1], near=0.1, far=100, light_intensity_ambient=0.5, light_intensity_directional=0.5,
                 light_color_ambient=[1, 1, 1], light_color_directional=[1, 1, 1], light_direction=[0, 1, 0]):
        super(Renderer, self).__init__()
        if camera_mode != 'projection':
            raise NotImplementedError("camera_mode=%r (the reference only uses 'projection')" % (camera_mode,))
        if anti_aliasing:
            raise NotImplementedError("anti_aliasing=True (the reference renders with anti_aliasing=False)")
        if dist_coeffs is not None and float(torch.as_tensor(dist_coeffs).abs().sum()) != 0.0:
            raise NotImplementedError("lens distortion is removed per the reference's README.md:13-18")
        if light_intensity_directional != 0.0 or light_intensity_ambient != 1.0 or list(light_color_ambient) != [1, 1, 1]:
            raise NotImplementedError("lighting other than ambient 1.0 / directional 0.0 (reference diff_render.py:361)")
        if list(background_color) != [0, 0, 0]:
            raise NotImplementedError("non-black background")
        self.image_size, self.anti_aliasing, self.fill_back = image_size, anti_aliasing, fill_back
        self.camera_mode, self.K, self.R, self.t, self.orig_size = camera_mode, K, R, t, orig_size
        self.near, self.far = near, far
        self.depth_near, self.depth_far = DEFAULT_NEAR, DEFAULT_FAR    # upstream: render_depth ignores the ctor's near/far
        self.texture_size = 2

    def _cfg(self, orig_size=None):
        return dict(orig_size=self.orig_size if orig_size is None else orig_size, image_size=self.image_size, fill_back=self.fill_back,
                    near=float(self.near), far=float(self.far), near_depth=float(self.depth_near), texture_size=self.texture_size)

    def forward(self, vertices, faces, textures=None, mode=None, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        K = self.K if K is None else K
        R = self.R if R is None else R
        t = self.t if t is None else t
        if mode == 'depth':
            cfg = self._cfg(orig_size)
            cfg["far"] = float(self.depth_far)
            return _DepthFn.apply(vertices, faces, K, R, t, cfg)
        if mode == 'rgb':
            if textures is None:
                raise ValueError("mode='rgb' needs textures")
            return _RgbFn.apply(vertices, faces, textures, K, R, t, self._cfg(orig_size))
        raise NotImplementedError("mode=%r (the reference calls mode='depth' and mode='rgb' only)" % (mode,))


def render_scene_classes(vertices, faces, face_cls, n_cls, K, R, t, image_size=256, orig_size=512, near=0.001, far=100.0,
                         fill_back=True, texture_size=2):
    """vertices [1,V,3], faces [1,F,3], face_cls [F] (class id per face, < n_cls) ->
    (depth [1,is,is] exactly as Renderer(mode='depth'), images [n_cls,is,is] where images[c] is exactly
    ``torch.sum(Renderer(mode='rgb')(v, f, textures_c), dim=1)[0] / 3`` for the 0/1 texture of class c)."""
    cfg = dict(orig_size=orig_size, image_size=image_size, fill_back=fill_back, near=float(near), far=float(far),
               near_depth=float(DEFAULT_NEAR), texture_size=texture_size)
    return _SceneFn.apply(vertices, faces, face_cls, n_cls, K, R, t, cfg)
