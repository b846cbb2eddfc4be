"""Drop-in for the hot part of the reference's ``models/diff_render.py``: ``get_cam_mat`` (:13-46) and ``mesh_render_func``
(:48-435) with the same signatures and return values.

What differs from the reference, by design:
  * the SUNCG mesh retrieval (`suncg_retrieve`, `load_suncg_obj`, models/misc.py — needs the SUNCG dataset, pywavefront,
    pymesh) is replaced by a resident ``MeshLibrary`` of synthetic meshes (data/synthetic_meshes.py); set another provider
    with ``set_mesh_library``.  Meshes stay on the device (the reference re-uploads every mesh on every call, misc.py:118);
  * the per-object Python loop of 4x4 transforms (:76-159) is one batched tensor expression per scene;
  * depth + the 32 per-class mask renders (:366-431, 33 renderer calls on identical geometry) come from ONE rasterization
    (neural_renderer.render_scene_classes) and the per-class loop of the compositing is vectorised over classes.
The result tensor has the reference's layout: [1, 1 + 40 + (n_classes - 3), 256, 256] = depth | one-hot(40) | depth planes.
"""
import ctypes
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from .. import neural_renderer as nr
from ..data.synthetic import OBJECT_NAMES
from ..data.synthetic_meshes import MeshLibrary, room_shell

nyu_class = ['wall', 'floor', 'cabinet', 'bed', 'chair', 'sofa', 'table', 'door', 'window', 'bookshelf', 'picture', 'counter', 'blinds',
             'desk', 'shelves', 'curtain', 'dresser', 'pillow', 'mirror', 'floor mat', 'clothes', 'ceiling', 'books', 'refridgerator',
             'television', 'paper', 'towel', 'shower curtain', 'box', 'whiteboard', 'person', 'night stand', 'toilet', 'sink', 'lamp',
             'bathtub', 'bag', 'otherstructure', 'otherfurniture', 'otherprop']
inter_out = 512
final_out = 256
object_idx_to_name = list(OBJECT_NAMES)
SKIPPED_TYPES = ["wall", "ceiling", "floor", "person", "door", "window", "curtain", "blinds"]   # reference :88

_LIBRARY = None


def set_mesh_library(lib):
    global _LIBRARY
    _LIBRARY = lib


def mesh_library(device):
    global _LIBRARY
    if _LIBRARY is None:
        _LIBRARY = MeshLibrary()
    return _LIBRARY.to(device)


def desired_classes():
    """Class order of the mask renders (reference :65-69,372-374): sorted(31 types + ceiling, floor, wall), wall first."""
    names = sorted(set(object_idx_to_name[1:] + ['ceiling', 'floor', 'wall']))
    names.remove("wall")
    names.insert(0, "wall")
    return names


def get_cam_mat(boxes):
    """K [1,3,3], R [1,3,3], t [1,1,3] on the device of ``boxes`` (reference :13-46)."""
    dev = boxes[-1].device
    theta_rot = -0.4
    fl_pix = 400
    int_mat = torch.tensor([[fl_pix * inter_out / 1024, 0, inter_out / 2.0], [0, fl_pix * inter_out / 1024, inter_out / 2.0],
                            [0, 0, 1.0]], dtype=torch.float32)[None]
    rot_w2c = torch.from_numpy(np.array([[1, 0, 0], [0, np.cos(theta_rot), np.sin(theta_rot)],
                                         [0, -np.sin(theta_rot), np.cos(theta_rot)]], dtype="float32"))
    room = boxes[-1].detach().float().cpu()
    cam = torch.zeros(3, 1)
    cam[0, 0] = room[3] / 2.0
    cam[1, 0] = room[4] / 2.0 + min(0.1, abs(float(room[4]) / 2.0))
    cam[2, 0] = room[5]
    t_w2c = torch.matmul(rot_w2c, -cam)
    cam2cv = torch.tensor([[1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=torch.float)
    R = torch.matmul(cam2cv, rot_w2c)
    t = torch.matmul(cam2cv, t_w2c)
    return int_mat.to(dev), R.reshape(1, 3, 3).to(dev), t.reshape(1, 1, 3).to(dev)


def assemble_scene(boxes, angles, objs, library, model_ids=None):
    """Per-object similarity transforms (reference :76-159) for all objects at once.  model_ids: optional mesh id per object row
    (the cached retrieval of an earlier iteration, reference :80-82); default = the class's canonical mesh.

    boxes: sequence of [6] tensors (objects normalised to the room; last row = room), angles: sequence of scalars (0..24),
    objs: class ids.  Returns vertices [1,V,3] (differentiable w.r.t. boxes and angles), faces [1,F,3] int32, face_cls [F]
    int32 (index into desired_classes()), valid object indices and their sizes."""
    dev = boxes[-1].device
    names = desired_classes()
    room = boxes[-1][3:]
    verts, faces, cls, kept, sizes = [], [], [], [], []
    off = 0
    for i in range(len(boxes) - 1):
        mtype = object_idx_to_name[int(objs[i])]
        if mtype in SKIPPED_TYPES:
            continue
        kept.append(i)
    if kept:
        B = torch.stack([boxes[i] for i in kept])                       # [n,6]
        A = torch.stack([torch.as_tensor(angles[i], device=dev, dtype=torch.float32).reshape(()) for i in kept])
        bmin, bmax = B[:, :3] * room, B[:, 3:] * room
        center, size = (bmax + bmin) / 2, bmax - bmin
        models = [library.get(object_idx_to_name[int(objs[i])] if model_ids is None else model_ids[i]) for i in kept]
        msize = torch.stack([m["size"] for m in models])
        mcent = torch.stack([m["center"] for m in models])
        scale = (size / msize).min(dim=1).values                          # :106
        theta = -A * (2 * math.pi / 24)
        c, s = torch.cos(theta), torch.sin(theta)
        zero, one = torch.zeros_like(c), torch.ones_like(c)
        rot = torch.stack([torch.stack([c, zero, s], -1), torch.stack([zero, one, zero], -1), torch.stack([-s, zero, c], -1)], 1)   # :107-113
        trans = center - scale[:, None] * torch.einsum("nij,nj->ni", rot, mcent)                                                  # :115
        for j, m in enumerate(models):
            v = scale[j] * torch.matmul(m["vertices"], rot[j].t()) + trans[j]
            verts.append(v)
            faces.append(m["faces"] + off)
            cls.append(torch.full((m["faces"].size(0),), names.index(object_idx_to_name[int(objs[kept[j]])]), dtype=torch.int32, device=dev))
            off += v.size(0)
            sizes.append(size[j])
    shell = room_shell(boxes[-1][3:].detach().cpu())
    for name in ("wall", "floor", "ceiling"):
        v, f = shell[name]
        verts.append(v.to(dev)); faces.append(f.to(dev) + off)
        cls.append(torch.full((f.size(0),), names.index(name), dtype=torch.int32, device=dev))
        off += v.size(0)
    vertices = torch.cat(verts)[None]
    face_buf = torch.cat(faces).to(torch.int32)[None]
    return vertices, face_buf, torch.cat(cls), kept, sizes


def cull_faces(vertices, face_buf, face_cls, R, t, eps=0.06):
    """Drop faces with any vertex closer than eps in front of the camera plane (reference :345-356)."""
    vc = torch.matmul(vertices, R.transpose(1, 2)) + t
    z = vc[0, :, 2]
    fz = z[face_buf[0].long()]
    valid = ~(fz < eps).any(dim=1)
    return face_buf[:, valid, :].detach(), face_cls[valid]


def composite(depth_data, images, names, index=None, keep=None):
    """Reference :366-434 vectorised over the classes.  depth_data [1,H,W], images [C,H,W] (class order `names`).
    index / keep: optional precomputed device tensors (one-hot channel of every class; classes that get a depth plane), so that
    the call contains no host->device transfer (CUDA-graph capture)."""
    depth_data = torch.where(depth_data > 15, torch.full_like(depth_data, -1.0), depth_data)           # :367
    C = images.size(0)
    hard = images.detach() > 0.1                                                                        # :401
    cnt = hard.sum(dim=(1, 2))
    sums = (depth_data * hard).sum(dim=(1, 2))
    # the reference's torch.mean over an empty selection is nan and is then replaced by wall_max (:411-419); dividing by
    # max(cnt, 1) gives the same forward value after the replacement without a 0/0 in the backward pass
    mean = sums / cnt.clamp(min=1)
    wall = names.index("wall")
    wall_depth = torch.where(hard[wall], depth_data[0], torch.full_like(depth_data[0], -float("inf")))
    wall_max = wall_depth.max().detach()
    wall_max = torch.where(cnt[wall] > 0, wall_max, torch.full_like(wall_max, 10.0))                    # :408-410
    mean = torch.where(cnt > 0, mean, wall_max.expand_as(mean))                                         # :411-419
    planes = torch.where(hard, depth_data / wall_max, (mean / wall_max)[:, None, None].expand(-1, *depth_data.shape[1:]))   # :420-421
    if keep is None:
        keep = torch.tensor([i for i, n in enumerate(names) if n not in ("wall", "floor", "ceiling")], device=depth_data.device)   # :422-425
    one_hot = torch.zeros(41, *depth_data.shape[1:], device=depth_data.device, dtype=depth_data.dtype)
    if index is None:
        index = torch.tensor([nyu_class.index(n.replace("_", " ")) + 1 for n in names], device=depth_data.device)
    one_hot = one_hot.index_copy(0, index, images)                                                      # :429-431
    return torch.cat((depth_data, one_hot[1:], planes.index_select(0, keep)), dim=0)[None]              # :433-434


class _CompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, images, st):
        lib = _lib.load()
        dev = depth.device
        d, im = depth.contiguous().float(), images.contiguous().float()
        C, P = im.size(0), im.size(-1) * im.size(-2)
        n_keep, n_onehot = st.keep32.numel(), st.inv_index32.numel()
        out = torch.empty(1, n_onehot + n_keep, im.size(-2), im.size(-1), device=dev, dtype=torch.float32)
        stats = torch.empty(2 * C + 1, device=dev, dtype=torch.float32)
        ws = torch.empty(lib.sln_composite_workspace_bytes(max(C, n_keep)), dtype=torch.uint8, device=dev)
        _lib.check(lib.sln_composite_fwd(d.data_ptr(), im.data_ptr(), C, P, st.wall, st.inv_index32.data_ptr(), n_onehot, _lib.ptr(st.keep32), n_keep,
                                         out.data_ptr(), stats.data_ptr(), ws.data_ptr(), ws.numel(), _lib.cur_stream(dev)), "composite_fwd")
        ctx.save_for_backward(d, im, stats)
        ctx.st, ctx.ws = st, ws
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        d, im, stats = ctx.saved_tensors
        st, dev = ctx.st, d.device
        C, P = im.size(0), im.size(-1) * im.size(-2)
        g = g_out.contiguous().float()
        g_depth, g_images = torch.empty_like(d), torch.empty_like(im)
        _lib.check(lib.sln_composite_bwd(d.data_ptr(), im.data_ptr(), g.data_ptr(), C, P, st.index32.data_ptr(), st.inv_index32.numel(), _lib.ptr(st.keep32),
                                         st.keep32.numel(), stats.data_ptr(), g_depth.data_ptr(), g_images.data_ptr(), ctx.ws.data_ptr(), ctx.ws.numel(),
                                         _lib.cur_stream(dev)), "composite_bwd")
        return g_depth, g_images, None


def composite_fused(depth_data, images, static):
    """composite() as one library call each way (csrc/scene.cu); `static` carries the class -> channel tables of the scene."""
    if depth_data.device.type != "cuda":
        raise RuntimeError("sln_b200 compositing runs on CUDA only (composite() is the torch restatement)")
    return _CompositeFn.apply(depth_data, images, static)


def room_metadata(room):
    """The wall / floor record of the room shell in the reference's wall_data_wfc.json format (what `wall_retrieve` /
    `floor_retrieve` return and `model_ids_return["wall"/"floor"]` cache, reference :171-175,243-247)."""
    X, Y, Z = [float(v) for v in room]
    return dict(house_id="synthetic", model_id="room", wall_bbox_min=[0.0, 0.0, 0.0], wall_bbox_max=[X, Y, Z],
                floor_bbox_min=[0.0, 0.0, 0.0], floor_bbox_max=[X, 0.0, Z])


def mesh_render_func(boxes, angles, objs, model_ids_old=None, obj_size_target=None):
    """Same contract as the reference: -> (final [1,70,256,256], model_ids_return, obj_size_return, size_loss).

    Like the reference (:56-57) a later iteration (`model_ids_old` given) overwrites the caller's ``boxes[-1]`` IN PLACE with the
    cached room box; `model_ids_return` holds `box_info`, one id per object row (recorded before the skip test, :84-89) and the
    `wall` / `floor` records (first iteration only); `model_ids_old[idx]` selects the mesh on later iterations."""
    dev = boxes[-1].device
    if dev.type != "cuda":
        raise RuntimeError("sln_b200 mesh_render_func runs on CUDA only (no CPU fallback)")
    model_ids_return, obj_size_return = {}, []
    size_loss = 0.0
    old_wall = boxes[-1].clone()
    if model_ids_old is not None:
        boxes[-1] = torch.from_numpy(model_ids_old["box_info"]).float().to(dev)     # :56-57 (in place, as the reference)
    else:
        model_ids_return["box_info"] = boxes[-1].detach().cpu().numpy()            # :60
    lib = mesh_library(dev)
    n_obj = len(boxes) - 1
    if model_ids_old is None:
        for i in range(n_obj):
            model_ids_return[i] = object_idx_to_name[int(objs[i])]                  # :84-89 (one canonical mesh per class)
        ids = None
    else:
        ids = [model_ids_old[i] for i in range(n_obj)]
    vertices, face_buf, face_cls, kept, sizes = assemble_scene(boxes, angles, objs, lib, ids)
    for j, i in enumerate(kept):
        if obj_size_target is not None:
            size_loss = size_loss + nn.functional.mse_loss(sizes[j], torch.from_numpy(obj_size_target[j]).float().to(dev))   # :98
        else:
            obj_size_return.append(sizes[j].detach().cpu().numpy())
    if obj_size_target is not None:
        size_loss = size_loss + nn.functional.mse_loss(old_wall, torch.from_numpy(obj_size_target[-1]).float().to(dev))      # :164
    else:
        obj_size_return.append(boxes[-1].detach().cpu().numpy())
    if model_ids_old is None:
        model_ids_return["wall"] = room_metadata(boxes[-1][3:].detach().cpu())      # :171-175
        model_ids_return["floor"] = dict(model_ids_return["wall"])                  # :243-247
    K, R, t = get_cam_mat(boxes)
    face_buf, face_cls = cull_faces(vertices, face_buf, face_cls, R, t)
    names = desired_classes()
    depth, images = nr.render_scene_classes(vertices, face_buf, face_cls, len(names), K, R, t, image_size=final_out, orig_size=inter_out,
                                            near=0.001)
    return composite(depth, images, names), model_ids_return, obj_size_return, size_loss


# ------------------------------------------------------------------------------------------------ static-scene fast path
class SceneStatic(object):
    """Everything of a scene that does not change during layout refinement (object classes, retrieved meshes, room box, camera):
    built once, then ``render_static`` is pure device-tensor arithmetic on the layout — no Python per-object loop, no host
    synchronisation, fixed shapes — so a whole refinement iteration can be captured in a CUDA graph (``RefineStep``)."""

    def __init__(self, objs, room_box, library, device):
        dev = torch.device(device)
        names = desired_classes()
        self.dev, self.names = dev, names
        self.objs = [int(o) for o in objs]
        self.kept = [i for i in range(len(self.objs) - 1) if object_idx_to_name[self.objs[i]] not in SKIPPED_TYPES]
        models = [library.get(object_idx_to_name[self.objs[i]]) for i in self.kept]
        self.n_kept = len(self.kept)
        self.kept_idx = torch.tensor(self.kept, dtype=torch.long, device=dev)
        faces, cls, off = [], [], 0
        mv, vobj = [], []
        for j, m in enumerate(models):
            nv = m["vertices"].size(0)
            mv.append(m["vertices"].to(dev)); vobj.append(torch.full((nv,), j, dtype=torch.long, device=dev))
            faces.append(m["faces"].to(dev) + off)
            cls.append(torch.full((m["faces"].size(0),), names.index(object_idx_to_name[self.objs[self.kept[j]]]), dtype=torch.int32, device=dev))
            off += nv
        self.mv = torch.cat(mv) if mv else torch.zeros(0, 3, device=dev)
        self.vobj = torch.cat(vobj) if vobj else torch.zeros(0, dtype=torch.long, device=dev)
        self.msize = torch.stack([m["size"].to(dev) for m in models]) if models else torch.zeros(0, 3, device=dev)
        self.mcent = torch.stack([m["center"].to(dev) for m in models]) if models else torch.zeros(0, 3, device=dev)
        room_box = room_box.detach().float().cpu()
        shell = room_shell(room_box[3:])
        sv = []
        for name in ("wall", "floor", "ceiling"):
            v, f = shell[name]
            sv.append(v.to(dev)); faces.append(f.to(dev) + off)
            cls.append(torch.full((f.size(0),), names.index(name), dtype=torch.int32, device=dev))
            off += v.size(0)
        self.shell_v = torch.cat(sv)
        self.faces = torch.cat(faces).to(torch.int32)[None].contiguous()
        self.face_cls = torch.cat(cls).contiguous()
        self.room = room_box.to(dev)
        K, R, t = get_cam_mat([self.room])
        self.K, self.R, self.t = K, R, t
        self.index = torch.tensor([nyu_class.index(n.replace("_", " ")) + 1 for n in names], device=dev)
        self.keep = torch.tensor([i for i, n in enumerate(names) if n not in ("wall", "floor", "ceiling")], device=dev)
        self.wall = names.index("wall")
        # class -> channel tables of the fused compositing (csrc/scene.cu)
        self.index32 = self.index.to(torch.int32).contiguous()
        inv = [-1] * 41
        for c, ch in enumerate(self.index.tolist()):
            inv[ch] = c
        self.inv_index32 = torch.tensor(inv, dtype=torch.int32, device=dev)
        self.keep32 = self.keep.to(torch.int32).contiguous()
        # int32 / host mirrors for the fused assembly (csrc/scene.cu)
        self.kept32 = self.kept_idx.to(torch.int32)
        self.vobj32 = self.vobj.to(torch.int32)
        counts = [m["vertices"].size(0) for m in models]
        self.vstart32 = torch.tensor([sum(counts[:j]) for j in range(len(counts) + 1)], dtype=torch.int32, device=dev)
        r2k = [-1] * len(self.objs)
        for j, r in enumerate(self.kept):
            r2k[r] = j
        self.row_to_kept32 = torch.tensor(r2k, dtype=torch.int32, device=dev)
        self.room3_c = (ctypes.c_float * 3)(*[float(v) for v in room_box[3:]])
        self.mv, self.shell_v, self.msize, self.mcent = [t.contiguous().float() for t in (self.mv, self.shell_v, self.msize, self.mcent)]
        self.R9, self.t3 = self.R.reshape(9).contiguous().float(), self.t.reshape(3).contiguous().float()

    def vertices(self, boxes, angles):
        """[1, V, 3] world-space vertices, differentiable w.r.t. boxes [n+1, 6] and angles [n+1] (reference diff_render.py:76-159)."""
        room = self.room[3:]
        B = boxes.index_select(0, self.kept_idx)
        A = angles.index_select(0, self.kept_idx).float()
        bmin, bmax = B[:, :3] * room, B[:, 3:] * room
        center, size = (bmax + bmin) / 2, bmax - bmin
        scale = (size / self.msize).min(dim=1).values
        theta = -A * (2 * math.pi / 24)
        c, s = torch.cos(theta), torch.sin(theta)
        zero, one = torch.zeros_like(c), torch.ones_like(c)
        rot = torch.stack([torch.stack([c, zero, s], -1), torch.stack([zero, one, zero], -1), torch.stack([-s, zero, c], -1)], 1)
        trans = center - scale[:, None] * torch.einsum("nij,nj->ni", rot, self.mcent)
        v = scale[self.vobj, None] * torch.einsum("vij,vj->vi", rot[self.vobj], self.mv) + trans[self.vobj]
        return torch.cat([v, self.shell_v])[None], size

    def assemble(self, boxes, angles, eps=0.06, refine_hooks=False):
        """vertices() + culled_faces() as one library call each way (csrc/scene.cu): -> (vertices [1,V,3], size [n_kept,3],
        faces [1,F,3] int32).  Differentiable w.r.t. boxes and angles.  refine_hooks: the backward kernel also applies the
        refinement loop's gradient hooks (fix_grad on the boxes, quad_grad on the angles; test_render_refine.py:220-230,288,297)."""
        return _AssembleFn.apply(boxes, angles, self, float(eps), bool(refine_hooks))

    def culled_faces(self, vertices, eps=0.06):
        """Reference :345-356 without a dynamic shape: a face with any vertex closer than eps keeps its slot but collapses to a
        zero-area triangle (vertex 0 three times), which the rasterizer never draws."""
        vc = torch.matmul(vertices.detach(), self.R.transpose(1, 2)) + self.t
        fz = vc[0, :, 2][self.faces[0].long()]
        valid = ~(fz < eps).any(dim=1)
        return torch.where(valid[None, :, None], self.faces, torch.zeros_like(self.faces))


class _AssembleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, boxes, angles, st, eps, refine_hooks=False):
        lib = _lib.load()
        dev = st.dev
        ctx.hooks = refine_hooks
        b, a = boxes.contiguous().float(), angles.contiguous().float()
        if b.device.type != "cuda":
            raise RuntimeError("sln_b200 scene assembly runs on CUDA only (no CPU fallback)")
        n_rows, n_obj, n_shell, F = b.size(0), st.mv.size(0), st.shell_v.size(0), st.faces.size(1)
        verts = torch.empty(1, n_obj + n_shell, 3, device=dev, dtype=torch.float32)
        size = torch.empty(st.n_kept, 3, device=dev, dtype=torch.float32)
        faces = torch.empty_like(st.faces)
        ws = torch.empty(lib.sln_scene_assemble_workspace_bytes(st.n_kept), dtype=torch.uint8, device=dev)
        _lib.check(lib.sln_scene_assemble_fwd(b.data_ptr(), a.data_ptr(), n_rows, _lib.ptr(st.kept32), st.n_kept, st.room3_c, _lib.ptr(st.mv),
                                              _lib.ptr(st.vobj32), n_obj, _lib.ptr(st.shell_v), n_shell, _lib.ptr(st.msize), _lib.ptr(st.mcent),
                                              _lib.ptr(st.faces), F, st.R9.data_ptr(), st.t3.data_ptr(), eps, verts.data_ptr(),
                                              _lib.ptr(size), _lib.ptr(faces), ws.data_ptr(), ws.numel(), _lib.cur_stream(dev)), "scene_assemble_fwd")
        ctx.st, ctx.ws, ctx.n_rows = st, ws, n_rows
        ctx.mark_non_differentiable(faces)
        return verts, size, faces

    @staticmethod
    def backward(ctx, g_verts, g_size, _g_faces):
        lib = _lib.load()
        st, dev = ctx.st, ctx.st.dev
        n_obj, n_shell = st.mv.size(0), st.shell_v.size(0)
        gv = torch.zeros(n_obj + n_shell, 3, device=dev) if g_verts is None else g_verts.reshape(-1, 3).contiguous().float()
        gs = None if g_size is None else g_size.contiguous().float()
        d_boxes = torch.empty(ctx.n_rows, 6, device=dev, dtype=torch.float32)
        d_angles = torch.empty(ctx.n_rows, device=dev, dtype=torch.float32)
        _lib.check(lib.sln_scene_assemble_bwd(gv.data_ptr(), gs.data_ptr() if gs is not None else None, ctx.n_rows, _lib.ptr(st.row_to_kept32),
                                              st.n_kept, st.room3_c, _lib.ptr(st.mv), _lib.ptr(st.vstart32), _lib.ptr(st.msize), _lib.ptr(st.mcent),
                                              ctx.ws.data_ptr(), ctx.ws.numel(), 1 if ctx.hooks else 0, 4.0 if ctx.hooks else 1.0, d_boxes.data_ptr(),
                                              d_angles.data_ptr(), _lib.cur_stream(dev)), "scene_assemble_bwd")
        return d_boxes, d_angles, None, None, None


def render_static(static, boxes, angles, fused=True, refine_hooks=False):
    """final [1, 70, 256, 256] of the scene described by `static` at layout (boxes, angles): scene assembly (one library call;
    fused=False: the torch-op restatement of the reference's arithmetic, kept as the checker) + one rasterization + compositing."""
    if fused:
        vertices, size, faces = static.assemble(boxes, angles, refine_hooks=refine_hooks)
    else:
        vertices, size = static.vertices(boxes, angles)
        faces = static.culled_faces(vertices)
    depth, images = nr.render_scene_classes(vertices, faces, static.face_cls, len(static.names), static.K, static.R, static.t,
                                            image_size=final_out, orig_size=inter_out, near=0.001)
    if fused:
        return composite_fused(depth, images, static), size
    return composite(depth, images, static.names, static.index, static.keep), size
