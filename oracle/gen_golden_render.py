"""TEST INFRASTRUCTURE ONLY — golden vectors of the reference's scene assembly / cull / compositing, produced by EXECUTING the
reference's own `get_cam_mat` and `mesh_render_func` (models/diff_render.py:13-46, 48-435) from where they lie.

    python oracle/gen_golden_render.py     # needs /root/reference; writes tests/golden/render_*.npz

models/diff_render.py cannot be imported (its `from models.misc import *` needs pywavefront, pymesh, neural_renderer, the SUNCG
metadata and `np.float`), so the two function defs are pulled out of the file's AST and executed unmodified in a namespace that
supplies what `models.misc` would have supplied:
  * `suncg_retrieve`, `wall_retrieve`, `floor_retrieve`, `get_bbox`: the reference's own defs, AST-extracted from models/misc.py
    (`np.float` is aliased to `float` for numpy >= 1.24);
  * `suncg_data`, `wall_data_json`, `load_suncg_obj`, `load_wall_obj_new`, `load_floor_obj`, `load_ceil_obj`: stubs serving the
    repository's synthetic MeshLibrary / room shell in the reference's metadata format (no SUNCG data exists offline);
  * `nr.Renderer`: a stub with the upstream constructor / call signature whose images and gradients come from the C rasterizer
    oracle (oracle/raster_oracle.c; the third-party core itself stays UNPINNED — see that file's header);
  * `Tensor.cuda()` / `.cpu()` are made plain copies for the duration of the run (the reference hard-codes `.cuda()`).
What this pins by execution: the camera (:13-46), the per-object scale / rotate / translate arithmetic (:76-159), the room shell
transforms (:167-342), the near-plane cull (:344-356), the 33-render compositing loop (:366-434), the return contract
(model_ids / sizes / size_loss, first and later iterations) and — through autograd over the reference's own statements — the
gradient of a seeded functional of the result w.r.t. boxes and angles.
"""
import ast
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import raster_oracle as ro  # noqa: E402
from sln_b200.data.synthetic import OBJECT_NAMES  # noqa: E402
from sln_b200.data.synthetic_meshes import MeshLibrary, room_shell, room_walls, synthetic_layout  # noqa: E402

REF_RENDER = "/root/reference/models/diff_render.py"
REF_MISC = "/root/reference/models/misc.py"
SAMPLE_STRIDE = 97


def _defs(path, names, also_assign=()):
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name in names) or
            (isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id in also_assign for t in n.targets))]
    assert {n.name for n in body if isinstance(n, ast.FunctionDef)} == set(names), path
    return compile(ast.Module(body=body, type_ignores=[]), path, "exec")


class _OracleDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertices, faces, orc):
        v, f = vertices[0].detach().numpy().astype(np.float32), faces[0].numpy().astype(np.int32)
        d, c = orc.depth(v, f)
        ctx.c, ctx.v, ctx.f, ctx.orc = c, v, f, orc
        return torch.from_numpy(d)[None].clone()

    @staticmethod
    def backward(ctx, g):
        return torch.from_numpy(ctx.orc.depth_bwd(ctx.v, ctx.f, ctx.c, g[0].contiguous().numpy()))[None], None, None


class _OracleRgb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vertices, faces, textures, orc):
        v, f = vertices[0].detach().numpy().astype(np.float32), faces[0].numpy().astype(np.int32)
        img, c = orc.rgb(v, f, textures[0].numpy())
        ctx.c, ctx.v, ctx.f, ctx.orc = c, v, f, orc
        return torch.from_numpy(img)[None].clone()

    @staticmethod
    def backward(ctx, g):
        return torch.from_numpy(ctx.orc.rgb_bwd(ctx.v, ctx.f, ctx.c, g[0].contiguous().numpy()))[None], None, None, None


class Recorder(object):
    def __init__(self):
        self.calls = []


def make_stub_nr(rec):
    class Renderer(object):
        """upstream nr.Renderer signature as the reference constructs it (models/diff_render.py:359-361)"""

        def __init__(self, camera_mode, image_size, K, R, t, anti_aliasing, orig_size, near, light_intensity_ambient,
                     light_intensity_directional):
            assert camera_mode == 'projection' and not anti_aliasing and light_intensity_ambient == 1.0 and light_intensity_directional == 0.0
            self.orc = ro.RendererOracle(image_size, K[0].numpy(), R[0].numpy(), t.reshape(3).numpy(), orig_size, near_rgb=near)
            rec.cam = (K.clone(), R.clone(), t.clone())

        def __call__(self, vertices, faces, textures, mode):
            rec.calls.append((mode, vertices.detach().clone(), faces.clone(), textures[0, :, 0, 0, 0, 0].clone()))
            if mode == 'depth':
                return _OracleDepth.apply(vertices, faces, self.orc)
            return _OracleRgb.apply(vertices, faces, textures, self.orc)
    return types.SimpleNamespace(Renderer=Renderer)


def reference_namespace(library, rec, image_size):
    np.float = float                                     # models/misc.py:128-147 (removed from numpy 1.24)
    shell_cache = {}

    def shell_for(data):
        key = tuple(data["wall_bbox_max"])
        if key not in shell_cache:
            shell_cache[key] = room_shell(torch.tensor(key))
        return shell_cache[key]

    def room_entry(room):
        X, Y, Z = [float(v) for v in room]
        return dict(house_id="synthetic", model_id="room", wall_bbox_min=[0.0, 0.0, 0.0], wall_bbox_max=[X, Y, Z],
                    floor_bbox_min=[0.0, 0.0, 0.0], floor_bbox_max=[X, 0.0, Z])

    ns = {"torch": torch, "np": np, "nn": torch.nn, "nr": make_stub_nr(rec), "inter_out": 512, "final_out": image_size,
          "object_idx_to_name": list(OBJECT_NAMES),
          "suncg_data": {name: [dict(library.meta[name])] for name in library.meta},
          "load_suncg_obj": lambda mid: (library.get(mid)["vertices"].clone(), library.get(mid)["faces"].to(torch.int32)),
          "wall_data_json": [],
          "load_wall_obj_new": lambda d: ([w[0].clone() for w in room_walls(d["wall_bbox_max"])], [w[1].to(torch.int32) for w in room_walls(d["wall_bbox_max"])]),
          "load_floor_obj": lambda d: (shell_for(d)["floor"][0].clone(), shell_for(d)["floor"][1].to(torch.int32)),
          "load_ceil_obj": lambda d: (shell_for(d)["ceiling"][0].clone(), shell_for(d)["ceiling"][1].to(torch.int32)),
          "print": lambda *a, **k: (sys.stderr.write("[ref] " + " ".join(str(x) for x in a) + "\n") if os.environ.get("REF_PRINT") else None)}
    exec(_defs(REF_MISC, {"suncg_retrieve", "wall_retrieve", "floor_retrieve", "get_bbox"}), ns)
    exec(_defs(REF_RENDER, {"get_cam_mat", "mesh_render_func"}, also_assign=("nyu_class",)), ns)
    ns["_room_entry"] = room_entry
    return ns


def run_reference(ns, boxes, angles, objs, **kw):
    """-> what mesh_render_func returns; the room's wall/floor metadata is the exact-fit entry of the synthetic shell."""
    ns["wall_data_json"][:] = [ns["_room_entry"](kw.pop("room"))]
    ns["wall_retrieve"].__defaults__ = (ns["wall_data_json"],)
    ns["floor_retrieve"].__defaults__ = (ns["wall_data_json"],)
    return ns["mesh_render_func"](boxes, angles, objs, **kw)


def case(name, layout, nu, nv, image_size, full):
    boxes, angles, objs = layout
    n = boxes.size(0)
    library = MeshLibrary(nu=nu, nv=nv)
    rec = Recorder()
    ns = reference_namespace(library, rec, image_size)
    cuda, cpu = torch.Tensor.cuda, torch.Tensor.cpu
    # device transfers give fresh storage in the reference (GPU <-> host); on this CPU-only run they must too: suncg_retrieve
    # (models/misc.py:35-42) scales the numpy view of `box.cpu()` in place and would otherwise corrupt the caller's boxes
    torch.Tensor.cuda = lambda self, *a, **k: self.clone()
    torch.Tensor.cpu = lambda self, *a, **k: self.clone()
    try:
        K, R, t = ns["get_cam_mat"]([boxes[i] for i in range(n)])
        b = [boxes[i].clone().requires_grad_(i < n - 1) for i in range(n)]
        a = [angles[i].clone().requires_grad_(i < n - 1) for i in range(n)]
        final, ids, sizes, size_loss = run_reference(ns, b, a, objs.tolist(), room=boxes[-1][3:])
        assert size_loss == 0.0
        g = torch.Generator().manual_seed(7)
        W = torch.randn(final.shape, generator=g)
        (final * W).sum().backward()
        gb = torch.stack([x.grad if x.grad is not None else torch.zeros(6) for x in b])
        ga = torch.stack([x.grad if x.grad is not None else torch.zeros(()) for x in a])
        calls1 = rec.calls
        # a later refinement iteration (test_render_refine.py:324): perturbed layout, cached ids + size targets, drifted room row
        rec.calls = []
        b2 = [(boxes[i] + (0.01 * torch.randn(6, generator=g) if i < n - 1 else 0.05)).clone().requires_grad_(True) for i in range(n)]
        a2 = [(angles[i] + (0.3 * torch.randn((), generator=g) if i < n - 1 else 0.0)).clone().requires_grad_(i < n - 1) for i in range(n)]
        b2_in = list(b2)
        final2, ids2, sizes2, size_loss2 = run_reference(ns, b2_in, a2, objs.tolist(), room=boxes[-1][3:], model_ids_old=ids,
                                                         obj_size_target=sizes)
        ((final2 * W).sum() + 2.0 * size_loss2).backward()
        gb2 = torch.stack([x.grad if x.grad is not None else torch.zeros(6) for x in b2])
        ga2 = torch.stack([x.grad if x.grad is not None else torch.zeros(()) for x in a2])
    finally:
        torch.Tensor.cuda, torch.Tensor.cpu = cuda, cpu
    depth_call = calls1[0]
    assert depth_call[0] == 'depth' and len(calls1) == 33
    face_cls = torch.full((depth_call[2].size(1),), -1, dtype=torch.int32)
    for c, call in enumerate(calls1[1:]):
        face_cls[call[3] > 0.5] = c
    assert (face_cls >= 0).all()
    blob = dict(
        meta=np.array(json.dumps(dict(name=name, nu=nu, nv=nv, image_size=image_size, n_rows=n, sample_stride=SAMPLE_STRIDE,
                                      ids_keys=[str(k) for k in ids.keys()], ids2_keys=[str(k) for k in ids2.keys()],
                                      ids_values={str(k): (v if isinstance(v, str) else None) for k, v in ids.items()},
                                      room_overwritten=bool(torch.equal(b2_in[-1], torch.from_numpy(ids["box_info"])))))),
        boxes=boxes.numpy(), angles=angles.numpy(), objs=objs.numpy(),
        K=K.numpy(), R=R.numpy(), t=t.numpy(),
        vertices=depth_call[1][0].numpy(), faces_culled=depth_call[2][0].numpy().astype(np.int32), face_cls=face_cls.numpy(),
        sizes=np.stack([np.asarray(s, dtype=np.float32).reshape(-1)[:3] for s in sizes[:-1]]) if len(sizes) > 1 else np.zeros((0, 3), np.float32),
        sizes_last=np.asarray(sizes[-1], dtype=np.float32), box_info=ids["box_info"],
        W_seed=np.int64(7), grad_boxes=gb.numpy(), grad_angles=ga.numpy(),
        boxes2=torch.stack([x.detach() for x in b2]).numpy(), angles2=torch.stack([x.detach() for x in a2]).numpy(),
        size_loss2=np.float64(float(size_loss2)), grad_boxes2=gb2.numpy(), grad_angles2=ga2.numpy(),
        vertices2=rec.calls[0][1][0].numpy(), n_sizes2=np.int64(len(sizes2)),
        depth=calls_depth(final), final_chan_sum=final.detach().double().sum(dim=(0, 2, 3)).numpy(),
        final_sample=final.detach().flatten()[::SAMPLE_STRIDE].numpy(),
        final2_chan_sum=final2.detach().double().sum(dim=(0, 2, 3)).numpy(), final2_sample=final2.detach().flatten()[::SAMPLE_STRIDE].numpy())
    if full:
        blob["final"] = final.detach().numpy()
        blob["final2"] = final2.detach().numpy()
        # raw class images of the first call (what the compositing consumes): sum(rgb)/3 per class == one_hot channel
    path = os.path.join(ROOT, "tests", "golden", "render_%s.npz" % name)
    np.savez_compressed(path, **blob)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "V=%d F=%d culled=%d" %
          (blob["vertices"].shape[0], len(face_cls), int(sum(library.get(OBJECT_NAMES[int(o)])["faces"].size(0) for o in objs[:-1]
                                                                if OBJECT_NAMES[int(o)] in library.meta
                                                                and OBJECT_NAMES[int(o)] not in ("wall", "ceiling", "floor", "person", "door", "window", "curtain", "blinds")) + 160 - len(face_cls))),
          "size_loss2=%.6g" % float(size_loss2))


def calls_depth(final):
    return final.detach()[0, 0].numpy()


def small_layout():
    """4 furniture objects + a skipped type (door) + one object pushed through the near plane (cull) + the room."""
    boxes, angles, objs = synthetic_layout(5, seed=3)
    objs[1] = OBJECT_NAMES.index("door")
    boxes[2, [2, 5]] += 0.62                       # towards the camera wall: some of its faces fall inside z_cam < 0.06
    angles = angles + torch.tensor([0.25, 0.0, -0.4, 0.1, 0.0, 0.0])      # fractional angles, as softargmax produces
    return boxes, angles, objs


def main():
    case("small", small_layout(), 2, 3, 64, full=True)
    case("config3", synthetic_layout(10, seed=13), 6, 7, 256, full=False)


if __name__ == "__main__":
    main()
