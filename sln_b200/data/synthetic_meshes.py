"""Synthetic meshes for the layout-refinement path (no SUNCG meshes are available offline; the reference loads them with
pywavefront in models/misc.py:66-121).  SURVEY.md 8(d) config 3: 10 objects, each a box mesh subdivided to 504 triangles
(6 faces x 6x7 quads x 2) = 5040 triangles, plus a room shell (floor, ceiling, 3 walls as 4x4-subdivided quads, 160 triangles).
"""
import math

import numpy as np
import torch

from .synthetic import OBJECT_NAMES


def grid_quad(p0, du, dv, nu, nv):
    """(nu x nv)-subdivided parallelogram p0 + a*du + b*dv -> vertices [(nu+1)(nv+1),3], faces [2*nu*nv,3]."""
    a = torch.linspace(0, 1, nu + 1)
    b = torch.linspace(0, 1, nv + 1)
    A, B = torch.meshgrid(a, b, indexing="ij")
    verts = p0[None, None, :] + A[..., None] * du[None, None, :] + B[..., None] * dv[None, None, :]
    verts = verts.reshape(-1, 3)
    idx = torch.arange((nu + 1) * (nv + 1)).view(nu + 1, nv + 1)
    v00, v10, v01, v11 = idx[:-1, :-1], idx[1:, :-1], idx[:-1, 1:], idx[1:, 1:]
    faces = torch.cat([torch.stack([v00, v10, v11], -1).reshape(-1, 3), torch.stack([v00, v11, v01], -1).reshape(-1, 3)])
    return verts.float(), faces.long()


def box_mesh(nu=6, nv=7, size=(1.0, 1.0, 1.0), center=(0.0, 0.0, 0.0)):
    """Axis-aligned box centred at `center`: 6 subdivided faces -> vertices [6(nu+1)(nv+1),3], faces [12 nu nv,3]."""
    sx, sy, sz = size
    c = torch.tensor(center, dtype=torch.float32)
    h = torch.tensor([sx, sy, sz], dtype=torch.float32) / 2
    ex, ey, ez = torch.tensor([sx, 0, 0.]), torch.tensor([0, sy, 0.]), torch.tensor([0, 0, sz])
    lo = c - h
    quads = [(lo, ex, ey), (lo + ez, ey, ex), (lo, ey, ez), (lo + ex, ez, ey), (lo, ez, ex), (lo + ey, ex, ez)]
    vs, fs, off = [], [], 0
    for p0, du, dv in quads:
        v, f = grid_quad(p0, du, dv, nu, nv)
        vs.append(v); fs.append(f + off); off += v.size(0)
    return torch.cat(vs), torch.cat(fs)


class MeshLibrary(object):
    """One canonical mesh per object class (the reference retrieves a SUNCG model per object by edge-length ratio,
    models/misc.py:34-64).  model_size / model_center play the role of suncg_data[...]['bbox_min'/'bbox_max']."""

    def __init__(self, nu=6, nv=7, device="cpu"):
        self.nu, self.nv = nu, nv
        g = torch.Generator().manual_seed(1234)
        self.models = {}
        self.meta = {}
        for name in OBJECT_NAMES[1:]:
            size = (0.5 + torch.rand(3, generator=g)).tolist()      # canonical model extents 0.5 .. 1.5 m
            center = ((torch.rand(3, generator=g) - 0.5) * 0.2).tolist()
            v, f = box_mesh(nu, nv, size, center)
            # the metadata the reference reads from suncg_data_many.json ("bbox_min"/"bbox_max", models/diff_render.py:100-105);
            # size / center are derived from it exactly as the reference does (float32 numpy arithmetic)
            bbox_min = [c - s / 2 for c, s in zip(center, size)]
            bbox_max = [c + s / 2 for c, s in zip(center, size)]
            mn, mx = np.array(bbox_min, dtype=np.float32), np.array(bbox_max, dtype=np.float32)
            self.models[name] = dict(vertices=v.to(device), faces=f.to(device), size=torch.from_numpy((mx - mn).astype("float32")).to(device),
                                     center=torch.from_numpy(((mn + mx) / 2.0).astype("float32")).to(device))
            self.meta[name] = dict(id=name, bbox_min=bbox_min, bbox_max=bbox_max)

    def to(self, device):
        for m in self.models.values():
            for k in m:
                m[k] = m[k].to(device)
        return self

    def get(self, name):
        return self.models[name]


def room_shell(room, n=4):
    """floor / ceiling / 3 walls of the room box [x,y,z] as n x n subdivided quads -> dict name -> (vertices, faces).
    (reference models/diff_render.py:167-342 builds walls/floor/ceiling from SUNCG wall meshes.)"""
    X, Y, Z = [float(v) for v in room]
    o = torch.zeros(3)
    ex, ey, ez = torch.tensor([X, 0, 0.]), torch.tensor([0, Y, 0.]), torch.tensor([0, 0, Z])
    floor = grid_quad(o, ez, ex, n, n)
    ceiling = grid_quad(o + ey, ex, ez, n, n)
    walls = [grid_quad(o, ex, ey, n, n), grid_quad(o, ey, ez, n, n), grid_quad(o + ex, ez, ey, n, n)]
    wv = torch.cat([w[0] for w in walls])
    wf = torch.cat([w[1] + i * walls[0][0].size(0) for i, w in enumerate(walls)])
    return {"floor": floor, "ceiling": ceiling, "wall": (wv, wf)}


def room_walls(room, n=4):
    """The three wall meshes of room_shell separately, as the reference's load_wall_obj_new returns them (models/misc.py:157-165)."""
    X, Y, Z = [float(v) for v in room]
    o = torch.zeros(3)
    ex, ey, ez = torch.tensor([X, 0, 0.]), torch.tensor([0, Y, 0.]), torch.tensor([0, 0, Z])
    return [grid_quad(o, ex, ey, n, n), grid_quad(o, ey, ez, n, n), grid_quad(o + ex, ez, ey, n, n)]


FURNITURE = ["bed", "chair", "sofa", "table", "desk", "cabinet", "dresser", "night_stand", "bookshelf", "television"]


def synthetic_layout(n_objects=10, seed=13, room=(4.0, 2.7, 4.0)):
    """boxes [n+1,6] (objects normalised to the room, room row last = [0,0,0,x,y,z]), angles [n+1] in 0..23, objs [n+1] ids.
    Non-overlapping boxes standing on the floor on a jittered grid."""
    g = torch.Generator().manual_seed(seed)
    cols = int(math.ceil(math.sqrt(n_objects)))
    boxes, angles, objs = [], [], []
    for i in range(n_objects):
        cx = (i % cols + 0.5) / cols + (torch.rand(1, generator=g).item() - 0.5) * 0.1 / cols
        cz = (i // cols + 0.5) / cols * 0.8 + (torch.rand(1, generator=g).item() - 0.5) * 0.1 / cols
        sx = (0.35 + 0.3 * torch.rand(1, generator=g).item()) / cols
        sz = (0.35 + 0.3 * torch.rand(1, generator=g).item()) / cols
        sy = 0.15 + 0.35 * torch.rand(1, generator=g).item()
        boxes.append([cx - sx / 2, 0.0, cz - sz / 2, cx + sx / 2, sy, cz + sz / 2])
        angles.append(int(torch.randint(0, 24, (1,), generator=g)))
        objs.append(OBJECT_NAMES.index(FURNITURE[i % len(FURNITURE)]))
    boxes.append([0.0, 0.0, 0.0, room[0], room[1], room[2]])
    angles.append(0)
    objs.append(0)
    return torch.tensor(boxes, dtype=torch.float32), torch.tensor(angles, dtype=torch.float32), torch.tensor(objs, dtype=torch.long)
