"""CPU self-consistency checks of the rasterizer oracle (oracle/raster_oracle.c).  PARITY UNPINNED: the third-party
`neural_renderer` is not available (see the oracle's header), so the restatement is pinned by independent derivations:
float64 brute-force rasterization away from ties, closed-form cases, and finite differences of the depth gradient."""
import importlib

import numpy as np
import pytest
import torch

from oracle import raster_oracle as ro

meshes = importlib.import_module("sln_b200.data.synthetic_meshes")


def scene(n_objects=4, seed=3, nu=2, nv=3):
    """A small synthetic room with boxes, world-space vertices + camera of the reference (get_cam_mat)."""
    boxes, angles, objs = meshes.synthetic_layout(n_objects, seed=seed)
    room = boxes[-1][3:]
    vs, fs, off = [], [], 0
    for i in range(n_objects):
        lo, hi = boxes[i][:3] * room, boxes[i][3:] * room
        v, f = meshes.box_mesh(nu, nv, size=(hi - lo).tolist(), center=((hi + lo) / 2).tolist())
        th = -float(angles[i]) * 2 * np.pi / 24
        rot = torch.tensor([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]], dtype=torch.float32)
        c = (hi + lo) / 2
        v = (v - c) @ rot.t() + c
        vs.append(v); fs.append(f + off); off += v.size(0)
    for name, (v, f) in meshes.room_shell(room, n=2).items():
        vs.append(v); fs.append(f + off); off += v.size(0)
    K, R, t = ro.get_cam_mat(boxes[-1])
    verts, faces = torch.cat(vs).numpy(), torch.cat(fs).numpy().astype(np.int32)
    zc = (verts @ R.T + t)[:, 2]
    faces = faces[~(zc[faces] < 0.06).any(axis=1)]          # the reference's culling (models/diff_render.py:345-356)
    return verts, faces, K, R, t


def raster64(fv, n, near, far):
    """float64 brute force of the same per-pixel rule.  Also returns `clear`: pixels where no decision is close — the
    distance to every front face's edge lines exceeds 1e-5 NDC units and the two nearest depths differ by > 1e-5."""
    fv = fv.astype(np.float64).reshape(-1, 3, 3)
    ys, xs = np.mgrid[0:n, 0:n]
    yp, xp = (2.0 * ys + 1 - n) / n, (2.0 * xs + 1 - n) / n
    best = np.full((n, n), far); second = np.full((n, n), np.inf); idx = np.full((n, n), -1)
    edge_margin = np.full((n, n), np.inf)
    for fn, f in enumerate(fv):
        if (f[2, 1] - f[0, 1]) * (f[1, 0] - f[0, 0]) < (f[1, 1] - f[0, 1]) * (f[2, 0] - f[0, 0]):
            continue
        inside = np.ones((n, n), bool)
        for k in range(3):
            a, b = f[k], f[(k + 1) % 3]
            e = (yp - a[1]) * (b[0] - a[0]) - (xp - a[0]) * (b[1] - a[1])
            inside &= e >= 0
            length = np.hypot(b[0] - a[0], b[1] - a[1])
            if length > 0:
                edge_margin = np.minimum(edge_margin, np.abs(e) / length)
        p = 0.5 * (f[:, :2] * n + n - 1)
        M = np.array([[p[0, 0], p[1, 0], p[2, 0]], [p[0, 1], p[1, 1], p[2, 1]], [1, 1, 1]])
        if abs(np.linalg.det(M)) < 1e-12:
            continue
        inv = np.linalg.inv(M)
        w = np.stack([inv[k, 0] * xs + inv[k, 1] * ys + inv[k, 2] for k in range(3)], -1)
        w = np.clip(w, 0, 1); w = w / w.sum(-1, keepdims=True)
        zp = 1.0 / (w[..., 0] / f[0, 2] + w[..., 1] / f[1, 2] + w[..., 2] / f[2, 2])
        ok = inside & (zp > near) & (zp < far)
        closer = ok & (zp < best)
        second = np.where(closer, best, np.where(ok, np.minimum(second, zp), second))
        best = np.where(closer, zp, best); idx = np.where(closer, fn, idx)
    clear = (edge_margin > 1e-5) & (np.abs(second - best) > 1e-5 * np.abs(best))
    return idx, best, clear


def test_projection_matches_float64_formula():
    verts, faces, K, R, t = scene()
    pv = ro.project(verts, K, R, t, 512)
    v = verts.astype(np.float64) @ R.astype(np.float64).T + t.astype(np.float64)
    x_, y_ = v[:, 0] / (v[:, 2] + 1e-9), v[:, 1] / (v[:, 2] + 1e-9)
    u = K[0, 0] * x_ + K[0, 1] * y_ + K[0, 2]
    w = 512 - (K[1, 0] * x_ + K[1, 1] * y_ + K[1, 2])
    want = np.stack([2 * (u - 256) / 512, 2 * (w - 256) / 512, v[:, 2]], -1)
    assert (np.abs(pv - want) <= 1e-5 * np.maximum(1.0, np.abs(want))).all()   # vertices near the camera plane have huge u,v


@pytest.mark.parametrize("n", [64, 96])
def test_face_index_map_matches_float64_bruteforce_away_from_ties(n):
    verts, faces, K, R, t = scene()
    fv = ro.gather_faces(ro.project(verts, K, R, t, 512), faces, True)
    maps = ro.face_index_map(fv, n, 0.1, 100.0)
    idx64, z64, clear = raster64(fv, n, 0.1, 100.0)
    assert clear.mean() > 0.9
    assert (maps["face_index"][clear] == idx64[clear]).all()
    cov = clear & (idx64 >= 0)
    assert cov.sum() > 0.4 * n * n                      # the room fills most of the view
    assert np.abs(maps["depth"][cov] - z64[cov]).max() < 1e-4
    assert (maps["depth"][maps["face_index"] < 0] == 100.0).all()


def test_single_triangle_closed_form_and_fill_back():
    # a camera-facing triangle at constant depth 2 covering the lower-left half of the view in NDC
    fv = np.array([[-0.9, -0.9, 2.0, 0.9, -0.9, 2.0, -0.9, 0.9, 2.0]], dtype=np.float32)
    for order in (fv, fv.reshape(1, 3, 3)[:, ::-1].reshape(1, 9).copy()):
        both = np.concatenate([order, order.reshape(1, 3, 3)[:, ::-1].reshape(1, 9)])     # fill_back pair
        maps = ro.face_index_map(both, 32, 0.1, 100.0)
        covered = maps["face_index"] >= 0
        ys, xs = np.mgrid[0:32, 0:32]
        xp, yp = (2 * xs + 1 - 32) / 32.0, (2 * ys + 1 - 32) / 32.0
        want = (xp > -0.9) & (yp > -0.9) & (xp + yp <= 0.0)   # pixel centres exactly on an edge count as inside
        assert (covered == want).mean() > 0.99           # only boundary pixels may differ
        assert len(np.unique(maps["face_index"][covered])) == 1      # exactly one of the two windings is front-facing
        assert np.allclose(maps["depth"][covered], 2.0, atol=1e-5)
        assert np.allclose(maps["weight"][covered].sum(-1), 1.0, atol=1e-5)
        rgb = ro.texture_sampling(both, np.ones((2, 2, 2, 2, 3), np.float32), maps, 2)
        assert np.allclose(rgb[covered], 1.0, atol=1e-5) and (rgb[~covered] == 0).all()


def test_depth_gradient_matches_finite_differences():
    verts, faces, K, R, t = scene(n_objects=2, seed=5)
    r = ro.RendererOracle(48, K, R, t, 512)
    depth, ctx = r.depth(verts, faces)
    rng = np.random.RandomState(0)
    G = rng.randn(48, 48).astype(np.float32) * (depth < 50)
    g = r.depth_bwd(verts, faces, ctx, G)
    # move whole objects rigidly along z: coverage changes only at silhouettes, interior depth changes smoothly
    for lo, hi in ((0, 72), (72, 144)):
        d = np.zeros_like(verts); d[lo:hi, 2] = 1.0
        h = 1e-3
        dp, _ = r.depth(verts + h * d, faces)
        dm, _ = r.depth(verts - h * d, faces)
        same = (ro.face_index_map(ro.gather_faces(ro.project(verts + h * d, K, R, t, 512), faces), 48, 0.1, 100.0)["face_index"] ==
                ro.face_index_map(ro.gather_faces(ro.project(verts - h * d, K, R, t, 512), faces), 48, 0.1, 100.0)["face_index"])[::-1]
        fd = (((dp - dm) / (2 * h)) * G * same).sum()
        # analytic directional derivative restricted to the same pixels
        g_same = r.depth_bwd(verts, faces, ctx, G * same)
        an = (g_same * d).sum()
        assert abs(fd - an) <= 0.05 * max(abs(fd), abs(an), 1e-3), (fd, an)


def test_composite_layout_and_gradient_flow():
    H = 16
    names = ["wall", "bed", "ceiling", "chair", "floor"]
    depth = (torch.rand(1, H, H) * 3 + 1).requires_grad_(True)
    labels = torch.randint(0, len(names), (H, H))
    images = [(labels == c).float()[None].requires_grad_(True) for c in range(len(names))]
    final = ro.composite(depth, images, names)
    assert final.shape == (1, 1 + 40 + 2, H, H)
    assert torch.equal(final[0, 0], depth[0].detach())
    assert torch.equal(final[0, 1 + ro.NYU_CLASS.index("bed")], images[1][0].detach())
    final.sum().backward()
    assert depth.grad.abs().sum() > 0 and images[1].grad.abs().sum() > 0
