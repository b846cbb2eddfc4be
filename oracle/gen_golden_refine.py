"""TEST INFRASTRUCTURE ONLY — golden vectors of the reference's refinement loss (testing/test_render_refine.py).

    python oracle/gen_golden_refine.py     # needs /root/reference; writes tests/golden/refine_loss.npz

testing/test_render_refine.py cannot be imported (module-level metadata/*.json loads, imageio, neural_renderer), so the pieces
on the path are EXECUTED FROM WHERE THEY LIE: the defs of softargmax, PSP_pool_new, fix_grad and quad_grad are pulled out of the
file's AST, and the loss statements of the loop body (:331-352) are compiled from the file's own lines — no source is copied.
Inputs are the seeded synthetic renders of tests/helpers.synthetic_render; stored: the three loss terms, the pooled pyramids'
checksums, a strided sample of d loss / d image, and fix_grad / quad_grad / softargmax on small seeded tensors.
"""
import ast
import os
import sys
import textwrap

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synthetic_render  # noqa: E402

REF = "/root/reference/testing/test_render_refine.py"
GRAD_STRIDE = 97


def reference_namespace():
    src = open(REF).read()
    tree = ast.parse(src)
    want = {"softargmax", "PSP_pool_new", "fix_grad", "quad_grad"}
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in want]
    assert {n.name for n in body} == want
    ns = {"torch": torch, "nn": torch.nn, "F": torch.nn.functional, "np": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    ns["depth_pooler"] = ns["PSP_pool_new"]()
    ns["semantic_pooler_novel"] = ns["PSP_pool_new"](use_max=False, output_list=True)
    ns["matching_loss_func"] = torch.nn.L1Loss()
    ns["ce_loss_func"] = torch.nn.CrossEntropyLoss()
    lines = src.splitlines()
    # the loss statements of the refinement loop: from the null-fill (:333) to the size-loss add (:354), prints dropped
    first = next(i for i, l in enumerate(lines) if "# Fill in null regions" in l)
    last = next(i for i, l in enumerate(lines) if "loss_val += size_loss * 2.0" in l)
    stmts = [l for l in lines[first:last + 1] if "print(" not in l]
    ns["_loss_code"] = compile(textwrap.dedent("\n".join(stmts)), REF + ":loss", "exec")
    return ns


def reference_loss(ns, iter_image, target_image, size_loss):
    env = dict(ns)
    env.update(iter_image=iter_image, target=target_image, target_mesh=None, size_infos=1, size_loss=size_loss, orig_scaler=0.5,
               long_dtype=torch.LongTensor)
    exec(ns["_loss_code"], env)
    return env["loss_val"], env["depth_loss"], env["semantic_loss"], env["scaled_input_depth"], env["train_labels_pooled"], env["target_container"]


def main():
    ns = reference_namespace()
    blob = {"grad_stride": np.int64(GRAD_STRIDE)}
    for case, (si, st) in enumerate([(11, 12), (21, 22), (21, 21)]):   # the last: identical renders (depth term = null-fill only)
        leaf = synthetic_render(si).requires_grad_(True)
        target = synthetic_render(st)
        size_loss = torch.tensor(0.03125)
        loss, dl, sl, pooled_d, pooled_s, labels = reference_loss(ns, leaf.clone(), target, size_loss)
        loss.backward()
        blob["c%d_seeds" % case] = np.array([si, st])
        blob["c%d_loss" % case] = np.array([loss.item(), dl.item(), float(sl)])
        blob["c%d_pooled_depth_sum" % case] = pooled_d.double().sum(dim=(0, 2, 3)).detach().numpy()
        blob["c%d_pooled_sem_sum" % case] = torch.stack([p.double().sum(dim=(0, 2, 3)) for p in pooled_s]).detach().numpy()
        blob["c%d_label_hist" % case] = np.stack([np.bincount(l.flatten().numpy() + 100, minlength=141) for l in labels])
        blob["c%d_grad_sample" % case] = leaf.grad.flatten()[::GRAD_STRIDE].numpy()
        blob["c%d_grad_abs_sum" % case] = leaf.grad.double().abs().sum(dim=(0, 2, 3)).numpy()
    g = torch.Generator().manual_seed(5)
    gv = torch.randn(7, 6, generator=g)
    blob["fix_grad_in"], blob["fix_grad_out"] = gv.numpy(), ns["fix_grad"](gv).numpy()
    blob["quad_grad_out"] = ns["quad_grad"](gv).numpy()
    sv = torch.randn(4, 24, generator=g)
    blob["softargmax_in"], blob["softargmax_out"] = sv.numpy(), ns["softargmax"](sv, 1).numpy()
    path = os.path.join(ROOT, "tests", "golden", "refine_loss.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, {k: (v.tolist() if v.size <= 3 else v.shape) for k, v in blob.items()})


if __name__ == "__main__":
    main()
