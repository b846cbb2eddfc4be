"""Refinement loss (reference testing/test_render_refine.py:20-25,192-230,332-352).

tests/golden/refine_loss.npz was produced by executing the reference's own statements (oracle/gen_golden_refine.py).  CPU: the
torch restatement in models/refine.py against it.  GPU: the fused CUDA loss (csrc/refine_loss.cu, through the C ABI) against the
golden values and against the restatement; tolerance 1e-4 relative (max-norm) on the loss terms and on d loss / d image."""
import importlib
import os

import numpy as np
import pytest
import torch

from helpers import synthetic_render

refine = importlib.import_module("sln_b200.models.refine")
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refine_loss.npz"))
STRIDE = int(Z["grad_stride"])
TOL = 1e-4


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _case(case, dev):
    si, st = [int(v) for v in Z["c%d_seeds" % case]]
    leaf = synthetic_render(si).to(dev).requires_grad_(True)
    target = synthetic_render(st).to(dev)
    return leaf, target, torch.tensor(0.03125, device=dev)


@pytest.mark.parametrize("case", [0, 1, 2])
def test_torch_restatement_matches_reference_statements(case):
    leaf, target, size_loss = _case(case, "cpu")
    t_depth, t_labels = refine.refine_targets(target)
    hist = np.stack([np.bincount(l.flatten().numpy() + 100, minlength=141) for l in t_labels])
    assert (hist == Z["c%d_label_hist" % case]).all()                    # argmax labels and the -100 fill are exact
    loss = refine.refine_loss(leaf, t_depth, t_labels, size_loss)
    loss.backward()
    assert _rel(loss.item(), Z["c%d_loss" % case][0]) < 1e-6
    assert _rel(leaf.grad.flatten()[::STRIDE].numpy(), Z["c%d_grad_sample" % case]) < 1e-6
    assert _rel(leaf.grad.double().abs().sum(dim=(0, 2, 3)).numpy(), Z["c%d_grad_abs_sum" % case]) < 1e-6
    pooled = refine.psp_pool(leaf.detach()[:, 1:41], output_list=True)
    assert _rel(torch.stack([p.double().sum(dim=(0, 2, 3)) for p in pooled]).numpy(), Z["c%d_pooled_sem_sum" % case]) < 1e-6


def test_gradient_hooks_and_softargmax_match_reference():
    g = torch.from_numpy(Z["fix_grad_in"])
    assert torch.equal(refine.fix_grad(g), torch.from_numpy(Z["fix_grad_out"]))
    assert torch.equal(refine.quad_grad(g), torch.from_numpy(Z["quad_grad_out"]))
    assert torch.allclose(refine.softargmax(torch.from_numpy(Z["softargmax_in"]), 1), torch.from_numpy(Z["softargmax_out"]), rtol=1e-6, atol=1e-6)


def test_fused_loss_refuses_cpu():
    with pytest.raises(RuntimeError):
        refine.FusedRefineLoss(torch.zeros(1, 116, 96, 96), [torch.zeros(1, 96, 96, dtype=torch.long)] * 4)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 1, 2])
def test_fused_loss_matches_reference_golden(case):
    leaf, target, size_loss = _case(case, "cuda:0")
    t_depth, t_labels = refine.refine_targets(target)
    f = refine.FusedRefineLoss(t_depth, t_labels)
    loss = f(leaf, size_loss)
    loss.backward()
    want = Z["c%d_loss" % case]
    terms = f.last_terms.cpu().numpy()
    assert _rel(loss.item(), want[0]) < TOL
    assert _rel(terms[1], want[1]) < TOL and _rel(terms[2], want[2]) < TOL
    g = leaf.grad.cpu()
    if case == 2:
        return      # identical renders: every pooled depth difference is +-1 ulp around the kink of |x|, its sign is rounding noise
    assert _rel(g.flatten()[::STRIDE].numpy(), Z["c%d_grad_sample" % case]) < TOL
    assert _rel(g.double().abs().sum(dim=(0, 2, 3)).numpy(), Z["c%d_grad_abs_sum" % case]) < TOL
    assert float(g[0, 0].abs().max()) == 0.0                             # the plain depth channel is not part of the loss


@pytest.mark.gpu
def test_fused_loss_matches_torch_restatement_everywhere_and_is_deterministic():
    dev = "cuda:0"
    a = synthetic_render(31).to(dev).requires_grad_(True)
    b = synthetic_render(31).to(dev).requires_grad_(True)
    t_depth, t_labels = refine.refine_targets(synthetic_render(32).to(dev))
    want = refine.refine_loss(a, t_depth, t_labels)
    want.backward()
    f = refine.FusedRefineLoss(t_depth, t_labels)
    got = f(b)
    got.backward()
    assert _rel(got.item(), want.item()) < TOL
    assert _rel(b.grad.cpu().numpy(), a.grad.cpu().numpy()) < TOL
    # null-filled pixels of the last plane are constants: no gradient there
    null = (a.detach()[:, 41:].sum(dim=1) < 0.5)[0]
    assert null.any() and float(b.grad[0, -1][null].abs().max()) == 0.0
    first = b.grad.clone()
    b.grad = None
    again = f(b)
    again.backward()
    assert torch.equal(again, got) and torch.equal(b.grad, first)       # gather-style backward: bit-reproducible
    with torch.no_grad():
        assert torch.equal(f(b.detach()), got)                           # forward-only call (no gradient buffers)


@pytest.mark.gpu
def test_refine_step_with_fused_loss_tracks_the_torch_loss():
    syn_m = importlib.import_module("sln_b200.data.synthetic_meshes")
    dev = torch.device("cuda:0")
    boxes, angles, objs = [t.to(dev) for t in syn_m.synthetic_layout(6, seed=13)]
    start = boxes.clone(); start[:-1, [0, 3]] += 0.03
    a = refine.RefineStep(start, angles, objs, boxes, angles, use_graph=False, fused_loss=True)
    b = refine.RefineStep(start, angles, objs, boxes, angles, use_graph=False, fused_loss=False, fused_scene=False)
    for _ in range(5):
        la, lb = a.step().item(), b.step().item()
        assert abs(la - lb) <= 2e-3 * abs(lb)
    assert torch.allclose(a.b, b.b, atol=2e-5)
