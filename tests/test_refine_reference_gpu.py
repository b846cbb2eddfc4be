"""The reference's own refinement iteration (testing/test_render_refine.py:279-359: z -> decoder -> hooks -> softargmax -> render ->
loss -> backward through the decoder -> re-created nesterov SGD) as ReferenceRefineStep, against the same iteration spelled out with
the reference's statements: tensor hooks (fix_grad / quad_grad), the torch-op scene assembly / compositing / loss restatements and a
real torch.optim.SGD constructed exactly as :286 does."""
import copy
import importlib

import pytest
import torch

from helpers import our_model, syn

pytestmark = pytest.mark.gpu
refine = importlib.import_module("sln_b200.models.refine")
dr = importlib.import_module("sln_b200.models.diff_render")
meshes = importlib.import_module("sln_b200.data.synthetic_meshes")
DEV = "cuda:0"


def _scene(n_obj=4, seed=3):
    boxes, angles, objs = meshes.synthetic_layout(n_obj, seed=seed)
    n = n_obj + 1
    triples = [[i, 0, n_obj] for i in range(n_obj)] + [[i, 1 + (i % 9), (i + 1) % n_obj] for i in range(n_obj)]
    return boxes.to(DEV), angles.to(DEV), objs.to(DEV), torch.tensor(triples, dtype=torch.long, device=DEV), torch.zeros(n, dtype=torch.long, device=DEV)


def _explicit_iteration(model, z, objs, triples, attrs, boxes_gt, angles_gt, static, t_depth, t_labels, size_target, lr_model, lr_z=2e-4):
    """test_render_refine.py:286-357 statement by statement (noise off), on the torch-op restatements of render and loss."""
    optimizer = torch.optim.SGD([{'params': [z]}, {'params': model.parameters(), 'lr': lr_model}], lr=lr_z, nesterov=True, momentum=0.1)
    boxes_pred, angles_pred = model.decoder(z, objs, triples, attrs)
    boxes_pred.register_hook(refine.fix_grad)
    boxes_pred = torch.cat([boxes_pred[:-1], boxes_gt[-1:]], 0)
    ang = refine.softargmax(angles_pred, sum_dim=1)
    ang.register_hook(refine.quad_grad)
    ang = torch.cat([ang[:-1], angles_gt[-1:]], 0)
    image, size = dr.render_static(static, boxes_pred, ang, fused=False)
    size_loss = ((size - size_target) ** 2).mean(dim=1).sum()
    loss = refine.refine_loss(image, t_depth, t_labels, size_loss)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return float(loss)


def test_reference_iteration_matches_the_spelled_out_loop_and_its_graph():
    boxes, angles, objs, triples, attrs = _scene()
    model_a = refine.bias_box_head(our_model(E=16, layers=2, norm="batch", device=DEV).eval())     # a decoder whose boxes render
    init_params = [p.detach().clone() for p in model_a.parameters()]
    with torch.no_grad():                      # non-trivial running statistics, as a trained checkpoint has
        for m in model_a.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    model_b, model_c = copy.deepcopy(model_a), copy.deepcopy(model_a)
    z0 = torch.randn(objs.size(0), 16, generator=torch.Generator().manual_seed(13)).to(DEV)
    lib = meshes.MeshLibrary(nu=2, nv=3).to(torch.device(DEV))
    # learning rates 1000x the reference's (2e-4 / 1e-5) so that three updates are far above fp32 resolution of the parameters
    LZ, LM = 0.2, 1e-2
    ours = refine.ReferenceRefineStep(model_a, z0, objs, triples, attrs, boxes, angles, lr_z=LZ, lr_model=LM, noise=False, use_graph=False, library=lib)
    graph = refine.ReferenceRefineStep(model_c, z0, objs, triples, attrs, boxes, angles, lr_z=LZ, lr_model=LM, noise=False, use_graph=True, library=lib)
    zb = z0.clone().requires_grad_(True)
    la, lb, lc = [], [], []
    for _ in range(3):
        la.append(float(ours.step()))
        lc.append(float(graph.step()))
        lb.append(_explicit_iteration(model_b, zb, objs, triples, attrs, boxes, angles, ours.static, ours.t_depth, ours.t_labels, ours.size_target, LM, LZ))
    assert la[0] > 0 and all(abs(a - b) <= 1e-3 * abs(b) for a, b in zip(la, lb)), (la, lb)
    assert all(abs(a - c) <= 1e-5 * abs(a) for a, c in zip(la, lc)), (la, lc)
    # the updates: z and every decoder parameter moved by -lr * 1.1 * grad exactly as the re-created nesterov SGD moves them
    dz_ours, dz_ref = (ours.z.detach() - z0), (zb.detach() - z0)
    assert dz_ref.abs().max() > 0
    assert (dz_ours - dz_ref).abs().max().item() <= 2e-2 * dz_ref.abs().max().item()
    moved = 0
    for (k, pa), (_, pb), p0 in zip(model_a.named_parameters(), model_b.named_parameters(), init_params):
        da, db = (pa.detach() - p0), (pb.detach() - p0)
        if db.abs().max().item() == 0.0:
            assert da.abs().max().item() == 0.0, k          # encoder parameters receive no gradient: untouched (SGD skips grad None)
            continue
        moved += 1
        assert (da - db).abs().max().item() <= 3e-2 * db.abs().max().item() + 1e-9, k
    assert moved > 10
    assert (ours.z.detach() - graph.z.detach()).abs().max().item() <= 1e-4 * max(1.0, dz_ref.abs().max().item())


def test_noise_and_frozen_model_options():
    boxes, angles, objs, triples, attrs = _scene(seed=5)
    model = refine.bias_box_head(our_model(E=16, layers=2, norm="none", device=DEV).eval())
    before = [p.detach().clone() for p in model.parameters()]
    z0 = torch.randn(objs.size(0), 16, generator=torch.Generator().manual_seed(2)).to(DEV)
    lib = meshes.MeshLibrary(nu=2, nv=3).to(torch.device(DEV))
    step = refine.ReferenceRefineStep(model, z0, objs, triples, attrs, boxes, angles, noise=True, use_graph=True, library=lib, update_model=False)
    a1 = None
    for _ in range(2):
        step.step()
        a2 = step.angles_pred.clone()
        assert a1 is None or not torch.equal(a1[:-1], a2[:-1])       # fresh N(0,1)/10 jitter every replay (:293)
        a1 = a2
    assert float(step.angles_pred[-1]) == float(angles[-1]) and torch.equal(step.boxes_pred[-1], boxes[-1])
    assert all(torch.equal(p, q) for p, q in zip(model.parameters(), before))
    with pytest.raises(RuntimeError, match="eval"):
        refine.ReferenceRefineStep(model.train(), z0, objs, triples, attrs, boxes, angles, library=lib)
