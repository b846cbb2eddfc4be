"""TEST INFRASTRUCTURE ONLY — generate tests/golden/* by running the UNMODIFIED reference modules (CPU) in this container.

    python oracle/gen_golden.py            # needs /root/reference; writes tests/golden/vae_*.npz, vae_kat.json

The GPU box has no /root/reference, so the vectors are committed.  Each .npz holds a seeded reference model's
state_dict, the inputs, the injected N(0,1) sample and what the reference computes from them in fp32 and fp64:
forward outputs, the three loss terms, every parameter gradient and the BatchNorm running statistics after the step.
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib  # noqa: E402

from oracle import ref_shim  # noqa: E402

syn = importlib.import_module("sln_b200.data.synthetic")
GOLD = os.path.join(ROOT, "tests", "golden")


def build_ref(norm, E, layers, mode='feedforward', use_AE=False, seed=42):
    torch.manual_seed(seed)
    Ref = ref_shim.vae_model_class()
    return Ref(syn.default_vocab(), embedding_dim=E, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode=mode,
               gconv_num_layers=layers, mlp_normalization=norm, vec_noise_dim=0, layout_noise_dim=32, use_AE=use_AE)


def run_ref(model, batch, eps, training, kl_weight=0.1):
    objs, triples, boxes, angles, attrs = batch
    model.train(training)
    dt = next(model.parameters()).dtype
    orig = torch.randn_like
    torch.randn_like = lambda t: eps.to(t.dtype)
    try:
        mu, lv, bp, ap = model(objs, triples, boxes.to(dt), angles, attrs, None)
    finally:
        torch.randn_like = orig
    total, parts = ref_shim.reference_losses()(types.SimpleNamespace(use_AE=False), model, boxes.to(dt), bp, angles, ap, mu=mu,
                                               logvar=lv, KL_weight=kl_weight)
    model.zero_grad()
    total.backward()
    out = {'mu': mu, 'logvar': lv, 'boxes_pred': bp, 'angles_pred': ap, 'total': total}
    out.update({'loss_' + k: torch.tensor(v, dtype=torch.float64) for k, v in parts.items()})
    out.update({'grad.' + k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()})
    out.update({'after.' + k: v for k, v in model.state_dict().items() if 'running' in k or 'num_batches' in k})
    return {k: v.detach().double().numpy() for k, v in out.items()}


def gen_small(name, norm, training, mode='feedforward', E=8, layers=2, scenes=3, nodes=6):
    import copy
    model = build_ref(norm, E, layers, mode)
    if not training and norm == 'batch':   # non-trivial running statistics for the eval-mode case
        g = torch.Generator().manual_seed(7)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    _, objs, boxes, triples, angles, attrs, _, _ = syn.synthetic_batch(scenes, nodes, seed=3)
    eps = torch.randn(objs.size(0), E, generator=torch.Generator().manual_seed(11))
    batch = (objs, triples, boxes, angles, attrs)
    blob = {'sd.' + k: v.detach().numpy() for k, v in model.state_dict().items()}
    blob.update({'in.objs': objs.numpy(), 'in.triples': triples.numpy(), 'in.boxes': boxes.numpy(), 'in.angles': angles.numpy(),
                 'in.attrs': attrs.numpy(), 'in.eps': eps.numpy()})
    blob['meta'] = np.array(json.dumps(dict(norm=norm, training=training, mode=mode, E=E, layers=layers, kl_weight=0.1)))
    r32 = run_ref(copy.deepcopy(model), batch, eps, training)
    r64 = run_ref(copy.deepcopy(model).double(), batch, eps.double(), training)
    blob.update({'f32.' + k: v.astype(np.float32) for k, v in r32.items()})
    blob.update({'f64.' + k: v for k, v in r64.items()})
    path = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(path, **blob)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def gen_kat():
    """seed 42, E=64 model, the 5-object fixture graph, use_AE=True, no_grad forward."""
    objs, triples, boxes, angles, attrs = syn.fixture_graph()
    out = {}
    for norm, training in (('batch', True), ('batch', False), ('none', True)):
        m = build_ref(norm, 64, 5, use_AE=True)
        m.train(training)
        with torch.no_grad():
            mu, lv, bp, ap = m(objs, triples, boxes, angles, attrs, None)
        out['%s/%s' % (norm, 'train' if training else 'eval')] = dict(
            mu_sum=mu.sum().item(), logvar_sum=lv.sum().item(), boxes_sum=bp.sum().item(), angles_sum=ap.sum().item(),
            argmax=ap.argmax(1).tolist(), boxes_pred=bp.tolist(), mu_row0=mu[0].tolist())
    with open(os.path.join(GOLD, 'vae_kat.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('wrote vae_kat.json')


if __name__ == '__main__':
    if not ref_shim.available():
        sys.exit("reference tree not available; golden vectors can only be regenerated in the build container")
    os.makedirs(GOLD, exist_ok=True)
    gen_small('vae_small_batch_train', 'batch', True)
    gen_small('vae_small_batch_eval', 'batch', False)
    gen_small('vae_small_none', 'none', True)
    gen_small('vae_small_recurrent', 'batch', True, mode='recurrent', layers=3)
    gen_kat()
