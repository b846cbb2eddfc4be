"""Shared helpers for the parity tests (test infrastructure)."""
import importlib
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

syn = importlib.import_module("sln_b200.data.synthetic")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    inp = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in.")}
    f32 = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("f32.")}
    f64 = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("f64.")}
    return meta, sd, inp, f32, f64


def our_model(E=64, layers=5, norm="batch", mode="feedforward", use_AE=False, seed=42, device=None):
    Model = importlib.import_module("sln_b200.models.Sg2ScVAE_model").Sg2ScVAEModel
    torch.manual_seed(seed)
    m = Model(syn.default_vocab(), embedding_dim=E, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode=mode,
              gconv_num_layers=layers, mlp_normalization=norm, vec_noise_dim=0, layout_noise_dim=32, use_AE=use_AE)
    return m.to(device) if device is not None else m


def max_norm_err(a, b):
    """max|a-b| / max(max|b|, tiny)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def check_close(name, new, truth64, ref32=None, tol=1e-4, slack=3.0, abs_floor=0.0):
    """err(new, fp64 truth) <= max(tol, slack * err(reference fp32, fp64 truth)) in max-norm (SURVEY App. F rule 1).

    Tensors whose true value is identically ~0 (e.g. Linear biases feeding a training-mode BatchNorm) are compared
    absolutely against the reference's own fp32 noise."""
    new, truth64 = new.detach().double().cpu(), truth64.detach().double().cpu()
    assert new.shape == truth64.shape, (name, new.shape, truth64.shape)
    scale = truth64.abs().max().item()
    noise = (ref32.detach().double().cpu() - truth64).abs().max().item() if ref32 is not None else 0.0
    err = (new - truth64).abs().max().item()
    bound = max(tol * scale, slack * noise, abs_floor)
    assert err <= bound, "%s: max|err|=%.3e > bound %.3e (scale %.3e, ref fp32 noise %.3e)" % (name, err, bound, scale, noise)
    return err / max(scale, 1e-30)


def with_eps(eps):
    """Context manager injecting a fixed N(0,1) sample into torch.randn_like (the model keeps torch's RNG call)."""
    import contextlib

    @contextlib.contextmanager
    def cm():
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: eps.to(device=t.device, dtype=t.dtype)
        try:
            yield
        finally:
            torch.randn_like = orig
    return cm()


def synthetic_render(seed, S=256, n_sem=40, n_dep=29):
    """A render-like [1, 1+n_sem+n_dep, S, S] image (what mesh_render_func returns): channel 0 depth, then 0/1 class masks made of
    random rectangles (with soft edges so that pooled logits are not all ties), then per-class depth planes that are non-zero only
    inside their class mask; a band of pixels is left empty so that the null-fill of the last plane triggers."""
    g = torch.Generator().manual_seed(seed)
    img = torch.zeros(1, 1 + n_sem + n_dep, S, S)
    img[0, 0] = torch.rand(S, S, generator=g)
    for c in range(n_sem):
        for _ in range(2):
            y0, x0 = [int(v) for v in torch.randint(0, S - S // 8, (2,), generator=g)]
            h, w = [int(v) for v in torch.randint(S // 16, S // 3, (2,), generator=g)]
            img[0, 1 + c, y0:y0 + h, x0:min(S - S // 10, x0 + w)] = 1.0
    img[0, 1:1 + n_sem] *= 0.75 + 0.25 * torch.rand(n_sem, S, S, generator=g)
    for d in range(n_dep):
        img[0, 1 + n_sem + d] = img[0, 1 + d] * torch.rand(S, S, generator=g)
    return img
