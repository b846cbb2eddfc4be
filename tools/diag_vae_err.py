"""Diagnostic: per-tensor error of the CUDA VAE path vs the fp64 oracle at config 2, for both contraction engines."""
import importlib, os, sys, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import our_model, syn, with_eps
from oracle import vae_oracle as vo
_lib = importlib.import_module("sln_b200._lib")
sutils = importlib.import_module("sln_b200.utils")
lib = _lib.load()
norm = sys.argv[1] if len(sys.argv) > 1 else "none"
_, objs, boxes, triples, angles, attrs, _, _ = syn.synthetic_batch(64, 32, seed=42)
m = our_model(E=64, layers=5, norm=norm)
sd0 = {k: v.clone() for k, v in m.state_dict().items()}
eps = torch.randn(2048, 64, generator=torch.Generator().manual_seed(11))
def oracle(dtype):
    sd = vo.leaf_state(sd0, dtype)
    mu, lv, bp, ap = vo.forward(sd, objs, triples, boxes.to(dtype), angles, attrs, eps.to(dtype), 5, True, False, {})
    total, _ = vo.losses(boxes.to(dtype), bp, angles, ap, mu, lv, 0.1)
    total.backward()
    out = {"mu": mu, "logvar": lv, "boxes_pred": bp, "angles_pred": ap}
    out.update({"grad." + k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd.items() if v.is_floating_point() and v.requires_grad})
    return {k: v.detach().double() for k, v in out.items()}
r64, r32 = oracle(torch.float64), oracle(torch.float32)
for eng in (0, 1):
    lib.sln_set_engine(eng)
    mm = our_model(E=64, layers=5, norm=norm); mm.load_state_dict(sd0); mm = mm.to("cuda").train()
    with with_eps(eps):
        mu, lv, bp, ap = mm(objs.cuda(), triples.cuda(), boxes.cuda(), angles.cuda(), attrs.cuda(), None)
    total, _ = sutils.calculate_model_losses(types.SimpleNamespace(use_AE=False), mm, boxes.cuda(), bp, angles.cuda(), ap, mu=mu, logvar=lv, KL_weight=0.1)
    total.backward()
    got = {"mu": mu, "logvar": lv, "boxes_pred": bp, "angles_pred": ap}
    got.update({"grad." + k: p.grad for k, p in mm.named_parameters()})
    rows = []
    for k, v in got.items():
        t = r64[k]; sc = t.abs().max().item()
        e = (v.detach().cpu().double() - t).abs().max().item(); n = (r32[k] - t).abs().max().item()
        rows.append((e / max(sc, 1e-30), n / max(sc, 1e-30), sc, k))
    rows.sort(reverse=True)
    print("== engine", eng, "norm", norm, ": worst tensors (rel err ours, rel err ref fp32, scale, name)")
    for r in rows[:12]:
        print("  %.3e  %.3e  %.3e  %s" % r)
    import statistics
    print("  median rel err ours %.3e, ref fp32 %.3e" % (statistics.median(r[0] for r in rows), statistics.median(r[1] for r in rows)))
