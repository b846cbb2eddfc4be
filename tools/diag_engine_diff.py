"""Diagnostic: which tcgen05 call site moves which gradient (engine masks vs the SIMT engine), config 2, norm none."""
import importlib, os, sys, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import our_model, syn, with_eps
_lib = importlib.import_module("sln_b200._lib")
sutils = importlib.import_module("sln_b200.utils")
lib = _lib.load()
norm = sys.argv[1] if len(sys.argv) > 1 else "none"
_, objs, boxes, triples, angles, attrs, _, _ = syn.synthetic_batch(64, 32, seed=42)
m = our_model(E=64, layers=5, norm=norm)
sd0 = {k: v.clone() for k, v in m.state_dict().items()}
eps = torch.randn(2048, 64, generator=torch.Generator().manual_seed(11))
def run(eng):
    lib.sln_set_engine(eng)
    mm = our_model(E=64, layers=5, norm=norm); mm.load_state_dict(sd0); mm = mm.to("cuda").train()
    with with_eps(eps):
        mu, lv, bp, ap = mm(objs.cuda(), triples.cuda(), boxes.cuda(), angles.cuda(), attrs.cuda(), None)
    total, _ = sutils.calculate_model_losses(types.SimpleNamespace(use_AE=False), mm, boxes.cuda(), bp, angles.cuda(), ap, mu=mu, logvar=lv, KL_weight=0.1)
    total.backward()
    got = {"mu": mu, "logvar": lv, "boxes_pred": bp, "angles_pred": ap}
    got.update({"grad." + k: p.grad for k, p in mm.named_parameters()})
    return {k: v.detach().double().cpu() for k, v in got.items()}
base = run(0)
for eng in (16 | 1, 16 | 2, 16 | 4, 16 | 8, 1):
    got = run(eng)
    rows = sorted(((got[k] - base[k]).abs().max().item() / max(base[k].abs().max().item(), 1e-30), k) for k in base)[::-1]
    print("engine mask", eng & 15 if eng >= 16 else "all", " ".join("%s=%.2e" % (k.replace("grad.", "g."), e) for e, k in rows[:4]))
