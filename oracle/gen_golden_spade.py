"""TEST INFRASTRUCTURE ONLY — tests/golden/spade_small.npz from the UNMODIFIED reference SPADEGenerator4 (CPU, this container).

    python oracle/gen_golden_spade.py      # needs /root/reference

A reduced generator (ngf=8, nz=16, crop 64 -> 2x2 latent grid, 41-channel 64x64 input, batch 2) with seeded weights.  The
file stores the input, z, the reference's fp32 and fp64 outputs (tanh image, pre-tanh conv_img output via a forward hook,
every block output) and a checksum of the seeded state_dict: the drop-in module built with the same seed must reproduce
the same state_dict (same creation order / initialisers), which is how the weights travel without being committed."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, spade_oracle  # noqa: E402

CFG = dict(semantic_nc=41, target_nc=3, nz=16, ngf=8, norm='spectralspadelayer3x3', crop_size=64, n_up='normal')


def state_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    import copy
    import warnings
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    Ref = ref_shim.spade_module().SPADEGenerator4
    ref = Ref(**CFG).eval()
    seg = spade_oracle.synthetic_input(2, S=64, seed=3)
    z = torch.randn(2, CFG["nz"], generator=torch.Generator().manual_seed(5))
    out = {"seg": seg.numpy(), "z": z.numpy()}
    out["state_sha256"] = np.frombuffer(state_checksum(ref.state_dict()).encode(), dtype=np.uint8)
    for tag, model in (("f32", ref), ("f64", copy.deepcopy(ref).double())):
        taps = {}
        hooks = [getattr(model, n).register_forward_hook(lambda m, i, o, n=n: taps.__setitem__(n, o.detach().clone())) for n in spade_oracle.BLOCKS]
        hooks.append(model.conv_img.register_forward_hook(lambda m, i, o: taps.__setitem__("pre_tanh", o.detach().clone())))
        with torch.no_grad():
            y = model(seg.to(next(model.parameters()).dtype), z.to(next(model.parameters()).dtype))
        for h in hooks:
            h.remove()
        out["out_" + tag] = y.numpy()
        for k, v in taps.items():
            out[k + "_" + tag] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "spade_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
