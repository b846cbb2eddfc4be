"""TEST INFRASTRUCTURE ONLY — tests/golden/spade_plain_*.npz from the UNMODIFIED reference SPADEGenerator (models/SPADE_related.py:151-250;
the plain SPADE / SPADEResnetBlock / SEResBlock2 classes), CPU, this container.

    python oracle/gen_golden_spade_plain.py      # needs /root/reference

Reduced generators (ngf=8, nz=16, crop 64, 41-channel input, batch 2, seeded weights) for both parameter-free norms the class accepts:
'spectralspadeinstance3x3' and 'spectralspadebatch3x3' (BatchNorm running statistics randomised so that eval mode is not the identity).
Stored: input, z, fp32 + fp64 outputs, the pre-tanh image and every block output, and the SHA-256 of the seeded state_dict."""
import copy
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, spade_oracle  # noqa: E402
from oracle.gen_golden_spade import state_checksum  # noqa: E402

BASE = dict(semantic_nc=41, target_nc=3, nz=16, ngf=8, crop_size=64, n_up='normal')


def randomise_bn(model, seed=11):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
            m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))


def main():
    warnings.filterwarnings("ignore")
    Ref = ref_shim.spade_module().SPADEGenerator
    for kind in ("instance", "batch"):
        torch.manual_seed(0)
        ref = Ref(norm='spectralspade%s3x3' % kind, **BASE).eval()
        randomise_bn(ref)
        seg = spade_oracle.synthetic_input(2, S=64, seed=3)
        z = torch.randn(2, BASE["nz"], generator=torch.Generator().manual_seed(5))
        out = {"seg": seg.numpy(), "z": z.numpy(), "state_sha256": np.frombuffer(state_checksum(ref.state_dict()).encode(), dtype=np.uint8)}
        for tag, model in (("f32", ref), ("f64", copy.deepcopy(ref).double())):
            taps = {}
            hooks = [getattr(model, n).register_forward_hook(lambda m, i, o, n=n: taps.__setitem__(n, o.detach().clone())) for n in spade_oracle.PLAIN_BLOCKS]
            hooks.append(model.conv_img_pre.register_forward_hook(lambda m, i, o: taps.__setitem__("conv_img_pre", o.detach().clone())))
            hooks.append(model.conv_img.register_forward_hook(lambda m, i, o: taps.__setitem__("pre_tanh", o.detach().clone())))
            dt = next(model.parameters()).dtype
            with torch.no_grad():
                y = model(seg.to(dt), z.to(dt))
            for h in hooks:
                h.remove()
            out["out_" + tag] = y.numpy()
            out["pre_tanh_" + tag] = taps["pre_tanh"].numpy()
            if tag == "f64":      # block outputs of the fp64 run, stored as float32 (compared at 1e-4)
                for k, v in taps.items():
                    if k != "pre_tanh":
                        out[k + "_f64"] = v.numpy().astype(np.float32)
        path = os.path.join(ROOT, "tests", "golden", "spade_plain_%s.npz" % kind)
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
