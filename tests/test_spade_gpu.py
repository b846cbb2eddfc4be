"""GPU parity of the SPADEGenerator4 path (csrc/spade.cu through the drop-in module) against (1) vectors computed by the
UNMODIFIED reference (tests/golden/spade_small.npz) and (2) the fp64 oracle at tensor-core-eligible and full sizes.
Asserted on every block output and on the PRE-tanh conv_img output (the tanh of a random-init generator
saturates), max-norm 1e-4 or 3x the fp32 reference's own distance from fp64."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import spade_oracle as so

pytestmark = pytest.mark.gpu
spade = importlib.import_module("sln_b200.models.SPADE_related")
_lib = importlib.import_module("sln_b200._lib")
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spade_small.npz")
NAMES = so.BLOCKS + ("pre_tanh",)


def _check(name, got, truth, ref32=None, tol=1e-4):
    got, truth = got.detach().double().cpu(), torch.as_tensor(truth).double()
    scale = truth.abs().max().item()
    err = (got - truth).abs().max().item() / max(scale, 1e-30)
    noise = 0.0 if ref32 is None else (torch.as_tensor(ref32).double() - truth).abs().max().item() / max(scale, 1e-30)
    assert err <= max(tol, 3 * noise), "%s: max-norm error %.3e (fp32 reference noise %.3e)" % (name, err, noise)
    return err


def _check_image(out, truth, ref32, pre_truth):
    """The tanh image: |d tanh| <= 1, so the pre-tanh bound (1e-4 of max|pre_tanh|) carries over; or 3x the fp32 reference noise."""
    truth, pre = torch.as_tensor(truth).double(), torch.as_tensor(pre_truth).double()
    err = (out.detach().double().cpu() - truth).abs().max().item()
    noise = (torch.as_tensor(ref32).double() - truth).abs().max().item()
    assert err <= max(1e-4 * pre.abs().max().item(), 3 * noise), "tanh image: max abs error %.3e (fp32 reference noise %.3e)" % (err, noise)


def _run(model, seg, z):
    model.taps = {}
    with torch.no_grad():
        out = model.to(DEV)(seg.to(DEV), z.to(DEV))
    torch.cuda.synchronize()
    taps, model.taps = model.taps, None
    return out, taps


def test_reference_golden_small_generator():
    """ngf=8: narrow layers take the FP32 SIMT implicit-GEMM / direct modulation paths; wide ones the tcgen05 path."""
    g = np.load(GOLD)
    torch.manual_seed(0)
    m = spade.SPADEGenerator4(semantic_nc=41, target_nc=3, nz=16, ngf=8, norm='spectralspadelayer3x3', crop_size=64, n_up='normal').eval()
    out, taps = _run(m, torch.from_numpy(g["seg"]), torch.from_numpy(g["z"]))
    for n in NAMES:
        _check(n, taps[n], g[n + "_f64"], g[n + "_f32"])
    _check_image(out, g["out_f64"], g["out_f32"], g["pre_tanh_f64"])


@pytest.mark.parametrize("ngf,crop,B", [(16, 64, 2), (32, 128, 1)])
def test_tensor_core_sizes_vs_fp64_oracle(ngf, crop, B):
    torch.manual_seed(1)
    m = spade.SPADEGenerator4(semantic_nc=41, target_nc=3, nz=32, ngf=ngf, norm='spectralspadelayer3x3', crop_size=crop, n_up='normal').eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    seg = so.synthetic_input(B, S=crop, seed=2)
    z = torch.randn(B, 32, generator=torch.Generator().manual_seed(3))
    t64, t32 = {}, {}
    with torch.no_grad():
        y64 = so.forward(sd, seg, z, ngf, crop // 32, torch.float64, t64)
        y32 = so.forward(sd, seg, z, ngf, crop // 32, torch.float32, t32)
    lib = _lib.load()
    n0 = lib.sln_launch_count()
    out, taps = _run(m, seg, z)
    assert lib.sln_launch_count() - n0 > 100
    for n in NAMES:
        _check(n, taps[n], t64[n], t32[n])
    _check_image(out, y64, y32, t64["pre_tanh"])
    # batch independence (every op of the eval-mode generator is per-sample): sample 0 alone == sample 0 in the batch, bit for bit
    if B > 1:
        out1, _ = _run(m, seg[:1], z[:1])
        assert torch.equal(out1[0], out[0])


def test_full_size_generator_vs_fp64_oracle():
    """BASELINE configs[3] architecture (ngf=64, 256x256, 41 channels, nz=256) at batch 1: 305 GFLOP through the tcgen05 path."""
    torch.manual_seed(0)
    m = spade.SPADEGenerator4(semantic_nc=41, target_nc=3, nz=256, ngf=64, norm='spectralspadelayer3x3', crop_size=256, n_up='normal').eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    seg = so.synthetic_input(1, S=256, seed=0)
    z = torch.randn(1, 256, generator=torch.Generator().manual_seed(0))
    t64, t32 = {}, {}
    with torch.no_grad():
        so.forward(sd, seg, z, 64, 8, torch.float64, t64)
        so.forward(sd, seg, z, 64, 8, torch.float32, t32)
    out, taps = _run(m, seg, z)
    for n in NAMES:
        _check(n, taps[n], t64[n], t32[n])
    assert torch.isfinite(out).all() and out.abs().max().item() <= 1.0


def test_errors_are_loud():
    torch.manual_seed(0)
    m = spade.SPADEGenerator4(semantic_nc=41, target_nc=3, nz=16, ngf=8, norm='spectralspadelayer3x3', crop_size=64, n_up='normal').eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 41, 64, 64), torch.zeros(1, 16))
    with pytest.raises(NotImplementedError):
        spade.SPADEGenerator4(semantic_nc=41, target_nc=3, nz=16, ngf=8, norm='spectralspadelayer3x3', crop_size=64, n_up='more')
    with pytest.raises(NotImplementedError):
        m.train().to(DEV)(torch.zeros(1, 41, 64, 64, device=DEV), torch.zeros(1, 16, device=DEV))


def test_graphed_forward_replays_the_eager_forward():
    """GraphedForward (one CUDA graph per input shape) returns bit-for-bit what the eager forward returns, for new inputs too."""
    torch.manual_seed(0)
    m = spade.SPADEGenerator4(semantic_nc=41, target_nc=3, nz=16, ngf=16, norm='spectralspadelayer3x3', crop_size=64, n_up='normal').eval().to(DEV)
    seg = so.synthetic_input(1, S=64, seed=5).to(DEV)
    z = torch.randn(1, 16, generator=torch.Generator().manual_seed(1)).to(DEV)
    run = spade.GraphedForward(m, seg, z)
    for seed in (2, 3):
        seg2 = so.synthetic_input(1, S=64, seed=10 + seed).to(DEV)
        z2 = torch.randn(1, 16, generator=torch.Generator().manual_seed(seed)).to(DEV)
        got = run(seg2, z2).clone()
        with torch.no_grad():
            want = m(seg2, z2)
        torch.cuda.synchronize()
        assert torch.equal(got, want)
    with pytest.raises(RuntimeError):
        spade.GraphedForward(m.train(), seg, z)
