"""Plain SPADEGenerator (reference models/SPADE_related.py:151-346) — the class BASELINE.json's north_star names: goldens from the
UNMODIFIED reference module (oracle/gen_golden_spade_plain.py), for both parameter-free norms ('instance', eval-mode 'batch').
CPU: the oracle port and the drop-in module's seeded state_dict against the golden.  GPU: the CUDA path against the golden."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import spade_oracle as so
from oracle.gen_golden_spade import state_checksum
from oracle.gen_golden_spade_plain import BASE, randomise_bn

sp = importlib.import_module("sln_b200.models.SPADE_related")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAPS = so.PLAIN_BLOCKS + ("conv_img_pre",)


def _load(kind):
    return np.load(os.path.join(GOLD, "spade_plain_%s.npz" % kind))


def _model(kind):
    torch.manual_seed(0)
    m = sp.SPADEGenerator(norm='spectralspade%s3x3' % kind, **BASE).eval()
    randomise_bn(m)
    return m


@pytest.mark.parametrize("kind", ["instance", "batch"])
def test_state_dict_and_oracle_match_the_reference(kind):
    z = _load(kind)
    m = _model(kind)
    assert state_checksum(m.state_dict()) == bytes(z["state_sha256"]).decode()        # same module tree, creation order and init as the reference
    taps = {}
    with torch.no_grad():
        y = so.forward_plain(m.state_dict(), torch.from_numpy(z["seg"]), torch.from_numpy(z["z"]), BASE["ngf"], 2, kind, torch.float64, taps)
    assert np.abs(y.numpy() - z["out_f64"]).max() <= 1e-10
    assert np.abs(taps["pre_tanh"].numpy() - z["pre_tanh_f64"]).max() <= 1e-9 * max(1.0, np.abs(z["pre_tanh_f64"]).max())
    for k in TAPS:
        assert np.abs(taps[k].numpy() - z[k + "_f64"]).max() <= 1e-5 * max(1.0, np.abs(z[k + "_f64"]).max()), k


def test_cpu_tensors_and_train_mode_raise():
    m = _model("instance")
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 41, 64, 64), torch.zeros(1, 16))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["instance", "batch"])
def test_cuda_path_matches_the_reference_golden(kind):
    z = _load(kind)
    m = _model(kind).cuda()
    m.taps = {}
    with torch.no_grad():
        y = m(torch.from_numpy(z["seg"]).cuda(), torch.from_numpy(z["z"]).cuda())
    torch.cuda.synchronize()
    noise = np.abs(z["pre_tanh_f32"].astype(np.float64) - z["pre_tanh_f64"]).max()    # the reference's own fp32 distance from fp64
    for k in TAPS:
        want = z[k + "_f64"].astype(np.float64)
        err = np.abs(m.taps[k].cpu().numpy() - want).max() / max(np.abs(want).max(), 1e-30)
        assert err <= 1e-4, (k, err)
    want = z["pre_tanh_f64"]
    err = np.abs(m.taps["pre_tanh"].cpu().numpy() - want).max()
    assert err <= max(1e-4 * np.abs(want).max(), 3 * noise), (err, noise)
    assert np.abs(y.cpu().numpy() - z["out_f64"]).max() <= 1e-4
    m.train()
    with pytest.raises(NotImplementedError):
        m(torch.from_numpy(z["seg"]).cuda(), torch.from_numpy(z["z"]).cuda())
