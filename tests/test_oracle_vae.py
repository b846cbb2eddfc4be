"""Pin the CPU oracle (oracle/vae_oracle.py) against vectors produced by the real reference modules."""
import json
import os

import pytest
import torch

from helpers import GOLD, check_close, load_golden, syn
from oracle import ref_shim, vae_oracle as vo

CASES = ["vae_small_batch_train", "vae_small_batch_eval", "vae_small_none", "vae_small_recurrent"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden_fp64(name):
    meta, sd0, inp, f32, f64 = load_golden(name)
    sd = vo.leaf_state(sd0, torch.float64)
    stats = {}
    mu, lv, bp, ap = vo.forward(sd, inp["objs"], inp["triples"], inp["boxes"].double(), inp["angles"], inp["attrs"], inp["eps"].double(),
                                meta["layers"], meta["training"], False, stats)
    total, parts = vo.losses(inp["boxes"].double(), bp, inp["angles"], ap, mu, lv, meta["kl_weight"])
    total.backward()
    for k, v in (("mu", mu), ("logvar", lv), ("boxes_pred", bp), ("angles_pred", ap), ("total", total)):
        check_close(k, v, f64[k], tol=1e-9)
    for k, v in parts.items():
        check_close(k, v, f64["loss_" + k], tol=1e-9)
    for k in [k for k in f64 if k.startswith("grad.")]:
        g = sd[k[5:]].grad
        g = torch.zeros_like(sd[k[5:]]) if g is None else g
        check_close(k, g, f64[k], tol=1e-7, abs_floor=1e-12)
    for k, v in stats.items():
        check_close("after." + k, v.double(), f64["after." + k], tol=1e-6)


def test_oracle_known_answer_appendix_d():
    """fixture graph, seed-42 E=64 weights; needs the same torch RNG stream as the build container."""
    from helpers import our_model
    with open(os.path.join(GOLD, "vae_kat.json")) as f:
        kat = json.load(f)
    objs, triples, boxes, angles, attrs = syn.fixture_graph()
    for key, want in kat.items():
        norm, mode = key.split("/")
        m = our_model(E=64, layers=5, norm=norm, use_AE=True)   # same ctor order + seed => same weights as the reference
        sd = vo.leaf_state(m.state_dict(), torch.float64, requires_grad=False)
        with torch.no_grad():
            mu, lv, bp, ap = vo.forward(sd, objs, triples, boxes.double(), angles, attrs, None, 5, mode == "train", True)
        tol = 2e-3 if norm == "batch" and mode == "train" else 1e-4   # 6-row BatchNorm amplifies fp32 noise (App. F)
        assert abs(mu.sum().item() - want["mu_sum"]) < tol * abs(want["mu_sum"]) + 1e-3
        assert abs(bp.sum().item() - want["boxes_sum"]) < max(tol * 50, 1e-3)
        if norm == "none":
            assert ap.argmax(1).tolist() == want["argmax"]


def test_oracle_matches_live_reference_fp64():
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    import types
    Ref = ref_shim.vae_model_class()
    torch.manual_seed(5)
    m = Ref(syn.default_vocab(), embedding_dim=16, batch_size=4, train_3d=True, decoder_cat=True, gconv_mode='feedforward',
            gconv_num_layers=3, mlp_normalization='batch', vec_noise_dim=0, layout_noise_dim=32, use_AE=False).double()
    _, objs, boxes, triples, angles, attrs, _, _ = syn.synthetic_batch(4, 8, seed=1)
    eps = torch.randn(objs.size(0), 16, dtype=torch.float64)
    orig = torch.randn_like
    torch.randn_like = lambda t: eps
    try:
        mu, lv, bp, ap = m(objs, triples, boxes.double(), angles, attrs, None)
    finally:
        torch.randn_like = orig
    tot, _ = ref_shim.reference_losses()(types.SimpleNamespace(use_AE=False), m, boxes.double(), bp, angles, ap, mu=mu, logvar=lv, KL_weight=0.1)
    sd = vo.leaf_state(m.state_dict(), torch.float64)
    mu2, lv2, bp2, ap2 = vo.forward(sd, objs, triples, boxes.double(), angles, attrs, eps, 3, True, False)
    tot2, _ = vo.losses(boxes.double(), bp2, angles, ap2, mu2, lv2, 0.1)
    for a, b in ((mu, mu2), (lv, lv2), (bp, bp2), (ap, ap2), (tot, tot2)):
        assert (a - b).abs().max().item() < 1e-10


def test_synthetic_batch_shape_matches_config2():
    ids, objs, boxes, triples, angles, attrs, o2i, t2i = syn.synthetic_batch(64, 32, seed=42)
    assert objs.shape == (2048,) and triples.shape == (3968, 3) and boxes.shape == (2048, 6)
    assert int(triples[:, [0, 2]].max()) < 2048 and int(triples[:, 1].max()) <= 10
    assert (objs.view(64, 32)[:, -1] == 0).all() and (objs.view(64, 32)[:, :-1] > 0).all()
    # graphs are block-diagonal: no edge crosses a scene (suncg_dataset.py:318-325)
    assert (o2i[triples[:, 0]] == o2i[triples[:, 2]]).all() and (o2i[triples[:, 0]] == t2i).all()
