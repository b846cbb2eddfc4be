"""The import switch of INTEGRATION.md section 1 against the reference's OWN caller code (build container only: /root/reference does not
exist on the GPU box), and the checkpoint format of train.py:30-57,93-98 both ways."""
import importlib
import os
import sys
import types

import pytest
import torch

from helpers import syn
from oracle import ref_shim

ckpt_mod = importlib.import_module("sln_b200.checkpoint")
sutils = importlib.import_module("sln_b200.utils")
OUR = importlib.import_module("sln_b200.models.Sg2ScVAE_model")

ARGS = dict(batch_size=128, train_3d=True, decoder_cat=True, embedding_dim=16, gconv_mode="feedforward", gconv_num_layers=2,
            mlp_normalization="batch", vec_noise_dim=0, layout_noise_dim=32, use_AE=False, multigpu=False, eval_mode_after=-1)


def _reference_modules_clean():
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k in ("build_dataset_model", "utils", "data", "data.suncg_dataset")]:
        del sys.modules[k]


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
def test_reference_build_model_runs_through_the_import_switch(tmp_path):
    """Alias the hot modules exactly as INTEGRATION.md section 1 says, then import the reference's build_dataset_model and call ITS
    build_model (build_dataset_model.py:39-56): the object train.py:13-15 goes on to use is this package's model, with the reference's
    state_dict keys, and the checkpoint train.py:93-98 would write restores through load_reference_checkpoint."""
    _reference_modules_clean()
    try:
        RefModel = ref_shim.vae_model_class()                                   # the unmodified reference class (for the key / shape check)
        torch.manual_seed(42)
        ref = RefModel(**{k: ARGS[k] for k in ARGS if k not in ("multigpu", "eval_mode_after")}, vocab=syn.default_vocab())
        _reference_modules_clean()
        sys.modules["models.graph"] = importlib.import_module("sln_b200.models.graph")
        sys.modules["models.Sg2ScVAE_model"] = OUR
        bdm = importlib.import_module("build_dataset_model")                    # the reference's file, from /root/reference
        assert os.path.realpath(bdm.__file__).startswith(os.path.realpath(ref_shim.REF_ROOT))
        torch.manual_seed(42)
        model, kwargs = bdm.build_model(types.SimpleNamespace(**ARGS), syn.default_vocab())
        assert type(model) is OUR.Sg2ScVAEModel
        sd, rsd = model.state_dict(), ref.state_dict()
        assert list(sd.keys()) == list(rsd.keys())
        assert all(torch.equal(sd[k], rsd[k]) for k in sd)                      # same creation order + init under the same seed
        # train.py:14-15 (minus .cuda()): parameters feed torch.optim.Adam; train.py:93-98: the checkpoint dict
        optimizer = torch.optim.Adam(model.float().parameters(), lr=1e-4)
        for p in model.parameters():
            p.grad = torch.full_like(p, 1e-3)
        optimizer.step()
        path = os.path.join(str(tmp_path), "x_with_model.pt")
        torch.save({"args": dict(ARGS), "vocab": syn.default_vocab(), "model_kwargs": kwargs, "counters": {"t": 7, "epoch": 1},
                    "model_state": model.state_dict(), "optim_state": optimizer.state_dict()}, path)
        m2, ck = ckpt_mod.load_reference_checkpoint(path)
        assert type(m2) is OUR.Sg2ScVAEModel and m2.training
        assert all(torch.equal(a, b) for a, b in zip(m2.state_dict().values(), model.state_dict().values()))
        ref.load_state_dict(ck["model_state"])                                  # and the reference class reads the same file
        fa = sutils.FusedAdam(m2.parameters(), lr=1e-4)
        assert ckpt_mod.restore_optimizer(fa, ck) and fa._import_pending
        assert all(float(st["step"]) == 1.0 for st in fa.state.values())
        # hot path is CUDA-only: the forward of the switched-in model refuses CPU tensors loudly instead of falling back
        objs, triples, boxes, angles, attrs = syn.fixture_graph()
        with pytest.raises(RuntimeError, match="CUDA"):
            model(objs, triples, boxes, angles, attrs, None)
    finally:
        _reference_modules_clean()


def test_checkpoint_loader_rejects_stripped_files_and_honours_eval_mode_after(tmp_path):
    torch.manual_seed(0)
    kw = {k: ARGS[k] for k in ARGS if k not in ("multigpu", "eval_mode_after")}
    m = OUR.Sg2ScVAEModel(vocab=syn.default_vocab(), **kw)
    ck = ckpt_mod.save_reference_checkpoint(os.path.join(str(tmp_path), "a.pt"), m, None, dict(kw, vocab=syn.default_vocab()), syn.default_vocab(), 120, 3,
                                            args=dict(ARGS, eval_mode_after=100))
    m2, _ = ckpt_mod.load_reference_checkpoint(os.path.join(str(tmp_path), "a.pt"))
    assert not m2.training                                                      # train.py:26-28
    assert not ckpt_mod.restore_optimizer(torch.optim.Adam(m2.parameters()), ck)
    with pytest.raises(ValueError, match="model_state"):
        ckpt_mod.load_reference_checkpoint({k: v for k, v in ck.items() if k != "model_state"})
    # kwargs missing from an older file are recovered from args + vocab
    old = dict(ck, model_kwargs={"embedding_dim": 16})
    m3, _ = ckpt_mod.load_reference_checkpoint(old)
    assert list(m3.state_dict().keys()) == list(m.state_dict().keys())


@pytest.mark.gpu
def test_train_py_body_through_the_switch_on_gpu(tmp_path):
    """train.py:69-84 statement by statement on the switched-in modules (tensor_aug, model(...), calculate_model_losses, isfinite guard,
    zero_grad / backward / step), then the train.py:93-98 checkpoint and a resume that continues the same trajectory."""
    import math
    dev = "cuda:0"
    kw = {k: ARGS[k] for k in ARGS if k not in ("multigpu", "eval_mode_after")}
    torch.manual_seed(42)
    model = OUR.Sg2ScVAEModel(vocab=syn.default_vocab(), **kw)
    model.float().cuda()
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)
    args = types.SimpleNamespace(use_AE=False, KL_linear_decay=False, KL_loss_weight=0.1)
    batches = [syn.synthetic_batch(8, 8, seed=s) for s in range(6)]
    gen = torch.Generator(device=dev).manual_seed(1)

    def one(batch, model, optimizer):
        ids, objs, boxes, triples, angles, attributes, obj_to_img, triple_to_img = sutils.tensor_aug(batch)
        eps = torch.randn(objs.size(0), 16, device=dev, generator=gen)
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: eps
        try:
            mu, logvar, boxes_pred, angles_pred = model(objs, triples, boxes, angles, attributes, obj_to_img)
        finally:
            torch.randn_like = orig
        total_loss, losses = sutils.calculate_model_losses(args, model, boxes, boxes_pred, angles, angles_pred, mu=mu, logvar=logvar, KL_weight=args.KL_loss_weight)
        losses['total_loss'] = total_loss.item()
        assert math.isfinite(losses['total_loss'])
        optimizer.zero_grad()
        total_loss.backward()
        optimizer.step()
        return losses['total_loss']
    first = [one(b, model, optimizer) for b in batches[:3]]
    path = os.path.join(str(tmp_path), "latest_with_model.pt")
    ckpt_mod.save_reference_checkpoint(path, model, optimizer.state_dict(), dict(kw, vocab=syn.default_vocab()), syn.default_vocab(), 3, 1, args=dict(ARGS))
    state = gen.get_state()
    straight = [one(b, model, optimizer) for b in batches[3:]]
    m2, ck = ckpt_mod.load_reference_checkpoint(path, device=dev)
    opt2 = sutils.FusedAdam(m2.parameters(), lr=1e-3)                            # resume under the fused optimizer
    ckpt_mod.restore_optimizer(opt2, ck)
    gen.set_state(state)
    resumed = [one(b, m2, opt2) for b in batches[3:]]
    assert first[0] > 0 and all(abs(a - b) <= 2e-4 * abs(a) for a, b in zip(straight, resumed)), (straight, resumed)
