"""Reference checkpoint compatibility.

The reference's train loop stores ``<name>_with_model.pt`` = a dict with ``model_state`` (model.state_dict()), ``optim_state``
(torch.optim.Adam.state_dict()), ``model_kwargs`` (the Sg2ScVAEModel constructor kwargs), ``vocab``, ``args`` and ``counters`` {t, epoch}
(reference train.py:30-57,93-98) and the test drivers restore it with ``build_model`` + ``load_state_dict(checkpoint['model_state'])``
(testing/test_VAE.py:19-25, test_render_refine.py:259-263).  The drop-in model has the reference's module tree, so the tensors load
unchanged; these helpers do the two steps around it: build the model from the stored kwargs, and hand the Adam moments to whichever
optimizer runs here (torch.optim.Adam, FusedAdam, or the captured VAETrainStep).
"""
import torch

from .models.Sg2ScVAE_model import Sg2ScVAEModel

MODEL_KWARGS = ("vocab", "batch_size", "train_3d", "decoder_cat", "embedding_dim", "gconv_mode", "gconv_num_layers", "mlp_normalization",
                "vec_noise_dim", "layout_noise_dim", "use_AE")       # build_dataset_model.py:40-52


def load_reference_checkpoint(path_or_dict, device=None, strict=True):
    """-> (model, checkpoint dict).  `path_or_dict`: a ``*_with_model.pt`` file written by the reference's train.py (or by
    save_reference_checkpoint below), or the already-loaded dict.  The model is built from checkpoint['model_kwargs'] (falling back to
    checkpoint['args'] + checkpoint['vocab'] for kwargs an older file lacks), loaded with checkpoint['model_state'] and put in train or
    eval mode the way train.py:26-30 does (eval once counters.t >= args.eval_mode_after >= 0)."""
    ckpt = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu", weights_only=False)
    if ckpt.get("model_state") is None:
        raise ValueError("checkpoint holds no 'model_state' (a *_no_model.pt file? train.py:107-113 strips it)")
    kwargs = dict(ckpt.get("model_kwargs") or {})
    args = ckpt.get("args") or {}
    for k in MODEL_KWARGS:
        if k not in kwargs:
            if k == "vocab" and ckpt.get("vocab") is not None:
                kwargs[k] = ckpt["vocab"]
            elif k in args:
                kwargs[k] = args[k]
    missing = [k for k in ("vocab", "embedding_dim") if k not in kwargs]
    if missing:
        raise ValueError("checkpoint lacks model_kwargs %s" % missing)
    model = Sg2ScVAEModel(**kwargs)
    model.load_state_dict(ckpt["model_state"], strict=strict)
    t = (ckpt.get("counters") or {}).get("t")
    after = args.get("eval_mode_after", -1)
    if t is not None and after is not None and 0 <= after <= t:
        model.eval()
    else:
        model.train()
    if device is not None:
        model = model.float().to(device)
    return model, ckpt


def restore_optimizer(optimizer_or_step, ckpt):
    """checkpoint['optim_state'] (torch.optim.Adam format over model.parameters() order, train.py:25,95) -> a torch optimizer / FusedAdam
    (load_state_dict) or a VAETrainStep (load_optim_state_dict)."""
    state = ckpt.get("optim_state")
    if state is None:
        return False
    if hasattr(optimizer_or_step, "load_optim_state_dict"):
        optimizer_or_step.load_optim_state_dict(state)
    else:
        optimizer_or_step.load_state_dict(state)
    return True


def save_reference_checkpoint(path, model, optim_state, model_kwargs, vocab, t, epoch, args=None, extra=None):
    """Write a file the reference's own drivers can restore (same keys as train.py:30-57,93-98)."""
    ckpt = {"args": dict(args or {}), "vocab": vocab, "model_kwargs": model_kwargs, "counters": {"t": t, "epoch": epoch},
            "model_state": model.state_dict(), "optim_state": optim_state}
    ckpt.update(extra or {})
    torch.save(ckpt, path)
    return ckpt
