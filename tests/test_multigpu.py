"""Multi-rank correctness ON HARDWARE: two NCCL ranks run VAETrainStep on the two halves of a global batch and must end,
after 3 optimizer steps, with the parameters of the single-GPU step on the whole batch — with mlp_normalization='none' (gradient
all-reduce only) and with 'batch' under bn_policy='sync' (BatchNorm statistics exchanged inside the finalising kernels over NVLink
peer memory; reference models/graph.py:14-15 normalises over ALL rows of the batch).  Skips below 2 GPUs."""
import importlib
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

syn = importlib.import_module("sln_b200.data.synthetic")
SCENES, NODES, E, L, STEPS = 16, 8, 16, 2, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(norm, dev):
    Model = importlib.import_module("sln_b200.models.Sg2ScVAE_model").Sg2ScVAEModel
    torch.manual_seed(42)
    m = Model(syn.default_vocab(), embedding_dim=E, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode='feedforward',
              gconv_num_layers=L, mlp_normalization=norm, vec_noise_dim=0, layout_noise_dim=32, use_AE=False)
    return m.float().to(dev).train()


def _run_steps(model, batch, eps_rows, dev, **kw):
    """STEPS train steps from the seeded initial state (capture()'s warm-up steps are undone first) -> (losses per step, state_dict)."""
    sutils = importlib.import_module("sln_b200.utils")
    init = {k: v.clone() for k, v in model.state_dict().items()}
    objs, triples, boxes, angles, attrs = batch
    step = sutils.VAETrainStep(model, objs.size(0), triples.size(0), lr=1e-3, kl_weight=0.1, sample_eps=False, **kw).capture()
    with torch.no_grad():
        for k, v in model.state_dict().items():
            v.copy_(init[k])
    step.m.zero_(); step.v.zero_(); step.step_count.zero_()
    losses, grads = [], None
    for it in range(STEPS):
        step.load_batch([t.to(dev) for t in batch])
        step.epsn.copy_(eps_rows[it])
        losses.append(step.run().tolist())
        if it == 0:       # the (all-reduced, still un-averaged) gradients of the first step
            scale = 1.0 / kw.get("world_size", 1)
            grads = {k: (p.grad.detach().cpu().clone() * scale) for k, p in model.named_parameters()}
    torch.cuda.synchronize(dev)
    return losses, {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, grads


def _global_inputs():
    full = syn.synthetic_batch(SCENES, NODES, seed=5)
    g = torch.Generator().manual_seed(9)
    eps = [torch.randn(full[1].size(0), E, generator=g) for _ in range(STEPS)]
    return full, eps


def _worker(rank, world, port, norm, policy, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        full, eps = _global_inputs()
        shard = syn.shard_batch(full, rank, world)
        o2i = full[6]
        per = SCENES // world
        rows = (o2i >= rank * per) & (o2i < (rank + 1) * per)
        losses, sd, grads = _run_steps(_model(norm, dev), shard, [e[rows] for e in eps], dev, process_group=dist.group.WORLD,
                                       world_size=world, bn_policy=policy)
        if rank == 0:
            torch.save({"losses": losses, "sd": sd, "grads": grads}, out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("norm,policy", [("none", "local"), ("batch", "sync")])
def test_two_rank_sharded_step_equals_single_gpu_step(tmp_path, norm, policy):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = os.path.join(str(tmp_path), "two_rank.pt")
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, norm, policy, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    two = torch.load(out)
    full, eps = _global_inputs()
    dev = torch.device("cuda", 0)
    batch = (full[1], full[3], full[2], full[4], full[5])
    losses, sd, grads = _run_steps(_model(norm, dev), batch, eps, dev)
    # (1) the averaged gradient of the first sharded step == the gradient of the single-GPU step on the whole batch.  Max-norm over
    # ALL tensors: under training-mode BatchNorm the gradient of every bias in front of a BatchNorm (and of box_embeddings.bias) is
    # analytically zero, so those tensors hold pure fp32 noise that only an absolute tolerance can judge.
    gmax = max(g.abs().max().item() for g in grads.values())
    degenerate = set()
    for k, g in grads.items():
        d = (g.double() - two["grads"][k].double()).abs().max().item()
        assert d <= 2e-5 * gmax, ("grad", k, d, gmax)
        if g.abs().max().item() < 1e-5 * gmax:
            degenerate.add(k)
    # (2) the parameters after STEPS Adam steps.  Adam turns the noise of the zero-gradient tensors into +-lr steps whose signs are
    # implementation noise (also between two single-GPU runs with different reduction orders): those tensors are skipped here.
    worst = 0.0
    for k, v in sd.items():
        if not v.is_floating_point():
            assert torch.equal(v, two["sd"][k]), k
            continue
        if k in degenerate or k.endswith("running_mean"):
            continue          # running_mean tracks mean(xW + b): it inherits the +-lr noise of the (zero-gradient) bias b; running_var does not
        diff = (v.double() - two["sd"][k].double()).abs()
        scale = max(v.abs().max().item(), 1e-3)
        tol = torch.full_like(diff, 2e-5 * scale + 2e-6)
        g = grads.get(k)
        if g is not None and g.shape == v.shape:
            # element-wise version of the same argument.  Adam moves an element by ~lr * g / |g| per step, so a gradient perturbation dg
            # (check (1) bounds it by 2e-5 * gmax; the split-K weight gradients are fp32 atomics, i.e. run-to-run noise) shifts the element
            # by ~lr * dg / |g| per step — up to ~2 lr where |g| is at the noise floor.  Elements with solid gradients stay tightly bound.
            tol = tol + STEPS * 1e-3 * torch.clamp(4e-5 * gmax / g.double().abs().clamp_min(1e-30), max=2.2)
        worst = max(worst, (diff / scale).max().item())
        bad = diff > tol
        assert not bool(bad.any()), (k, diff[bad].max().item(), scale, int(bad.sum()))
    assert len(degenerate) < len(grads) // 2
    assert all(torch.isfinite(torch.tensor(l)).all() for l in two["losses"])
    print("max relative parameter difference 2 ranks vs 1 GPU (%s, %s): %.2e" % (norm, policy, worst))


@pytest.mark.gpu
def test_local_batchnorm_policy_differs_from_global_statistics(tmp_path):
    """Sanity of the test above: with per-rank statistics (P2, DDP semantics) the sharded BatchNorm run does NOT reproduce the
    single-GPU parameters — the equality under 'sync' is due to the exchange, not to a coincidence of the data."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = os.path.join(str(tmp_path), "two_rank_local.pt")
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, "batch", "local", out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs)
    two = torch.load(out)
    full, eps = _global_inputs()
    dev = torch.device("cuda", 0)
    _, sd, grads = _run_steps(_model("batch", dev), (full[1], full[3], full[2], full[4], full[5]), eps, dev)
    gmax = max(g.abs().max().item() for g in grads.values())
    diff = max((grads[k].double() - two["grads"][k].double()).abs().max().item() for k in grads)
    assert diff > 1e-3 * gmax
