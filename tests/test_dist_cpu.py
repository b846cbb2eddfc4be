"""Host-side logic of the N > 1 path on CPU: world-size-2 gloo processes exercise (1) the scene sharding of a global batch
(block-diagonal graph: ranks own contiguous scenes, node ids re-based), (2) the gradient all-reduce + 1/world scaling that
VAETrainStep applies between the backward graph and the Adam launch, checked against the single-process oracle: the averaged
per-rank gradients of per-rank losses equal the gradient of the mean of the two rank losses.  No CUDA involved."""
import importlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vae_oracle as vo

syn = importlib.import_module("sln_b200.data.synthetic")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sd0, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    objs, triples, boxes, angles, attrs = syn.shard_batch(syn.synthetic_batch(4, 6, seed=5), rank, world)
    sd = vo.leaf_state(sd0, torch.float64)
    eps = torch.randn(objs.size(0), 8, generator=torch.Generator().manual_seed(100 + rank)).double()
    mu, lv, bp, ap = vo.forward(sd, objs, triples, boxes.double(), angles, attrs, eps, 2, True, False, {})
    total, _ = vo.losses(boxes.double(), bp, angles, ap, mu, lv, 0.1)
    total.backward()
    keys = [k for k, v in sd.items() if v.is_floating_point() and v.requires_grad]
    flat = torch.cat([(sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])).reshape(-1) for k in keys])
    dist.all_reduce(flat)                 # what VAETrainStep._allreduce does with the flat gradient arena
    flat *= 1.0 / world                   # grad_scale = 1/world inside the Adam kernel
    if rank == 0:
        ret["flat"] = flat.clone()
        ret["loss"] = float(total)
    ret["n%d" % rank] = (objs.size(0), triples.size(0), int(triples[:, [0, 2]].max()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_average_matches_single_process():
    Model = importlib.import_module("sln_b200.models.Sg2ScVAE_model").Sg2ScVAEModel
    torch.manual_seed(42)
    m = Model(syn.default_vocab(), embedding_dim=8, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode='feedforward',
              gconv_num_layers=2, mlp_normalization='none', vec_noise_dim=0, layout_noise_dim=32, use_AE=False)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sd0, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # single process: mean over the two shards' losses, same per-rank eps
    full = syn.synthetic_batch(4, 6, seed=5)
    sd = vo.leaf_state(sd0, torch.float64)
    tot = 0.0
    for r in range(2):
        objs, triples, boxes, angles, attrs = syn.shard_batch(full, r, 2)
        assert ret["n%d" % r] == (objs.size(0), triples.size(0), objs.size(0) - 1)     # node ids re-based to the shard
        eps = torch.randn(objs.size(0), 8, generator=torch.Generator().manual_seed(100 + r)).double()
        mu, lv, bp, ap = vo.forward(sd, objs, triples, boxes.double(), angles, attrs, eps, 2, True, False, {})
        tot = tot + vo.losses(boxes.double(), bp, angles, ap, mu, lv, 0.1)[0] / 2
    tot.backward()
    keys = [k for k, v in sd.items() if v.is_floating_point() and v.requires_grad]
    want = torch.cat([(sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])).reshape(-1) for k in keys])
    assert torch.allclose(ret["flat"], want, rtol=1e-10, atol=1e-12)


def test_shard_batch_partitions_scenes_contiguously():
    full = syn.synthetic_batch(6, 5, seed=1)
    seen = 0
    for r in range(3):
        objs, triples, boxes, angles, attrs = syn.shard_batch(full, r, 3)
        assert objs.size(0) == 10 and boxes.shape == (10, 6) and triples[:, [0, 2]].min() >= 0 and triples[:, [0, 2]].max() < 10
        seen += objs.size(0)
    assert seen == full[1].size(0)
