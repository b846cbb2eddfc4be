"""Drop-ins for the hot-path helpers of the reference's ``utils.py`` plus the fused train step.

* ``calculate_model_losses`` (reference utils.py:12-33) — same signature and return value; one fused kernel computes the
  three loss terms and their gradient seeds, and the ``losses`` dict is filled from ONE device->host read (the reference
  does three ``.item()`` syncs, utils.py:141).
* ``tensor_aug`` (utils.py:114-124), ``add_loss`` (:139-146), ``get_model_attr`` (:149-153) — unchanged behaviour.
* ``FusedAdam`` — ``torch.optim.Adam``-compatible step over flat arenas (train.py:15,82-84): 1 launch instead of a
  multi-tensor foreach sequence.
* ``VAETrainStep`` — the whole ``train.py:69-84`` loop body (H2D, forward, losses, backward, Adam) as a replayable CUDA graph.
"""
import os

import torch
import torch.nn as nn

from . import _lib


# ---------------------------------------------------------------------------------------------- reference helpers
def tensor_aug(tensors, volatile=False, use_gpu=True):
    out = []
    for t in tensors:
        v = t.cuda(non_blocking=True) if use_gpu else t
        if volatile:
            v.requires_grad = False
        out.append(v)
    return tuple(out)


def add_loss(total_loss, curr_loss, loss_dict, loss_name, weight=1):
    weighted = curr_loss * weight
    loss_dict[loss_name] = weighted.item()
    return weighted if total_loss is None else total_loss + weighted


def get_model_attr(_object, attr):
    if isinstance(_object, nn.DataParallel):
        return getattr(_object.module, attr)
    return getattr(_object, attr)


def _loss_scratch(dev, O):
    n = 16 + 12 * ((O + 63) // 64)
    return torch.zeros((n + 3) // 4, dtype=torch.int32, device=dev)


class _FusedLossFn(torch.autograd.Function):
    """total = L1(bbox) + NLL(angle) + KL_weight * KLD, with the gradient seeds produced by the same kernel."""

    @staticmethod
    def forward(ctx, bbox_pred, bbox, angles_pred, angles, mu, logvar, kl_weight, holder):
        lib = _lib.load()
        dev = bbox_pred.device
        bbox_pred, bbox, angles_pred = bbox_pred.contiguous().float(), bbox.contiguous().float(), angles_pred.contiguous().float()
        angles = angles.contiguous().long()
        O, BD, NA = bbox_pred.size(0), bbox_pred.size(1), angles_pred.size(1)
        has_kl = mu is not None
        if has_kl:
            mu, logvar = mu.contiguous().float(), logvar.contiguous().float()
        losses = torch.empty(4, device=dev, dtype=torch.float32)
        d_boxes = torch.empty_like(bbox_pred)
        d_angles = torch.empty_like(angles_pred)
        d_mu = torch.empty_like(mu) if has_kl else None
        d_logvar = torch.empty_like(logvar) if has_kl else None
        scratch = _loss_scratch(dev, O)
        _lib.check(lib.sln_vae_loss(bbox_pred.data_ptr(), bbox.data_ptr(), BD, angles_pred.data_ptr(), angles.data_ptr(), NA,
                                    _lib.ptr(mu), _lib.ptr(logvar), mu.size(1) if has_kl else 0, float(kl_weight if has_kl else 0.0), O,
                                    losses.data_ptr(), d_boxes.data_ptr(), d_angles.data_ptr(), 0, _lib.ptr(d_mu), _lib.ptr(d_logvar),
                                    scratch.data_ptr(), scratch.numel() * 4, _lib.cur_stream(dev)), "vae_loss")
        holder.append(losses)
        ctx.save_for_backward(d_boxes, d_angles, *([d_mu, d_logvar] if has_kl else []))
        ctx.has_kl = has_kl
        return losses[3].clone()

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        d_boxes, d_angles = saved[0] * g, saved[1] * g
        d_mu = saved[2] * g if ctx.has_kl else None
        d_logvar = saved[3] * g if ctx.has_kl else None
        return d_boxes, None, d_angles, None, d_mu, d_logvar, None, None


def calculate_model_losses(args, model, bbox, bbox_pred, angles, angles_pred, mu=None, logvar=None, KL_weight=None):
    """Same contract as the reference: returns (total_loss tensor, {'bbox_pred','angle_pred','KLD_Gauss'} floats)."""
    if not bbox_pred.is_cuda:
        raise RuntimeError("sln_b200.calculate_model_losses runs on CUDA only (no CPU fallback)")
    use_kl = not args.use_AE
    holder = []
    total = _FusedLossFn.apply(bbox_pred, bbox, angles_pred, angles, mu if use_kl else None, logvar if use_kl else None,
                               KL_weight if use_kl else 0.0, holder)
    vals = holder[0].tolist()   # the single device->host synchronisation of the loss computation
    losses = {'bbox_pred': vals[0], 'angle_pred': vals[1]}
    if use_kl:
        losses['KLD_Gauss'] = vals[2]
    return total, losses


# ---------------------------------------------------------------------------------------------- fused Adam
class FusedAdam(torch.optim.Optimizer):
    """Adam with torch.optim.Adam's update rule, run by sln_adam_step over flat parameter / gradient / state arenas.

    On the first ``step()`` the parameters are re-homed (``p.data`` re-pointed, values preserved) into one arena ordered
    like their gradients, so that a model whose gradients live in the Sg2ScVAEModel gradient arena is updated by a single
    launch.  Parameters whose gradients are not laid out contiguously are updated run by run.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super(FusedAdam, self).__init__(params, defaults)
        self.grad_scale = grad_scale
        self._plans = {}
        self._import_pending = False

    # -- torch.optim.Adam-compatible checkpoints (reference train.py:25,95: optimizer.state_dict() / load_state_dict) --------------
    def _export_state(self):
        """Publish the arenas as per-parameter {'step', 'exp_avg', 'exp_avg_sq'} entries of Optimizer.state (views, no copy)."""
        for plan in self._plans.values():
            step = torch.tensor(float(plan['step'].item()))
            for p in plan['params']:
                o, k = plan['offsets'][id(p)], p.numel()
                self.state[p] = {'step': step.clone(), 'exp_avg': plan['m'][o:o + k].view_as(p), 'exp_avg_sq': plan['v'][o:o + k].view_as(p)}

    def state_dict(self):
        self._export_state()
        return super(FusedAdam, self).state_dict()

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.Adam (or FusedAdam) state_dict: moments and step are imported into the arenas — at once for
        parameters that are already planned, at the first step() otherwise."""
        super(FusedAdam, self).load_state_dict(state_dict)
        self._import_pending = True
        for plan in self._plans.values():
            self._import_into(plan)

    def _import_into(self, plan):
        steps = []
        for p in plan['params']:
            st = self.state.get(p)
            if not st or 'exp_avg' not in st:
                continue
            o, k = plan['offsets'][id(p)], p.numel()
            if st['exp_avg'].data_ptr() != plan['m'][o:o + k].data_ptr():
                plan['m'][o:o + k].copy_(st['exp_avg'].reshape(-1).to(plan['m']))
                plan['v'][o:o + k].copy_(st['exp_avg_sq'].reshape(-1).to(plan['v']))
            steps.append(int(float(st['step'])))
        if steps:
            if len(set(steps)) != 1:
                raise RuntimeError("FusedAdam keeps one step counter per parameter group; the loaded state has steps %s" % sorted(set(steps)))
            plan['step'].fill_(steps[0])

    @staticmethod
    def _key(ps):
        return tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)

    def _plan(self, gi, group):
        ps = [p for p in group['params'] if p.grad is not None]
        if not ps:
            return None
        old = self._plans.get(gi)
        if old is not None and len(old['params']) == len(ps) and old['key'] == self._key(old['params']):
            return old
        for p in ps:
            if not (p.is_cuda and p.dtype == torch.float32 and p.grad.is_contiguous()):
                raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters and gradients")
        ps.sort(key=lambda p: p.grad.data_ptr())
        n = sum((p.numel() + 3) // 4 * 4 for p in ps)
        dev = ps[0].device
        arena = torch.zeros(n, device=dev, dtype=torch.float32)
        m = torch.zeros(n, device=dev, dtype=torch.float32)
        v = torch.zeros(n, device=dev, dtype=torch.float32)
        off, runs, offsets = 0, [], {}
        for p in ps:
            k, g0 = p.numel(), p.grad.data_ptr()
            cont = bool(runs) and runs[-1][2] + runs[-1][1] * 4 == g0 and runs[-1][0] + runs[-1][1] == off
            if not cont:
                off = (off + 3) // 4 * 4      # a new run starts 16-byte aligned
            arena[off:off + k].copy_(p.data.reshape(-1))
            if old is not None and id(p) in old['offsets']:   # carry optimizer state across a re-plan
                o = old['offsets'][id(p)]
                m[off:off + k].copy_(old['m'][o:o + k])
                v[off:off + k].copy_(old['v'][o:o + k])
            p.data = arena[off:off + k].view_as(p)
            offsets[id(p)] = off
            if cont:
                runs[-1][1] += k
            else:
                runs.append([off, k, g0])
            off += k
        step = old['step'] if old is not None else torch.zeros(1, device=dev, dtype=torch.int64)
        plan = dict(key=self._key(ps), arena=arena, m=m, v=v, runs=runs, step=step, offsets=offsets, params=ps)
        self._plans[gi] = plan
        if self._import_pending:
            self._import_into(plan)
        return plan

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            plan = self._plan(gi, group)
            if plan is None:
                continue
            b1, b2 = group['betas']
            st = _lib.cur_stream(plan['arena'].device)
            if gi == len(self.param_groups) - 1:
                self._import_pending = False
            for ri, (off, k, gptr) in enumerate(plan['runs']):
                if gptr % 16:
                    raise RuntimeError("FusedAdam: gradient run not 16-byte aligned")
                _lib.check(lib.sln_adam_step(plan['arena'].data_ptr() + off * 4, gptr, plan['m'].data_ptr() + off * 4,
                                             plan['v'].data_ptr() + off * 4, k, group['lr'], b1, b2, group['eps'],
                                             group['weight_decay'], self.grad_scale, plan['step'].data_ptr(), int(ri == 0), st),
                           "adam_step")
        return loss


# ---------------------------------------------------------------------------------------------- SyncBatchNorm (multi-GPU policy P1)
class BnSyncTable(object):
    """Peer-mapped buffers + the device table (``sln_bn_sync``, include/sln_b200.h) of the in-kernel SyncBatchNorm exchange.

    Every rank allocates one symmetric buffer [receive area | arrival flags] (torch.distributed._symmetric_memory: CUDA VMM
    allocations mapped into every peer over NVLink), zeroes it, and records all ranks' addresses in a small device struct that the
    BatchNorm-finalising kernels read.  No NCCL call happens on the data path afterwards."""

    def __init__(self, device, group=None):
        import ctypes
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        lib = _lib.load()
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.BN_SYNC_MAX_WORLD:
            raise ValueError("in-kernel SyncBatchNorm supports up to %d ranks of one NVLink domain" % _lib.BN_SYNC_MAX_WORLD)
        nrecv, nflag = lib.sln_bn_sync_recv_bytes(self.world), lib.sln_bn_sync_flag_bytes()
        self.buf = symm_mem.empty(nrecv + nflag, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.use = torch.zeros(nflag // 4, dtype=torch.int32, device=device)
        t = _lib.BnSync()
        t.world, t.rank = self.world, self.rank
        for r in range(self.world):
            t.recv[r] = ptrs[r]
            t.flag[r] = ptrs[r] + nrecv
        t.use = self.use.data_ptr()
        self.table = torch.frombuffer(bytearray(bytes(t)), dtype=torch.uint8).clone().to(device)
        torch.cuda.synchronize(device)
        dist.barrier(group)               # every rank's buffer is zeroed before any peer can write into it


def enable_sync_batchnorm(model, group=None):
    """Training-mode BatchNorm statistics of `model` (Sg2ScVAEModel) become global over the process group — SURVEY 8e policy P1: the
    N-GPU sharded step equals the 1-GPU step at the global batch (the reference's nn.BatchNorm1d normalises over all rows of the
    batch, models/graph.py:14-15).  Returns the BnSyncTable (keep it alive as long as the model trains)."""
    dev = next(model.parameters()).device
    model._bn_sync = BnSyncTable(dev, group)
    return model._bn_sync


# ---------------------------------------------------------------------------------------------- fused train step
class VAETrainStep(object):
    """The reference train-loop body (train.py:69-84) for a fixed batch shape, replayed as a CUDA graph.

    step(batch) : batch = (objs, triples, boxes, angles, attributes) HOST (pinned) or device tensors of the captured
                  shape -> device tensor losses[4] = {bbox, angle, KL_weight*KLD, total}
    The timed body is: H2D copies into static buffers, encoder fwd, eps ~ N(0,1), reparameterise, decoder fwd, fused
    losses + gradient seeds, decoder bwd, reparam bwd, encoder bwd, [all-reduce], Adam.
    """

    def __init__(self, model, O, T, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, kl_weight=0.1, use_graph=True, process_group=None,
                 world_size=1, sample_eps=True, pack_weights=True, wire_meta=None, bn_policy="local"):
        """bn_policy: 'local' = per-rank BatchNorm statistics (DDP semantics, SURVEY 8e P2: no extra communication);
        'sync' = statistics over all ranks, exchanged inside the finalising kernels over NVLink peer memory (P1: the sharded step
        equals the single-GPU step on the global batch)."""
        self.lib = _lib.load()
        self.model = model
        # overlap_allreduce: two gradient buckets (decoder | encoder), the first in flight during the encoder's backward pass, and the
        # whole step (kernels + NCCL) captured as ONE graph; False = one all-reduce between a forward/backward graph and an Adam graph
        self.overlap_allreduce = True
        self._pending = []
        if bn_policy not in ("local", "sync"):
            raise ValueError("bn_policy must be 'local' or 'sync'")
        self.bn_policy = bn_policy
        if bn_policy == "sync" and world_size > 1 and getattr(model, "_bn_sync", None) is None:
            enable_sync_batchnorm(model, process_group)
        dev = next(model.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError("VAETrainStep needs a CUDA model")
        self.dev, self.O, self.T = dev, O, T
        self.kl_weight, self.lr, self.betas, self.eps = kl_weight, lr, betas, eps
        self.world_size, self.pg = world_size, process_group
        self.sample_eps = sample_eps   # False: the caller writes the N(0,1) draw into self.epsn before run() (parity tests)
        E, BD, NA = model.embedding_dim, model.box_dim, model.Nangle
        model._tables()
        cache = model._cache
        sink = model._grad_sink()
        # flat parameter arena in the gradient arena's order -> one Adam launch, one all-reduce bucket
        n = sink.flat.numel()
        self.p_arena = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for i in sink.order:
                p = cache['params'][i]
                k, off = p.numel(), sink.views[i].storage_offset()     # same (16-byte aligned, zero-padded) slots as the gradient arena
                self.p_arena[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.p_arena[off:off + k].view_as(p)
        model._cache = None
        model._tables()
        model._sink, model._gtable = sink, None
        sink.params = model._cache['params']
        for g in ('enc', 'dec'):
            sink.publish(g)
        self.sink = sink
        self.m = torch.zeros(n, device=dev); self.v = torch.zeros(n, device=dev)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int64)
        i64 = dict(device=dev, dtype=torch.int64)
        f32 = dict(device=dev, dtype=torch.float32)
        self.objs = torch.zeros(O, **i64); self.triples = torch.zeros(T, 3, **i64); self.boxes = torch.zeros(O, BD, **f32)
        self.angles = torch.zeros(O, **i64); self.attrs = torch.zeros(O, **i64)
        # wire_meta = (B, O, T, box_dim, offsets10) of data.collate.packed_batch: the static inputs become views of ONE device buffer in
        # the batch-assembly wire layout (include/sln_b200.h, sln_collate_layout), so that step_wire() feeds a step with one H2D copy +
        # sln_collate_finish instead of suncg_collate_fn + five copies (reference train.py:69, utils.py:114-124)
        self.wire_meta = wire_meta
        if wire_meta is not None:
            wB, wO, wT, wbd, lay = wire_meta
            if (wO, wT, wbd) != (O, T, BD):
                raise ValueError("VAETrainStep: wire layout is for O=%d T=%d box_dim=%d, the step for O=%d T=%d box_dim=%d" % (wO, wT, wbd, O, T, BD))
            self.wire_dev = torch.zeros(lay[9], dtype=torch.uint8, device=dev)

            def dv(off, n, dt):
                return self.wire_dev[off: off + n * dt.itemsize].view(dt)
            self.objs, self.angles, self.attrs = dv(lay[4], O, torch.int64), dv(lay[5], O, torch.int64), dv(lay[6], O, torch.int64)
            self.triples = dv(lay[7], 3 * T, torch.int64).view(T, 3)
            self.boxes = dv(lay[8], O * BD, torch.float32).view(O, BD)
            self.obj_to_img = torch.zeros(O, **i64); self.triple_to_img = torch.zeros(T, **i64)
        self.mu = torch.zeros(O, E, **f32); self.logvar = torch.zeros(O, E, **f32); self.epsn = torch.zeros(O, E, **f32)
        self.z = torch.zeros(O, E, **f32); self.boxes_pred = torch.zeros(O, BD, **f32); self.angles_pred = torch.zeros(O, NA, **f32)
        self.d_boxes = torch.zeros(O, BD, **f32); self.d_logits = torch.zeros(O, NA, **f32)
        self.d_mu = torch.zeros(O, E, **f32); self.d_logvar = torch.zeros(O, E, **f32); self.d_z = torch.zeros(O, E, **f32)
        self.losses = torch.zeros(4, **f32)
        # [lr, kl_weight] on the device: read by k_adam / k_vae_loss at launch time, so set_lr() / set_kl_weight() (the reference's
        # KL_linear_decay schedule, train.py:73-74) take effect on the next replay of the captured graph
        self.hyper = torch.tensor([lr, kl_weight], **f32)
        self.skip_nonfinite = True       # train.py:78-80: a non-finite total loss skips the update (device-side predicate in k_adam)
        self._captured_training = None
        self.loss_scratch = _loss_scratch(dev, O)
        self.ws_enc = torch.empty(self.lib.sln_vae_workspace_bytes(model._desc(), O, T, 0), dtype=torch.uint8, device=dev)
        self.ws_dec = torch.empty(self.lib.sln_vae_workspace_bytes(model._desc(), O, T, 1), dtype=torch.uint8, device=dev)
        # pre-split / pre-tiled weight images, refreshed from the parameter arena at the start of every step (one launch)
        nbytes = self.lib.sln_vae_packed_bytes(model._desc()) if pack_weights else 0
        self.packed = torch.empty(max(nbytes // 4, 4), device=dev, dtype=torch.float32) if nbytes else None
        self.launches_per_step = None
        self.graph_fb = None
        self.graph_opt = None
        self.use_graph = use_graph
        self._static_inputs = (self.objs, self.triples, self.boxes, self.angles, self.attrs)

    # -- the launch sequences -------------------------------------------------------------------
    def _fwd_bwd(self):
        lib, m, O, T = self.lib, self.model, self.O, self.T
        st = _lib.cur_stream(self.dev)
        desc = m._desc()
        params, bufs = m._tables()
        grads = m._grad_table()
        E = m.embedding_dim
        use_kl = not m.use_AE
        self.sink.flat.zero_()
        if self.packed is not None:
            _lib.check(lib.sln_vae_pack_weights(desc, params, self.packed.data_ptr(), self.packed.numel() * 4, st), "vae_pack_weights")
            desc.packed_weights = self.packed.data_ptr()
        _lib.check(lib.sln_vae_encoder_fwd(desc, params, bufs, self.objs.data_ptr(), self.triples.data_ptr(), self.boxes.data_ptr(),
                                           self.angles.data_ptr(), self.attrs.data_ptr(), O, T, self.mu.data_ptr(), self.logvar.data_ptr(),
                                           self.ws_enc.data_ptr(), self.ws_enc.numel(), st), "encoder_fwd")
        if use_kl:
            if self.sample_eps:
                self.epsn.normal_()
            _lib.check(lib.sln_reparam_fwd(self.mu.data_ptr(), self.logvar.data_ptr(), self.epsn.data_ptr(), O * E, self.z.data_ptr(), st), "reparam_fwd")
            z = self.z
        else:
            z = self.mu
        desc.graph_ws = self.ws_enc.data_ptr()     # same objs / triples / attributes: the decoder reuses the encoder's CSR and index arrays
        _lib.check(lib.sln_vae_decoder_fwd(desc, params, bufs, z.data_ptr(), self.objs.data_ptr(), self.triples.data_ptr(),
                                           self.attrs.data_ptr(), O, T, self.boxes_pred.data_ptr(), self.angles_pred.data_ptr(),
                                           self.ws_dec.data_ptr(), self.ws_dec.numel(), st), "decoder_fwd")
        _lib.check(lib.sln_vae_loss_dyn(self.boxes_pred.data_ptr(), self.boxes.data_ptr(), m.box_dim, self.angles_pred.data_ptr(),
                                        self.angles.data_ptr(), m.Nangle, self.mu.data_ptr() if use_kl else None,
                                        self.logvar.data_ptr() if use_kl else None, E if use_kl else 0, self.hyper.data_ptr() + 4, O,
                                        self.losses.data_ptr(), self.d_boxes.data_ptr(), self.d_logits.data_ptr(), 1,
                                        self.d_mu.data_ptr() if use_kl else None, self.d_logvar.data_ptr() if use_kl else None,
                                        self.loss_scratch.data_ptr(), self.loss_scratch.numel() * 4, st), "vae_loss")
        _lib.check(lib.sln_vae_decoder_bwd(desc, params, grads, self.d_boxes.data_ptr(), self.d_logits.data_ptr(), 1, self.d_z.data_ptr(),
                                           O, T, self.ws_dec.data_ptr(), self.ws_dec.numel(), st), "decoder_bwd")
        self._pending = []
        if self.world_size > 1 and self.overlap_allreduce:
            # the decoder's gradients are final half a backward pass before the encoder's: their bucket goes out now, on NCCL's own
            # stream, and crosses NVLink while sln_vae_encoder_bwd runs (a parallel branch of the captured graph)
            import torch.distributed as dist
            lo, hi = self.sink.ranges['dec']
            self._pending.append(dist.all_reduce(self.sink.flat[lo:hi], group=self.pg, async_op=True))
        if use_kl:
            _lib.check(lib.sln_reparam_bwd(self.d_z.data_ptr(), self.logvar.data_ptr(), self.epsn.data_ptr(), O * E, self.d_mu.data_ptr(),
                                           self.d_logvar.data_ptr(), st), "reparam_bwd")
            d_mu, d_lv = self.d_mu, self.d_logvar
        else:
            self.d_logvar.zero_()
            d_mu, d_lv = self.d_z, self.d_logvar
        _lib.check(lib.sln_vae_encoder_bwd(desc, params, grads, self.boxes.data_ptr(), d_mu.data_ptr(), d_lv.data_ptr(), O, T,
                                           self.ws_enc.data_ptr(), self.ws_enc.numel(), st), "encoder_bwd")

    def _opt(self):
        st = _lib.cur_stream(self.dev)
        _lib.check(self.lib.sln_adam_step_dyn(self.p_arena.data_ptr(), self.sink.flat.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                              self.p_arena.numel(), self.hyper.data_ptr(), self.betas[0], self.betas[1], self.eps, 0.0,
                                              1.0 / self.world_size, self.step_count.data_ptr(), 1,
                                              self.losses.data_ptr() + 12 if self.skip_nonfinite else None, st), "adam_step")

    # -- schedules / validation / checkpoints ----------------------------------------------------
    def set_lr(self, lr):
        self.lr = float(lr)
        self.hyper[0:1].fill_(self.lr)

    def set_kl_weight(self, w):
        """reference train.py:73-74 (KL_linear_decay): takes effect on the next step, captured graph or not."""
        self.kl_weight = float(w)
        self.hyper[1:2].fill_(self.kl_weight)

    def check_indices(self):
        """Raise IndexError if the last step saw an out-of-range id (the kernels remap it to row 0 and set a flag; the reference raises
        at the embedding lookup).  One 8-byte D2H read + sync: called automatically after the first step of a captured shape."""
        from .models.Sg2ScVAE_model import _IDX_BITS
        desc = self.model._desc()
        for which, ws, what in ((0, self.ws_enc, "encoder"), (1, self.ws_dec, "decoder")):
            off = self.lib.sln_vae_index_flag_offset(desc, self.O, self.T, which)
            flag = int(ws[off:off + 4].view(torch.int32).item())
            if flag:
                raise IndexError("index out of range in VAETrainStep (%s): %s" % (what, ", ".join(n for b, n in _IDX_BITS if flag & b)))

    def _param_slots(self):
        """[(parameter, arena offset, numel)] in model.parameters() order (= the index order of torch.optim.Adam(model.parameters()))."""
        cache = self.model._cache
        off = {id(cache['params'][i]): self.sink.views[i].storage_offset() for i in self.sink.order}
        return [(p, off[id(p)], p.numel()) for p in self.model.parameters()]

    def optim_state_dict(self):
        """Adam state in torch.optim.Adam(model.parameters()).state_dict() format — what the reference stores as checkpoint
        ['optim_state'] (train.py:95) — so runs can resume in either implementation."""
        step = float(self.step_count.item())
        state = {}
        for i, (p, o, k) in enumerate(self._param_slots()):
            state[i] = {'step': torch.tensor(step), 'exp_avg': self.m[o:o + k].view_as(p).clone(), 'exp_avg_sq': self.v[o:o + k].view_as(p).clone()}
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None, capturable=False,
                     differentiable=False, fused=None, decoupled_weight_decay=False, params=list(range(len(state))))
        return {'state': state, 'param_groups': [group]}

    def load_optim_state_dict(self, sd):
        slots = self._param_slots()
        steps = set()
        with torch.no_grad():
            for i, (p, o, k) in enumerate(slots):
                st = sd['state'].get(i)
                if st is None:
                    self.m[o:o + k].zero_(); self.v[o:o + k].zero_()
                    continue
                self.m[o:o + k].copy_(st['exp_avg'].reshape(-1).to(self.m))
                self.v[o:o + k].copy_(st['exp_avg_sq'].reshape(-1).to(self.v))
                steps.add(int(float(st['step'])))
        if len(steps) > 1:
            raise RuntimeError("VAETrainStep keeps one Adam step counter; the loaded state has steps %s" % sorted(steps))
        self.step_count.fill_(steps.pop() if steps else 0)
        if sd.get('param_groups'):
            self.set_lr(sd['param_groups'][0].get('lr', self.lr))

    def state_dict(self):
        return {'model_state': self.model.state_dict(), 'optim_state': self.optim_state_dict(), 'kl_weight': self.kl_weight}

    def load_state_dict(self, sd):
        """Parameters are copied INTO the step's arena (load_state_dict copies in place), moments and step into m / v / step_count."""
        self.model.load_state_dict(sd['model_state'])
        self.load_optim_state_dict(sd['optim_state'])
        if 'kl_weight' in sd:
            self.set_kl_weight(sd['kl_weight'])

    def _allreduce(self):
        if self.world_size > 1:
            import torch.distributed as dist
            if self.overlap_allreduce:
                lo, hi = self.sink.ranges['enc']
                self._pending.append(dist.all_reduce(self.sink.flat[lo:hi], group=self.pg, async_op=True))
                for w in self._pending:
                    w.wait()              # the current stream waits for NCCL's stream (an event edge, also under capture)
                self._pending = []
            else:
                dist.all_reduce(self.sink.flat, group=self.pg)

    def capture(self):
        """Warm up on a side stream, then capture forward+backward (and Adam) into CUDA graphs."""
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._fwd_bwd(); self._allreduce(); self._opt()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self._captured_training = self.model.training
        self.check_indices_pending = True
        if not self.use_graph:
            return self
        # The chain is captured on a HIGH-priority stream: its kernel nodes inherit the priority, the library's side / leaf streams
        # (weight gradients, parallel branches) have the default one, so the block scheduler places a ready chain kernel before the
        # pending CTAs of a weight-gradient grid instead of behind them.
        # (2.595 -> 2.576 ms at configs[1]; SLN_CHAIN_PRIORITY=0 restores the default capture stream.  Single-GPU steps only: the
        # multi-rank graphs, which also hold NCCL's kernels, stay on the capture stream they were validated with.)
        cap = torch.cuda.Stream(self.dev, priority=-1) if (self.world_size == 1 and os.environ.get("SLN_CHAIN_PRIORITY", "1") != "0") else None
        if self.world_size > 1 and self.overlap_allreduce:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap):
                    self._fwd_bwd(); self._allreduce(); self._opt()
                self.graph_fb, self.graph_opt, self.single_graph = g, None, True
                return self
            except RuntimeError as e:          # NCCL capture unavailable in this build: fall back to graph | all-reduce | graph
                import warnings
                warnings.warn("VAETrainStep: capturing NCCL inside the step graph failed (%s); using the two-graph path" % (str(e)[:200],))
                torch.cuda.synchronize(self.dev)
                self.overlap_allreduce = False
        self.single_graph = self.world_size == 1
        self.graph_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_fb, stream=cap):
            self._fwd_bwd()
            if self.world_size == 1:
                self._opt()
        if self.world_size > 1:
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, stream=cap):
                self._opt()
        return self

    def load_batch(self, batch):
        for dst, src in zip(self._static_inputs, batch):
            dst.copy_(src, non_blocking=True)

    def run(self):
        if self._captured_training is not None and self.model.training != self._captured_training:
            # the train/eval flag (BatchNorm batch vs running statistics; reference train.py:64-66 eval_mode_after) is baked into the
            # captured launches: re-capture instead of silently running the old mode
            if self.graph_fb is not None:
                self.graph_fb = self.graph_opt = None
                self.capture()
            self._captured_training = self.model.training
        out = self._run()
        if getattr(self, "check_indices_pending", False):
            self.check_indices_pending = False
            self.check_indices()
        return out

    def _run(self):
        if self.graph_fb is not None:
            self.graph_fb.replay()
            if self.graph_opt is not None:
                self._allreduce()
                self.graph_opt.replay()
        else:
            self._fwd_bwd(); self._allreduce(); self._opt()
        return self.losses

    def step(self, batch):
        self.load_batch(batch)
        return self.run()

    def step_wire(self, pinned_wire, meta=None):
        """One train step from a host batch in wire layout (data.collate.packed_batch): one async H2D copy, the device half of the
        batch assembly (global triple ids, obj_to_img / triple_to_img), then the step."""
        if self.wire_meta is None:
            raise RuntimeError("VAETrainStep.step_wire needs wire_meta= at construction")
        B, O, T, bd, lay = self.wire_meta
        if meta is not None and (tuple(meta[:4]) != (B, O, T, bd) or tuple(meta[4]) != tuple(lay)):
            raise ValueError("VAETrainStep.step_wire: batch layout %r differs from the captured one %r" % (tuple(meta[:4]), (B, O, T, bd)))
        if pinned_wire.numel() < lay[9]:
            raise ValueError("VAETrainStep.step_wire: wire buffer holds %d bytes, the captured layout needs %d" % (pinned_wire.numel(), lay[9]))
        self.wire_dev.copy_(pinned_wire[:lay[9]], non_blocking=True)
        _lib.check(self.lib.sln_collate_finish(self.wire_dev.data_ptr(), self.wire_dev.numel(), B, O, T, bd, self.triples.data_ptr(),
                                               self.obj_to_img.data_ptr(), self.triple_to_img.data_ptr(), None, _lib.cur_stream(self.dev)),
                   "collate_finish")
        return self.run()
