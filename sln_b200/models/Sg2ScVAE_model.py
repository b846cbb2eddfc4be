"""Drop-in for the reference's ``models/Sg2ScVAE_model.py`` (Sg2ScVAEModel :6-188).

Same constructor kwargs (build_dataset_model.py:40-52), same sub-module names / creation order / initialisation (so a
reference checkpoint's ``model_state`` loads unchanged and ``torch.manual_seed`` gives identical weights), same
``encoder`` / ``decoder`` / ``forward`` signatures and return values — but encoder and decoder each run as ONE call into
libsln_b200.so (sln_vae_encoder_fwd/bwd, sln_vae_decoder_fwd/bwd): ~65 fused launches instead of ~2500 aten ops.
"""
import torch
import torch.nn as nn

from .. import _lib
from .graph import make_mlp, GraphTripleConvNet, _init_weights, mlp_blocks, block_params, require_cuda, _NORMS


_IDX_BITS = ((1, "objs"), (2, "attributes"), (4, "angles"), (8, "triples subject/object"), (16, "triples predicate"))


def _check_index_flag(model, ws, desc, O, T, which, what):
    """The reference raises IndexError for an out-of-range id (nn.Embedding / obj_vecs[s_idx]); the kernels remap it to row 0 and set a
    bit in the workspace's index flag (include/sln_b200.h SLN_IDX_*).  Reading the flag is a 4-byte D2H copy + sync, so it is done on
    the first call of every (which, O, T) shape, on every call when ``model.check_indices`` is True, never when it is False, and
    never while a CUDA graph is being captured."""
    mode = getattr(model, "check_indices", "first")
    if mode is False or torch.cuda.is_current_stream_capturing():
        return
    key = (which, O, T)
    if mode == "first":
        if key in model._idx_checked:
            return
        model._idx_checked.add(key)
    off = _lib.load().sln_vae_index_flag_offset(desc, O, T, which)
    if off < 0:
        return
    flag = int(ws[off:off + 4].view(torch.int32).item())
    if flag:
        bad = ", ".join(name for bit, name in _IDX_BITS if flag & bit)
        raise IndexError("index out of range in %s: %s" % (what, bad))


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, anchor, objs, triples, boxes_gt, angles_gt, attributes):
        lib = _lib.load()
        dev = objs.device
        objs, triples, angles_gt, attributes = [t.contiguous().long() for t in (objs, triples, angles_gt, attributes)]
        boxes_gt = boxes_gt.contiguous().float()
        O, T = objs.size(0), triples.size(0)
        desc = model._desc()
        params, bufs = model._tables()
        E = model.embedding_dim
        ws_bytes = lib.sln_vae_workspace_bytes(desc, O, T, 0)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        mu = torch.empty(O, E, device=dev, dtype=torch.float32)
        logvar = torch.empty(O, E, device=dev, dtype=torch.float32)
        _lib.check(lib.sln_vae_encoder_fwd(desc, params, bufs, objs.data_ptr(), _lib.ptr(triples) if T else None,
                                           boxes_gt.data_ptr(), angles_gt.data_ptr(), attributes.data_ptr(), O, T,
                                           mu.data_ptr(), logvar.data_ptr(), ws.data_ptr(), ws_bytes, _lib.cur_stream(dev)),
                   "vae_encoder_fwd")
        _check_index_flag(model, ws, desc, O, T, 0, "Sg2ScVAEModel.encoder")
        ctx.model, ctx.desc, ctx.ws, ctx.dims, ctx.boxes = model, desc, ws, (O, T), boxes_gt
        return mu, logvar

    @staticmethod
    def backward(ctx, d_mu, d_logvar):
        lib = _lib.load()
        model, (O, T) = ctx.model, ctx.dims
        dev = ctx.ws.device
        d_mu = torch.zeros(O, model.embedding_dim, device=dev) if d_mu is None else d_mu.contiguous().float()
        d_logvar = torch.zeros(O, model.embedding_dim, device=dev) if d_logvar is None else d_logvar.contiguous().float()
        params, _ = model._tables()
        sink = model._grad_sink()
        sink.prepare('enc')
        _lib.check(lib.sln_vae_encoder_bwd(ctx.desc, params, model._grad_table(), ctx.boxes.data_ptr(), d_mu.data_ptr(),
                                           d_logvar.data_ptr(), O, T, ctx.ws.data_ptr(), ctx.ws.numel(), _lib.cur_stream(dev)),
                   "vae_encoder_bwd")
        sink.publish('enc')
        return (None,) * 7


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, anchor, z, objs, triples, attributes):
        lib = _lib.load()
        dev = z.device
        objs, triples, attributes = [t.contiguous().long() for t in (objs, triples, attributes)]
        z = z.contiguous().float()
        O, T = objs.size(0), triples.size(0)
        desc = model._desc()
        params, bufs = model._tables()
        ws_bytes = lib.sln_vae_workspace_bytes(desc, O, T, 1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        boxes_pred = torch.empty(O, model.box_dim, device=dev, dtype=torch.float32)
        angles_pred = torch.empty(O, model.Nangle, device=dev, dtype=torch.float32)
        _lib.check(lib.sln_vae_decoder_fwd(desc, params, bufs, z.data_ptr(), objs.data_ptr(), _lib.ptr(triples) if T else None,
                                           attributes.data_ptr(), O, T, boxes_pred.data_ptr(), angles_pred.data_ptr(),
                                           ws.data_ptr(), ws_bytes, _lib.cur_stream(dev)), "vae_decoder_fwd")
        _check_index_flag(model, ws, desc, O, T, 1, "Sg2ScVAEModel.decoder")
        ctx.model, ctx.desc, ctx.ws, ctx.dims = model, desc, ws, (O, T)
        ctx.z_needs_grad = z.requires_grad
        return boxes_pred, angles_pred

    @staticmethod
    def backward(ctx, d_boxes, d_angles):
        lib = _lib.load()
        model, (O, T) = ctx.model, ctx.dims
        dev = ctx.ws.device
        d_boxes = torch.zeros(O, model.box_dim, device=dev) if d_boxes is None else d_boxes.contiguous().float()
        d_angles = torch.zeros(O, model.Nangle, device=dev) if d_angles is None else d_angles.contiguous().float()
        params, _ = model._tables()
        sink = model._grad_sink()
        sink.prepare('dec')
        d_z = torch.empty(O, model.embedding_dim, device=dev, dtype=torch.float32)
        _lib.check(lib.sln_vae_decoder_bwd(ctx.desc, params, model._grad_table(), d_boxes.data_ptr(), d_angles.data_ptr(), 0,
                                           d_z.data_ptr(), O, T, ctx.ws.data_ptr(), ctx.ws.numel(), _lib.cur_stream(dev)),
                   "vae_decoder_bwd")
        sink.publish('dec')
        return None, None, d_z, None, None, None


class _ModelGradSink(object):
    """One flat fp32 gradient arena for the whole model, laid out [encoder parameters | decoder parameters].

    The kernels ACCUMULATE into it.  ``prepare(group)`` makes the group's range hold the parameters' current ``.grad``
    (one memset when they are all None, the common case after ``optimizer.zero_grad()``); ``publish(group)`` points each
    ``.grad`` at its view of the arena, which reproduces autograd's accumulate-into-.grad semantics without one
    elementwise kernel per parameter.  The arena doubles as the single NCCL all-reduce bucket for multi-GPU training.
    """

    def __init__(self, params, groups):
        self.params = list(params)
        self.groups = groups
        dev = self.params[0].device
        # every slot starts 16-byte aligned (float4 / red.v4 / tensor-core loader paths need it; an unaligned weight silently
        # falls back to the scalar FP32 kernel): slots are padded to a multiple of 4 floats, the pads stay zero
        n = sum((p.numel() + 3) // 4 * 4 for p in self.params)
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.views = [None] * len(self.params)
        self.ranges, off = {}, 0
        self.order = []
        for name, slots in groups.items():
            lo = off
            for i in slots:
                k = self.params[i].numel()
                self.views[i] = self.flat[off:off + k].view_as(self.params[i])
                self.order.append(i)
                off += (k + 3) // 4 * 4
            self.ranges[name] = (lo, off)
        assert off == n and all(v is not None for v in self.views)

    def matches(self, params):
        return len(params) == len(self.params) and all(a is b for a, b in zip(params, self.params)) and \
            self.flat.device == params[0].device

    def prepare(self, group):
        slots = self.groups[group]
        if all(self.params[i].grad is None for i in slots):
            lo, hi = self.ranges[group]
            self.flat[lo:hi].zero_()
            return
        for i in slots:
            p, v = self.params[i], self.views[i]
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)

    def publish(self, group):
        for i in self.groups[group]:
            p, v = self.params[i], self.views[i]
            if p.requires_grad and (p.grad is None or p.grad.data_ptr() != v.data_ptr()):
                p.grad = v

    def pointers(self):
        return [v if p.requires_grad else None for p, v in zip(self.params, self.views)]


class Sg2ScVAEModel(nn.Module):
    def __init__(self, vocab, embedding_dim=128, batch_size=32,
                 train_3d=True,
                 decoder_cat=False,
                 Nangle=24,
                 gconv_mode='feedforward',
                 gconv_pooling='avg', gconv_num_layers=5,
                 mlp_normalization='none',
                 vec_noise_dim=0,
                 layout_noise_dim=0,
                 use_AE=False,
                 use_attr=True):
        super(Sg2ScVAEModel, self).__init__()
        E = embedding_dim
        hidden = 4 * E
        box_e, ang_e = int(E * 3 / 4), int(E / 4)
        obj_e, attr_e = (int(E * 3 / 4), int(E / 4)) if use_attr else (E, 0)

        self.embedding_dim, self.Nangle = E, Nangle
        self.box_dim = 6 if train_3d else 4
        self.mlp_normalization = mlp_normalization
        self.gconv_mode, self.gconv_num_layers = gconv_mode, gconv_num_layers
        self.use_attr = use_attr
        self.batch_size = batch_size
        self.train_3d = train_3d
        self.decoder_cat = decoder_cat
        self.vocab = vocab
        self.vec_noise_dim = vec_noise_dim
        self.layout_noise_dim = layout_noise_dim
        self.use_AE = use_AE

        n_obj = len(vocab['object_idx_to_name'])
        n_pred = len(vocab['pred_idx_to_name'])
        n_attr = len(vocab['attrib_idx_to_name'])

        # Module creation order == reference order: it fixes both the state_dict key order and the RNG stream of the init.
        self.obj_embeddings_ec = nn.Embedding(n_obj + 1, obj_e)
        self.pred_embeddings_ec = nn.Embedding(n_pred, 2 * E)
        self.obj_embeddings_dc = nn.Embedding(n_obj + 1, obj_e)
        self.pred_embeddings_dc = nn.Embedding(n_pred, E)
        if use_attr:
            self.attr_embedding_ec = nn.Embedding(n_attr, attr_e)
            self.attr_embedding_dc = nn.Embedding(n_attr, attr_e)
        if decoder_cat:
            self.pred_embeddings_dc = nn.Embedding(n_pred, 2 * E)
        self.box_embeddings = nn.Linear(self.box_dim, box_e)
        self.angle_embeddings = nn.Embedding(Nangle, ang_e)
        norm = mlp_normalization
        self.box_mean_var = make_mlp([2 * E, hidden, 2 * E], batch_norm=norm)
        self.box_mean = make_mlp([2 * E, box_e], batch_norm=norm, norelu=True)
        self.box_var = make_mlp([2 * E, box_e], batch_norm=norm, norelu=True)
        self.angle_mean_var = make_mlp([2 * E, hidden, 2 * E], batch_norm=norm)
        self.angle_mean = make_mlp([2 * E, ang_e], batch_norm=norm, norelu=True)
        self.angle_var = make_mlp([2 * E, ang_e], batch_norm=norm, norelu=True)
        self.gconv_net_ec = None
        self.gconv_net_dc = None
        if gconv_num_layers > 0:
            common = dict(hidden_dim=hidden, pooling=gconv_pooling, num_layers=gconv_num_layers, mode=gconv_mode,
                          mlp_normalization=norm)
            self.gconv_net_ec = GraphTripleConvNet(input_dim=2 * E, **common)
            self.gconv_net_dc = GraphTripleConvNet(input_dim=2 * E if decoder_cat else E, **common)
        self.box_net = make_mlp([2 * E + attr_e, hidden, self.box_dim], batch_norm=norm, norelu=True)
        self.angle_net = make_mlp([2 * E, hidden, Nangle], batch_norm=norm, norelu=True)

        for m in (self.box_embeddings, self.box_mean_var, self.box_mean, self.box_var, self.angle_mean_var,
                  self.angle_mean, self.angle_var, self.box_net):
            m.apply(_init_weights)

        self._bn_sync = None              # utils.enable_sync_batchnorm(): cross-rank BatchNorm statistics (SyncBatchNorm)
        self.check_indices = "first"      # "first" | True | False: see _check_index_flag
        self._idx_checked = set()
        self._cache = None   # (params ptr table, bn ptr table, param list, enc/dec slot lists)
        self._sink = None
        self._gtable = None
        self._anchor = None
        self._bn_cfg = None

    # ------------------------------------------------------------------ binding to the C ABI
    def _supported(self):
        if not (self.use_attr and self.decoder_cat and self.gconv_num_layers > 0):
            raise NotImplementedError("sln_b200 implements the released configuration only: use_attr=True, "
                                      "decoder_cat=True, gconv_num_layers>0")
        if self.mlp_normalization not in _NORMS:
            raise NotImplementedError("mlp_normalization=%r" % (self.mlp_normalization,))

    def _param_list(self):
        """Parameters / BN buffers in the canonical order of include/sln_b200.h, plus which slots the encoder and the
        decoder own."""
        self._supported()
        want_bn = self.mlp_normalization == 'batch'
        ps = [self.obj_embeddings_ec.weight, self.attr_embedding_ec.weight, self.angle_embeddings.weight,
              self.pred_embeddings_ec.weight, self.obj_embeddings_dc.weight, self.attr_embedding_dc.weight,
              self.pred_embeddings_dc.weight]
        enc, dec = [0, 1, 2, 3], [4, 5, 6]
        bufs = []

        def add(blocks, owner):
            p, b = block_params(blocks, want_bn)
            owner.extend(range(len(ps), len(ps) + len(p)))
            ps.extend(p)
            bufs.extend(b)

        add([(self.box_embeddings, None, False)], enc)
        for net, owner in ((self.gconv_net_ec, enc), (self.gconv_net_dc, dec)):
            for g in net.gconvs:
                add(mlp_blocks(g.net1) + mlp_blocks(g.net2), owner)
        for seq in (self.box_mean_var, self.angle_mean_var, self.box_mean, self.box_var, self.angle_mean, self.angle_var):
            add(mlp_blocks(seq), enc)
        add(mlp_blocks(self.box_net), dec)
        add(mlp_blocks(self.angle_net), dec)
        return ps, bufs, enc, dec

    def _build_cache(self):
        ps, bufs, enc, dec = self._param_list()
        require_cuda(*ps)
        for t in ps + bufs:
            if not t.is_contiguous():
                raise RuntimeError("sln_b200 needs contiguous parameters")
        if any(p.dtype != torch.float32 for p in ps):
            raise RuntimeError("sln_b200 computes in fp32; call model.float()")
        desc = self._desc()
        lib = _lib.load()
        assert lib.sln_vae_num_params(desc) == len(ps), (lib.sln_vae_num_params(desc), len(ps))
        assert lib.sln_vae_num_bn(desc) * 3 == len(bufs)
        self._cache = dict(params=ps, bufs=bufs, enc=enc, dec=dec, ptable=_lib.ptr_array(ps), btable=_lib.ptr_array(bufs),
                           key=tuple(p.data_ptr() for p in ps), bkey=tuple(b.data_ptr() for b in bufs))
        return self._cache

    def _tables(self):
        c = self._cache
        if c is None or tuple(p.data_ptr() for p in c['params']) != c['key'] or tuple(b.data_ptr() for b in c['bufs']) != c['bkey']:
            c = self._build_cache()      # any re-homed parameter / buffer (load_state_dict(assign=True), p.data = ...) drops the tables
        return c['ptable'], c['btable']

    def _grad_sink(self):
        ps = self._cache['params']
        if self._sink is None or not self._sink.matches(ps):
            self._sink = _ModelGradSink(ps, {'enc': self._cache['enc'], 'dec': self._cache['dec']})
            self._gtable = None
        return self._sink

    def _grad_table(self):
        if self._gtable is None:
            self._gtable = _lib.ptr_array(self._sink.pointers())
            self._gtable_req = tuple(p.requires_grad for p in self._sink.params)
        elif self._gtable_req != tuple(p.requires_grad for p in self._sink.params):
            self._gtable = None
            return self._grad_table()
        return self._gtable

    def _apply(self, fn, *a, **kw):   # .cuda() / .float() / .to(): parameter storage moves -> drop pointer caches
        self._cache, self._sink, self._anchor = None, None, None
        return super(Sg2ScVAEModel, self)._apply(fn, *a, **kw)

    def load_state_dict(self, state_dict, *a, **kw):   # assign=True swaps Parameter objects: drop every pointer cache
        out = super(Sg2ScVAEModel, self).load_state_dict(state_dict, *a, **kw)
        if kw.get("assign", False):
            self._cache, self._sink, self._gtable = None, None, None
        return out

    def _desc(self):
        if self._bn_cfg is None:
            bn = [m for m in self.modules() if isinstance(m, nn.BatchNorm1d)]
            self._bn_cfg = (bn[0].eps, bn[0].momentum or 0.1) if bn else (1e-5, 0.1)
        return _lib.VaeDesc(embedding_dim=self.embedding_dim, n_layers=self.gconv_num_layers,
                            recurrent=int(self.gconv_mode == 'recurrent'), norm=_NORMS[self.mlp_normalization],
                            training=int(self.training), box_dim=self.box_dim, n_angle=self.Nangle,
                            num_objs=self.obj_embeddings_ec.num_embeddings, num_preds=self.pred_embeddings_ec.num_embeddings,
                            num_attrs=self.attr_embedding_ec.num_embeddings,
                            bn_eps=self._bn_cfg[0], bn_momentum=self._bn_cfg[1],
                            gconv_dim_override=0, gconv_hidden_override=0,
                            bn_sync=self._bn_sync.table.data_ptr() if getattr(self, "_bn_sync", None) is not None else None)

    def _get_anchor(self, dev):
        if self._anchor is None or self._anchor.device != dev:
            self._anchor = torch.zeros((), device=dev, requires_grad=True)
        return self._anchor

    def _check_bn_rows(self, O, T):
        if self.training and self.mlp_normalization == 'batch' and (O < 2 or T < 2):
            raise ValueError("Expected more than 1 value per channel when training, got O=%d, T=%d" % (O, T))

    # ------------------------------------------------------------------ reference surface
    def encoder(self, objs, triples, boxes_gt, angles_gt, attributes):
        require_cuda(objs, triples, boxes_gt, angles_gt, attributes)
        self._tables()
        self._check_bn_rows(objs.size(0), triples.size(0))
        return _EncoderFn.apply(self, self._get_anchor(objs.device), objs, triples, boxes_gt, angles_gt, attributes)

    def decoder(self, z, objs, triples, attributes):
        require_cuda(z, objs, triples, attributes)
        self._tables()
        self._check_bn_rows(objs.size(0), triples.size(0))
        return _DecoderFn.apply(self, self._get_anchor(z.device), z, objs, triples, attributes)

    def forward(self, objs, triples, boxes_gt, angles_gt, attributes, obj_to_img):
        mu, logvar = self.encoder(objs, triples, boxes_gt, angles_gt, attributes)
        if self.use_AE:
            z = mu
        else:
            std = torch.exp(0.5 * logvar)
            eps = torch.randn_like(std)      # torch's Philox stream, as in the reference
            z = eps.mul(std).add_(mu)
        boxes_pred, angles_pred = self.decoder(z, objs, triples, attributes)
        return mu, logvar, boxes_pred, angles_pred
