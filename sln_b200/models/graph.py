"""Drop-in for the reference's ``models/graph.py`` (make_mlp :10-27, GraphTripleConv :36-111, GraphTripleConvNet :114-143).

The modules own ordinary ``nn.Linear`` / ``nn.BatchNorm1d`` parameters with the reference's names, creation order and
initialisation (so ``state_dict`` keys, ``torch.manual_seed`` reproducibility, ``.cuda()``, ``.train()/.eval()`` behave
identically), but ``forward`` runs the fused sm_100a kernels of libsln_b200.so:

    gather(s,o) + concat + Linear + BN statistics  ->  one contraction kernel   (graph.py:78-84)
    BN + ReLU                                       ->  folded into the next kernel's operand load
    scatter_add + count + divide                    ->  one deterministic CSR gather-reduce kernel (graph.py:92-108)

There is no PyTorch fallback: CPU tensors raise.
"""
import torch
import torch.nn as nn

from .. import _lib

_NORMS = {"none": 0, "batch": 1}


def make_mlp(dim_list, activation='relu', batch_norm='none', dropout=0, norelu=False):
    """[Linear, (BatchNorm1d), act, (Dropout)] per consecutive dim pair; ``norelu`` drops the trailing act (and its BN)."""
    mods = []
    for d_in, d_out in zip(dim_list[:-1], dim_list[1:]):
        stage = [nn.Linear(d_in, d_out)]
        if batch_norm == 'batch':
            stage.append(nn.BatchNorm1d(d_out))
        if activation == 'relu':
            stage.append(nn.ReLU())
        elif activation == 'leakyrelu':
            stage.append(nn.LeakyReLU())
        if dropout > 0:
            stage.append(nn.Dropout(p=dropout))
        mods.extend(stage)
    if norelu:
        mods = mods[:-1] if batch_norm == 'none' else mods[:-2]
    return nn.Sequential(*mods)


def _init_weights(module):
    if isinstance(module, nn.Linear) and hasattr(module, 'weight'):
        nn.init.kaiming_normal_(module.weight)


def mlp_blocks(seq):
    """Split an ``nn.Sequential`` built by make_mlp into (linear, bn_or_None, has_relu) blocks."""
    blocks, cur = [], None
    for m in seq:
        if isinstance(m, nn.Linear):
            if cur is not None:
                blocks.append(tuple(cur))
            cur = [m, None, False]
        elif isinstance(m, nn.BatchNorm1d):
            cur[1] = m
        elif isinstance(m, nn.ReLU):
            cur[2] = True
        elif isinstance(m, nn.Dropout) and m.p == 0:
            pass
        else:
            raise NotImplementedError("sln_b200: unsupported MLP stage %r (only Linear/BatchNorm1d/ReLU are fused)" % (m,))
    if cur is not None:
        blocks.append(tuple(cur))
    return blocks


def block_params(blocks, want_bn):
    """Flat [W, b, (gamma, beta)] list and [rm, rv, nbt] list in the canonical table order (include/sln_b200.h)."""
    ps, bufs = [], []
    for lin, bn, _ in blocks:
        ps += [lin.weight, lin.bias]
        if want_bn and bn is not None:
            ps += [bn.weight, bn.bias]
            bufs += [bn.running_mean, bn.running_var, bn.num_batches_tracked]
    return ps, bufs


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sln_b200 runs on CUDA (sm_100a) only; got a %s tensor. There is no CPU fallback." % t.device)


class GradSink:
    """Owns one flat fp32 gradient buffer for a fixed list of parameters and hands the kernels per-parameter pointers.

    Kernels ACCUMULATE into the buffer.  ``prepare()`` makes the buffer hold each parameter's current ``.grad``
    (zeros when it is None) and ``publish()`` points ``.grad`` at the per-parameter views, which gives the usual
    autograd accumulate-into-.grad semantics without one elementwise kernel per parameter.
    """

    def __init__(self, params):
        self.params = list(params)
        dev = self.params[0].device
        n = sum((p.numel() + 3) // 4 * 4 for p in self.params)     # 16-byte aligned slots (vector reductions into dW)
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += (p.numel() + 3) // 4 * 4
        self.key = tuple(p.data_ptr() for p in self.params)

    def matches(self, params):
        return len(params) == len(self.params) and all(a is b for a, b in zip(params, self.params)) and \
            self.flat.device == params[0].device

    def prepare(self):
        if all(p.grad is None for p in self.params):
            self.flat.zero_()
            return
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)

    def publish(self):
        for p, v in zip(self.params, self.views):
            if p.requires_grad:
                if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                    p.grad = v

    def pointers(self):
        return [v if p.requires_grad else None for p, v in zip(self.params, self.views)]


class _GconvLayerFn(torch.autograd.Function):
    """One GraphTripleConv layer through sln_gconv_layer_fwd / sln_gconv_layer_bwd."""

    @staticmethod
    def forward(ctx, layer, anchor, obj_vecs, pred_vecs, edges):
        lib = _lib.load()
        obj_vecs = obj_vecs.contiguous().float()
        pred_vecs = pred_vecs.contiguous().float()
        edges = edges.contiguous()
        O, T = obj_vecs.size(0), pred_vecs.size(0)
        desc = layer._desc()
        params, bufs = layer._tables()
        ws_bytes = lib.sln_vae_workspace_bytes(desc, O, T, 2)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=obj_vecs.device)
        new_obj = torch.empty(O, layer.output_dim, device=obj_vecs.device, dtype=torch.float32)
        new_pred = torch.empty(T, layer.output_dim, device=obj_vecs.device, dtype=torch.float32)
        _lib.check(lib.sln_gconv_layer_fwd(desc, _lib.ptr_array(params), _lib.ptr_array(bufs), obj_vecs.data_ptr(),
                                           pred_vecs.data_ptr(), edges.data_ptr(), O, T, new_obj.data_ptr(),
                                           new_pred.data_ptr(), ws.data_ptr(), ws_bytes, _lib.cur_stream(obj_vecs.device)),
                   "gconv_layer_fwd")
        ctx.layer, ctx.desc, ctx.ws, ctx.dims = layer, desc, ws, (O, T)
        ctx.save_for_backward(obj_vecs, pred_vecs)
        return new_obj, new_pred

    @staticmethod
    def backward(ctx, d_new_obj, d_new_pred):
        lib = _lib.load()
        layer, (O, T) = ctx.layer, ctx.dims
        obj_vecs, pred_vecs = ctx.saved_tensors
        d_new_obj = d_new_obj.contiguous().float()
        d_new_pred = d_new_pred.contiguous().float() if d_new_pred is not None else None
        params, _ = layer._tables()
        sink = layer._grad_sink(params)
        sink.prepare()
        d_obj = torch.empty_like(obj_vecs)
        d_pred = torch.empty_like(pred_vecs)
        _lib.check(lib.sln_gconv_layer_bwd(ctx.desc, _lib.ptr_array(params), _lib.ptr_array(sink.pointers()), obj_vecs.data_ptr(),
                                           pred_vecs.data_ptr(), d_new_obj.data_ptr(), _lib.ptr(d_new_pred), O, T,
                                           d_obj.data_ptr(), d_pred.data_ptr(), ctx.ws.data_ptr(), ctx.ws.numel(),
                                           _lib.cur_stream(obj_vecs.device)), "gconv_layer_bwd")
        sink.publish()
        return None, None, d_obj, d_pred, None


class GraphTripleConv(nn.Module):
    """A single layer of scene graph convolution (same ctor and call signature as the reference)."""

    def __init__(self, input_dim, output_dim=None, hidden_dim=512, pooling='avg', mlp_normalization='none'):
        super(GraphTripleConv, self).__init__()
        output_dim = input_dim if output_dim is None else output_dim
        self.input_dim, self.output_dim, self.hidden_dim = input_dim, output_dim, hidden_dim
        assert pooling in ['avg'], 'Invalid pooling "%s"' % pooling
        self.pooling = pooling
        self.mlp_normalization = mlp_normalization
        self.net1 = make_mlp([3 * input_dim, hidden_dim, 2 * hidden_dim + output_dim], batch_norm=mlp_normalization)
        self.net1.apply(_init_weights)
        self.net2 = make_mlp([hidden_dim, hidden_dim, output_dim], batch_norm=mlp_normalization)
        self.net2.apply(_init_weights)
        self._sink = None
        self._anchor = None

    # ---- binding helpers
    def _blocks(self):
        return mlp_blocks(self.net1) + mlp_blocks(self.net2)

    def _tables(self):
        return block_params(self._blocks(), self.mlp_normalization == 'batch')

    def _grad_sink(self, params):
        if self._sink is None or not self._sink.matches(params):
            self._sink = GradSink(params)
        return self._sink

    def _desc(self):
        if self.mlp_normalization not in _NORMS:
            raise NotImplementedError("mlp_normalization=%r" % (self.mlp_normalization,))
        if self.input_dim != self.output_dim:
            raise NotImplementedError("sln_b200 GraphTripleConv kernels require output_dim == input_dim")
        bn = [b for _, b, _ in self._blocks() if b is not None]
        return _lib.VaeDesc(embedding_dim=4, n_layers=1, recurrent=0, norm=_NORMS[self.mlp_normalization],
                            training=int(self.training), box_dim=6, n_angle=24, num_objs=1, num_preds=1, num_attrs=1,
                            bn_eps=bn[0].eps if bn else 1e-5, bn_momentum=(bn[0].momentum if bn else 0.1) or 0.1,
                            gconv_dim_override=self.input_dim, gconv_hidden_override=self.hidden_dim)

    def forward(self, obj_vecs, pred_vecs, edges):
        """obj_vecs (O, D), pred_vecs (T, D), edges (T, 2) int64 -> new_obj_vecs (O, D), new_pred_vecs (T, D)."""
        require_cuda(obj_vecs, pred_vecs, edges, self.net1[0].weight)
        if self.training and self.mlp_normalization == 'batch' and min(obj_vecs.size(0), pred_vecs.size(0)) < 2:
            raise ValueError("Expected more than 1 value per channel when training, got input size %s" % (list(obj_vecs.size()),))
        if self._anchor is None or self._anchor.device != obj_vecs.device:
            self._anchor = torch.zeros((), device=obj_vecs.device, requires_grad=True)
        return _GconvLayerFn.apply(self, self._anchor, obj_vecs, pred_vecs, edges)


class GraphTripleConvNet(nn.Module):
    """A sequence of scene graph convolution layers (same ctor and call signature as the reference)."""

    def __init__(self, input_dim, num_layers=5, hidden_dim=512, pooling='avg', mode='recurrent', mlp_normalization='none'):
        super(GraphTripleConvNet, self).__init__()
        self.num_layers = num_layers
        self.mode = mode
        self.gconvs = nn.ModuleList()
        kw = dict(input_dim=input_dim, hidden_dim=hidden_dim, pooling=pooling, mlp_normalization=mlp_normalization)
        if mode == 'recurrent':
            self.gconvs.append(GraphTripleConv(**kw))
        elif mode == 'feedforward':
            for _ in range(self.num_layers):
                self.gconvs.append(GraphTripleConv(**kw))
        else:
            raise ValueError('Invalid mode "%s"' % mode)

    def layer(self, i):
        return self.gconvs[0] if self.mode == 'recurrent' else self.gconvs[i]

    def forward(self, obj_vecs, pred_vecs, edges):
        for i in range(self.num_layers):
            obj_vecs, pred_vecs = self.layer(i)(obj_vecs, pred_vecs, edges)
        return obj_vecs, pred_vecs
