"""TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of the reference's VAE-graph hot path.

Nothing under ``sln_b200/`` imports this file.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker / CPU baseline — never as the product.

It restates, as plain functional torch code over a ``state_dict``-style mapping (reference key names), what these
reference functions compute:

    make_mlp / BatchNorm1d / ReLU stacks      models/graph.py:10-27
    GraphTripleConv.forward                   models/graph.py:57-111
    GraphTripleConvNet.forward                models/graph.py:136-143
    Sg2ScVAEModel.encoder / decoder / forward models/Sg2ScVAE_model.py:115-188
    calculate_model_losses                    utils.py:12-33
    Adam step                                 train.py:15,82-84 (torch.optim.Adam defaults)

Pinning: ``tests/test_oracle_vae.py`` checks it against (a) golden vectors generated from the real reference modules by
``oracle/gen_golden.py`` (committed under tests/golden/), (b) the known-answer sums of SURVEY.md App. D and (c) — when
``/root/reference`` is present — the live reference modules.  dtype follows the tensors in ``sd`` (fp32 or fp64).
"""
import math

import torch

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _seq_layout(sd, prefix):
    """[(linear_idx, bn_idx or None)] of an nn.Sequential built by make_mlp (graph.py:10-27), read off the key names."""
    idx = sorted({int(k[len(prefix) + 1:].split('.')[0]) for k in sd if k.startswith(prefix + '.') and k.endswith('.weight')})
    out = []
    for i in idx:
        if sd['%s.%d.weight' % (prefix, i)].dim() == 2:
            out.append([i, None])
        else:
            out[-1][1] = i
    return out


def batch_norm(x, sd, key, training, stats_out=None):
    """nn.BatchNorm1d over rows (graph.py:14-15): batch statistics (biased var) in training, running stats in eval."""
    w, b = sd[key + '.weight'], sd[key + '.bias']
    if training:
        mean = x.mean(dim=0)
        var = ((x - mean) ** 2).mean(dim=0)
        if stats_out is not None:
            n = x.shape[0]
            # chained through stats_out so that shared ('recurrent') layers update their statistics once per use
            cur_rm = stats_out.get(key + '.running_mean', sd[key + '.running_mean'])
            cur_rv = stats_out.get(key + '.running_var', sd[key + '.running_var'])
            cur_n = stats_out.get(key + '.num_batches_tracked', sd[key + '.num_batches_tracked'])
            stats_out[key + '.running_mean'] = (1 - BN_MOMENTUM) * cur_rm + BN_MOMENTUM * mean.detach().to(cur_rm.dtype)
            stats_out[key + '.running_var'] = (1 - BN_MOMENTUM) * cur_rv + BN_MOMENTUM * (var.detach() * n / max(n - 1, 1)).to(cur_rv.dtype)
            stats_out[key + '.num_batches_tracked'] = cur_n + 1
    else:
        mean, var = sd[key + '.running_mean'].to(x.dtype), sd[key + '.running_var'].to(x.dtype)
    return (x - mean) / torch.sqrt(var + BN_EPS) * w + b


def mlp(sd, prefix, x, training, norelu=False, stats_out=None):
    """make_mlp stack: Linear -> (BatchNorm1d) -> ReLU per stage; ``norelu`` stages have neither BN nor ReLU (graph.py:22-26)."""
    for lin, bn in _seq_layout(sd, prefix):
        x = x @ sd['%s.%d.weight' % (prefix, lin)].t() + sd['%s.%d.bias' % (prefix, lin)]
        if norelu:
            continue
        if bn is not None:
            x = batch_norm(x, sd, '%s.%d' % (prefix, bn), training, stats_out)
        x = torch.clamp(x, min=0)
    return x


def gconv_layer(sd, prefix, obj_vecs, pred_vecs, edges, training, stats_out=None):
    """GraphTripleConv.forward (graph.py:57-111)."""
    O, H = obj_vecs.shape[0], sd[prefix + '.net1.0.weight'].shape[0]
    last = _seq_layout(sd, prefix + '.net2')[-1][0]
    Dout = sd['%s.net2.%d.weight' % (prefix, last)].shape[0]
    s_idx, o_idx = edges[:, 0], edges[:, 1]
    t_in = torch.cat([obj_vecs.index_select(0, s_idx), pred_vecs, obj_vecs.index_select(0, o_idx)], dim=1)   # :78-83
    t_out = mlp(sd, prefix + '.net1', t_in, training, stats_out=stats_out)                                  # :84
    new_s, new_p, new_o = t_out[:, :H], t_out[:, H:H + Dout], t_out[:, H + Dout:2 * H + Dout]                # :88-90
    pooled = torch.zeros(O, H, dtype=obj_vecs.dtype, device=obj_vecs.device)
    pooled = pooled.index_add(0, s_idx, new_s).index_add(0, o_idx, new_o)                                   # :93-100
    counts = torch.zeros(O, dtype=obj_vecs.dtype, device=obj_vecs.device).index_add(0, s_idx, torch.ones_like(s_idx, dtype=obj_vecs.dtype))
    counts = counts.index_add(0, o_idx, torch.ones_like(o_idx, dtype=obj_vecs.dtype)).clamp(min=1)          # :102-107
    pooled = pooled / counts[:, None]                                                                       # :108
    return mlp(sd, prefix + '.net2', pooled, training, stats_out=stats_out), new_p                           # :109


def gconv_net(sd, prefix, obj_vecs, pred_vecs, edges, num_layers, training, stats_out=None):
    """GraphTripleConvNet.forward (graph.py:136-143); 'recurrent' mode has a single gconvs.0 reused by every layer."""
    n_w = len({k.split('.')[len(prefix.split('.')) + 1] for k in sd if k.startswith(prefix + '.gconvs.')})
    for i in range(num_layers):
        obj_vecs, pred_vecs = gconv_layer(sd, '%s.gconvs.%d' % (prefix, i if n_w > 1 else 0), obj_vecs, pred_vecs, edges,
                                          training, stats_out)
    return obj_vecs, pred_vecs


def encoder(sd, objs, triples, boxes_gt, angles_gt, attributes, num_layers=5, training=True, stats_out=None):
    """Sg2ScVAEModel.encoder (Sg2ScVAE_model.py:115-143)."""
    edges = torch.stack([triples[:, 0], triples[:, 2]], dim=1)
    obj_vecs = torch.cat([sd['obj_embeddings_ec.weight'][objs], sd['attr_embedding_ec.weight'][attributes]], dim=1)
    boxes_vecs = boxes_gt.to(sd['box_embeddings.weight'].dtype) @ sd['box_embeddings.weight'].t() + sd['box_embeddings.bias']
    obj_vecs = torch.cat([obj_vecs, boxes_vecs, sd['angle_embeddings.weight'][angles_gt]], dim=1)
    pred_vecs = sd['pred_embeddings_ec.weight'][triples[:, 1]]
    obj_vecs, _ = gconv_net(sd, 'gconv_net_ec', obj_vecs, pred_vecs, edges, num_layers, training, stats_out)
    hb = mlp(sd, 'box_mean_var', obj_vecs, training, stats_out=stats_out)
    ha = mlp(sd, 'angle_mean_var', obj_vecs, training, stats_out=stats_out)
    mu = torch.cat([mlp(sd, 'box_mean', hb, training, norelu=True), mlp(sd, 'angle_mean', ha, training, norelu=True)], dim=1)
    logvar = torch.cat([mlp(sd, 'box_var', hb, training, norelu=True), mlp(sd, 'angle_var', ha, training, norelu=True)], dim=1)
    return mu, logvar


def _head(sd, prefix, x, training, stats_out):
    """make_mlp([in, hidden, out], norelu=True): Linear-(BN)-ReLU-Linear."""
    stages = _seq_layout(sd, prefix)
    (l0, bn0), (l1, _) = stages
    x = x @ sd['%s.%d.weight' % (prefix, l0)].t() + sd['%s.%d.bias' % (prefix, l0)]
    if bn0 is not None:
        x = batch_norm(x, sd, '%s.%d' % (prefix, bn0), training, stats_out)
    x = torch.clamp(x, min=0)
    return x @ sd['%s.%d.weight' % (prefix, l1)].t() + sd['%s.%d.bias' % (prefix, l1)]


def decoder(sd, z, objs, triples, attributes, num_layers=5, training=True, stats_out=None):
    """Sg2ScVAEModel.decoder with decoder_cat=True, use_attr=True (Sg2ScVAE_model.py:145-172)."""
    edges = torch.stack([triples[:, 0], triples[:, 2]], dim=1)
    attr_vecs = sd['attr_embedding_dc.weight'][attributes]
    obj_vecs = torch.cat([sd['obj_embeddings_dc.weight'][objs], attr_vecs, z], dim=1)
    pred_vecs = sd['pred_embeddings_dc.weight'][triples[:, 1]]
    obj_vecs, _ = gconv_net(sd, 'gconv_net_dc', obj_vecs, pred_vecs, edges, num_layers, training, stats_out)
    boxes_pred = _head(sd, 'box_net', torch.cat([obj_vecs, attr_vecs], dim=1), training, stats_out)
    angles_pred = torch.log_softmax(_head(sd, 'angle_net', obj_vecs, training, stats_out), dim=1)
    return boxes_pred, angles_pred


def forward(sd, objs, triples, boxes_gt, angles_gt, attributes, eps=None, num_layers=5, training=True, use_AE=False, stats_out=None):
    """Sg2ScVAEModel.forward (Sg2ScVAE_model.py:174-188) with the N(0,1) sample ``eps`` injected."""
    mu, logvar = encoder(sd, objs, triples, boxes_gt, angles_gt, attributes, num_layers, training, stats_out)
    z = mu if use_AE else eps.to(mu.dtype) * torch.exp(0.5 * logvar) + mu
    boxes_pred, angles_pred = decoder(sd, z, objs, triples, attributes, num_layers, training, stats_out)
    return mu, logvar, boxes_pred, angles_pred


def losses(boxes_gt, boxes_pred, angles_gt, angles_pred, mu=None, logvar=None, kl_weight=0.1, use_AE=False):
    """calculate_model_losses (utils.py:12-33): returns (total, {bbox_pred, angle_pred, KLD_Gauss})."""
    l_box = (boxes_pred - boxes_gt.to(boxes_pred.dtype)).abs().mean()
    l_ang = -angles_pred.gather(1, angles_gt[:, None]).mean()
    total = l_box + l_ang
    out = {'bbox_pred': l_box, 'angle_pred': l_ang}
    if not use_AE:
        l_kl = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()) / mu.shape[0]
        out['KLD_Gauss'] = l_kl * kl_weight
        total = total + l_kl * kl_weight
    return total, out


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
    """One torch.optim.Adam update (defaults of train.py:15) on lists of tensors, in place; ``step`` is 1-based."""
    b1, b2 = betas
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        p.addcdiv_(m, v.sqrt() / math.sqrt(bc2) + eps, value=-lr / bc1)


def leaf_state(state_dict, dtype=torch.float64, requires_grad=True, device="cpu"):
    """Detached copy of a model state_dict as oracle input: float tensors cast to ``dtype`` (parameters become leaves).
    device: "cpu" for the oracle proper; bench.py also runs the same plain-torch restatement on the GPU as the "PyTorch eager on
    the same B200" baseline (what the reference's train.py would execute after model.cuda())."""
    sd = {}
    for k, v in state_dict.items():
        v = v.detach().to(device)
        if v.is_floating_point():
            v = v.to(dtype).clone()
            if requires_grad and not (k.endswith('running_mean') or k.endswith('running_var')):
                v.requires_grad_(True)
        else:
            v = v.clone()
        sd[k] = v
    return sd


def train_step(sd, batch, eps, opt_state, step, num_layers=5, kl_weight=0.1, lr=1e-4):
    """train.py:70-84 body on the oracle: forward, losses, backward, Adam, BN running-stat update.  Returns loss dict."""
    objs, triples, boxes, angles, attrs = batch
    for v in sd.values():
        if v.is_floating_point() and v.requires_grad:
            v.grad = None
    stats = {}
    mu, logvar, bp, ap = forward(sd, objs, triples, boxes, angles, attrs, eps, num_layers, True, False, stats)
    total, parts = losses(boxes, bp, angles, ap, mu, logvar, kl_weight)
    total.backward()
    keys = [k for k, v in sd.items() if v.is_floating_point() and v.requires_grad]
    with torch.no_grad():
        ps = [sd[k] for k in keys]
        gs = [sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k]) for k in keys]
        if not opt_state:
            opt_state['m'] = [torch.zeros_like(p) for p in ps]
            opt_state['v'] = [torch.zeros_like(p) for p in ps]
        adam_step(ps, gs, opt_state['m'], opt_state['v'], step, lr)
        for k, v in stats.items():
            sd[k] = v
    return float(total.detach()), {k: float(v.detach()) for k, v in parts.items()}
