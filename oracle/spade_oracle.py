"""TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of the reference's SPADEGenerator4 forward.

Nothing under ``sln_b200/`` imports this file.  Functional form over a ``state_dict`` (keys as in the reference:
``head_0.conv_0.1.weight_orig`` ...), plain torch CPU ops, any dtype (fp64 = truth, fp32 = the reference's noise floor).
Follows models/SPADE_related.py: LayerNorm2D :139-149, SEBlock2 :81-85, SPADE4 :1438-1454, SPADEResnetBlock4 :1487-1505,
SPADEGenerator4.forward :1563-1605 (eval mode: spectral norm = W / (u^T W v), no power iteration).
Pinned against the reference itself by tests/golden/spade_small.npz (oracle/gen_golden_spade.py).
"""
import torch
import torch.nn.functional as F

BLOCKS = ("head_0", "G_middle_0", "G_middle_1", "up_0", "up_1", "up_2", "up_3")


def _w(sd, key, dt):
    return sd[key].detach().to(dt)


def _sn(sd, prefix, dt):
    w, u, v = _w(sd, prefix + ".weight_orig", dt), _w(sd, prefix + ".weight_u", dt), _w(sd, prefix + ".weight_v", dt)
    sigma = torch.dot(u, torch.mv(w.reshape(w.size(0), -1), v))
    return w / sigma


def layernorm2d(x, eps=1e-5):
    flat = x.reshape(x.size(0), -1)
    mean = flat.mean(1).view(-1, 1, 1, 1)
    std = flat.std(1).view(-1, 1, 1, 1)          # unbiased, as torch.std
    return (x - mean) / (std + eps)


def rpad_conv(x, w, b):
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), w, b)


def spade4(sd, p, x, segmap, dt):
    normalized = layernorm2d(x)
    segmap = F.interpolate(segmap, size=x.shape[2:], mode='bilinear', align_corners=False)
    depth = F.leaky_relu(rpad_conv(segmap[:, 0:1], _w(sd, p + ".mlp_preshared_depth.1.weight", dt), _w(sd, p + ".mlp_preshared_depth.1.bias", dt)), 0.01)
    actv = F.relu(rpad_conv(torch.cat((depth, segmap[:, 1:]), 1), _w(sd, p + ".mlp_shared.1.weight", dt), _w(sd, p + ".mlp_shared.1.bias", dt)))
    gamma = rpad_conv(actv, _w(sd, p + ".mlp_gamma.1.weight", dt), _w(sd, p + ".mlp_gamma.1.bias", dt))
    beta = rpad_conv(actv, _w(sd, p + ".mlp_beta.1.weight", dt), _w(sd, p + ".mlp_beta.1.bias", dt))
    return normalized * (1 + gamma) + beta


def resblock4(sd, p, x, seg, dt):
    learned = (p + ".conv_s.weight_orig") in sd
    x_s = F.conv2d(spade4(sd, p + ".norm_s", x, seg, dt), _sn(sd, p + ".conv_s", dt)) if learned else x
    dx = rpad_conv(F.leaky_relu(spade4(sd, p + ".norm_0", x, seg, dt), 0.2), _sn(sd, p + ".conv_0.1", dt), _w(sd, p + ".conv_0.1.bias", dt))
    dx = rpad_conv(F.leaky_relu(spade4(sd, p + ".norm_1", dx, seg, dt), 0.2), _sn(sd, p + ".conv_1.1", dt), _w(sd, p + ".conv_1.1.bias", dt))
    y = dx.mean(dim=(2, 3))
    y = torch.sigmoid(F.linear(F.relu(F.linear(y, _w(sd, p + ".se.fc.0.weight", dt))), _w(sd, p + ".se.fc.2.weight", dt)))
    return x_s + dx * y[:, :, None, None]


def forward(sd, seg, z, nf, sh, dtype=torch.float64, taps=None):
    """-> tanh image [B, 3, S, S]; taps (dict) receives every block output and the pre-tanh conv_img output."""
    dt = dtype
    seg, z = seg.to(dt), z.to(dt)
    x = F.linear(z, _w(sd, "fc.weight", dt), _w(sd, "fc.bias", dt)).view(-1, 16 * nf, sh, sh)
    seg_1 = F.interpolate(seg, size=[sh, sh])           # nearest (:1579)
    x = resblock4(sd, "head_0", x, seg_1, dt)
    if taps is not None: taps["head_0"] = x
    x = F.interpolate(x, scale_factor=2, mode='nearest')
    for name in ("G_middle_0", "G_middle_1"):
        x = resblock4(sd, name, x, seg, dt)
        if taps is not None: taps[name] = x
    for name, mode in (("up_0", 'nearest'), ("up_1", 'nearest'), ("up_2", 'nearest'), ("up_3", 'bilinear')):
        x = F.interpolate(x, scale_factor=2, mode=mode) if mode == 'nearest' else F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
        x = resblock4(sd, name, x, seg, dt)
        if taps is not None: taps[name] = x
    pre = F.conv2d(F.leaky_relu(x, 0.2), _w(sd, "conv_img.weight", dt), _w(sd, "conv_img.bias", dt), padding=2)
    if taps is not None: taps["pre_tanh"] = pre
    return torch.tanh(pre)


def synthetic_input(B, S=256, n_labels=40, seed=0):
    """channel 0 = smooth depth in [-1, 1], channels 1..40 = one-hot of a random Voronoi segmentation."""
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(B, 1, 8, 8, generator=g)
    depth = torch.tanh(F.interpolate(low, size=(S, S), mode='bicubic', align_corners=False))
    sites = torch.rand(B, 24, 2, generator=g) * S
    labels = torch.randint(0, n_labels, (B, 24), generator=g)
    yy, xx = torch.meshgrid(torch.arange(S, dtype=torch.float32), torch.arange(S, dtype=torch.float32), indexing="ij")
    d2 = (yy[None, None] - sites[:, :, 0, None, None]) ** 2 + (xx[None, None] - sites[:, :, 1, None, None]) ** 2
    owner = d2.argmin(1)
    lab = torch.gather(labels, 1, owner.view(B, -1)).view(B, S, S)
    onehot = F.one_hot(lab, n_labels).permute(0, 3, 1, 2).float()
    return torch.cat([depth, onehot], 1).contiguous()


# ---------------------------------------------------------------------------------------------------------------------------------
# Plain SPADEGenerator (reference models/SPADE_related.py:151-346; Conv2dBlock :16-68, SEResBlock2 :87-101), eval mode.
PLAIN_BLOCKS = ("head_0", "G_middle_0", "G_middle_1", "up_0", "up_1", "up_2", "up_3")


def _cw(sd, prefix, dt):
    """conv weight: spectral-normalised (eval: W / u^T W v) when the checkpoint holds weight_orig, else plain."""
    return _sn(sd, prefix, dt) if (prefix + ".weight_orig") in sd else _w(sd, prefix + ".weight", dt)


def spade_plain(sd, p, x, segmap, dt, kind):
    if kind == "instance":
        normalized = F.instance_norm(x, eps=1e-5)
    else:      # eval-mode BatchNorm2d(affine=False)
        normalized = F.batch_norm(x, _w(sd, p + ".param_free_norm.running_mean", dt), _w(sd, p + ".param_free_norm.running_var", dt), training=False, eps=1e-5)
    segmap = F.interpolate(segmap, size=x.shape[2:], mode='bilinear', align_corners=False)
    actv = F.relu(F.conv2d(segmap, _w(sd, p + ".mlp_shared.0.weight", dt), _w(sd, p + ".mlp_shared.0.bias", dt), padding=1))
    gamma = F.conv2d(actv, _w(sd, p + ".mlp_gamma.weight", dt), _w(sd, p + ".mlp_gamma.bias", dt), padding=1)
    beta = F.conv2d(actv, _w(sd, p + ".mlp_beta.weight", dt), _w(sd, p + ".mlp_beta.bias", dt), padding=1)
    return normalized * (1 + gamma) + beta


def resblock_plain(sd, p, x, seg, dt, kind):
    learned = any(k.startswith(p + ".conv_s.") for k in sd)
    x_s = F.conv2d(spade_plain(sd, p + ".norm_s", x, seg, dt, kind), _cw(sd, p + ".conv_s", dt)) if learned else x
    dx = F.conv2d(F.leaky_relu(spade_plain(sd, p + ".norm_0", x, seg, dt, kind), 0.2), _cw(sd, p + ".conv_0", dt), _w(sd, p + ".conv_0.bias", dt), padding=1)
    dx = F.conv2d(F.leaky_relu(spade_plain(sd, p + ".norm_1", dx, seg, dt, kind), 0.2), _cw(sd, p + ".conv_1", dt), _w(sd, p + ".conv_1.bias", dt), padding=1)
    return x_s + dx


def forward_plain(sd, seg, z, nf, sh, kind="instance", dtype=torch.float64, taps=None):
    """SPADEGenerator.forward (:209-250) with n_up='normal' -> tanh image [B, 3, S, S]."""
    dt = dtype
    seg, z = seg.to(dt), z.to(dt)
    x = F.linear(z, _w(sd, "fc.weight", dt), _w(sd, "fc.bias", dt)).view(-1, 16 * nf, sh, sh)
    x = resblock_plain(sd, "head_0", x, F.interpolate(seg, size=[sh, sh]), dt, kind)
    if taps is not None: taps["head_0"] = x
    for name, up in (("G_middle_0", True), ("G_middle_1", False), ("up_0", True), ("up_1", True), ("up_2", True), ("up_3", True)):
        if up:
            x = F.interpolate(x, scale_factor=2, mode='nearest')
        x = resblock_plain(sd, name, x, seg, dt, kind)
        if taps is not None: taps[name] = x
    y = x                                                                  # conv_img_pre = SEResBlock2
    for i, act in ((0, True), (1, False)):
        y = F.conv2d(F.pad(y, (1, 1, 1, 1), mode='reflect'), _w(sd, "conv_img_pre.model.%d.conv.weight" % i, dt), _w(sd, "conv_img_pre.model.%d.conv.bias" % i, dt))
        y = F.instance_norm(y, eps=1e-5)
        if act:
            y = F.relu(y)
    s = y.mean(dim=(2, 3))
    s = torch.sigmoid(F.linear(F.relu(F.linear(s, _w(sd, "conv_img_pre.model.2.fc.0.weight", dt))), _w(sd, "conv_img_pre.model.2.fc.2.weight", dt)))
    y = y * s[:, :, None, None] + x
    if taps is not None: taps["conv_img_pre"] = y
    pre = F.conv2d(F.leaky_relu(y, 0.2), _w(sd, "conv_img.weight", dt), _w(sd, "conv_img.bias", dt), padding=2)
    if taps is not None: taps["pre_tanh"] = pre
    return torch.tanh(pre)
