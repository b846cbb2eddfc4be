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
