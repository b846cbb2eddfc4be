"""Drop-in for the live part of the reference's ``models/SPADE_related.py``: ``SEBlock2`` (:70-85), ``LayerNorm2D``
(:128-149), ``SPADE4`` (:1404-1454), ``SPADEResnetBlock4`` (:1457-1505) and ``SPADEGenerator4`` (:1507-1605) — the generator
that ``testing/test_SPADE_shade.py:9`` instantiates as ``SPADEGenerator4(semantic_nc=41, target_nc=3, nz=256, ngf=64,
norm='spectralspadelayer3x3', crop_size=256, n_up='normal')`` and calls as ``model(total, color_z)`` in eval mode.

The module tree (names, creation order, spectral-norm wrappers) is the reference's, so ``state_dict`` keys
(``head_0.conv_0.1.weight_orig``, ``...weight_u``, ``fc.weight`` ...) and seeded initial weights are identical and a
``latest_net_G_AB.pth`` checkpoint loads unchanged.  ``forward`` does not run these torch modules: it hands packed
weights (spectral norm folded: sigma = u^T W v in eval mode, no power iteration; convolution kernels re-ordered to
[Cout][ky][kx][Cin] for the NHWC implicit GEMM) to libsln_b200.so (csrc/spade.cu): every convolution is a tcgen05
3xTF32 implicit GEMM, the gamma/beta convolutions of each SPADE4 are ONE contraction whose epilogue applies
``lrelu(x_hat * (1 + gamma) + beta)`` so gamma and beta never reach HBM.  Inference (eval) only, CUDA only.
"""
import ctypes
import re

import torch
import torch.nn as nn
import torch.nn.utils.spectral_norm as spectral_norm

from .. import _lib

NHIDDEN = 128   # SPADE4's hard-coded embedding width (:1427)


class SEBlock2(nn.Module):
    def __init__(self, channel, reduction=4):
        super(SEBlock2, self).__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())


class LayerNorm2D(nn.Module):
    def __init__(self, num_features, eps=1e-5, affine=True):
        super(LayerNorm2D, self).__init__()
        self.num_features, self.affine, self.eps = num_features, affine, eps
        if self.affine:
            self.gamma = nn.Parameter(torch.Tensor(num_features).uniform_())
            self.beta = nn.Parameter(torch.zeros(num_features))


class SPADE4(nn.Module):
    def __init__(self, config_text, norm_nc, label_nc):
        super().__init__()
        parsed = re.search(r'spade(\D+)(\d)x\d', config_text)
        if not config_text.startswith('spade') or parsed is None:
            raise ValueError("not a SPADE norm specification: %r" % (config_text,))
        if str(parsed.group(1)) != 'layer' or int(parsed.group(2)) != 3:
            raise NotImplementedError("sln_b200 implements the configuration the reference runs: 'spadelayer3x3' (got %r)" % (config_text,))
        self.param_free_norm = LayerNorm2D(norm_nc, affine=False)
        self.norm_nc = norm_nc
        self.mlp_preshared_depth = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(1, NHIDDEN // 8, kernel_size=3, padding=0), nn.LeakyReLU(inplace=True))
        self.mlp_shared = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(NHIDDEN // 8 + label_nc - 1, NHIDDEN, kernel_size=3, padding=0), nn.ReLU(inplace=True))
        self.mlp_gamma = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(NHIDDEN, norm_nc, kernel_size=3, padding=0))
        self.mlp_beta = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(NHIDDEN, norm_nc, kernel_size=3, padding=0))


class SPADEResnetBlock4(nn.Module):
    def __init__(self, fin, fout, norm, semantic_nc):
        super().__init__()
        self.learned_shortcut = (fin != fout)
        self.semantic_nc = semantic_nc
        self.fin, self.fout, self.fmiddle = fin, fout, min(fin, fout)
        self.conv_0 = nn.Conv2d(fin, self.fmiddle, kernel_size=3, padding=0)
        self.conv_1 = nn.Conv2d(self.fmiddle, fout, kernel_size=3, padding=0)
        self.se = SEBlock2(fout, reduction=8)
        if self.learned_shortcut:
            self.conv_s = nn.Conv2d(fin, fout, kernel_size=1, bias=False)
        if 'spectral' not in norm:
            raise NotImplementedError("sln_b200 implements the reference's 'spectralspadelayer3x3' blocks only")
        self.conv_0 = nn.Sequential(nn.ReflectionPad2d(1), spectral_norm(self.conv_0))
        self.conv_1 = nn.Sequential(nn.ReflectionPad2d(1), spectral_norm(self.conv_1))
        if self.learned_shortcut:
            self.conv_s = spectral_norm(self.conv_s)
        cfg = norm.replace('spectral', '')
        self.norm_0 = SPADE4(cfg, fin, semantic_nc)
        self.norm_1 = SPADE4(cfg, self.fmiddle, semantic_nc)
        if self.learned_shortcut:
            self.norm_s = SPADE4(cfg, fin, semantic_nc)


def _sn_weight(conv):
    """Eval-mode spectral-norm weight: W / sigma with sigma = u^T W v from the stored vectors (torch spectral_norm does no
    power iteration in eval mode)."""
    w = conv.weight_orig.detach()
    u, v = conv.weight_u.detach(), conv.weight_v.detach()
    sigma = torch.dot(u, torch.mv(w.reshape(w.size(0), -1), v))
    return w / sigma


def _pack_conv(w):
    """[Cout, Cin, kh, kw] -> [Cout, kh*kw*Cin] (tap-major, channel-minor: the K order of the NHWC implicit GEMM)."""
    return w.permute(0, 2, 3, 1).reshape(w.size(0), -1).contiguous().float()


class _Packed(object):
    pass


def _pretile(w):
    """Pre-split (TF32 hi | lo) and pre-tile a packed weight matrix [N, K] on its device for the bulk-copy operand path
    (sln_pack_weights); None when the matrix is too narrow for the tensor-core path anyway."""
    if not w.is_cuda or w.size(0) < 32 or w.size(1) < 32:
        return None
    lib = _lib.load()
    out = torch.empty(lib.sln_packed_weights_bytes(w.size(0), w.size(1)) // 4, device=w.device, dtype=torch.float32)
    _lib.check(lib.sln_pack_weights(w.data_ptr(), w.size(0), w.size(1), out.data_ptr(), _lib.cur_stream(w.device)), "pack_weights")
    return out


class GraphedForward(object):
    """One CUDA graph of ``generator(seg, z)`` for a fixed input shape: the ~190 launches of a forward become one replay.  The
    reference draws 50 z vectors at batch 1 (testing/test_SPADE_shade.py:77-79), where a forward is launch-bound, not GPU-bound.

        run = GraphedForward(netG, seg, z)        # warms up (packs the weights), captures
        img = run(seg, z)                          # copies the inputs into the captured buffers, replays; ``img`` is overwritten by the next call

    Built for the generator's CURRENT weights (their packed images are baked into the graph): rebuild after loading a checkpoint."""

    def __init__(self, generator, seg, z, warmup=2):
        if generator.training or getattr(generator, "taps", None) is not None:
            raise RuntimeError("GraphedForward needs an eval-mode generator without taps")
        dev = seg.device
        self.generator = generator
        self.seg = seg.detach().contiguous().float().clone()
        self.z = z.detach().contiguous().float().clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):               # first-use initialisation (function attributes, weight packing) must not be captured
            for _ in range(max(1, warmup)):
                generator(self.seg, self.z)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = generator(self.seg, self.z)

    def __call__(self, seg, z):
        self.seg.copy_(seg)
        self.z.copy_(z)
        self.graph.replay()
        return self.out


class SPADEGenerator4(nn.Module):
    def __init__(self, semantic_nc, target_nc, nz, ngf, norm, crop_size, n_up):
        super().__init__()
        if n_up != 'normal':
            raise NotImplementedError("only n_up='normal' is runnable in the reference (self.up is undefined, SPADE_related.py:1587,1600)")
        if nz <= 0:
            raise NotImplementedError("the reference instantiates the generator with a latent vector (nz=256)")
        nf = ngf
        self.nf, self.n_up, self.nz, self.has_z = ngf, n_up, nz, True
        self.semantic_nc, self.target_nc = semantic_nc, target_nc
        self.sw = self.sh = crop_size // 32
        self.fc = nn.Linear(nz, 16 * nf * self.sw * self.sh)
        self.head_0 = SPADEResnetBlock4(16 * nf, 16 * nf, norm, semantic_nc)
        self.G_middle_0 = SPADEResnetBlock4(16 * nf, 16 * nf, norm, semantic_nc)
        self.G_middle_1 = SPADEResnetBlock4(16 * nf, 16 * nf, norm, semantic_nc)
        self.up_0 = SPADEResnetBlock4(16 * nf, 8 * nf, norm, semantic_nc)
        self.up_1 = SPADEResnetBlock4(8 * nf, 4 * nf, norm, semantic_nc)
        self.up_2 = SPADEResnetBlock4(4 * nf, 2 * nf, norm, semantic_nc)
        self.up_3 = SPADEResnetBlock4(2 * nf, 1 * nf, norm, semantic_nc)
        self.conv_img = nn.Conv2d(nf, target_nc, 5, padding=2)
        self.up_b = nn.Upsample(scale_factor=2, mode='bilinear')
        self.up_n = nn.Upsample(scale_factor=2, mode='nearest')
        self._packed = None
        self.taps = None      # optional dict: block name -> NCHW copy of its output (parity tests)

    # ---------------------------------------------------------------------------------------- weight packing (once per weight version)
    def _pack_spade(self, sp, dev):
        P = _Packed()
        C = sp.norm_nc
        P.C = C
        P.dw = sp.mlp_preshared_depth[1].weight.detach().reshape(NHIDDEN // 8, 9).contiguous().float().to(dev)
        P.db = sp.mlp_preshared_depth[1].bias.detach().contiguous().float().to(dev)
        P.ws = _pack_conv(sp.mlp_shared[1].weight.detach()).to(dev)
        P.bs = sp.mlp_shared[1].bias.detach().contiguous().float().to(dev)
        P.ws_t = _pretile(P.ws)
        wg, wb = _pack_conv(sp.mlp_gamma[1].weight.detach()), _pack_conv(sp.mlp_beta[1].weight.detach())
        # one contraction for gamma and beta: output tile of `pair` columns = [gamma of pair/2 channels | beta of the same channels]
        P.pair = min(128, 2 * C)
        half = P.pair // 2
        K = wg.size(1)
        P.wgb = torch.stack([wg.view(C // half, half, K), wb.view(C // half, half, K)], dim=1).reshape(2 * C, K).contiguous().to(dev)
        P.wgb_t = _pretile(P.wgb)
        P.bg = sp.mlp_gamma[1].bias.detach().contiguous().float().to(dev)
        P.bb = sp.mlp_beta[1].bias.detach().contiguous().float().to(dev)
        P.eps = float(sp.param_free_norm.eps)
        return P

    def _pack_block(self, blk, dev):
        P = _Packed()
        P.fin, P.fout, P.fmiddle, P.learned = blk.fin, blk.fout, blk.fmiddle, blk.learned_shortcut
        P.w0 = _pack_conv(_sn_weight(blk.conv_0[1])).to(dev); P.w0_t = _pretile(P.w0); P.b0 = blk.conv_0[1].bias.detach().contiguous().float().to(dev)
        P.w1 = _pack_conv(_sn_weight(blk.conv_1[1])).to(dev); P.w1_t = _pretile(P.w1); P.b1 = blk.conv_1[1].bias.detach().contiguous().float().to(dev)
        P.n0, P.n1 = self._pack_spade(blk.norm_0, dev), self._pack_spade(blk.norm_1, dev)
        if blk.learned_shortcut:
            P.ws = _pack_conv(_sn_weight(blk.conv_s)).to(dev)
            P.ws_t = _pretile(P.ws)
            P.ns = self._pack_spade(blk.norm_s, dev)
        P.se1 = blk.se.fc[0].weight.detach().contiguous().float().to(dev)      # [C/8, C]
        P.se2 = blk.se.fc[2].weight.detach().contiguous().float().to(dev)      # [C, C/8]
        return P

    def _version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + tuple((b.data_ptr(), b._version) for b in self.buffers())

    def _pack(self, dev):
        ver = self._version()
        if self._packed is not None and self._packed.ver == ver and self._packed.dev == dev:
            return self._packed
        P = _Packed()
        P.ver, P.dev = ver, dev
        C0, hw = 16 * self.nf, self.sh * self.sw
        # fc output (B, C0*sh*sw) is viewed as NCHW (B, C0, sh, sw) by the reference: permute the rows so that it comes out NHWC
        w = self.fc.weight.detach().view(C0, hw, self.nz).permute(1, 0, 2).reshape(C0 * hw, self.nz)
        P.fcw = w.contiguous().float().to(dev)
        P.fcw_t = _pretile(P.fcw)
        P.fcb = self.fc.bias.detach().view(C0, hw).t().reshape(-1).contiguous().float().to(dev)
        P.blocks = {n: self._pack_block(getattr(self, n), dev) for n in ("head_0", "G_middle_0", "G_middle_1", "up_0", "up_1", "up_2", "up_3")}
        P.wimg = self.conv_img.weight.detach().permute(0, 2, 3, 1).contiguous().float().to(dev)     # [3, 5, 5, nf]
        P.bimg = self.conv_img.bias.detach().contiguous().float().to(dev)
        self._packed = P
        return P

    # ---------------------------------------------------------------------------------------- kernels
    @staticmethod
    def _conv(lib, st, x, B, H, W, Cin, ks, relu_in, w, b, Cout, wt=None):
        out = torch.empty(B, H, W, Cout, device=x.device, dtype=torch.float32)
        _lib.check(lib.sln_spade_conv(x.data_ptr(), B, H, W, Cin, ks, int(relu_in), w.data_ptr(), _lib.ptr(wt), _lib.ptr(b), Cout, out.data_ptr(), st),
                   "spade_conv")
        return out

    def _spade(self, lib, st, P, x, seg, seg_mode, B, H, W, slope, scratch):
        """SPADE4.forward (+ the block's leaky_relu when slope != 1): x NHWC [B,H,W,C] -> NHWC [B,H,W,C]."""
        dev = x.device
        C = P.C
        mean = torch.empty(B, device=dev); inv = torch.empty(B, device=dev)
        _lib.check(lib.sln_spade_ln_stats(x.data_ptr(), B, H * W * C, P.eps, scratch.data_ptr(), mean.data_ptr(), inv.data_ptr(), st), "ln_stats")
        feat = torch.empty(B, H, W, NHIDDEN // 8 + self.semantic_nc - 1, device=dev)
        _lib.check(lib.sln_spade_seg_features(seg.data_ptr(), B, self.semantic_nc, seg.size(2), seg_mode, H, W, P.dw.data_ptr(), P.db.data_ptr(),
                                              NHIDDEN // 8, feat.data_ptr(), st), "seg_features")
        actv = self._conv(lib, st, feat, B, H, W, feat.size(3), 3, False, P.ws, P.bs, NHIDDEN, P.ws_t)      # ReLU applied lazily by the consumer
        out = torch.empty(B, H, W, C, device=dev)
        _lib.check(lib.sln_spade_modulate(actv.data_ptr(), B, H, W, NHIDDEN, P.wgb.data_ptr(), _lib.ptr(P.wgb_t), P.bg.data_ptr(), P.bb.data_ptr(), C, P.pair,
                                          x.data_ptr(), mean.data_ptr(), inv.data_ptr(), float(slope), out.data_ptr(), st), "spade_modulate")
        return out

    def _block(self, lib, st, P, x, seg, seg_mode, B, H, W, scratch):
        """SPADEResnetBlock4.forward (:1487-1495)."""
        if P.learned:
            xs = self._conv(lib, st, self._spade(lib, st, P.ns, x, seg, seg_mode, B, H, W, 1.0, scratch), B, H, W, P.fin, 1, False, P.ws, None, P.fout, P.ws_t)
        else:
            xs = x
        dx = self._conv(lib, st, self._spade(lib, st, P.n0, x, seg, seg_mode, B, H, W, 0.2, scratch), B, H, W, P.fin, 3, False, P.w0, P.b0, P.fmiddle, P.w0_t)
        dx = self._conv(lib, st, self._spade(lib, st, P.n1, dx, seg, seg_mode, B, H, W, 0.2, scratch), B, H, W, P.fmiddle, 3, False, P.w1, P.b1, P.fout, P.w1_t)
        out = torch.empty(B, H, W, P.fout, device=x.device)
        _lib.check(lib.sln_spade_se_residual(dx.data_ptr(), xs.data_ptr(), B, H, W, P.fout, P.se1.data_ptr(), P.se2.data_ptr(), P.se1.size(0),
                                             scratch.data_ptr(), scratch.numel() * 4, out.data_ptr(), st), "se_residual")
        return out

    def _up(self, lib, st, x, B, H, W, C, bilinear):
        out = torch.empty(B, 2 * H, 2 * W, C, device=x.device)
        _lib.check(lib.sln_spade_upsample2x(x.data_ptr(), B, H, W, C, int(bilinear), out.data_ptr(), st), "upsample2x")
        return out

    def forward(self, input, z=None):
        if not input.is_cuda:
            raise RuntimeError("sln_b200 SPADEGenerator4 runs on CUDA (sm_100a) only; no CPU fallback")
        if self.training:
            raise NotImplementedError("SPADEGenerator4 is inference-only in the reference pipeline (test_SPADE_shade.py:12 .eval()); call .eval()")
        lib = _lib.load()
        dev = input.device
        st = _lib.cur_stream(dev)
        seg = input.contiguous().float()
        B, S = seg.size(0), seg.size(2)
        if seg.size(1) != self.semantic_nc or seg.size(3) != S:
            raise ValueError("input must be [B, %d, S, S]" % self.semantic_nc)
        if z is None:
            print("Missing z vector, sampling from normal")
            z = torch.randn(B, self.nz, dtype=torch.float32, device=dev)
        z = z.contiguous().float()
        with torch.no_grad():
            P = self._pack(dev)
            nf = self.nf
            scratch = torch.zeros(max(4096, 2 * B * 16 * nf * 64 + 64), device=dev, dtype=torch.float32)
            taps = self.taps
            x = self._conv(lib, st, z, B, 1, 1, self.nz, 1, False, P.fcw, P.fcb, 16 * nf * self.sh * self.sw, P.fcw_t).view(B, self.sh, self.sw, 16 * nf)
            H = W = self.sh

            def tap(name, t):
                if taps is not None:
                    taps[name] = t.permute(0, 3, 1, 2).contiguous()
            # head_0 sees the nearest-downsampled map seg_1 (:1579); every other block sees the full map, bilinearly resized inside SPADE4
            x = self._block(lib, st, P.blocks["head_0"], x, seg, 1, B, H, W, scratch); tap("head_0", x)
            x = self._up(lib, st, x, B, H, W, 16 * nf, False); H, W = 2 * H, 2 * W
            x = self._block(lib, st, P.blocks["G_middle_0"], x, seg, 0, B, H, W, scratch); tap("G_middle_0", x)
            x = self._block(lib, st, P.blocks["G_middle_1"], x, seg, 0, B, H, W, scratch); tap("G_middle_1", x)
            for name, cin, bil in (("up_0", 16 * nf, False), ("up_1", 8 * nf, False), ("up_2", 4 * nf, False), ("up_3", 2 * nf, True)):
                x = self._up(lib, st, x, B, H, W, cin, bil); H, W = 2 * H, 2 * W
                x = self._block(lib, st, P.blocks[name], x, seg, 0, B, H, W, scratch); tap(name, x)
            out = torch.empty(B, self.target_nc, H, W, device=dev)
            pre = torch.empty(B, self.target_nc, H, W, device=dev) if taps is not None else None
            _lib.check(lib.sln_spade_to_rgb(x.data_ptr(), B, H, W, nf, P.wimg.data_ptr(), P.bimg.data_ptr(), self.target_nc, 5, 0.2, _lib.ptr(pre),
                                            out.data_ptr(), st), "to_rgb")
            if taps is not None:
                taps["pre_tanh"] = pre
        return out


# =====================================================================================================================================
# Plain SPADE generator (reference SPADE_related.py:151-346: SPADEGenerator / SPADEResnetBlock / SPADE, + Conv2dBlock :16-68 and
# SEResBlock2 :87-101 for conv_img_pre).  The reference never instantiates it (test_SPADE_shade.py builds SPADEGenerator4), but it is the
# class BASELINE.json's north_star names; it runs on the same engine: every 3x3 convolution is the tcgen05 implicit GEMM with ZERO
# padding (nn.Conv2d(padding=1)), gamma and beta are one contraction whose epilogue applies x_hat * (1 + gamma) + beta with the
# parameter-free normalisation's statistics per (sample, channel) (nn.InstanceNorm2d) or per channel (eval-mode nn.BatchNorm2d).
class Conv2dBlock(nn.Module):
    def __init__(self, input_dim, output_dim, kernel_size, stride, padding=0, norm='none', activation='relu', pad_type='zero', use_bias=True):
        super(Conv2dBlock, self).__init__()
        self.use_bias = use_bias
        if pad_type == 'reflect':
            self.pad = nn.ReflectionPad2d(padding)
        elif pad_type == 'zero':
            self.pad = nn.ZeroPad2d(padding)
        else:
            raise ValueError("Unsupported padding type: {}".format(pad_type))
        self.conv = nn.Conv2d(input_dim, output_dim, kernel_size, stride, bias=self.use_bias)
        if norm == 'inst':
            self.norm = nn.InstanceNorm2d(output_dim, track_running_stats=False)
        elif norm == 'none':
            self.norm = None
        else:
            raise NotImplementedError("Conv2dBlock norm=%r (SEResBlock2 uses 'inst')" % (norm,))
        if activation == 'relu':
            self.activation = nn.ReLU(inplace=True)
        elif activation == 'none':
            self.activation = None
        else:
            raise NotImplementedError("Conv2dBlock activation=%r (SEResBlock2 uses 'relu' / 'none')" % (activation,))
        if kernel_size != 3 or stride != 1 or padding != 1:
            raise NotImplementedError("Conv2dBlock: 3x3, stride 1, padding 1 (as SEResBlock2 builds it)")
        self.pad_type, self.norm_type, self.act_type = pad_type, norm, activation


class SEResBlock2(nn.Module):
    def __init__(self, dim, norm='inst', activation='relu', pad_type='reflect', nz=0):
        super(SEResBlock2, self).__init__()
        if nz != 0:
            raise NotImplementedError("SEResBlock2 with nz > 0 is not used by SPADEGenerator")
        model = [Conv2dBlock(dim + nz, dim, 3, 1, 1, norm=norm, activation=activation, pad_type=pad_type),
                 Conv2dBlock(dim, dim + nz, 3, 1, 1, norm=norm, activation='none', pad_type=pad_type),
                 SEBlock2(dim + nz, reduction=4)]
        self.model = nn.Sequential(*model)


class SPADE(nn.Module):
    def __init__(self, config_text, norm_nc, label_nc):
        super().__init__()
        assert config_text.startswith('spade')
        parsed = re.search(r'spade(\D+)(\d)x\d', config_text)
        kind, ks = str(parsed.group(1)), int(parsed.group(2))
        if kind == 'instance':
            self.param_free_norm = nn.InstanceNorm2d(norm_nc, affine=False)
        elif kind == 'batch':
            self.param_free_norm = nn.BatchNorm2d(norm_nc, affine=False)
        else:
            raise ValueError('%s is not a recognized param-free norm type in SPADE' % kind)
        if ks != 3:
            raise NotImplementedError("SPADE kernel size %d (the engine implements 3x3)" % ks)
        self.kind, self.norm_nc, self.label_nc = kind, norm_nc, label_nc
        self.mlp_shared = nn.Sequential(nn.Conv2d(label_nc, NHIDDEN, kernel_size=ks, padding=ks // 2), nn.ReLU(inplace=True))
        self.mlp_gamma = nn.Conv2d(NHIDDEN, norm_nc, kernel_size=ks, padding=ks // 2)
        self.mlp_beta = nn.Conv2d(NHIDDEN, norm_nc, kernel_size=ks, padding=ks // 2)


class SPADEResnetBlock(nn.Module):
    def __init__(self, fin, fout, norm, semantic_nc):
        super().__init__()
        self.learned_shortcut = (fin != fout)
        self.semantic_nc = semantic_nc
        self.fin, self.fout, self.fmiddle = fin, fout, min(fin, fout)
        self.conv_0 = nn.Conv2d(fin, self.fmiddle, kernel_size=3, padding=1)
        self.conv_1 = nn.Conv2d(self.fmiddle, fout, kernel_size=3, padding=1)
        if self.learned_shortcut:
            self.conv_s = nn.Conv2d(fin, fout, kernel_size=1, bias=False)
        self.spectral = 'spectral' in norm
        if self.spectral:
            self.conv_0 = spectral_norm(self.conv_0)
            self.conv_1 = spectral_norm(self.conv_1)
            if self.learned_shortcut:
                self.conv_s = spectral_norm(self.conv_s)
        cfg = norm.replace('spectral', '')
        self.norm_0 = SPADE(cfg, fin, semantic_nc)
        self.norm_1 = SPADE(cfg, self.fmiddle, semantic_nc)
        if self.learned_shortcut:
            self.norm_s = SPADE(cfg, fin, semantic_nc)


def _conv_weight(conv, spectral):
    return _sn_weight(conv) if spectral else conv.weight.detach()


class SPADEGenerator(nn.Module):
    """reference SPADE_related.py:151-250, e.g. SPADEGenerator(41, 3, 256, 64, 'spectralspadeinstance3x3', 256, 'normal').
    Inference (eval) only, CUDA only; n_up 'normal' / 'more' / 'most' as in the reference."""

    def __init__(self, semantic_nc, target_nc, nz, ngf, norm, crop_size, n_up):
        super().__init__()
        if nz <= 0:
            raise NotImplementedError("SPADEGenerator without a latent vector (fc = Conv2d on the downsampled map) is not implemented")
        nf = ngf
        self.nf, self.n_up = ngf, n_up
        self.sw, self.sh = self.compute_latent_vector_size(n_up, crop_size)
        self.has_z, self.nz = True, nz
        self.semantic_nc, self.target_nc = semantic_nc, target_nc
        self.fc = nn.Linear(self.nz, 16 * nf * self.sw * self.sh)
        self.head_0 = SPADEResnetBlock(16 * nf, 16 * nf, norm, semantic_nc)
        self.G_middle_0 = SPADEResnetBlock(16 * nf, 16 * nf, norm, semantic_nc)
        self.G_middle_1 = SPADEResnetBlock(16 * nf, 16 * nf, norm, semantic_nc)
        self.up_0 = SPADEResnetBlock(16 * nf, 8 * nf, norm, semantic_nc)
        self.up_1 = SPADEResnetBlock(8 * nf, 4 * nf, norm, semantic_nc)
        self.up_2 = SPADEResnetBlock(4 * nf, 2 * nf, norm, semantic_nc)
        self.up_3 = SPADEResnetBlock(2 * nf, 1 * nf, norm, semantic_nc)
        final_nc = nf
        if n_up == 'most':
            self.up_4 = SPADEResnetBlock(1 * nf, nf // 2, norm, semantic_nc)
            final_nc = nf // 2
        self.final_nc = final_nc
        self.conv_img_pre = SEResBlock2(final_nc)
        self.conv_img = nn.Conv2d(final_nc, target_nc, 5, padding=2)
        self.up = nn.Upsample(scale_factor=2)
        self._packed = None
        self.taps = None

    def compute_latent_vector_size(self, n_up, crop_size):
        if n_up == 'normal':
            num_up_layers = 5
        elif n_up == 'more':
            num_up_layers = 6
        elif n_up == 'most':
            num_up_layers = 7
        else:
            raise ValueError('opt.num_upsampling_layers [%s] not recognized' % n_up)
        sw = crop_size // (2 ** num_up_layers)
        return sw, sw

    # ---------------------------------------------------------------------------------------- packing
    def _pack_spade(self, sp, dev):
        P = _Packed()
        C, nc = sp.norm_nc, sp.label_nc
        P.C, P.kind, P.eps = C, sp.kind, float(sp.param_free_norm.eps)
        P.cpad = (nc + 3) // 4 * 4
        w = sp.mlp_shared[0].weight.detach()                                   # [128, nc, 3, 3] -> channels zero-padded to cpad
        wp = torch.zeros(w.size(0), P.cpad, 3, 3, dtype=w.dtype, device=w.device)
        wp[:, :nc] = w
        P.ws = _pack_conv(wp).to(dev); P.ws_t = _pretile(P.ws)
        P.bs = sp.mlp_shared[0].bias.detach().contiguous().float().to(dev)
        wg, wb = _pack_conv(sp.mlp_gamma.weight.detach()), _pack_conv(sp.mlp_beta.weight.detach())
        P.pair = min(128, 2 * C)
        half, K = P.pair // 2, wg.size(1)
        P.wgb = torch.stack([wg.view(C // half, half, K), wb.view(C // half, half, K)], dim=1).reshape(2 * C, K).contiguous().to(dev)
        P.wgb_t = _pretile(P.wgb)
        P.bg = sp.mlp_gamma.bias.detach().contiguous().float().to(dev)
        P.bb = sp.mlp_beta.bias.detach().contiguous().float().to(dev)
        if sp.kind == 'batch':            # eval-mode BatchNorm2d(affine=False): (x - running_mean) / sqrt(running_var + eps), per channel
            P.bn_mean = sp.param_free_norm.running_mean.detach().contiguous().float().to(dev)
            P.bn_inv = torch.rsqrt(sp.param_free_norm.running_var.detach().double() + P.eps).float().contiguous().to(dev)
        return P

    def _pack_block(self, blk, dev):
        P = _Packed()
        P.fin, P.fout, P.fmiddle, P.learned = blk.fin, blk.fout, blk.fmiddle, blk.learned_shortcut
        P.w0 = _pack_conv(_conv_weight(blk.conv_0, blk.spectral)).to(dev); P.w0_t = _pretile(P.w0); P.b0 = blk.conv_0.bias.detach().contiguous().float().to(dev)
        P.w1 = _pack_conv(_conv_weight(blk.conv_1, blk.spectral)).to(dev); P.w1_t = _pretile(P.w1); P.b1 = blk.conv_1.bias.detach().contiguous().float().to(dev)
        P.n0, P.n1 = self._pack_spade(blk.norm_0, dev), self._pack_spade(blk.norm_1, dev)
        if blk.learned_shortcut:
            P.ws = _pack_conv(_conv_weight(blk.conv_s, blk.spectral)).to(dev); P.ws_t = _pretile(P.ws)
            P.ns = self._pack_spade(blk.norm_s, dev)
        return P

    def _version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + tuple((b.data_ptr(), b._version) for b in self.buffers())

    def _block_names(self):
        return ["head_0", "G_middle_0", "G_middle_1", "up_0", "up_1", "up_2", "up_3"] + (["up_4"] if self.n_up == 'most' else [])

    def _pack(self, dev):
        ver = self._version()
        if self._packed is not None and self._packed.ver == ver and self._packed.dev == dev:
            return self._packed
        P = _Packed()
        P.ver, P.dev = ver, dev
        C0, hw = 16 * self.nf, self.sh * self.sw
        w = self.fc.weight.detach().view(C0, hw, self.nz).permute(1, 0, 2).reshape(C0 * hw, self.nz)
        P.fcw = w.contiguous().float().to(dev); P.fcw_t = _pretile(P.fcw)
        P.fcb = self.fc.bias.detach().view(C0, hw).t().reshape(-1).contiguous().float().to(dev)
        P.blocks = {n: self._pack_block(getattr(self, n), dev) for n in self._block_names()}
        pre = self.conv_img_pre.model
        P.pre = []
        for cb in (pre[0], pre[1]):
            q = _Packed()
            q.w = _pack_conv(cb.conv.weight.detach()).to(dev); q.w_t = _pretile(q.w); q.b = cb.conv.bias.detach().contiguous().float().to(dev)
            q.pad = 0 if cb.pad_type == 'reflect' else 1
            q.act = 1 if cb.act_type == 'relu' else 0
            q.eps = float(cb.norm.eps)
            P.pre.append(q)
        P.se1 = pre[2].fc[0].weight.detach().contiguous().float().to(dev)
        P.se2 = pre[2].fc[2].weight.detach().contiguous().float().to(dev)
        P.wimg = self.conv_img.weight.detach().permute(0, 2, 3, 1).contiguous().float().to(dev)
        P.bimg = self.conv_img.bias.detach().contiguous().float().to(dev)
        self._packed = P
        return P

    # ---------------------------------------------------------------------------------------- kernels
    @staticmethod
    def _conv(lib, st, x, B, H, W, Cin, ks, relu_in, pad_mode, w, b, Cout, wt=None):
        out = torch.empty(B, H, W, Cout, device=x.device, dtype=torch.float32)
        _lib.check(lib.sln_conv2d_nhwc(x.data_ptr(), B, H, W, Cin, ks, int(relu_in), int(pad_mode), w.data_ptr(), _lib.ptr(wt), _lib.ptr(b), Cout,
                                       out.data_ptr(), st), "conv2d_nhwc")
        return out

    def _spade(self, lib, st, P, x, segs, B, H, W, slope):
        """SPADE.forward (:328-341) (+ the block's leaky_relu when slope != 1) on NHWC x."""
        dev, C = x.device, P.C
        if P.kind == 'instance':
            mean = torch.empty(B, C, device=dev); inv = torch.empty(B, C, device=dev)
            _lib.check(lib.sln_instnorm_stats(x.data_ptr(), B, H * W, C, P.eps, mean.data_ptr(), inv.data_ptr(), st), "instnorm_stats")
            sb, sc = C, 1
        else:
            mean, inv, sb, sc = P.bn_mean, P.bn_inv, 0, 1
        seg = segs[(H, W)]                                                     # label map at this resolution, NHWC, computed once per forward
        actv = self._conv(lib, st, seg, B, H, W, P.cpad, 3, False, 1, P.ws, P.bs, NHIDDEN, P.ws_t)   # ReLU applied lazily by the consumer
        out = torch.empty(B, H, W, C, device=dev)
        _lib.check(lib.sln_spade_modulate_ex(actv.data_ptr(), B, H, W, NHIDDEN, P.wgb.data_ptr(), _lib.ptr(P.wgb_t), P.bg.data_ptr(), P.bb.data_ptr(), C,
                                             P.pair, x.data_ptr(), mean.data_ptr(), inv.data_ptr(), sb, sc, 1, float(slope), out.data_ptr(), st),
                   "spade_modulate_ex")
        return out

    def _block(self, lib, st, P, x, segs, B, H, W):
        """SPADEResnetBlock.forward (:281-296)."""
        xs = x
        if P.learned:
            xs = self._conv(lib, st, self._spade(lib, st, P.ns, x, segs, B, H, W, 1.0), B, H, W, P.fin, 1, False, 1, P.ws, None, P.fout, P.ws_t)
        dx = self._conv(lib, st, self._spade(lib, st, P.n0, x, segs, B, H, W, 0.2), B, H, W, P.fin, 3, False, 1, P.w0, P.b0, P.fmiddle, P.w0_t)
        dx = self._conv(lib, st, self._spade(lib, st, P.n1, dx, segs, B, H, W, 0.2), B, H, W, P.fmiddle, 3, False, 1, P.w1, P.b1, P.fout, P.w1_t)
        return xs + dx

    def forward(self, input, z=None):
        if not input.is_cuda:
            raise RuntimeError("sln_b200 SPADEGenerator runs on CUDA (sm_100a) only; no CPU fallback")
        if self.training:
            raise NotImplementedError("SPADEGenerator is inference-only here (the reference releases no SPADE training code, SPADE_related.py:7); call .eval()")
        lib = _lib.load()
        dev = input.device
        st = _lib.cur_stream(dev)
        seg_in = input.contiguous().float()
        B, S = seg_in.size(0), seg_in.size(2)
        if seg_in.size(1) != self.semantic_nc or seg_in.size(3) != S:
            raise ValueError("input must be [B, %d, S, S]" % self.semantic_nc)
        if z is None:
            print("Missing z vector, sampling from normal")
            z = torch.randn(B, self.nz, dtype=torch.float32, device=dev)
        z = z.contiguous().float()
        with torch.no_grad():
            P = self._pack(dev)
            nf, nc = self.nf, self.semantic_nc
            cpad = (nc + 3) // 4 * 4
            taps = self.taps

            def tap(name, t):
                if taps is not None:
                    taps[name] = t.permute(0, 3, 1, 2).contiguous()

            def resize(h, w, nearest):
                out = torch.empty(B, h, w, cpad, device=dev)
                _lib.check(lib.sln_seg_resize_nhwc(seg_in.data_ptr(), B, nc, S, int(nearest), h, w, cpad, out.data_ptr(), st), "seg_resize")
                return out
            x = SPADEGenerator4._conv(lib, st, z, B, 1, 1, self.nz, 1, False, P.fcw, P.fcb, 16 * nf * self.sh * self.sw, P.fcw_t).view(B, self.sh, self.sw, 16 * nf)
            H, W = self.sh, self.sw
            # head_0 sees seg_1 = F.interpolate(seg, size=[sh, sw]) (nearest, :226), which SPADE then resizes bilinearly to the same size (identity);
            # every later block sees the full map resized bilinearly (align_corners=False) inside SPADE (:330): one resize per resolution
            segs = {(H, W): resize(H, W, True)}
            x = self._block(lib, st, P.blocks["head_0"], x, segs, B, H, W); tap("head_0", x)

            def up(x, H, W, C):
                out = torch.empty(B, 2 * H, 2 * W, C, device=dev)
                _lib.check(lib.sln_spade_upsample2x(x.data_ptr(), B, H, W, C, 0, out.data_ptr(), st), "upsample2x")     # nn.Upsample(scale_factor=2): nearest
                return out, 2 * H, 2 * W
            plan = [("G_middle_0", True)]
            plan.append(("G_middle_1", self.n_up in ('more', 'most')))
            plan += [("up_0", True), ("up_1", True), ("up_2", True), ("up_3", True)]
            if self.n_up == 'most':
                plan.append(("up_4", True))
            C = 16 * nf
            for name, do_up in plan:
                if do_up:
                    x, H, W = up(x, H, W, C)
                    segs = {(H, W): resize(H, W, False)}
                elif (H, W) in segs and name == "G_middle_1" and len(segs) == 1 and H == self.sh:
                    segs = {(H, W): resize(H, W, False)}      # (unreachable for sh < S: G_middle_0 always upsamples first)
                x = self._block(lib, st, P.blocks[name], x, segs, B, H, W); tap(name, x)
                C = P.blocks[name].fout
            # conv_img_pre = SEResBlock2 (:87-101): conv (reflect) -> InstanceNorm -> ReLU -> conv (reflect) -> InstanceNorm -> SE -> + residual
            y = x
            mean = torch.empty(B, C, device=dev); inv = torch.empty(B, C, device=dev)
            for q in P.pre:
                y = self._conv(lib, st, y, B, H, W, C, 3, False, q.pad, q.w, q.b, C, q.w_t)
                _lib.check(lib.sln_instnorm_stats(y.data_ptr(), B, H * W, C, q.eps, mean.data_ptr(), inv.data_ptr(), st), "instnorm_stats")
                yn = torch.empty_like(y)
                _lib.check(lib.sln_norm_act(y.data_ptr(), B, H * W, C, mean.data_ptr(), inv.data_ptr(), C, 1, q.act, yn.data_ptr(), st), "norm_act")
                y = yn
            scratch = torch.zeros(max(4096, 2 * B * C * 64 + 64), device=dev, dtype=torch.float32)
            pre_out = torch.empty(B, H, W, C, device=dev)
            _lib.check(lib.sln_spade_se_residual(y.data_ptr(), x.data_ptr(), B, H, W, C, P.se1.data_ptr(), P.se2.data_ptr(), P.se1.size(0),
                                                 scratch.data_ptr(), scratch.numel() * 4, pre_out.data_ptr(), st), "se_residual")
            tap("conv_img_pre", pre_out)
            out = torch.empty(B, self.target_nc, H, W, device=dev)
            pre = torch.empty(B, self.target_nc, H, W, device=dev) if taps is not None else None
            _lib.check(lib.sln_spade_to_rgb(pre_out.data_ptr(), B, H, W, C, P.wimg.data_ptr(), P.bimg.data_ptr(), self.target_nc, 5, 0.2, _lib.ptr(pre),
                                            out.data_ptr(), st), "to_rgb")
            if taps is not None:
                taps["pre_tanh"] = pre
        return out
