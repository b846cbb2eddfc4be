"""Decoder-only sampling loops.

Reference: testing/test_heatmap.py:52-62 (20 000 sequential batch-1 ``model.decoder`` calls, each preceded by a CPU
``np.random.multivariate_normal(mean_est, cov_est, O)`` draw and a host->device copy), testing/test_VAE.py:79-83 (4 draws per
validation batch), testing/test_acc_mean_std.py:47-51.

In eval mode nothing couples two decoder calls: BatchNorm applies running statistics row by row and the graph convolution never
crosses scenes (the batched graph is block-diagonal, data/suncg_dataset.py:318-325).  So K draws for one scene graph ARE one
decoder call on K replicas of the graph — the same kernels as the training path at a useful batch size, instead of K
latency-bound launches.  ``decode_samples`` does that; z is drawn on the device (``MVNSampler``: mean + L n, L = chol(cov)), or
taken from the caller (parity tests feed the reference's own draws).
"""
import torch


class MVNSampler(object):
    """z ~ N(mean, cov) on the device.  mean [Z], cov [Z,Z] (numpy / tensors, as unpickled from mean_cov.pkl)."""

    def __init__(self, mean, cov, device, jitter=1e-6):
        mean = torch.as_tensor(mean, dtype=torch.float64)
        cov = torch.as_tensor(cov, dtype=torch.float64)
        if mean.dim() != 1 or cov.shape != (mean.numel(), mean.numel()):
            raise ValueError("MVNSampler: mean [Z] and cov [Z,Z] expected, got %s / %s" % (tuple(mean.shape), tuple(cov.shape)))
        cov = 0.5 * (cov + cov.t())
        # an estimated covariance can be semi-definite: fall back to the symmetric square root (what numpy's SVD route computes)
        L, info = torch.linalg.cholesky_ex(cov + jitter * torch.eye(cov.size(0), dtype=torch.float64))
        if int(info) != 0:
            w, v = torch.linalg.eigh(cov)
            L = v * w.clamp(min=0).sqrt()
        self.mean = mean.to(device=device, dtype=torch.float32)
        self.Lt = L.t().contiguous().to(device=device, dtype=torch.float32)

    def sample(self, n, generator=None):
        eps = torch.randn(n, self.mean.numel(), device=self.mean.device, dtype=torch.float32, generator=generator)
        return torch.addmm(self.mean.unsqueeze(0), eps, self.Lt)


def replicate_graph(objs, triples, attributes, k):
    """K block-diagonal copies of one scene graph (node ids of copy i offset by i*O, as suncg_collate_fn does for scenes)."""
    O, T = objs.size(0), triples.size(0)
    off = (torch.arange(k, device=triples.device, dtype=triples.dtype) * O).view(k, 1, 1)
    tr = triples.unsqueeze(0).repeat(k, 1, 1)
    tr[:, :, 0:1] += off
    tr[:, :, 2:3] += off
    return objs.repeat(k), tr.view(k * T, 3), attributes.repeat(k)


@torch.no_grad()
def decode_samples(model, objs, triples, attributes, num_samples, sampler=None, z=None, chunk=2048, generator=None):
    """num_samples decoder draws for ONE scene graph -> (boxes [num_samples, O, box_dim], angles [num_samples, O, Nangle]).

    Equivalent to ``for k in range(num_samples): model.decoder(z[k], objs, triples, attributes)`` (test_heatmap.py:56-62) with the
    model in eval mode; runs ceil(num_samples / chunk) decoder calls on `chunk` replicas of the graph.
    z: optional [num_samples, O, Z] latent draws; otherwise drawn from `sampler` (MVNSampler) or N(0, I).
    """
    if model.training:
        raise RuntimeError("decode_samples needs model.eval(): training-mode BatchNorm couples the rows of a batch")
    dev = objs.device
    O = objs.size(0)
    Z = model.embedding_dim
    if z is not None and tuple(z.shape) != (num_samples, O, Z):
        raise ValueError("decode_samples: z must be [num_samples, O, %d]" % Z)
    boxes = torch.empty(num_samples, O, model.box_dim, device=dev, dtype=torch.float32)
    angles = torch.empty(num_samples, O, model.Nangle, device=dev, dtype=torch.float32)
    rep = None
    for lo in range(0, num_samples, chunk):
        k = min(chunk, num_samples - lo)
        if rep is None or rep[0] != k:
            rep = (k,) + replicate_graph(objs, triples, attributes, k)
        if z is not None:
            zk = z[lo:lo + k].reshape(k * O, Z)
        elif sampler is not None:
            zk = sampler.sample(k * O, generator)
        else:
            zk = torch.randn(k * O, Z, device=dev, dtype=torch.float32, generator=generator)
        b, a = model.decoder(zk, rep[1], rep[2], rep[3])
        boxes[lo:lo + k] = b.view(k, O, -1)
        angles[lo:lo + k] = a.view(k, O, -1)
    return boxes, angles
