"""Decoder-only sampling loops (reference testing/test_heatmap.py:52-62, testing/test_VAE.py:79-83): K draws as one
decoder call on K replicas of the scene graph == K sequential batch-1 decoder calls, bit for bit."""
import importlib

import numpy as np
import pytest
import torch

from helpers import our_model, syn

samp = importlib.import_module("sln_b200.models.sampling")
DEV = "cuda:0"


def test_replicate_graph_is_block_diagonal():
    objs, triples, _, _, attrs = syn.fixture_graph()
    o, t, a = samp.replicate_graph(objs, triples, attrs, 3)
    assert o.tolist() == objs.tolist() * 3 and a.tolist() == attrs.tolist() * 3
    for i in range(3):
        blk = t[i * 9:(i + 1) * 9]
        assert torch.equal(blk[:, 1], triples[:, 1])
        assert torch.equal(blk[:, [0, 2]] - 6 * i, triples[:, [0, 2]])
    assert torch.equal(triples, syn.fixture_graph()[1])     # input untouched


def test_mvn_sampler_moments_cpu():
    g = torch.Generator().manual_seed(0)
    A = torch.randn(8, 8, generator=g, dtype=torch.float64)
    cov, mean = A @ A.t() / 8 + 0.1 * torch.eye(8, dtype=torch.float64), torch.arange(8, dtype=torch.float64)
    s = samp.MVNSampler(mean.numpy(), cov.numpy(), "cpu")
    z = s.sample(200000, generator=torch.Generator().manual_seed(1)).double()
    assert torch.allclose(z.mean(0), mean, atol=2e-2)
    assert torch.allclose(torch.cov(z.t()), cov, atol=3e-2)
    # a singular covariance (rank 3) must not fail
    B = torch.randn(8, 3, generator=g, dtype=torch.float64)
    z = samp.MVNSampler(mean, B @ B.t(), "cpu", jitter=0.0).sample(50000, generator=torch.Generator().manual_seed(2)).double()
    assert torch.allclose(torch.cov(z.t()), B @ B.t(), atol=0.15, rtol=0.03)   # sampling noise of 50k draws, entries up to ~9


@pytest.mark.gpu
@pytest.mark.parametrize("norm", ["batch", "none"])
def test_batched_draws_equal_sequential_decoder_calls(norm):
    objs, triples, _, _, attrs = [t.to(DEV) for t in syn.fixture_graph()]
    m = our_model(E=64, layers=5, norm=norm).to(DEV).eval()
    K, O = 37, objs.size(0)
    z = torch.randn(K, O, 64, generator=torch.Generator().manual_seed(3)).to(DEV)
    boxes, angles = samp.decode_samples(m, objs, triples, attrs, K, z=z, chunk=16)     # chunks of 16, 16, 5
    with torch.no_grad():
        for k in range(K):
            b, a = m.decoder(z[k], objs, triples, attrs)                                  # test_heatmap.py:59
            assert torch.equal(boxes[k], b) and torch.equal(angles[k], a)


@pytest.mark.gpu
def test_device_mvn_draws_feed_the_decoder():
    objs, triples, _, _, attrs = [t.to(DEV) for t in syn.fixture_graph()]
    m = our_model(E=64, layers=5, norm="batch").to(DEV).eval()
    rs = np.random.RandomState(0)
    A = rs.randn(64, 64)
    s = samp.MVNSampler(rs.randn(64), A @ A.T / 64, DEV)
    boxes, angles = samp.decode_samples(m, objs, triples, attrs, 1000, sampler=s, chunk=512)
    assert boxes.shape == (1000, 6, 6) and angles.shape == (1000, 6, 24)
    assert torch.isfinite(boxes).all() and torch.allclose(angles.exp().sum(-1), torch.ones(1000, 6, device=DEV), atol=1e-4)
    assert boxes.std(0).min() > 0          # draws differ
    m.train()
    with pytest.raises(RuntimeError):
        samp.decode_samples(m, objs, triples, attrs, 2)


def test_decode_samples_chunking_with_a_stub_decoder_cpu():
    """Host logic only (no CUDA): the chunked replica batches cover every draw exactly once, in order, whatever the chunk size."""
    class Stub(object):
        training, embedding_dim, box_dim, Nangle = False, 4, 6, 24
        calls = []

        def decoder(self, z, objs, triples, attributes):
            self.calls.append((z.size(0), triples.size(0)))
            assert int(triples[:, [0, 2]].max()) < objs.numel()                 # replica offsets stay inside the replicated graph
            return z[:, :1].expand(-1, 6) + objs[:, None].float(), z[:, 1:2].expand(-1, 24)
    objs, triples, _, _, attrs = syn.fixture_graph()
    O = objs.numel()
    z = torch.arange(11 * O * 4, dtype=torch.float32).view(11, O, 4)
    for chunk in (1, 4, 11, 64):
        m = Stub(); m.calls = []
        boxes, angles = samp.decode_samples(m, objs, triples, attrs, 11, z=z, chunk=chunk)
        assert boxes.shape == (11, O, 6) and angles.shape == (11, O, 24)
        assert torch.equal(boxes[..., 0], z[..., 0] + objs[None].float()) and torch.equal(angles[..., 0], z[..., 1])
        assert sum(c[0] for c in m.calls) == 11 * O and len(m.calls) == -(-11 // chunk)
    with pytest.raises(ValueError):
        samp.decode_samples(Stub(), objs, triples, attrs, 3, z=z)              # z has 11 draws, 3 requested
