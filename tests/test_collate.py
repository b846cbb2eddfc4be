"""Batch assembly: the reference's suncg_collate_fn (data/suncg_dataset.py:295-337) vs the host restatement
(CPU) and vs the one-copy + one-kernel device path (GPU, through the C ABI)."""
import importlib
import os

import numpy as np
import pytest
import torch

syn = importlib.import_module("sln_b200.data.synthetic")
col = importlib.import_module("sln_b200.data.collate")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collate.npz")
NAMES = ("ids", "objs", "boxes", "triples", "angles", "attributes", "obj_to_img", "triple_to_img")


def _golden():
    z = np.load(GOLD)
    samples = []
    for i in range(int(z["n"])):
        samples.append(tuple(torch.from_numpy(np.asarray(z["in%d_%s" % (i, k)])) if k != "id" else int(z["in%d_id" % i])
                             for k in ("id", "objs", "boxes", "triples", "angles", "attributes")))
    want = [torch.from_numpy(z["out_" + k]) for k in NAMES]
    return samples, want


def _same(got, want):
    assert len(got) == len(want) == 8
    for name, g, w in zip(NAMES, got, want):
        g = g.cpu()
        assert g.dtype == w.dtype, name
        assert g.shape == w.shape, name
        assert torch.equal(g, w), name


def test_host_collate_matches_reference_golden():
    samples, want = _golden()
    _same(col.suncg_collate_fn(samples), want)


def test_host_collate_matches_synthetic_batch_builder():
    got = col.suncg_collate_fn(syn.synthetic_samples(6, 8, seed=3))
    want = syn.synthetic_batch(6, 8, seed=3)
    for g, w in zip(got[1:], want[1:]):
        assert torch.equal(g, w)


def test_host_collate_matches_live_reference_when_available():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref = ref_shim._import("data.suncg_dataset").suncg_collate_fn
    samples = syn.synthetic_samples(21, 12, seed=11, ragged=True, empty_every=4)
    _same(col.suncg_collate_fn(samples), ref(samples))


def test_device_collator_refuses_cpu():
    with pytest.raises(RuntimeError):
        col.DeviceCollator("cpu")


@pytest.mark.gpu
def test_device_collate_matches_reference_golden():
    samples, want = _golden()
    c = col.DeviceCollator("cuda:0", check_ids=True)
    _same(c(samples), want)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nodes,ragged,empty", [(1, 2, False, 0), (64, 32, False, 0), (300, 40, True, 7), (512, 32, False, 0)])
def test_device_collate_matches_host_collate(n, nodes, ragged, empty):
    samples = syn.synthetic_samples(n, nodes, seed=n, ragged=ragged, empty_every=empty)
    c = col.DeviceCollator("cuda:0", slots=2, check_ids=True)
    for _ in range(3):                      # slot reuse
        got = [t.clone() for t in c(samples)]
    _same(got, col.suncg_collate_fn(samples))


@pytest.mark.gpu
def test_device_collate_flags_out_of_scene_ids():
    samples = syn.synthetic_samples(4, 6, seed=1)
    bad = list(samples[2]); tr = bad[3].clone(); tr[0, 2] = 6; bad[3] = tr; samples[2] = tuple(bad)
    with pytest.raises(ValueError):
        col.DeviceCollator("cuda:0", check_ids=True)(samples)


@pytest.mark.gpu
def test_prefetcher_yields_every_batch_in_order():
    batches = [syn.synthetic_samples(8, 10, seed=s, ragged=True) for s in range(5)]
    got = [[t.clone() for t in b] for b in col.DevicePrefetcher(batches, "cuda:0")]
    assert len(got) == 5
    for g, b in zip(got, batches):
        _same(g, col.suncg_collate_fn(b))


@pytest.mark.gpu
def test_train_step_from_wire_buffer_equals_step_from_collated_tensors():
    """VAETrainStep.step_wire (one H2D copy of the packed batch + sln_collate_finish + the step) == VAETrainStep.step on the tensors
    the reference collate produces, bit for bit."""
    from helpers import our_model
    sutils = importlib.import_module("sln_b200.utils")
    samples = syn.synthetic_samples(8, 12, seed=21, ragged=True)
    ids, objs, boxes, triples, angles, attrs, o2i, t2i = col.suncg_collate_fn(samples)
    wire, meta = col.packed_batch(samples)
    losses = []
    for use_wire in (False, True):
        m = our_model(E=16, layers=2, norm="batch", seed=5).to("cuda:0").train()
        st = sutils.VAETrainStep(m, objs.size(0), triples.size(0), use_graph=False, sample_eps=False, wire_meta=meta if use_wire else None)
        st.epsn.copy_(torch.randn(st.epsn.shape, generator=torch.Generator().manual_seed(9)).to("cuda:0"))
        out = st.step_wire(wire) if use_wire else st.step((objs, triples, boxes, angles, attrs))
        losses.append(out.clone().cpu())
        if use_wire:
            assert torch.equal(st.triples.cpu(), triples) and torch.equal(st.obj_to_img.cpu(), o2i) and torch.equal(st.triple_to_img.cpu(), t2i)
    assert torch.isfinite(losses[0]).all() and torch.equal(losses[0], losses[1])


def test_host_wire_packing_holds_the_collated_tensors():
    """pack_wire (the host half of the batch assembly, CPU only): every section of the wire buffer equals what the reference collate
    produces, except that triples still carry scene-local ids (the device kernel adds the object offsets)."""
    samples = syn.synthetic_samples(9, 7, seed=4, ragged=True, empty_every=4)
    ids, objs, boxes, triples, angles, attrs, o2i, t2i = col.suncg_collate_fn(samples)
    wire, (B, O, T, bd, lay) = col.packed_batch(samples)
    assert wire.is_pinned() or not torch.cuda.is_available()
    assert (B, O, T, bd) == (ids.numel(), objs.numel(), triples.size(0), 6) and wire.numel() == lay[9]

    def sec(k, n, dt):
        return wire[lay[k]: lay[k] + n * dt.itemsize].view(dt)
    i64 = torch.int64
    kept = [i for i, s in enumerate(samples) if s[1].dim() > 0]
    assert sec(0, B, i64).tolist() == kept                                     # positions in the DataLoader batch
    obj_off, tri_off = sec(1, B + 1, i64), sec(2, B + 1, i64)
    assert obj_off[0] == 0 and obj_off[-1] == O and tri_off[-1] == T
    assert torch.equal(sec(3, B, i64), ids) and torch.equal(sec(4, O, i64), objs)
    assert torch.equal(sec(5, O, i64), angles) and torch.equal(sec(6, O, i64), attrs)
    assert torch.equal(sec(8, O * bd, torch.float32).view(O, bd), boxes)
    local = sec(7, 3 * T, i64).view(T, 3).clone()
    shift = torch.repeat_interleave(obj_off[:-1], tri_off[1:] - tri_off[:-1])
    local[:, 0] += shift; local[:, 2] += shift
    assert torch.equal(local, triples)
    assert torch.equal(torch.repeat_interleave(sec(0, B, i64), obj_off[1:] - obj_off[:-1]), o2i)
