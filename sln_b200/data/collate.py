"""Batch assembly for the VAE-graph path.

Reference: ``suncg_collate_fn`` (data/suncg_dataset.py:295-337) builds the flat batch on the CPU with eight ``torch.cat`` calls
and ``tensor_aug`` (utils.py:114-124) then issues eight separate ``.cuda()`` copies per step (train.py:69).

* ``suncg_collate_fn(batch)``     — the reference function, same 8-tuple of CPU tensors (host logic only; for DataLoader workers).
* ``DeviceCollator(device)(batch)`` — same 8-tuple, but ON THE DEVICE: the scenes are packed into one pinned wire buffer
  (include/sln_b200.h, ``sln_collate_layout``), moved with ONE async H2D copy, and ``sln_collate_finish`` (csrc/collate.cu)
  offsets the triple ids and writes obj_to_img / triple_to_img.  Drop it in as ``collate_fn`` is not possible (workers must not
  touch CUDA), so it wraps the *un-collated* list: ``DataLoader(..., collate_fn=list)`` + ``DeviceCollator``, or use
  ``DevicePrefetcher`` which also overlaps the copy of batch i+1 with the step of batch i.
"""
import ctypes

import torch

from .. import _lib


def _keep(batch):
    """(position in the batch, sample) of the scenes the reference keeps (suncg_dataset.py:311-312)."""
    return [(i, s) for i, s in enumerate(batch) if not (s[1].dim() == 0 or s[3].dim() == 0)]


def suncg_collate_fn(batch):
    """CPU restatement of the reference collate (same outputs, same dtypes): (ids, objs, boxes, triples, angles, attributes,
    obj_to_img, triple_to_img)."""
    kept = _keep(batch)
    ids = torch.LongTensor([int(s[0]) for _, s in kept])
    objs = torch.cat([s[1] for _, s in kept])
    boxes = torch.cat([s[2] for _, s in kept])
    angles = torch.cat([s[4] for _, s in kept])
    attrs = torch.cat([s[5] for _, s in kept])
    n_obj = torch.tensor([s[1].size(0) for _, s in kept], dtype=torch.int64)
    n_tri = torch.tensor([s[3].size(0) for _, s in kept], dtype=torch.int64)
    pos = torch.tensor([i for i, _ in kept], dtype=torch.int64)
    off = torch.cumsum(n_obj, 0) - n_obj
    triples = torch.cat([s[3] for _, s in kept]).clone()
    shift = torch.repeat_interleave(off, n_tri)
    triples[:, 0] += shift
    triples[:, 2] += shift
    return (ids, objs, boxes, triples, angles, attrs, torch.repeat_interleave(pos, n_obj), torch.repeat_interleave(pos, n_tri))


def wire_layout(lib, kept):
    """(B, O, T, box_dim, offsets10) of the wire buffer for the kept scenes (include/sln_b200.h, sln_collate_layout)."""
    B = len(kept)
    O = sum(s[1].size(0) for _, s in kept)
    T = sum(s[3].size(0) for _, s in kept)
    box_dim = kept[0][1][2].size(1)
    off = (ctypes.c_int64 * 10)()
    _lib.check(lib.sln_collate_layout(B, O, T, box_dim, off), "collate_layout")
    return B, O, T, box_dim, list(off)


def pack_wire(kept, lay, pin):
    """Host side of the batch assembly: write the kept scenes into the (pinned) uint8 buffer `pin` in wire layout.  Pure CPU work —
    what a DataLoader worker does instead of suncg_collate_fn; the device finishes it with sln_collate_finish."""
    o_si, o_oo, o_to, o_ids, o_objs, o_ang, o_att, o_tri, o_box, total = lay
    B = len(kept)
    n_obj = [s[1].size(0) for _, s in kept]
    n_tri = [s[3].size(0) for _, s in kept]
    O, T = sum(n_obj), sum(n_tri)
    box_dim = kept[0][1][2].size(1)
    i64, f32 = torch.int64, torch.float32

    def hv(off, n, dt):
        return pin[off: off + n * dt.itemsize].view(dt)
    hv(o_si, B, i64).copy_(torch.tensor([i for i, _ in kept], dtype=i64))
    oo = hv(o_oo, B + 1, i64); oo[0] = 0; torch.cumsum(torch.tensor(n_obj, dtype=i64), 0, out=oo[1:])
    to = hv(o_to, B + 1, i64); to[0] = 0; torch.cumsum(torch.tensor(n_tri, dtype=i64), 0, out=to[1:])
    hv(o_ids, B, i64).copy_(torch.tensor([int(s[0]) for _, s in kept], dtype=i64))
    torch.cat([s[1] for _, s in kept], out=hv(o_objs, O, i64))
    torch.cat([s[4] for _, s in kept], out=hv(o_ang, O, i64))
    torch.cat([s[5] for _, s in kept], out=hv(o_att, O, i64))
    torch.cat([s[3] for _, s in kept], out=hv(o_tri, 3 * T, i64).view(T, 3))
    torch.cat([s[2] for _, s in kept], out=hv(o_box, O * box_dim, f32).view(O, box_dim))
    return pin


def packed_batch(batch, lib=None):
    """un-collated samples -> (pinned wire buffer [total] uint8, (B, O, T, box_dim, offsets10)): the host-resident form of a batch that
    VAETrainStep.step_wire() consumes with ONE host->device copy."""
    lib = lib if lib is not None else _lib.load()
    kept = _keep(batch)
    if not kept:
        raise ValueError("packed_batch: every scene of the batch is empty")
    meta = wire_layout(lib, kept)
    pin = torch.empty(meta[4][9], dtype=torch.uint8)
    if torch.cuda.is_available():          # host-side packing also runs in loader workers / CPU tests, where there is nothing to pin for
        pin = pin.pin_memory()
    pack_wire(kept, meta[4], pin)
    return pin, meta


class DeviceCollator(object):
    """list of (room_id, objs, boxes, triples, angles, attributes) CPU samples -> the collated 8-tuple as CUDA tensors.

    One pinned staging buffer + one device wire buffer per slot (``slots`` >= 2 lets batch i+1 be packed while the copy of
    batch i is still in flight); buffers grow on demand.  The returned objs / boxes / angles / attributes / ids are views of
    the slot's device buffer: they stay valid until the slot is reused (``slots`` calls later).
    """

    def __init__(self, device="cuda", slots=2, check_ids=False):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("DeviceCollator needs a CUDA device (there is no CPU fallback; use suncg_collate_fn on the host)")
        self.lib = _lib.load()
        self.slots = [dict(pin=None, dev=None, evt=None) for _ in range(max(1, slots))]
        self.turn = 0
        self.check_ids = check_ids
        self.h2d_bytes = 0      # bytes of the last call's single host->device copy

    def __call__(self, batch):
        kept = _keep(batch)
        if not kept:
            raise ValueError("DeviceCollator: every scene of the batch is empty")
        B, O, T, box_dim, lay = wire_layout(self.lib, kept)
        o_si, o_oo, o_to, o_ids, o_objs, o_ang, o_att, o_tri, o_box, total = lay
        slot = self.slots[self.turn]
        self.turn = (self.turn + 1) % len(self.slots)
        if slot["pin"] is None or slot["pin"].numel() < total:
            cap = max(total * 2, 1 << 16)
            slot["pin"] = torch.empty(cap, dtype=torch.uint8).pin_memory()
            slot["dev"] = torch.empty(cap, dtype=torch.uint8, device=self.dev)
        elif slot["evt"] is not None:
            slot["evt"].synchronize()          # the previous copy out of this pinned buffer must have finished
        pin, dbuf = slot["pin"], slot["dev"]
        pack_wire(kept, lay, pin)

        def dv(off, n, dt):
            return dbuf[off: off + n * dt.itemsize].view(dt)
        i64, f32 = torch.int64, torch.float32
        dbuf[:total].copy_(pin[:total], non_blocking=True)
        slot["evt"] = torch.cuda.Event()
        slot["evt"].record(torch.cuda.current_stream(self.dev))
        self.h2d_bytes = total
        triples = dv(o_tri, 3 * T, i64).view(T, 3)                 # fixed up in place
        obj_to_img = torch.empty(O, dtype=i64, device=self.dev)
        triple_to_img = torch.empty(T, dtype=i64, device=self.dev)
        err = torch.zeros(1, dtype=torch.int32, device=self.dev) if self.check_ids else None
        _lib.check(self.lib.sln_collate_finish(dbuf.data_ptr(), dbuf.numel(), B, O, T, box_dim, triples.data_ptr(), obj_to_img.data_ptr(),
                                               triple_to_img.data_ptr(), err.data_ptr() if err is not None else None,
                                               _lib.cur_stream(self.dev)), "collate_finish")
        if err is not None and int(err.item()):
            raise ValueError("DeviceCollator: %d triples reference objects outside their scene" % int(err.item()))
        return (dv(o_ids, B, i64), dv(o_objs, O, i64), dv(o_box, O * box_dim, f32).view(O, box_dim), triples, dv(o_ang, O, i64),
                dv(o_att, O, i64), obj_to_img, triple_to_img)


class DevicePrefetcher(object):
    """Iterate a loader of un-collated sample lists; batch i+1 is packed, copied and finished on a side stream while the caller
    trains on batch i (the reference blocks on eight synchronous copies per step, train.py:69)."""

    def __init__(self, loader, device="cuda"):
        self.loader = loader
        self.collate = DeviceCollator(device, slots=3)
        self.stream = torch.cuda.Stream(self.collate.dev)

    def __iter__(self):
        dev = self.collate.dev
        nxt = None
        for samples in self.loader:
            # a slot's device buffer is reused three batches later: everything the caller has enqueued so far (the consumers of
            # the batch that last lived in it) must precede the new copy, which still overlaps the step enqueued next
            self.stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self.stream):
                cur = self.collate(samples)
                done = torch.cuda.Event(); done.record(self.stream)
            if nxt is not None:
                yield self._hand_over(nxt, dev)
            nxt = (cur, done)
        if nxt is not None:
            yield self._hand_over(nxt, dev)

    @staticmethod
    def _hand_over(item, dev):
        batch, done = item
        torch.cuda.current_stream(dev).wait_event(done)
        for t in batch:
            t.record_stream(torch.cuda.current_stream(dev))
        return batch
