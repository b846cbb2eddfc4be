"""Synthetic SUNCG-shaped scene-graph batches (no dataset is available offline).

Mirrors what ``SuncgDataset.__getitem__`` + ``suncg_collate_fn`` emit (reference data/suncg_dataset.py:110-337) without
the SUNCG metadata:  per scene ``n`` nodes, the last one is the room (``objs[-1] == 0``); object boxes are normalised to
the room, the room box is ``[0,0,0,x,y,z]`` un-normalised (:134-141,216-225); one random relation per object with a random
direction (:193-205) and one ``__in_room__`` triple per object (:208-212); node ids are offset per scene (:318-325).
"""
import random

import torch

# reference testing/test_utils.py:44-64 and data/suncg_dataset.py:64-70
PRED_NAMES = ['__in_room__', 'left of', 'right of', 'behind', 'in front of', 'inside', 'surrounding', 'left touching',
              'right touching', 'front touching', 'behind touching', 'front left', 'front right', 'back left', 'back right', 'on']
OBJECT_NAMES = ["__room__", "curtain", "shower_curtain", "dresser", "counter", "bookshelf", "picture", "mirror", "floor_mat",
                "chair", "sink", "desk", "table", "lamp", "door", "clothes", "person", "toilet", "cabinet", "floor", "window",
                "blinds", "wall", "pillow", "whiteboard", "bathtub", "television", "night_stand", "sofa", "refridgerator", "bed",
                "shelves"]
ATTRIB_NAMES = ['none', 'tall', 'short', 'large', 'small']


def default_vocab():
    return {
        'object_idx_to_name': list(OBJECT_NAMES),
        'pred_idx_to_name': list(PRED_NAMES),
        'attrib_idx_to_name': list(ATTRIB_NAMES),
    }


def synthetic_scene(n_nodes, gen, rng):
    """One scene: (objs[n], boxes[n,6], triples[2(n-1),3], angles[n], attributes[n]) with the room as last node."""
    n_obj = n_nodes - 1
    objs = torch.cat([torch.randint(1, len(OBJECT_NAMES), (n_obj,), generator=gen), torch.zeros(1, dtype=torch.long)])
    lo_hi = torch.rand(n_obj, 3, 2, generator=gen).sort(dim=2).values
    boxes = torch.cat([lo_hi[:, :, 0], lo_hi[:, :, 1]], dim=1)
    room = torch.tensor([[0.0, 0.0, 0.0, 2 + 4 * rng.random(), 2 + rng.random(), 2 + 4 * rng.random()]])
    boxes = torch.cat([boxes, room], dim=0).float()
    angles = torch.cat([torch.randint(0, 24, (n_obj,), generator=gen), torch.zeros(1, dtype=torch.long)])
    attributes = torch.cat([torch.randint(0, len(ATTRIB_NAMES), (n_obj,), generator=gen), torch.zeros(1, dtype=torch.long)])
    triples = []
    for cur in range(n_obj):
        if n_obj > 1:
            other = rng.choice([k for k in range(n_obj) if k != cur])
            s, o = (cur, other) if rng.random() > 0.5 else (other, cur)
            triples.append([s, rng.randint(1, 10), o])
    for i in range(n_obj):
        triples.append([i, 0, n_obj])
    return objs, boxes, torch.tensor(triples, dtype=torch.long).view(-1, 3), angles, attributes


def synthetic_batch(n_scenes, nodes_per_scene=32, seed=42):
    """Collated batch in the order train.py:69 unpacks it:
    (ids, objs, boxes, triples, angles, attributes, obj_to_img, triple_to_img), all CPU tensors."""
    gen = torch.Generator().manual_seed(seed)
    rng = random.Random(seed)
    cols = [[] for _ in range(7)]
    off = 0
    for i in range(n_scenes):
        objs, boxes, triples, angles, attrs = synthetic_scene(nodes_per_scene, gen, rng)
        triples = triples.clone()
        triples[:, 0] += off
        triples[:, 2] += off
        for c, t in zip(cols, (objs, boxes, triples, angles, attrs, torch.full((objs.size(0),), i, dtype=torch.long),
                               torch.full((triples.size(0),), i, dtype=torch.long))):
            c.append(t)
        off += objs.size(0)
    objs, boxes, triples, angles, attrs, o2i, t2i = [torch.cat(c) for c in cols]
    return torch.arange(n_scenes), objs, boxes, triples, angles, attrs, o2i, t2i


def synthetic_samples(n_scenes, nodes_per_scene=32, seed=42, ragged=False, empty_every=0):
    """Un-collated samples as SuncgDataset.__getitem__ returns them (data/suncg_dataset.py:292):
    [(room_id, objs, boxes, triples, angles, attributes)], triples with scene-local ids.  ragged: node counts vary in
    [2, nodes_per_scene]; empty_every = k > 0: every k-th sample is a degenerate scene (0-dim objs) that the collate drops."""
    gen = torch.Generator().manual_seed(seed)
    rng = random.Random(seed)
    out = []
    for i in range(n_scenes):
        if empty_every and i % empty_every == empty_every - 1:
            out.append((1000 + i, torch.tensor(0), torch.zeros(0, 6), torch.tensor(0), torch.tensor(0), torch.tensor(0)))
            continue
        n = rng.randint(2, nodes_per_scene) if ragged else nodes_per_scene
        objs, boxes, triples, angles, attrs = synthetic_scene(n, gen, rng)
        out.append((1000 + i, objs, boxes, triples, angles, attrs))
    return out


def fixture_graph():
    """Config-1 fixture: the 5-object graph of reference testing/test_heatmap.py:41-43 with the layout of test.py:46-50."""
    objs = torch.tensor([30, 11, 18, 9, 13, 0])
    triples = torch.tensor([[0, 3, 1], [2, 1, 0], [3, 1, 1], [4, 15, 1], [0, 0, 5], [1, 0, 5], [2, 0, 5], [3, 0, 5], [4, 0, 5]])
    boxes = torch.tensor([
        [0.31150928139686584, 0.3127100169658661, 0.003096628002822399, 0.7295752763748169, 0.8262581825256348, 0.054250866174697876],
        [-0.06599953025579453, 0.017223943024873734, 0.2885378897190094, 0.2573782205581665, 0.7553179860115051, 0.42857787013053894],
        [0.5567594766616821, 0.017786923795938492, 0.142490953207016, 0.9046159982681274, 0.31667089462280273, 0.6691973209381104],
        [0.6205720901489258, 0.018211644142866135, 0.8416993021965027, 0.8348240852355957, 0.3893248736858368, 0.963701605796814],
        [0.171146959066391, 0.017671708017587662, 0.8085968494415283, 0.4601595997810364, 0.5026606321334839, 0.9657217264175415],
        [0.0, 0.0, 0.0, 1.0, 0.7327236533164978, 0.9278678297996521]], dtype=torch.float32)
    angles = torch.tensor([0, 18, 6, 12, 12, 0])
    attributes = torch.zeros(6, dtype=torch.long)
    return objs, triples, boxes, angles, attributes


def shard_batch(batch, rank, world):
    """Scene-shard a collated batch (the 8-tuple of synthetic_batch / suncg_collate_fn, data/suncg_dataset.py:295-337) for data
    parallelism: rank r keeps a contiguous block of scenes; the batched graph is block-diagonal (no edge crosses scenes,
    :318-325), so the shard is self-contained once node ids are re-based.  -> (objs, triples, boxes, angles, attributes)."""
    _, objs, boxes, triples, angles, attrs, o2i, t2i = batch
    n_scenes = int(o2i.max().item()) + 1
    per = (n_scenes + world - 1) // world
    lo, hi = rank * per, min(n_scenes, (rank + 1) * per)
    om = (o2i >= lo) & (o2i < hi)
    tm = (t2i >= lo) & (t2i < hi)
    first = int(torch.nonzero(om)[0]) if om.any() else 0
    tr = triples[tm].clone()
    tr[:, 0] -= first
    tr[:, 2] -= first
    return objs[om], tr, boxes[om], angles[om], attrs[om]
