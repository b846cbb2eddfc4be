"""TEST INFRASTRUCTURE ONLY — golden vectors of the reference's batch assembly (data/suncg_dataset.py:295-337).

    python oracle/gen_golden_collate.py     # needs /root/reference; writes tests/golden/collate.npz

Runs the UNMODIFIED reference ``suncg_collate_fn`` on seeded un-collated samples (ragged scene sizes, every 5th sample a
degenerate 0-dim scene that the reference drops) and stores the per-sample inputs together with its eight outputs.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

syn = importlib.import_module("sln_b200.data.synthetic")
NAMES = ("ids", "objs", "boxes", "triples", "angles", "attributes", "obj_to_img", "triple_to_img")


def main():
    ref = ref_shim._import("data.suncg_dataset").suncg_collate_fn
    samples = syn.synthetic_samples(13, nodes_per_scene=9, seed=7, ragged=True, empty_every=5)
    out = ref(samples)
    blob = {"n": np.int64(len(samples))}
    for i, s in enumerate(samples):
        for k, t in zip(("id", "objs", "boxes", "triples", "angles", "attributes"), s):
            blob["in%d_%s" % (i, k)] = np.asarray(t)
    for k, t in zip(NAMES, out):
        blob["out_" + k] = t.numpy()
    path = os.path.join(ROOT, "tests", "golden", "collate.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, {k: tuple(v.shape) for k, v in blob.items() if k.startswith("out_")})


if __name__ == "__main__":
    main()
