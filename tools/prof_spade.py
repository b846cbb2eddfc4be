"""One SPADEGenerator4 forward at the bench configuration (for ncu: `ncu -k regex:tc_gemm -s 47 -c 1 ... python tools/prof_spade.py [batch]`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench_spade  # noqa: E402
from oracle import spade_oracle as so  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
m = bench_spade._model(dev)
seg = so.synthetic_input(B, S=256, seed=100).to(dev)
z = torch.randn(B, 256, generator=torch.Generator().manual_seed(7)).to(dev)
out = m(seg, z)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.abs().mean()))
