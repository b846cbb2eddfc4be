"""HBM roofline sweep of the north-star "scatter" stage (sln_csr_build + sln_gconv_pool_fwd: reference graph.py:92-108) over batch
sizes: algorithmic bytes (read new_s,new_o + CSR, write pooled) / CUDA-event time vs the measured HBM peak.
usage: python tools/bench_pool.py [scenes ...]"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

_lib = importlib.import_module("sln_b200._lib")
syn = importlib.import_module("sln_b200.data.synthetic")
lib = _lib.load()
dev = torch.device("cuda:0")
st = _lib.cur_stream(dev)
peak = 6538.6
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
H, D = 256, 128
sizes = [int(a) for a in sys.argv[1:]] or [64, 512, 2048, 8192]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
for B in sizes:
    _, objs, boxes, triples, angles, attrs, _, _ = syn.synthetic_batch(B, 32, seed=1)
    O, T = objs.size(0), triples.size(0)
    edges = triples[:, [0, 2]].contiguous().to(dev)
    x = torch.randn(T, 2 * H + D, device=dev)
    pooled = torch.empty(O, H, device=dev)
    ws = torch.empty(lib.sln_gconv_pool_workspace_bytes(O, T), dtype=torch.uint8, device=dev)
    _lib.check(lib.sln_csr_build(edges.data_ptr(), 2, O, T, ws.data_ptr(), ws.numel(), st), "csr_build")

    def run():
        _lib.check(lib.sln_gconv_pool_fwd(x.data_ptr(), O, T, H, D, pooled.data_ptr(), ws.data_ptr(), ws.numel(), st), "pool_fwd")
    for _ in range(3):
        run()
    reps = 10
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        if not os.environ.get("SLN_POOL_WARM"):      # SLN_POOL_WARM=1: operands stay L2-resident, as inside the train step
            flush.zero_()          # cold L2 for every timed launch ...
            flush_rd.sum()         # ... and clean: a 256 MB read pass forces the dirty lines of the write out before the timed region
        a.record(); run(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[reps // 2]
    nbytes = 4.0 * (2.0 * T * H + 2.0 * T + 2.0 * O + 1.0 + O * H)
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"SLN_POOL": os.environ.get("SLN_POOL", "auto"), "scenes": B, "O": O, "T": T, "algorithmic_MB": round(nbytes / 1e6, 2), "us": round(ms * 1e3, 2), "GB/s": round(gbs, 1),
                      "frac_of_measured_hbm_peak": round(gbs / peak, 3)}))
