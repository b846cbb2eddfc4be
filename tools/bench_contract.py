"""Micro-benchmark of sln_contract (the contraction primitive): isolated launches, CUDA-event timed.
usage: python tools/bench_contract.py [engine]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_lib = importlib.import_module("sln_b200._lib")
lib = _lib.load()
eng = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
st = _lib.cur_stream(dev)
shapes = [(3968, 640, 256, 1, 1, 0), (3968, 384, 256, 1, 0, 0), (2048, 256, 128, 1, 0, 0), (3968, 640, 32, 1, 1, 0), (3968, 640, 128, 1, 1, 0), (3968, 256, 384, 1, 1, 0), (2048, 256, 256, 1, 1, 0), (2048, 128, 256, 1, 1, 0),
          (2048, 128, 32, 1, 1, 0), (128, 128, 32, 1, 1, 0), (128, 128, 256, 1, 1, 0), (3968, 256, 640, 1, 0, 0), (640, 256, 3968, 0, 0, 1), (256, 384, 3968, 0, 0, 1), (256, 256, 2048, 0, 0, 1)]
for (M, N, K, arc, brc, acc) in shapes:
    A = torch.randn((M, K) if arc else (K, M), device=dev)
    B = torch.randn((N, K) if brc else (K, N), device=dev)
    C = torch.zeros(M, N, device=dev)
    def run():
        _lib.check(lib.sln_contract(A.data_ptr(), A.stride(0), arc, B.data_ptr(), B.stride(0), brc, C.data_ptr(), N, M, N, K, acc, eng, st), "contract")
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    reps = 50
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    import ctypes
    tr = (ctypes.c_longlong * 48)()
    lib.sln_debug_tc_trace(tr)
    t = list(tr)
    if any(t):
        w = t[1]      # producer warp 0: time of griddepcontrol.wait return = origin
        rel = lambda x: (x - w) if x else None
        print("   [cycles after the dependency wait]  producer: init done %s, first fetches issued %s | chunk c begin/stored: %s | main loop done %s, staged %s, barrier %s, tile applied %s, stats %s, end %s" % (
            rel(t[24]), rel(t[25]), " ".join("%s/%s" % (rel(t[16 + 2 * c]), rel(t[17 + 2 * c])) for c in range(4)), rel(t[4]), rel(t[26]), rel(t[27]), rel(t[5]), rel(t[6]), rel(t[7])))
        print("   MMA warp: full[c] observed at %s | first chunk issued %s, all issued %s   (entry->wait %d)" % (
            " ".join(str(rel(t[32 + c])) for c in range(8)), rel(t[10]), rel(t[11]), t[1] - t[0]))
    print("M=%5d N=%4d K=%5d a_rc=%d b_rc=%d acc=%d : %8.2f us  %7.2f TFLOP/s" % (M, N, K, arc, brc, acc, us, 2.0 * M * N * K / us / 1e6))
