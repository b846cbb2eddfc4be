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
    tr = (ctypes.c_longlong * 16)()
    lib.sln_debug_tc_trace(tr)
    t = list(tr)
    print("   producer warp0: alloc+sync %d | first-fetch issued %d | chunk0 produced %d | mainloop+drain %d | staging+apply %d | stats/finalize %d | dealloc %d   MMA warp: first chunk issued at %d, all issued at %d" % (
        t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6], t[10] - t[8], t[11] - t[8]))
    print("M=%5d N=%4d K=%5d a_rc=%d b_rc=%d acc=%d : %8.2f us  %7.2f TFLOP/s" % (M, N, K, arc, brc, acc, us, 2.0 * M * N * K / us / 1e6))
