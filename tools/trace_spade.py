"""Per-phase SM-clock trace of CTA (0,0,0) of the generator's convolution kernels (tuning build: tools/build_trace.sh spade,
SLN_LIB_PATH=sln_b200/libsln_b200_trace.so) plus isolated CUDA-event timings of the same launches.
usage: python tools/trace_spade.py"""
import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_lib = importlib.import_module("sln_b200._lib")
sp = importlib.import_module("sln_b200.models.SPADE_related")
lib = _lib.load()
dev = torch.device("cuda:0")
st = _lib.cur_stream(dev)
B = 16
#        H   Cin  ks  Cout
shapes = [(256, 44, 3, 128), (256, 128, 3, 64), (256, 64, 3, 64), (256, 128, 1, 64), (128, 256, 3, 128), (128, 128, 3, 128), (32, 1024, 3, 512), (16, 1024, 3, 1024)]
try:
    tracefn = lib.sln_debug_tc_trace_spade
except AttributeError:
    tracefn = None
for (H, Cin, ks, Cout) in shapes:
    x = torch.randn(B, H, H, Cin, device=dev)
    w = torch.randn(Cout, ks * ks * Cin, device=dev) * 0.05
    wt = sp._pretile(w)
    bias = torch.zeros(Cout, device=dev)
    out = torch.empty(B, H, H, Cout, device=dev)
    def run():
        _lib.check(lib.sln_spade_conv(x.data_ptr(), B, H, H, Cin, ks, 0, w.data_ptr(), _lib.ptr(wt), bias.data_ptr(), Cout, out.data_ptr(), st), "conv")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    reps = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    M, K = B * H * H, ks * ks * Cin
    print("H=%3d Cin=%4d ks=%d Cout=%4d  M=%7d K=%5d : %9.1f us  %6.1f TFLOP/s  (%.0f CTA-slots of %.1f us)" % (
        H, Cin, ks, Cout, M, K, us, 2.0 * M * Cout * K / us / 1e6, M / 128 * max(1, Cout // 128) / 148, us / (M / 128 * max(1, Cout // 128) / 148)))
    if tracefn is not None:
        tr = (ctypes.c_longlong * 48)()
        tracefn(tr)
        t = list(tr)
        w0 = t[1]
        rel = lambda v: (v - w0) if v else None
        print("   producer: init %s, fetches issued %s | chunk begin/stored: %s | main loop done %s, staged %s, barrier %s, tile applied %s, stats %s, end %s" % (
            rel(t[24]), rel(t[25]), " ".join("%s/%s" % (rel(t[16 + 2 * c]), rel(t[17 + 2 * c])) for c in range(4)), rel(t[4]), rel(t[26]), rel(t[27]), rel(t[5]), rel(t[6]), rel(t[7])))
        print("   MMA warp: full[c] at %s | first chunk issued %s, all issued %s   (entry->wait %d)" % (
            " ".join(str(rel(t[32 + c])) for c in range(8)), rel(t[10]), rel(t[11]), t[1] - t[0]))
