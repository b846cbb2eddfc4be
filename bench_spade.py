"""bench.py --workload spade : BASELINE.json configs[3] — SPADEGenerator4 256x256 semantic+depth -> RGB, batch 16 per GPU.

One "step" = one forward of the generator over a batch of 16 images (41-channel 256x256 input + z[256] -> [16,3,256,256]),
random-init seeded weights, eval mode.  metric = images/s (whole job).  Batch-sharded: every GPU runs its own 16 images
with a replica of the 443 MB weights; no collective on the data path -> "scaling": "weak".
"""
import importlib
import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC = "SPADEGenerator4 images/sec (256x256, batch 16 per GPU)"
FLOP_PER_IMAGE = 305.23e9      # algorithmic forward FLOPs per image (measured with hooks on the reference)
BATCH = 16


def _config(n):
    return {"workload": "BASELINE configs[3]: SPADEGenerator4(semantic_nc=41, target_nc=3, nz=256, ngf=64, 'spectralspadelayer3x3', crop 256, 'normal'), "
                        "batch %d per GPU, eval, seeded random-init weights, synthetic depth + 40-class Voronoi one-hot input" % BATCH,
            "global_batch": BATCH * n, "parallelism": "dp%d (image-sharded, weights replicated, no collective)" % n,
            "l2": "per-step activations (up to 537 MB per tensor) far exceed the 126 MB L2; no extra flush needed"}


def _model(dev=None, ngf=64, crop=256):
    spade = importlib.import_module("sln_b200.models.SPADE_related")
    torch.manual_seed(0)
    m = spade.SPADEGenerator4(semantic_nc=41, target_nc=3, nz=256, ngf=ngf, norm='spectralspadelayer3x3', crop_size=crop, n_up='normal').eval()
    return m.to(dev) if dev is not None else m


def run(args):
    import bench as B
    from oracle import spade_oracle as so   # synthetic input generator only (shared with the tests)
    rank, local_rank, world = B.dist_env()
    n = max(args.gpus, 1)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib = importlib.import_module("sln_b200._lib")
    lib = _lib.load()
    m = _model(dev)
    seg_h = so.synthetic_input(BATCH, S=256, seed=100 + rank).pin_memory()
    z_h = torch.randn(BATCH, 256, generator=torch.Generator().manual_seed(7 + rank)).pin_memory()
    seg, z = seg_h.to(dev), z_h.to(dev)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    n0 = lib.sln_launch_count()
    out = m(seg, z)
    launches = int(lib.sln_launch_count() - n0)
    for _ in range(max(args.warmup, 3) - 1):
        out = m(seg, z)
    barrier()
    sampler = B.ClockSampler(local_rank).start() if rank == 0 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in evs:
        e0.record()
        out = m(seg, z)
        e1.record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    out_h = torch.empty(BATCH, 3, 256, 256).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o = m(seg_h.to(dev, non_blocking=True), z_h.to(dev, non_blocking=True))
        out_h.copy_(o, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    prof = {}
    if rank == 0:
        import ctypes
        lib.sln_prof_enable(1)
        m(seg, z)
        torch.cuda.synchronize(dev)
        for ci, cname in enumerate(B.PROF_CLASSES):
            ms, work, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
            _lib.check(lib.sln_prof_read(ci, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(cnt)), "prof_read")
            if cnt.value:
                prof[cname] = dict(ms=ms.value, work=work.value, launches=cnt.value)
        lib.sln_prof_enable(0)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dev_ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = t.tolist()
        dist.barrier()
    if rank != 0:
        return None
    peaks = B.measured_peaks()
    ms_per_step = dev_ms / args.steps
    conv = prof.get("spade_conv")
    roofline = None
    if conv:
        tf = conv["work"] / (conv["ms"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "tc_gemm_kernel<Im2col, MatView, TcEpiStore|TcEpiSpade> (3xTF32 implicit-GEMM convolutions)",
                    "achieved": tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops_sustained"],
                    **B.profile_traffic("tc_spade", "tc_gemm_kernel", pick="max"),
                    "algorithmic_flops_per_step": conv["work"], "launches_per_step": conv["launches"], "ms_per_step": conv["ms"],
                    "share_of_step": conv["ms"] / max(sum(v["ms"] for v in prof.values()), 1e-9),
                    "peak_source": peaks["source"] + ", sustained bf16",
                    "note": "useful (algorithmic) FLOPs; the 3xTF32 scheme issues 3x as many TF32 MMAs, and the TF32 MMA rate is half the "
                            "bf16 rate, so 1/6 of the bf16 peak (%.0f TFLOP/s) is the ceiling of this arithmetic" % (peaks["bf16_tflops_sustained"] / 6.0)}
    cpu = None
    if n == 1 and not args.no_cpu_baseline:
        cpu = cpu_spade_baseline(budget_s=25.0)
    eager, batch1 = None, None
    if n == 1:
        try:
            eager = eager_gpu_spade(dev, m, seg, z)
            batch1 = batch1_latency(dev, m, seg, z)
        except Exception as e:       # comparators only: never fail the bench line because of them
            eager = {"unavailable": repr(e)[:300]}
    return {"eager_gpu_baseline": eager, "batch1": batch1, "metric": METRIC, "value": BATCH * n * 1e3 / ms_per_step, "unit": "images/s", "n_gpus": n, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(n),
            "e2e": {"value": BATCH * n * args.steps / e2e_s, "unit": "images/s", "h2d_bytes_per_step": seg_h.numel() * 4 + z_h.numel() * 4,
                    "d2h_bytes_per_step": out_h.numel() * 4},
            "gpu_launches": launches * args.steps, "launches_per_step": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "kernel_classes_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
            "useful_tflops": FLOP_PER_IMAGE * BATCH * n / (ms_per_step * 1e-3) / 1e12}


def batch1_latency(dev, m, seg, z, reps=10):
    """What testing/test_SPADE_shade.py:77-79 actually runs: 50 z draws at batch 1.  ms per image of this path at batch 1."""
    s1, z1 = seg[:1].contiguous(), z[:1].contiguous()
    for _ in range(3):
        m(s1, z1)
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        m(s1, z1)
    b.record()
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / reps
    out = {"ms_per_image": ms, "images_per_s": 1e3 / ms, "what": "batch 1 (the reference's own call pattern, test_SPADE_shade.py:77-79), CUDA events, %d forwards" % reps}
    try:     # the same forward as one CUDA graph (models/SPADE_related.py GraphedForward): launch-bound at batch 1
        import importlib
        run = importlib.import_module("sln_b200.models.SPADE_related").GraphedForward(m, s1, z1)
        for _ in range(3):
            run(s1, z1)
        torch.cuda.synchronize(dev)
        a.record()
        for _ in range(reps):
            run(s1, z1)
        b.record()
        torch.cuda.synchronize(dev)
        out["graphed_ms_per_image"] = a.elapsed_time(b) / reps
    except Exception as e:      # context only
        out["graphed_ms_per_image"] = None
        out["graphed_error"] = repr(e)[:200]
    return out


def eager_gpu_spade(dev, m, seg, z, reps=3):
    """Comparators on the SAME GPU: the oracle port (the reference's forward as plain torch ops -> stock cuDNN / cuBLAS kernels, what
    SPADEGenerator4.forward executes after .cuda()) with cudnn.allow_tf32 False (fp32 parity class of this path) and True (torch's
    default for convolutions: faster, ~1e-3 relative error), at batch 16 and batch 1.  Baselines only."""
    from oracle import spade_oracle as so
    sd = {k: v.detach().to(dev) for k, v in m.state_dict().items()}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    out = {}
    try:
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for nb in (BATCH, 1):
                s_, z_ = seg[:nb].contiguous(), z[:nb].contiguous()
                with torch.no_grad():
                    for _ in range(2):
                        so.forward(sd, s_, z_, 64, 8, torch.float32)
                    torch.cuda.synchronize(dev)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(reps):
                        so.forward(sd, s_, z_, 64, 8, torch.float32)
                    b.record()
                    torch.cuda.synchronize(dev)
                ms = a.elapsed_time(b) / reps
                out["tf32_%s_batch%d" % ("on" if tf32 else "off", nb)] = {"ms_per_step": ms, "images_per_s": nb * 1e3 / ms,
                                                                          "useful_tflops": FLOP_PER_IMAGE * nb / (ms * 1e-3) / 1e12}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    out["kind"] = "port, torch eager (cuDNN/cuBLAS) on the same GPU; tf32_off = fp32 arithmetic class of this path, tf32_on = torch's conv default"
    return out


def cpu_spade_baseline(budget_s=25.0):
    """The oracle port (oracle/spade_oracle.py: the reference's forward as plain torch CPU ops, fp32, all host threads) on a bounded
    sample: batch 1 (the reference's own call pattern, test_SPADE_shade.py:77-79), as many forwards as fit the budget."""
    from oracle import spade_oracle as so
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    m = _model()
    sd = m.state_dict()
    seg = so.synthetic_input(1, S=256, seed=100)
    z = torch.randn(1, 256, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        so.forward(sd, seg, z, 64, 8, torch.float32)
        t0 = time.perf_counter()
        k = 0
        while k < 1 or (time.perf_counter() - t0) < budget_s * 0.6:
            so.forward(sd, seg, z, 64, 8, torch.float32)
            k += 1
        dt = (time.perf_counter() - t0) / k
    return {"value": 1.0 / dt, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": "%d forwards at batch 1 (of the 16-image step), fp32 torch CPU ops (oracle/spade_oracle.py)" % k}


def run_reference(args):
    cpu = cpu_spade_baseline(budget_s=60.0)
    n = max(args.gpus, 1)
    return {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "images/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": BATCH * 1e3 / cpu["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": _config(n), "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
