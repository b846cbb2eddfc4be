#!/usr/bin/env python
"""bench.py — the measurement contract of the 3D_SLN B200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload vae|render|spade] [--impl reference]

Default workload = BASELINE.json configs[1]: the VAE-graph train step (reference train.py:69-84: forward, losses, backward,
Adam) on synthetic SUNCG-shaped scene graphs, batch 64 x 32 nodes per GPU (O=2048, T=3968), E=64, BatchNorm MLPs, fp32.
One "step" = one train step over one batch.  N>1 (torchrun, one rank per GPU): every rank steps its own 64 scenes and the
flat gradient arena is all-reduced over NCCL (weak scaling; BatchNorm statistics stay per rank = DDP semantics).

Printed JSON (rank 0, one line):
  value        scene-graphs/s, inputs resident in HBM, CUDA-event time of K steps, max over ranks
  e2e          same metric through the public API (VAETrainStep.step(host batch) + loss read-back) with pinned-host inputs
  roofline     dominant kernel class, timed live with CUDA events around every launch of an un-graphed step
  cpu_baseline the oracle port (oracle/vae_oracle.py, plain torch CPU ops restating the reference) on this box's host cores
  --impl reference : that CPU path alone, on the same config/metric (the reference itself is Python and cannot travel to
                     the GPU box: /root/reference only exists in the build container).
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

SCENES_PER_GPU = 64
NODES_PER_SCENE = 32
FP32_SIMT_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 148 SMs x 128 FMA lanes x 2 flop x max SM clock (nominal)
PROF_CLASSES = ["gemm_fwd", "gemm_bwd_x", "gemm_bwd_w", "pool", "prep", "misc", "raster_fwd", "raster_bwd", "spade_conv", "spade_misc"]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def profile_traffic(tag, kernel_substr, pick="mean"):
    """DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel, READ from the newest committed ncu summary
    profiles/r<NN>_prof_<tag>.csv (first row = metric names, second row = units) — not a literal typed into this file, so it cannot go
    stale silently: the source file is named next to the number, and a missing capture yields null."""
    import csv
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_prof_%s.csv" % tag)),
                   key=lambda f: int(re.search(r"r(\d+)_prof_", os.path.basename(f)).group(1)))
    if not files:
        return {"traffic": None, "traffic_source": "no ncu capture profiles/r*_prof_%s.csv committed" % tag}
    path = files[-1]
    try:
        rows = list(csv.reader(open(path)))
        head, units = rows[0], rows[1]
        ir, iw, ik = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum"), head.index("Kernel Name")
        vals = []
        for r in rows[2:]:
            if len(r) > max(ir, iw) and kernel_substr in r[ik]:
                vals.append(float(r[ir]) * _UNIT.get(units[ir], 1.0) + float(r[iw]) * _UNIT.get(units[iw], 1.0))
        if not vals:
            return {"traffic": None, "traffic_source": "%s holds no launch of %s" % (os.path.relpath(path, ROOT), kernel_substr)}
        t = max(vals) if pick == "max" else sum(vals) / len(vals)
        return {"traffic": t, "traffic_source": "%s of dram__bytes_read+write over the %d captured %s launches in %s (ncu --set full)" % (
            pick, len(vals), kernel_substr, os.path.relpath(path, ROOT))}
    except Exception as e:     # a malformed summary must not fail the bench line
        return {"traffic": None, "traffic_source": "could not parse %s: %r" % (os.path.relpath(path, ROOT), e)}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ================================================================================================ CPU reference arm / baseline
def cpu_vae_steps(n_scenes, steps, warmup, budget_s):
    """The reference train-step body (train.py:70-84) restated in oracle/vae_oracle.py, fp32, all host threads.
    Returns (scene-graphs/s, seconds per step, scenes per step actually used, threads)."""
    from oracle import vae_oracle as vo
    syn = importlib.import_module("sln_b200.data.synthetic")
    Model = importlib.import_module("sln_b200.models.Sg2ScVAE_model").Sg2ScVAEModel
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(42)
    m = Model(syn.default_vocab(), embedding_dim=64, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode='feedforward',
              gconv_num_layers=5, mlp_normalization='batch', vec_noise_dim=0, layout_noise_dim=32, use_AE=False)
    sd = vo.leaf_state(m.state_dict(), torch.float32)

    def make(n):
        _, objs, boxes, triples, angles, attrs, _, _ = syn.synthetic_batch(n, NODES_PER_SCENE, seed=42)
        return (objs, triples, boxes, angles, attrs)

    batch = make(n_scenes)
    opt = {}
    gen = torch.Generator().manual_seed(0)
    t0 = time.perf_counter()
    vo.train_step(sd, batch, torch.randn(batch[0].size(0), 64, generator=gen), opt, 1)
    first = time.perf_counter() - t0
    per_step_budget = budget_s / max(steps + warmup, 1)
    used = n_scenes
    if first > per_step_budget and n_scenes > 8:        # bounded sample: fewer scenes per step, same per-scene shape
        used = max(8, int(n_scenes * per_step_budget / first))
        batch = make(used)
    it = 1
    for _ in range(max(warmup - 1, 0)):
        it += 1
        vo.train_step(sd, batch, torch.randn(batch[0].size(0), 64, generator=gen), opt, it)
    t0 = time.perf_counter()
    for _ in range(steps):
        it += 1
        vo.train_step(sd, batch, torch.randn(batch[0].size(0), 64, generator=gen), opt, it)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return used / dt, dt, used, threads


def eager_gpu_vae_steps(dev, n_scenes, steps=10, warmup=3):
    """Baseline only: the oracle's plain-torch restatement of the reference train step (train.py:70-84) run EAGERLY ON THE SAME GPU —
    the launch-per-aten-op execution the reference itself has after model.cuda() (the reference cannot travel to the GPU box)."""
    from oracle import vae_oracle as vo
    syn = importlib.import_module("sln_b200.data.synthetic")
    Model = importlib.import_module("sln_b200.models.Sg2ScVAE_model").Sg2ScVAEModel
    torch.manual_seed(42)
    m = Model(syn.default_vocab(), embedding_dim=64, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode='feedforward',
              gconv_num_layers=5, mlp_normalization='batch', vec_noise_dim=0, layout_noise_dim=32, use_AE=False)
    sd = vo.leaf_state(m.state_dict(), torch.float32, device=dev)
    _, objs, boxes, triples, angles, attrs, _, _ = syn.synthetic_batch(n_scenes, NODES_PER_SCENE, seed=42)
    batch = tuple(t.to(dev) for t in (objs, triples, boxes, angles, attrs))
    opt = {}
    for it in range(warmup):
        vo.train_step(sd, batch, torch.randn(batch[0].size(0), 64, device=dev), opt, it + 1)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for it in range(steps):
        vo.train_step(sd, batch, torch.randn(batch[0].size(0), 64, device=dev), opt, warmup + it + 1)   # returns python floats: syncs like train.py:78
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / steps
    return {"value": n_scenes / dt, "unit": "scene-graphs/s", "ms_per_step": dt * 1e3, "kind": "port, torch eager on the same GPU (fp32, TF32 matmul off)",
            "sample": "%d scenes x %d nodes, %d timed steps of oracle/vae_oracle.py train_step on cuda" % (n_scenes, NODES_PER_SCENE, steps)}


def scatter_sweep(dev, lib, _lib, syn, hbm_peak, sizes=(512, 8192), reps=7):
    """The north-star "scatter" stage (sln_gconv_pool_fwd = graph.py:92-108) alone at batch sizes where it is a bandwidth problem:
    algorithmic bytes / CUDA-event time of one launch, L2 flushed (and cleaned by a read pass) before every launch."""
    H, D = 256, 128
    st = _lib.cur_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    out = []
    for B in sizes:
        _, objs, _, triples, _, _, _, _ = syn.synthetic_batch(B, NODES_PER_SCENE, seed=1)
        O, T = objs.size(0), triples.size(0)
        edges = triples[:, [0, 2]].contiguous().to(dev)
        x = torch.randn(T, 2 * H + D, device=dev)
        pooled = torch.empty(O, H, device=dev)
        ws = torch.empty(lib.sln_gconv_pool_workspace_bytes(O, T), dtype=torch.uint8, device=dev)
        _lib.check(lib.sln_csr_build(edges.data_ptr(), 2, O, T, ws.data_ptr(), ws.numel(), st), "csr_build")
        evs = []
        for i in range(reps + 2):
            flush.zero_(); flush_rd.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.sln_gconv_pool_fwd(x.data_ptr(), O, T, H, D, pooled.data_ptr(), ws.data_ptr(), ws.numel(), st), "pool_fwd")
            b.record()
            if i >= 2:
                evs.append((a, b))
        torch.cuda.synchronize(dev)
        ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
        nbytes = 4.0 * (2.0 * T * H + 2.0 * T + 2.0 * O + 1.0 + O * H)
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"scenes": B, "O": O, "T": T, "algorithmic_bytes": nbytes, "us": ms * 1e3, "achieved": gbs, "unit": "GB/s", "frac": gbs / hbm_peak})
        del x, pooled, ws, edges
    return out


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    n = max(args.gpus, 1)
    if args.workload != "vae":
        mod = importlib.import_module("bench_%s" % args.workload)
        print(json.dumps(mod.run_reference(args)), flush=True)
        return
    scenes = SCENES_PER_GPU * n
    val, dt, used, threads = cpu_vae_steps(scenes, args.steps, args.warmup, budget_s=150.0)
    sample = "%d scenes x %d nodes per step (%s of the %d-scene workload), %d steps, fp32 torch CPU ops" % (
        used, NODES_PER_SCENE, "all" if used == scenes else "a bounded sample", scenes, args.steps)
    line = {
        "impl": "reference", "metric": "scene-graphs/sec VAE train step (batch64, 32obj)", "value": val, "unit": "scene-graphs/s",
        "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * scenes / used, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": vae_config(n),
        "cpu_baseline": {"value": val, "unit": "scene-graphs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "scene-graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def vae_config(n, bn_policy="local"):
    bn_note = "per-rank BatchNorm statistics" if bn_policy == "local" or n == 1 else "SyncBatchNorm inside the finalising kernels over NVLink peer memory"
    return {"workload": "BASELINE configs[1]: VAE-graph train step, %d scenes x %d nodes per GPU (O=%d, T=%d per GPU), embedding_dim=64, "
                        "5+5 GraphTripleConv layers, mlp_normalization=batch, Adam lr 1e-4" % (
                            SCENES_PER_GPU, NODES_PER_SCENE, SCENES_PER_GPU * NODES_PER_SCENE, SCENES_PER_GPU * (NODES_PER_SCENE - 1) * 2),
            "global_batch_scenes": SCENES_PER_GPU * n, "parallelism": "dp%d (scene-sharded; NCCL all-reduce of the 15.5 MB gradient arena in two buckets, the decoder's in flight "
                                                                     "during the encoder's backward pass, captured inside the step graph; %s)" % (
                                                                         n, bn_note),
            "l2": "L2 flushed (256 MiB write) before every timed step"}


# ================================================================================================ VAE train step on the GPU
def run_vae(args):
    rank, local_rank, world = dist_env()
    n = max(args.gpus, 1)
    if world != n and world > 1:
        n = world
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    pg = None
    if world > 1:
        import torch.distributed as dist
        pg = dist.group.WORLD              # created by main()
    _lib = importlib.import_module("sln_b200._lib")
    lib = _lib.load()
    syn = importlib.import_module("sln_b200.data.synthetic")
    Model = importlib.import_module("sln_b200.models.Sg2ScVAE_model").Sg2ScVAEModel
    sutils = importlib.import_module("sln_b200.utils")

    torch.manual_seed(42)   # reference options/options.py:59 — identical initial weights on every rank
    model = Model(syn.default_vocab(), embedding_dim=64, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode='feedforward',
                  gconv_num_layers=5, mlp_normalization='batch', vec_noise_dim=0, layout_noise_dim=32, use_AE=False).float().to(dev).train()
    _, objs, boxes, triples, angles, attrs, _, _ = syn.synthetic_batch(SCENES_PER_GPU, NODES_PER_SCENE, seed=42 + rank)
    host = [t.pin_memory() for t in (objs, triples, boxes, angles, attrs)]
    O, T = objs.size(0), triples.size(0)
    # e2e input: the same scenes as un-collated samples, packed into one pinned wire buffer (what a DataLoader worker hands over)
    collate = importlib.import_module("sln_b200.data.collate")
    wire, wire_meta = collate.packed_batch(syn.synthetic_samples(SCENES_PER_GPU, NODES_PER_SCENE, seed=42 + rank), lib)
    h2d = int(wire_meta[4][9])
    step = sutils.VAETrainStep(model, O, T, lr=1e-4, kl_weight=0.1, use_graph=True, process_group=pg, world_size=world, wire_meta=wire_meta,
                               bn_policy=args.bn_policy if world > 1 else "local")
    step.load_batch(host)
    n0 = lib.sln_launch_count()
    step._fwd_bwd(); step._allreduce(); step._opt()
    launches_per_step = int(lib.sln_launch_count() - n0)
    step.capture()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    losses_host = torch.empty(4, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step.step(host)
    barrier()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    # ---- value: device-resident inputs, CUDA events around each step, L2 flushed between steps
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in evs:
        flush.zero_()
        a.record()
        step.run()
        b.record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # ---- e2e: public API with pinned-host inputs: H2D of the (wire-packed) batch, device batch assembly, the step, D2H of the losses, every step
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step.step_wire(wire)     # one H2D copy of the packed batch + device-side batch assembly + the step
        losses_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    final_losses = losses_host.tolist()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dev_ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = t.tolist()
    # ---- roofline: one un-graphed step with an event pair around every launch of the library
    prof = None
    sync_bn = world > 1 and args.bn_policy == "sync"
    if rank == 0 or sync_bn:     # under SyncBatchNorm the kernels of every rank wait for each other: all ranks run the pass
        barrier_local = lambda: torch.cuda.synchronize(dev)   # noqa: E731
        rows = {}
        reps = 3
        lib.sln_prof_enable(1 if rank == 0 else 0)
        step.overlap_allreduce = False       # no NCCL collective is issued from this pass (rank 0 may be alone in it)
        for _ in range(reps):
            flush.zero_()
            step._fwd_bwd(); step._opt()
        barrier_local()
    if rank == 0:
        total_ms = 0.0
        for ci, cname in enumerate(PROF_CLASSES):
            ms, work, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
            _lib.check(lib.sln_prof_read(ci, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(cnt)), "prof_read")
            if cnt.value:
                rows[cname] = dict(ms=ms.value / reps, work=work.value / reps, launches=cnt.value // reps)
                total_ms += ms.value / reps
        lib.sln_prof_enable(0)
        prof = dict(rows=rows, total_ms=total_ms)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank != 0:
        return None
    peaks = measured_peaks()
    scenes = SCENES_PER_GPU * n
    ms_per_step = dev_ms / args.steps
    value = scenes / (ms_per_step * 1e-3)
    e2e = scenes * args.steps / e2e_s
    gemm = {k: v for k, v in prof["rows"].items() if k.startswith("gemm")}
    g_ms = sum(v["ms"] for v in gemm.values()); g_fl = sum(v["work"] for v in gemm.values()); g_n = sum(v["launches"] for v in gemm.values())
    # the event-pair pass runs un-graphed (host launch gaps inflate every interval), so it only provides the contraction kernels' SHARE
    # of the step; their time inside the timed (graph-replayed) region = share x ms_per_step
    share = g_ms / prof["total_ms"] if prof["total_ms"] else 0.0
    achieved_events = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    achieved = g_fl / (share * ms_per_step * 1e-3) / 1e12 if share > 0 else 0.0
    tf32_ceiling = peaks["bf16_tflops_sustained"] / 6.0   # TF32 MMA rate = 1/2 bf16; 3xTF32 issues 3 MMAs per useful product
    roofline = {
        "bound": "tensor", "kernel": "tc::tc_gemm_kernel (tcgen05 kind::tf32 3xTF32 contraction of the graph-conv MLPs: fwd + bwd-data + bwd-weight)",
        "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops_sustained"],
        **profile_traffic("tc_vae", "tc_gemm_kernel"),
        "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a step)",
        "algorithmic_flops_per_launch": g_fl / max(g_n, 1), "launches_per_step": g_n, "avg_launch_us": g_ms * 1e3 / max(g_n, 1),
        "share_of_step": share, "achieved_event_pairs_ungraphed": achieved_events,
        "note": "achieved = useful (algorithmic) 2MNK FLOPs of all contraction launches of a step / (their share of the step x ms_per_step); the "
                "share comes from CUDA-event pairs around every launch of an un-graphed step on the launching streams (weight-gradient launches "
                "overlap the chain on a side stream, so shares are of summed kernel time). fp32-parity "
                "arithmetic is 3xTF32: its ceiling is bf16_peak/6 = %.0f TFLOP/s -> frac of that ceiling %.3f; nominal FP32-SIMT peak %.1f TFLOP/s" % (
                    tf32_ceiling, achieved / tf32_ceiling, FP32_SIMT_PEAK_TFLOPS),
    }
    pool = prof["rows"].get("pool")
    roofline_scatter = None
    if pool:
        gbs = pool["work"] / (pool["ms"] * 1e-3) / 1e9
        t64 = profile_traffic("pool64", "k_pool_fwd")      # the warp-per-node kernel captured inside a configs[1] step
        roofline_scatter = {"bound": "hbm", "kernel": "k_pool_fwd_node / k_pool_fwd (scatter_add+count+divide as a CSR gather-reduce)", "achieved": gbs,
                            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": t64["traffic"],
                            "traffic_source": t64["traffic_source"],
                            "traffic_large_batch": profile_traffic("pool", "k_pool_fwd"),
                            "algorithmic_bytes_per_launch": pool["work"] / pool["launches"], "avg_launch_us": pool["ms"] * 1e3 / pool["launches"],
                            "note": "10.3 MB per launch at this config (1.6 us at the HBM peak): latency-bound; large_batch = the same entry point alone at "
                                    "512 / 8192 scenes, where it is a bandwidth kernel (ncu: profiles/*_prof_pool.csv)"}
    if roofline_scatter is not None:
        try:
            roofline_scatter["large_batch"] = scatter_sweep(dev, lib, _lib, syn, peaks["hbm_gbs"])
        except Exception as e:   # context only: never fail the bench line because of it
            roofline_scatter["large_batch"] = {"unavailable": repr(e)[:200]}
    cpu = None
    if n == 1 and not args.no_cpu_baseline:
        val, dt, used, threads = cpu_vae_steps(SCENES_PER_GPU, steps=8, warmup=2, budget_s=25.0)
        cpu = {"value": val, "unit": "scene-graphs/s", "cores": threads, "kind": "port",
               "sample": "%d scenes x %d nodes per step, 8 timed steps, fp32 torch CPU ops (oracle/vae_oracle.py train_step)" % (used, NODES_PER_SCENE)}
    eager = None
    if n == 1 and not args.no_cpu_baseline:
        try:
            eager = eager_gpu_vae_steps(dev, SCENES_PER_GPU)
        except Exception as e:   # baseline only: never fail the bench line because of it
            eager = {"unavailable": repr(e)[:200]}
    line = {
        "metric": "scene-graphs/sec VAE train step (batch64, 32obj)", "value": value, "unit": "scene-graphs/s", "n_gpus": n,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": vae_config(n, args.bn_policy),
        "e2e": {"value": e2e, "unit": "scene-graphs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "roofline": roofline, "roofline_scatter": roofline_scatter, "cpu_baseline": cpu, "eager_gpu_baseline": eager, "clocks": clocks,
        "kernel_classes_ms": {k: round(v["ms"], 4) for k, v in prof["rows"].items()},
        "final_losses": {"bbox": final_losses[0], "angle": final_losses[1], "kld_weighted": final_losses[2], "total": final_losses[3]},
    }
    if n == 1 and not args.no_cpu_baseline:
        for key, fn in (("drop_in_path", lambda: drop_in_path(dev, syn, Model, sutils)), ("config0_fixture", lambda: config0_latency(dev, syn, Model, sutils)),
                        ("eager_gpu_graphed_baseline", lambda: eager_gpu_vae_graphed(dev, SCENES_PER_GPU))):
            try:
                line[key] = fn()
            except Exception as e:       # context only: never fail the bench line because of it
                line[key] = {"unavailable": repr(e)[:300]}
    return line


def _time_loop(dev, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize(dev)
    return (time.perf_counter() - t0) / steps * 1e3


def drop_in_path(dev, syn, Model, sutils, steps=30, warmup=5):
    """ms/step of what an UNCHANGED train.py executes through the import switch: model(...) ->
    calculate_model_losses -> optimizer.zero_grad -> backward -> optimizer.step (train.py:70-84), eager (no CUDA graph, Python between the
    four library calls), with torch.optim.Adam as train.py:15 constructs it and with sln_b200.utils.FusedAdam."""
    import types
    _, objs, boxes, triples, angles, attrs, _, _ = syn.synthetic_batch(SCENES_PER_GPU, NODES_PER_SCENE, seed=42)
    batch = [t.to(dev) for t in (objs, triples, boxes, angles, attrs)]
    out = {}
    for name, make_opt in (("torch.optim.Adam", lambda ps: torch.optim.Adam(ps, lr=1e-4)), ("FusedAdam", lambda ps: sutils.FusedAdam(ps, lr=1e-4))):
        torch.manual_seed(42)
        m = Model(syn.default_vocab(), embedding_dim=64, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode='feedforward',
                  gconv_num_layers=5, mlp_normalization='batch', vec_noise_dim=0, layout_noise_dim=32, use_AE=False).float().to(dev).train()
        opt = make_opt(m.parameters())
        ns = types.SimpleNamespace(use_AE=False)

        def step():
            mu, logvar, bp, ap = m(batch[0], batch[1], batch[2], batch[3], batch[4], None)
            total, parts = sutils.calculate_model_losses(ns, m, batch[2], bp, batch[3], ap, mu=mu, logvar=logvar, KL_weight=0.1)
            opt.zero_grad()
            total.backward()
            opt.step()
        ms = _time_loop(dev, step, steps, warmup)
        out[name] = {"ms_per_step": ms, "scene_graphs_per_s": SCENES_PER_GPU / (ms * 1e-3)}
    out["what"] = "reference train.py:70-84 body, unchanged caller code, 64 scenes x 32 nodes, BatchNorm, device-resident batch; wall clock incl. the loss .item() syncs"
    return out


def config0_latency(dev, syn, Model, sutils, steps=50, warmup=5):
    """BASELINE configs[0] on the GPU: the reference's embedded 5-object scene graph (testing/test_heatmap.py:41-43), forward only and a
    full train step through the drop-in path (mlp_normalization='none': training-mode BatchNorm needs more than a handful of rows)."""
    import types
    objs, triples, boxes, angles, attrs = [t.to(dev) for t in syn.fixture_graph()]
    torch.manual_seed(42)
    m = Model(syn.default_vocab(), embedding_dim=64, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode='feedforward',
              gconv_num_layers=5, mlp_normalization='none', vec_noise_dim=0, layout_noise_dim=32, use_AE=False).float().to(dev).train()
    opt = sutils.FusedAdam(m.parameters(), lr=1e-4)
    ns = types.SimpleNamespace(use_AE=False)

    def fwd():
        with torch.no_grad():
            m(objs, triples, boxes, angles, attrs, None)

    def step():
        mu, logvar, bp, ap = m(objs, triples, boxes, angles, attrs, None)
        total, _ = sutils.calculate_model_losses(ns, m, boxes, bp, angles, ap, mu=mu, logvar=logvar, KL_weight=0.1)
        opt.zero_grad()
        total.backward()
        opt.step()
    return {"forward_ms": _time_loop(dev, fwd, steps, warmup), "train_step_ms": _time_loop(dev, step, steps, warmup),
            "what": "1 scene, 6 nodes, 9 triples (O=6, T=9), eager drop-in path, wall clock"}


def eager_gpu_vae_graphed(dev, n_scenes, steps=20, warmup=3):
    """Baseline only: the same plain-torch restatement as eager_gpu_vae_steps, but forward + losses + backward + torch.optim.Adam
    (capturable) captured into ONE CUDA graph — the strongest 'stock PyTorch' comparator (aten/cuBLAS kernels, no launch overhead,
    no host syncs).  cuBLAS fp32 (TF32 off), as the reference's defaults."""
    from oracle import vae_oracle as vo
    syn = importlib.import_module("sln_b200.data.synthetic")
    Model = importlib.import_module("sln_b200.models.Sg2ScVAE_model").Sg2ScVAEModel
    torch.manual_seed(42)
    m = Model(syn.default_vocab(), embedding_dim=64, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode='feedforward',
              gconv_num_layers=5, mlp_normalization='batch', vec_noise_dim=0, layout_noise_dim=32, use_AE=False)
    sd = vo.leaf_state(m.state_dict(), torch.float32, device=dev)
    params = [v for v in sd.values() if v.is_floating_point() and v.requires_grad]
    _, objs, boxes, triples, angles, attrs, _, _ = syn.synthetic_batch(n_scenes, NODES_PER_SCENE, seed=42)
    objs, triples, boxes, angles, attrs = [t.to(dev) for t in (objs, triples, boxes, angles, attrs)]
    eps = torch.randn(objs.size(0), 64, device=dev)
    opt = torch.optim.Adam(params, lr=1e-4, capturable=True)

    def body():
        eps.normal_()
        mu, logvar, bp, ap = vo.forward(sd, objs, triples, boxes, angles, attrs, eps, 5, True, False, {})
        total, _ = vo.losses(boxes, bp, angles, ap, mu, logvar, 0.1)
        total.backward()
        opt.step()
        return total
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            opt.zero_grad(set_to_none=True)
            body()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    with torch.cuda.graph(g):
        total = body()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        g.replay()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        flush.zero_()
        a.record(); g.replay(); b.record()
    torch.cuda.synchronize(dev)
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    return {"value": n_scenes / (ms * 1e-3), "unit": "scene-graphs/s", "ms_per_step": ms, "loss": float(total),
            "kind": "port, plain torch ops (aten + cuBLAS fp32, TF32 off) + torch.optim.Adam(capturable) replayed as one CUDA graph on the same GPU",
            "sample": "%d scenes x %d nodes, %d timed replays, L2 flushed before each, CUDA events" % (n_scenes, NODES_PER_SCENE, steps)}


def sub_record(workload, args, steps, warmup):
    """A secondary workload's bench line (BASELINE's second metric 'diff-render iters/sec', configs[3] SPADE) as a sub-record of the default
    line, so the driver's one `bench.py --gpus N` run covers all three hot paths; their CPU legs are skipped here (time budget) — run
    `bench.py --workload render|spade` for the full stand-alone line."""
    sub = argparse.Namespace(**vars(args))
    sub.steps, sub.warmup, sub.no_cpu_baseline, sub.workload = steps, warmup, True, workload
    try:
        return importlib.import_module("bench_%s" % workload).run(sub)
    except Exception as e:
        rank = dist_env()[0]
        if dist_env()[2] > 1:
            raise                      # a rank that fails alone would dead-lock the others at the next barrier
        return {"unavailable": repr(e)[:300]} if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300, help="timed steps (default sized for a >= 1 s timed region of the VAE step)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="vae", choices=["vae", "render", "spade"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bn-policy", default="local", choices=["local", "sync"],
                    help="N > 1: per-rank BatchNorm statistics (P2, default) or SyncBatchNorm inside the kernels over NVLink (P1)")
    ap.add_argument("--no-extra", action="store_true", help="vae workload: skip the render / spade sub-records")
    ap.add_argument("--no-graph", action="store_true", help="render workload: run the refinement iteration eagerly instead of as a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.workload == "vae":
            line = run_vae(args)
            if not args.no_extra:
                # BASELINE.json's second metric and configs[3], measured in the same driver-visible run
                r = sub_record("render", args, steps=200, warmup=5)
                sp = sub_record("spade", args, steps=5, warmup=3)
                if line is not None:
                    line["render"], line["spade"] = r, sp
        else:
            line = importlib.import_module("bench_%s" % args.workload).run(args)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
