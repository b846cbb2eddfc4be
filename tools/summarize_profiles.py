"""Turn the ncu outputs that a gpurun call left in gpurun_out/ into the committed summaries under profiles/.
usage: python tools/summarize_profiles.py <round tag, e.g. r1>"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def read_launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    m = re.match(r"(void )?((sln|at)::)?((tc|native)::)?(\w+)", name)
    base = m.group(6) if m else name[:30]
    t = re.search(r"<(.*)>\(", name)
    targs = ""
    if t and base in ("tc_gemm_kernel", "gemm_kernel", "k_prep", "k_skinny_fwd", "k_skinny_bwd_w", "k_to_rgb"):
        targs = re.sub(r"sln::(tc::)?|\((int|bool)\)|\(anonymous namespace\)::", "", t.group(1))
    return base, targs


def launch_table(rows, sep_kernel, fname, title):
    """one step = the launches between the last two occurrences of `sep_kernel`"""
    idx = [i for i, r in enumerate(rows) if short(r["Kernel Name"])[0] == sep_kernel]
    a, b = (idx[-2], idx[-1]) if len(idx) >= 2 else (-1, len(rows) - 1)
    step = rows[a + 1:b + 1]
    with open(os.path.join(OUT, fname + ".csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["#", "kernel", "template", "grid", "block", "gpu__time_duration_us"])
        for i, r in enumerate(step):
            base, targs = short(r["Kernel Name"])
            w.writerow([i, base, targs, r["Grid Size"], r["Block Size"], "%.2f" % (float(r["Metric Value"].replace(",", "")) / 1e3)])
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in step:
        base, targs = short(r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", "")) / 1e3
        agg[base][0] += 1; agg[base][1] += v; tot += v
    lines = ["### %s" % title, "", "%d launches, %.1f us summed (ncu: serialised, cold caches — compare SHARES, not absolutes)" % (len(step), tot), "",
             "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        lines.append("| `%s` | %d | %.1f | %.2f | %.3f |" % (k, v[0], v[1], v[1] / v[0], v[1] / tot))
    return "\n".join(lines) + "\n"


def full_capture(rep, fname, title):
    """rep: an .ncu-rep, or the `ncu -i <rep> --page raw --csv` dump of one (tools/gpu_profile.sh writes those on the GPU box because big
    reports do not fit gpurun's copy-back limit)"""
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    cols = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(os.path.join(OUT, fname + ".csv"), "w") as f:
        w = csv.writer(f)
        w.writerow([c[0] for c in cols]); w.writerow([rows[1][c[1]] for c in cols])
        for r in rows[2:]:
            w.writerow([re.sub(r"sln::(tc::)?|\((int|bool)\)", "", r[c[1]]) if c[0] == "Kernel Name" else r[c[1]] for c in cols])
    lines = ["### %s" % title, "", "`ncu --set full --clock-control none --import-source on` (values per launch; file `%s.csv`)" % fname, "",
             "| kernel | grid | regs | us | DRAM rd MB | DRAM wr MB | L2 MB | tensor pipe % | issue % |", "|---|---|---|---|---|---|---|---|---|"]
    def g(r, k):
        return r[hdr.index(k)] if k in hdr else ""
    def mb(r, k):
        v, u = g(r, k), rows[1][hdr.index(k)] if k in hdr else ""
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            return v
        return "%.2f" % (x * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0))
    def us(r):
        k = "gpu__time_duration.sum"
        v, u = g(r, k), rows[1][hdr.index(k)]
        try:
            return "%.1f" % (float(v.replace(",", "")) * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}.get(u, 1.0))
        except ValueError:
            return v
    for r in rows[2:]:
        t = re.search(r"tc_gemm_kernel<(.*?)>\(", g(r, "Kernel Name"))
        name = re.sub(r"sln::(tc::)?|\((int|bool)\)|\(anonymous namespace\)::", "", t.group(1)) if t else g(r, "Kernel Name")[:40]
        lines.append("| `%s` | %s | %s | %s | %s | %s | %s | %s | %s |" % (name, g(r, "launch__grid_size"), g(r, "launch__registers_per_thread"),
                     us(r), mb(r, "dram__bytes_read.sum"), mb(r, "dram__bytes_write.sum"), mb(r, "lts__t_bytes.sum"),
                     g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")[:6], g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")[:6]))
    return "\n".join(lines) + "\n"


parts = []
for wl, sep in (("vae", "k_adam"), ("render", "k_project_bwd"), ("spade", "k_to_rgb")):
    p = os.path.join(SRC, "launches_%s.csv" % wl)
    if os.path.exists(p):
        parts.append(launch_table(read_launches(p), sep, "%s_launches_%s" % (tag, wl), "%s: one step of `bench.py --workload %s` (launch list)" % (tag, wl)))
p = os.path.join(SRC, "launches_vae_warm.csv")
if os.path.exists(p):
    parts.append(launch_table(read_launches(p), "k_adam", "%s_launches_vae_warm" % tag,
                              "%s: the same VAE step with `--cache-control none` (caches NOT flushed between kernels: closer to the in-graph times)" % tag))
p = os.path.join(SRC, "launches_render_warm.csv")
if os.path.exists(p):
    parts.append(launch_table(read_launches(p), "k_project_bwd", "%s_launches_render_warm" % tag,
                              "%s: the same render iteration with `--cache-control none`" % tag))
for name in sorted(os.listdir(SRC)):
    if name.endswith("_raw.csv"):
        parts.append(full_capture(os.path.join(SRC, name), "%s_%s" % (tag, name[:-8]), "%s: %s.ncu-rep" % (tag, name[:-8])))
    elif name.endswith(".ncu-rep") and not os.path.exists(os.path.join(SRC, name[:-8] + "_raw.csv")):
        parts.append(full_capture(os.path.join(SRC, name), "%s_%s" % (tag, name[:-8]), "%s: %s" % (tag, name)))
with open(os.path.join(OUT, "%s_summary.md" % tag), "w") as f:
    f.write("\n".join(parts))
print("\n".join(parts))
