"""Turn the ncu artefacts of one GPU visit into the committed summaries under profiles/.

    python tools/ncu_summary.py <tag> [round-prefix]

    gpurun_out/<tag>_launches.csv      -> profiles/<prefix>_launches_<tag>.txt  (per-kernel launch count, total, share of the step)
    gpurun_out/<tag>_prof_bench.ncu-rep -> profiles/<prefix>_ncu_full_<tag>.txt (key metrics per captured kernel)
                                          profiles/traffic.json                 (dram bytes per launch, read by bench.py)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
prefix = sys.argv[2] if len(sys.argv) > 2 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
]


def launches():
    src = os.path.join(G, f"{tag}_launches.csv")
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 3   (tag {tag})",
           "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes", f"# total kernel time {tot/1e3:.2f} ms",
           f"{'total us':>12} {'launches':>8} {'us/launch':>10} {'share':>7}  kernel"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{a[1]:12.1f} {a[0]:8d} {a[1]/a[0]:10.1f} {a[1]/tot*100:6.1f}%  {k[:110]}")
    open(os.path.join(P, f"{prefix}_launches_{tag}.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:14]))


def full(kind="bench"):
    rep = os.path.join(G, f"{tag}_prof_{kind}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    what = "kernels of one bench.py step" if kind == "bench" else "verification-stage kernels of tools/prof_verify.py (296 pairs x 4096 correspondences, 1024 hypotheses)"
    out = [f"# ncu --set full --clock-control none --import-source on : {what} (tag {tag})"]
    traffic = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        out.append(f"--- {name[:120]}")
        vals = {}
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k:
                    out.append(f"    {k} [{units[i]}] = {r[i]}")
                    vals[k] = (r[i], units[i])
        try:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd, ru = vals["dram__bytes_read.sum"]
            wr, wu = vals["dram__bytes_write.sum"]
            short = next(n for n in ("match_tc_kernel", "refine_kernel", "ransac_f_kernel", "filter_kernel") if n in name)
            traffic[short] = float(rd.replace(",", "")) * scale[ru] + float(wr.replace(",", "")) * scale[wu]
        except (KeyError, StopIteration, ValueError):
            pass
    open(os.path.join(P, f"{prefix}_ncu_{'full' if kind == 'bench' else kind}_{tag}.txt"), "w").write("\n".join(out) + "\n")
    if traffic and kind == "bench":
        traffic["source"] = f"profiles/{prefix}_ncu_full_{tag}.txt (dram__bytes_read.sum + dram__bytes_write.sum per launch, one bench.py step)"
        json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print("\n".join(out))


launches()
full()
full("verify")
