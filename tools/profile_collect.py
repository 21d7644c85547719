"""Turns the raw artefacts of tools/profile_round.sh (gpurun_out/) into the committed summaries under profiles/:
launch list (kernel share of the step), `ncu --set full` tables, and traffic.json (DRAM bytes per launch of the
dominant kernel, read back by bench.py's roofline.traffic).  Usage: python tools/profile_collect.py r01"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def short(name):
    return re.sub(r"void |<unnamed>::|\(.*", "", name)


# ---- 1. launch list -> per-kernel share of the profiled run ---------------------------------------------------
lp = os.path.join(G, f"launches_{R}.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    ix = {h: i for i, h in enumerate(rows[hi])}
    agg = collections.OrderedDict()
    per_launch = []
    for r in rows[hi + 1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = short(r[ix["Kernel Name"]])
        ns = float(r[ix["Metric Value"]].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        per_launch.append((r[ix["ID"]], k, r[ix["Grid Size"]], r[ix["Block Size"]], ns))
    total = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("k_"))
    with open(os.path.join(P, f"{R}_launches_bench.md"), "w") as f:
        f.write(f"# ncu launch list, `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e` ({R})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and serialised; "
                "what must agree with bench.py is each kernel's SHARE of the engine's launches.\n\n")
        f.write("| kernel | launches | total us | avg us | share of all GPU time | share of engine kernels |\n|---|---|---|---|---|---|\n")
        for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            sh2 = f"{100 * ns / ours:.1f} %" if k.startswith("k_") and ours else "-"
            f.write(f"| `{k[:90]}` | {c} | {ns / 1e3:.1f} | {ns / c / 1e3:.1f} | {100 * ns / total:.1f} % | {sh2} |\n")
        f.write("\n(torch kernels in the list are the synthetic-input generator and the round-trip error check outside the timed region.)\n")
    with open(os.path.join(P, f"{R}_launches_bench.csv"), "w") as f:
        f.write("id,kernel,grid,block,duration_ns\n")
        for t in per_launch:
            f.write(",".join(str(x).replace(",", " ") for x in t) + "\n")

# ---- 2. full captures -> markdown tables + traffic.json --------------------------------------------------------
traffic_path = os.path.join(P, "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for tag, title in ((f"prof_{R}_bench", "default bench workload (4096 x 4096, db4, J = 4, PERIODIC)"),
                   (f"prof_{R}_coif5", "one 2^26-sample signal, coif5, J = 10, PERIODIC (config #4 kernels at 1/4 length)")):
    rep = os.path.join(G, tag + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{tag[5:]}_ncu_full.md"), "w") as f:
        f.write(f"# `ncu --set full --clock-control none` -- {title} ({R})\n\n{md}\n")
        f.write("\nColumns: DRAM bytes are `dram__bytes_read.sum` / `dram__bytes_write.sum` per launch; fp64 pipe = "
                "`sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active`; issue = `sm__issue_active.avg.pct_of_peak_sustained_elapsed`; "
                "stalls = `smsp__average_warps_issue_stalled_*_per_issue_active.ratio`.\n")
    if tag.endswith("_bench"):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        ix = {n: i for i, n in enumerate(rows[0])}
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if "k_fused_analysis" in r[ix["Kernel Name"]]:
                t = sum(float(r[ix[m]].replace(",", "")) * mult[rows[1][ix[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                traffic["batch4096x4096_db4_J4"] = t
                break
json.dump(traffic, open(traffic_path, "w"), indent=1)
print("profiles/:", sorted(os.listdir(P)))

# ---- 4. plain copies: the bench lines and the per-config table ------------------------------------------------
import shutil
for src, dst in ((f"bench_{R}.json", f"{R}_bench_line.json"), (f"bench_ref_{R}.json", f"{R}_bench_reference_line.json"),
                 (f"results_{R}.jsonl", f"{R}_config_results.jsonl")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
