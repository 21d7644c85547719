"""gpurun_out/ artefacts of tools/profile_round2.sh -> committed summaries under profiles/ (round 2):
launch list with each kernel's share, the `ncu --set full` table of the bench kernels, per-config DRAM traffic tables and
profiles/traffic.json (read back by bench.py's roofline.traffic).  Usage: python tools/profile_collect2.py r02"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}


def short(name):
    return re.sub(r"void |<unnamed>::|\(anonymous namespace\)::|\(.*", "", name)


def metric_rows(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    ix = {h: i for i, h in enumerate(rows[hi])}
    out = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= ix["Metric Value"]:
            continue
        d = out.setdefault(r[ix["ID"]], {"kernel": short(r[ix["Kernel Name"]]), "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
        d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * MULT.get(r[ix["Metric Unit"]], 1)
    return list(out.values())


# ---- launch list ----------------------------------------------------------------------------------------------
lp = os.path.join(G, f"launches_{R}.csv")
if os.path.exists(lp):
    agg = collections.OrderedDict()
    launches = metric_rows(lp)
    for d in launches:
        a = agg.setdefault(d["kernel"], [0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
    total = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("k_") and "probe" not in k)
    with open(os.path.join(P, f"{R}_launches_bench.md"), "w") as f:
        f.write(f"# ncu launch list, `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra` ({R})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and serialised; what must "
                "agree with bench.py is each kernel's SHARE of the engine's launches.\n\n"
                "| kernel | launches | total us | avg us | share of all GPU time | share of engine kernels |\n|---|---|---|---|---|---|\n")
        for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            sh2 = f"{100 * ns / ours:.1f} %" if k.startswith("k_") and "probe" not in k and ours else "-"
            f.write(f"| `{k[:90]}` | {c} | {ns / 1e3:.1f} | {ns / c / 1e3:.1f} | {100 * ns / total:.1f} % | {sh2} |\n")
        f.write("\n(torch kernels are the synthetic-input generator and the round-trip check outside the timed region; "
                "k_probe_dfma is the FP64 roof measurement bench.py makes once per run, outside the timed region: not an engine kernel.)\n")
    with open(os.path.join(P, f"{R}_launches_bench.csv"), "w") as f:
        f.write("kernel,grid,block,duration_ns\n")
        for d in launches:
            f.write(f"{d['kernel'].replace(',', ' ')},{d['grid'].replace(',', ' ')},{d['block'].replace(',', ' ')},{d.get('gpu__time_duration.sum', 0):.0f}\n")

# ---- full capture of the bench kernels --------------------------------------------------------------------------
rep = os.path.join(G, f"prof_{R}_bench.ncu-rep")
if os.path.exists(rep):
    md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{R}_bench_ncu_full.md"), "w") as f:
        f.write(f"# `ncu --set full --clock-control none` -- default bench workload (4096 x 4096, db4, J = 4, PERIODIC) ({R})\n\n{md}\n"
                "\nColumns: DRAM bytes are `dram__bytes_read.sum` / `dram__bytes_write.sum` per launch; fp64 pipe = "
                "`sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active`; issue = `sm__issue_active.avg.pct_of_peak_sustained_elapsed`; "
                "stalls = `smsp__average_warps_issue_stalled_*_per_issue_active.ratio`.\n")

rep = os.path.join(G, f"prof_{R}_coif5.ncu-rep")
if os.path.exists(rep):
    md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{R}_coif5_ncu_full.md"), "w") as f:
        f.write(f"# `ncu --set full --clock-control none` -- one 2^26-sample signal, coif5, J = 10, PERIODIC (config #4's kernels at 1/4 length) ({R})\n\n"
                "Levels 1-2: direct-form tile kernels; levels 3-10: lattice column kernels, two levels per launch "
                f"(`csrc/vw_column.cu`, `csrc/vw_lattice.cu`).\n\n{md}\n")

# ---- per-config traffic ------------------------------------------------------------------------------------------
traffic_path = os.path.join(P, "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
traffic["_meaning"] = ("DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum, ncu) -- for the bench default workload: of ONE "
                       "launch of the dominant kernel (the analysis tile kernel); for every other key: of ALL engine kernels of one "
                       "forward + inverse (keys c5_* and *denoise*: see the table in profiles/" + R + "_traffic.md)")
tables = []
for fn in sorted(os.listdir(G)):
    m = re.match(rf"traffic_{R}_(.+)\.csv$", fn)
    if not m:
        continue
    key = m.group(1)
    ks = [d for d in metric_rows(os.path.join(G, fn)) if d["kernel"].startswith("k_")]
    if not ks:
        continue
    tot = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in ks)
    tables.append((key, ks, tot))
    if key == "batch4096x4096_db4_J4":
        an = [d for d in ks if "analysis" in d["kernel"]][0]
        traffic[key] = an.get("dram__bytes_read.sum", 0) + an.get("dram__bytes_write.sum", 0)
        traffic[key + "_step"] = tot
    elif key.startswith("c5_"):
        # launches: forward, inverse, then the denoise call (forward, select, inverse)
        nfi = len(ks)
        traffic[key] = None
        traffic[key + "_all_launches"] = tot
    else:
        traffic[key] = tot
if tables:
    with open(os.path.join(P, f"{R}_traffic.md"), "w") as f:
        f.write(f"# DRAM traffic and duration of every engine kernel of one forward + inverse per BASELINE config ({R})\n\n"
                "`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none` on "
                "`tools/prof_once.py --warm 0` (first launches: cold caches, serialised).  c5_* rows also contain the SWT denoise call "
                "(decompose, universal-threshold selection, reconstruct with threshold-on-load) after the plain forward + inverse.\n")
        for key, ks, tot in tables:
            f.write(f"\n## {key}: {tot / 1e6:.0f} MB over {len(ks)} launches\n\n| kernel | grid x block | us | dram rd MB | dram wr MB | GB/s |\n|---|---|---|---|---|---|\n")
            for d in ks:
                rd, wr, ns = d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0), d.get("gpu__time_duration.sum", 1)
                f.write(f"| `{d['kernel'][:70]}` | {d['grid']} x {d['block']} | {ns / 1e3:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / ns:.0f} |\n")
# split the c5 captures into plain and denoise parts: kernels before the first k_select / k_nonfinite belong to fwd + inv
for key, ks, tot in tables:
    if key.startswith("c5_"):
        cut = next((i for i, d in enumerate(ks) if "select" in d["kernel"] or "universal" in d["kernel"]), None)
        if cut is not None:
            # the denoise call starts with its own forward: as many launches as the plain forward had
            nf = sum(1 for d in ks[:cut] if "analysis" in d["kernel"]) // 2
            first_den = next(i for i in range(cut - 1, -1, -1) if "analysis" not in ks[i]["kernel"]) + 1
            plain = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in ks[:first_den])
            den = tot - plain
            traffic[key] = plain
            traffic["c5_denoise_" + key[3:]] = den
            traffic.pop(key + "_all_launches", None)
json.dump(traffic, open(traffic_path, "w"), indent=1)

for src, dst in ((f"bench_{R}.json", f"{R}_bench_line.json"), (f"bench_ref_{R}.json", f"{R}_bench_reference_line.json")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
print("profiles/:", sorted(os.listdir(P)))
