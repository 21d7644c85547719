"""Summarise an .ncu-rep (ncu --set full) into one compact table per kernel launch: duration, DRAM traffic and
throughput, FP64-pipe and issue utilisation, occupancy and the top warp stall reasons.
Usage: python tools/ncu_summary.py report.ncu-rep [--md]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {n: i for i, n in enumerate(hdr)}


def val(r, name, default=float("nan")):
    i = ix.get(name)
    if i is None or r[i] == "":
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


def unit(name):
    i = ix.get(name)
    return units[i] if i is not None else ""


def to_bytes(r, name):
    v, u = val(r, name), unit(name).lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(r, name):
    v, u = val(r, name), unit(name).lower()
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


stall_cols = [n for n in hdr if re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio", n)]
print("| # | kernel | grid x block | us | dram rd MB | dram wr MB | dram GB/s | dram % | fp64 pipe % | issue % | warps act % | regs | smem KB | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = re.sub(r"void |<unnamed>::|\(.*", "", r[ix["Kernel Name"]])
    us = to_us(r, "gpu__time_duration.sum")
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    stalls = sorted(((val(r, c, 0.0), re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active", c).group(1))
                     for c in stall_cols), reverse=True)[:4]
    st = ", ".join(f"{n} {v:.2f}" for v, n in stalls)
    print(f"| {r[ix['ID']]} | {name} | {r[ix['Grid Size']]} x {r[ix['Block Size']]} | {us:.1f} | {rd/1e6:.1f} | {wr/1e6:.1f} | "
          f"{(rd+wr)/us/1e3:.0f} | {val(r,'dram__bytes_read.sum.pct_of_peak_sustained_elapsed', 0.0) + val(r,'dram__bytes_write.sum.pct_of_peak_sustained_elapsed', 0.0):.1f} | "
          f"{val(r,'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{val(r,'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{val(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {val(r,'launch__registers_per_thread'):.0f} | "
          f"{to_bytes(r,'launch__shared_mem_per_block_dynamic')/1e3 if unit('launch__shared_mem_per_block_dynamic') else float('nan'):.1f} | {st} |")
