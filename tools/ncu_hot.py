"""Developer helper: per-opcode and top-instruction stall sampling from `ncu --page source --csv` output (stdin)."""
import csv
import sys
from collections import Counter, defaultdict

rows = list(csv.reader(sys.stdin))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ix = {n: i for i, n in enumerate(hdr)}
stall_names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
per_op = defaultdict(lambda: [0, 0, Counter()])
tot = 0
insts = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Address", "Kernel Name"):
        break   # first kernel of the export only
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    s = int(r[ix["# Samples"]] or 0)
    ex = int(r[ix["Instructions Executed"]] or 0)
    per_op[op][0] += s
    per_op[op][1] += ex
    for n in stall_names:
        v = int(r[ix[n]] or 0)
        if v:
            per_op[op][2][n] += v
    tot += s
    insts.append((s, ex, src, {n: int(r[ix[n]] or 0) for n in stall_names if int(r[ix[n]] or 0)}))
print(f"total samples {tot}")
print("opcode        samples   share   executed   samples/exec  top stalls")
for op, (s, ex, c) in sorted(per_op.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"{op:12s} {s:8d} {100*s/tot:6.1f}% {ex:10d} {s/max(ex,1):10.4f}   " + ", ".join(f"{k[6:]}={v}" for k, v in c.most_common(4)))
print("top instructions:")
for s, ex, src, st in sorted(insts, key=lambda t: -t[0])[:int(sys.argv[1]) if len(sys.argv) > 1 else 12]:
    print(f"{s:7d} {ex:9d}  {src[:70]:70s} " + ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3]))
