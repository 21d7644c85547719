"""Developer helper: warp-stall samples of one kernel aggregated by CUDA source line.
ncu's CSV source page carries SASS only; the line of every SASS instruction comes from `nvdisasm --print-line-info` of
the same build, aligned by instruction index.
Usage: python tools/ncu_lines.py <source.csv> <cubin> <mangled-function-substring> [kernel-index-in-csv] [top]"""
import csv
import re
import subprocess
import sys
from collections import Counter, defaultdict

src_csv, cubin, fun = sys.argv[1], sys.argv[2], sys.argv[3]
kidx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
lines_of = []   # per instruction: (file line, inlined-at chain)
inside = False
cur = None
for ln in dis:
    if ln.startswith("//-") and ".text." in ln:
        inside = fun in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (int(m.group(2)), m.group(3).strip())
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', ln):
        lines_of.append(cur)

rows = list(csv.reader(open(src_csv)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = heads[kidx]
end = heads[kidx + 1] - 1 if kidx + 1 < len(heads) else len(rows)
hdr = rows[hi]
ix = {n: i for i, n in enumerate(hdr)}
stall_names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[hi + 1:end] if len(r) >= len(hdr)]
print(f"kernel {rows[hi - 1][1][:90]}: {len(body)} SASS rows, {len(lines_of)} disassembled")
per = defaultdict(lambda: [0, 0, Counter()])
tot = 0
for i, r in enumerate(body):
    key = lines_of[i] if i < len(lines_of) and lines_of[i] else (-1, "")
    s = int(r[ix["# Samples"]] or 0)
    per[key][0] += s
    per[key][1] += int(r[ix["Instructions Executed"]] or 0)
    for n in stall_names:
        v = int(r[ix[n]] or 0)
        if v:
            per[key][2][n[6:]] += v
    tot += s
src = open("" + (sys.argv[6] if len(sys.argv) > 6 else "vectorwave_b200/csrc/vw_fused.cu") + "").read().split("\n")
print(f"total samples {tot}")
for (line, ctx), (s, ex, c) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[line - 1].strip()[:70] if 0 < line <= len(src) else "?"
    print(f"{s:6d} {100*s/tot:5.1f}% ex={ex:9d} L{line:<5d} {text:70s} | " + ", ".join(f"{k}={v}" for k, v in c.most_common(3)))
