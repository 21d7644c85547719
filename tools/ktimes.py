"""Summarise an ncu --csv launch list (gpu__time_duration.sum [+ other metrics]) per kernel launch."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
recs = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    key = (int(r[ix["ID"]]), r[ix["Kernel Name"]])
    recs.setdefault(key, {})[r[ix["Metric Name"]]] = r[ix["Metric Value"]]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for (i, k), m in recs.items():
    if i < skip:
        continue
    name = k.replace("void <unnamed>::", "").split("(")[0]
    t = float(m.get("gpu__time_duration.sum", "0").replace(",", ""))
    extra = " ".join(f"{kk.split('__')[-1][:28]}={vv}" for kk, vv in m.items() if kk != "gpu__time_duration.sum")
    print(f"{i:4d} {name:34s} {t/1e3:10.1f} us  {extra}")
