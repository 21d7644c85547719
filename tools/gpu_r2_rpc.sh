python -m pytest tests/test_lattice.py tests/test_abi_v2.py -m gpu -x -q 2>&1 | tail -2
for cfg in c4e_coif5 c4_coif5; do
  python tools/quickbench.py --configs $cfg --reps 8 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print(d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'], d['rt_err'])
"
done
python tools/prof_once.py --warm 0 --log2n 25 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_2p25.csv python tools/prof_once.py --warm 0 --log2n 25 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_2p25.csv")) if len(r)>10]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value"); ig=h.index("Grid Size")
for r in rows[1:]:
    if "k_" in r[ik]: print(r[ik][:60], r[ig], r[iv])
PY
