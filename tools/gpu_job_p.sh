for t in 128 192 256; do python tools/quickbench.py --configs c3_sym8,c5_db8,c2_db4 --reps 5 --threads $t; done > gpurun_out/thr_sweep2.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/thr_sweep2.jsonl"):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d["config"], d["opts"]["threads"], d["opts"]["tile"], d["fwd_ms"], d["inv_ms"], d["fwdinv_gsamples"], d["fwd_launches"], d["inv_launches"], d["rt_err"])
PY
