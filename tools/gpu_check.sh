python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/quickbench.py --configs c2_haar,c2_db4,c3_sym8,c5_db8,c4_coif5 --reps 5 > gpurun_out/quick_r1u.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/quick_r1u.jsonl"):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d["config"], d["fwd_ms"], d["inv_ms"], d["fwdinv_gsamples"], d["fwd_launches"], d["inv_launches"], d["rt_err"])
PY
