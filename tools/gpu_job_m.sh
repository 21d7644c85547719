python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r1t.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_r1t.log
tail -12 gpurun_out/gpu_tests_r1t.log
python tools/quickbench.py --configs c2_haar,c2_db4,c3_sym8,c5_db8,c4_coif5 --reps 5 > gpurun_out/quick_r1t.jsonl 2>&1
for t in 684 1024 1366 2048; do python tools/quickbench.py --configs c2_db4 --reps 10 --tile $t; done >> gpurun_out/quick_r1t.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/quick_r1t.jsonl"):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(d["config"], d["opts"]["tile"], d["fwd_ms"], d["inv_ms"], d["fwdinv_gsamples"], d["rt_err"])
PY
