for t in 512 684 1024 1366 2048 4096; do python tools/quickbench.py --configs c2_db4,c2_haar --reps 10 --tile $t; done > gpurun_out/tilesweep_c2.jsonl 2>&1
for t in 1024 2048 3072 4096 6144 8192; do python tools/quickbench.py --configs c5_db8 --reps 3 --tile $t; done > gpurun_out/tilesweep_c5.jsonl 2>&1
for th in 128 192 256; do python tools/quickbench.py --configs c2_db4,c5_db8 --reps 5 --threads $th; done > gpurun_out/threadsweep.jsonl 2>&1
python - <<'PY'
import json
for f in ("tilesweep_c2","tilesweep_c5","threadsweep"):
    for l in open(f"gpurun_out/{f}.jsonl"):
        try: d=json.loads(l)
        except Exception: print(l[:200]); continue
        print(f, d["config"], d["opts"]["tile"], d["opts"]["threads"], d["fwd_ms"], d["inv_ms"], d["fwdinv_gsamples"])
PY
