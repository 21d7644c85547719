# A/B of one engine option on the quick configs: bash tools/gpu_ab.sh <option> <configs>
opt=${1:-fpipe}; cfgs=${2:-c2_haar,c2_db4,c3_sym8,c5_db8}
python -m pytest tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -2
for v in 0 1 0 1; do
  python tools/quickbench.py --configs $cfgs --reps 10 --$opt $v > gpurun_out/ab_$v.jsonl 2>&1
  python - $v <<'PY'
import json, sys
for l in open(f"gpurun_out/ab_{sys.argv[1]}.jsonl"):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(sys.argv[1], d["config"], d["fwd_ms"], d["inv_ms"], d["fwdinv_gsamples"], d["fwd_launches"], d["rt_err"])
PY
done
