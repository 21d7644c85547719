O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
python tools/quickbench.py --configs c2_haar,c2_db4,c3_sym8,c5_db8 --reps 20 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('default', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['rt_err'])"
done
tools/_build/latency 0 2000 zc
