python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in 0 1 2; do
python tools/quickbench.py --configs c5_db8 --reps 5 --mode $m --denoise 1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('mode $m', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d.get('denoise_ms'), d['fwd_launches'], d['inv_launches'])"
done
python tools/quickbench.py --configs c2_haar,c2_db4,c3_sym8 --reps 20 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('default', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['rt_err'])"
