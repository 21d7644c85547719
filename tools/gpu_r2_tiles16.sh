for tl in 0 1024 1536 2048; do
  python tools/quickbench.py --configs c3_sym8,c5_db8 --reps 10 --tile $tl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tile $tl', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'])"
done
for tl in 1024 1536; do
  python tools/quickbench.py --configs c3_sym8,c5_db8 --reps 10 --tile $tl --threads 256 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tile $tl thr 256', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'])"
done
