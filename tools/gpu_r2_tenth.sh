for cm in 0 6 7 8; do
  python tools/quickbench.py --configs c3_sym8,c5_db8 --reps 10 --colmin $cm | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('colmin $cm', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'])"
done
for fu in 2 3; do
  python tools/quickbench.py --configs c3_sym8,c5_db8 --reps 10 --fuse $fu | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('fuse $fu', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'])"
done
