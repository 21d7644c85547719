for cm in 0 1 2 0 1; do
  python tools/quickbench.py --configs c4_coif5 --reps 4 --colmin $cm 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print('colmin $cm', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'], d['rt_err'])
"
done
