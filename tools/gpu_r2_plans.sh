# plan alternatives for the 16-tap configs (column threshold x fuse cap), forward / inverse ms
for cfg in c3_sym8 c5_db8; do
for cm in 0 7 8 9; do
 for fu in 0 2 3; do
  python tools/quickbench.py --configs $cfg --reps 6 --colmin $cm --fuse $fu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:200]); continue
    print('colmin $cm fuse $fu', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'])
"
 done
done
done
