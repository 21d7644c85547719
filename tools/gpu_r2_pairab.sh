# A/B of pair-kernel builds (libvwmodwt_<x>.so made with EXTRA=-DVW_PAIR_*): config #4 forward / inverse
for v in "$@"; do
  export VW_LIB_PATH=$PWD/vectorwave_b200/libvwmodwt_$v.so
  python tools/quickbench.py --configs c4_coif5 --reps 4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print('$v', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'], d['rt_err'])
"
done
