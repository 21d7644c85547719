q() {
  python tools/quickbench.py --configs c4_coif5,c4e_coif5 --reps 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print('$LABEL', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['rt_err'])
"
}
for i in 1 2; do
unset VW_LIB_PATH; LABEL="RA=8" q
export VW_LIB_PATH=$PWD/vectorwave_b200/libvwmodwt_b.so; LABEL="RA=16" q
done
