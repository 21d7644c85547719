# shared-memory ring prefetch of the pair kernels: parity on the default build, then builds A (2 CTAs/SM, depth 4), B (3 CTAs, depth 2), C (3 CTAs, analysis depth 4)
python -m pytest tests/test_lattice.py -m gpu -x -q 2>&1 | tail -3
q() {
  python tools/quickbench.py --configs c4_coif5,c4e_coif5,c3_sym8,c5_db8 --reps 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print('$LABEL', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['rt_err'])
"
}
unset VW_LIB_PATH; LABEL=A q
for v in b c; do export VW_LIB_PATH=$PWD/vectorwave_b200/libvwmodwt_$v.so; LABEL=$v q; done
unset VW_LIB_PATH; LABEL=A q
