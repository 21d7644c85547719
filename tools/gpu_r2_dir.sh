# direct-form synthesis pairs (16-20 taps): parity, then c3 / c5 inverse by first pair level
python -m pytest tests/test_lattice.py -m gpu -x -q 2>&1 | tail -3
q() {
  python tools/quickbench.py --configs c3_sym8,c5_db8 --reps 6 "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print('$LABEL', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'])
"
}
for cm in 0 5 6 7; do LABEL="pairs from level $cm" q --colmin $cm; done
LABEL="pairs off" q --lattice 3
LABEL="pairs from level 0" q
