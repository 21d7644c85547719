# direct-form ANALYSIS pairs (16-20 taps, bit 3 of the lattice option): parity, then config #3 / #5 with them on and off
python -m pytest tests/test_lattice.py -m gpu -x -q 2>&1 | tail -4
python -m pytest tests -m gpu -x -q -k "full_sizes or seeded or specialised" 2>&1 | tail -2
for lat in 15 7 15; do
  python tools/quickbench.py --configs c3_sym8,c5_db8 --reps 6 --lattice $lat 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print('lattice $lat', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'], d['rt_err'])
"
done
