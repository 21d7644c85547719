# lattice column kernels (coif5): parity first, then A/B against the direct form and the column threshold
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_lattice.py -m gpu -x -q 2>&1 | tail -8
python -m pytest tests -m gpu -x -q -k "coif5 or specialised or seeded or full_sizes or stream or soa" 2>&1 | tail -4
for opt in "--lattice 1" "--lattice 0" "--lattice 1 --colmin 2" "--lattice 1 --colmin 4"; do
  python tools/quickbench.py --configs c4_coif5 --reps 4 $opt 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print('$opt', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'], d['rt_err'])
"
done
