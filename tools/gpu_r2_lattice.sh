# lattice column kernels (coif5): parity first, then A/B of the pair kernels (2 vs 3 CTAs per SM) and against single levels
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_lattice.py -m gpu -x -q 2>&1 | tail -8
python -m pytest tests -m gpu -x -q -k "coif5 or specialised or seeded or full_sizes or stream or soa or shard or span" 2>&1 | tail -4
q() {
  python tools/quickbench.py --configs c4_coif5 --reps 4 "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print('$LABEL', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'], d['rt_err'])
"
}
LABEL="pairs(2 CTAs)" q --lattice 3
LABEL="singles" q --lattice 1
export VW_LIB_PATH=$PWD/vectorwave_b200/libvwmodwt_b.so
LABEL="pairs(3 CTAs, spills)" q --lattice 3
unset VW_LIB_PATH
LABEL="pairs(2 CTAs) again" q --lattice 3
