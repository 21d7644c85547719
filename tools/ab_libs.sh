# A/B of two engine builds on the quick configs: bash tools/ab_libs.sh <lib_b.so> "<configs>" [extra quickbench args]
LB=${1}; CFG=${2:-c2_haar,c2_db4,c3_sym8,c5_db8}; shift 2
for i in 1 2; do
  for lib in A B; do
    if [ $lib = B ]; then export VW_LIB_PATH=$LB; else unset VW_LIB_PATH; fi
    python tools/quickbench.py --configs $CFG --reps 20 "$@" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$lib', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['rt_err'])"
  done
done
