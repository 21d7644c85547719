# A/B of two engine builds over forced tile sizes: bash tools/ab_tiles.sh <lib_b.so> "<tiles>" "<configs>"
LB=${1}; TILES=${2:-"0 1024 1366 1536 2048"}; CFG=${3:-c2_haar,c2_db4}
for t in $TILES; do
  for lib in A B; do
    if [ $lib = B ]; then export VW_LIB_PATH=$LB; else unset VW_LIB_PATH; fi
    python tools/quickbench.py --configs $CFG --reps 20 --tile $t | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$lib', $t, d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'])"
  done
done
