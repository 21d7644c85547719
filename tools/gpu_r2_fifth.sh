O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for rep in 1 2; do
  python tools/quickbench.py --configs c2_haar,c2_db4,c3_sym8,c5_db8 --reps 20 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('default', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'], d['rt_err'])"
done
for tt in "128 1024" "128 2048" "256 1024"; do set -- $tt
  python tools/quickbench.py --configs c2_haar,c2_db4 --reps 20 --threads $1 --tile $2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('thr/tile $1/$2', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'])"
done
python tools/quickbench.py --configs c4_coif5 --reps 3 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('default', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'], d['rt_err'])"
python tools/prof_once.py --wavelet db4 --batch 4096 --log2n 12 --levels 4 --warm 1 > $O/plain_db4.log 2>&1 &&
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_lean" -s 2 -c 2 -o $O/p4_db4 -f \
   python tools/prof_once.py --wavelet db4 --batch 4096 --log2n 12 --levels 4 --warm 1 > $O/ncu_db4.log 2>&1
ncu -i $O/p4_db4.ncu-rep --page source --csv > $O/p4_db4_source.csv 2>/dev/null
