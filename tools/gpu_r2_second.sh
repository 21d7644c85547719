# round-2 second box call: where does the db4 tile kernel's time go?  (compute-only variant, tile/thread sweep, ncu source pages)
O=gpurun_out; mkdir -p $O
echo "== compute-only variant (B = no global stores)"
bash tools/ab_libs.sh /root/repo/vectorwave_b200/libvwmodwt_b.so c2_haar,c2_db4
echo "== threads/tile sweep"
for tt in "256 2048" "128 1024" "128 2048" "256 1366" "256 4096"; do set -- $tt
  python tools/quickbench.py --configs c2_haar,c2_db4 --reps 20 --threads $1 --tile $2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('thr/tile $1/$2', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'])"
done
echo "== ncu"
for wv in db4 haar; do
python tools/prof_once.py --wavelet $wv --batch 4096 --log2n 12 --levels 4 --warm 1 > $O/plain_$wv.log 2>&1 &&
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_fused_(analysis|synthesis)" -s 2 -c 2 -o $O/p2_$wv -f \
   python tools/prof_once.py --wavelet $wv --batch 4096 --log2n 12 --levels 4 --warm 1 > $O/ncu_$wv.log 2>&1
ncu -i $O/p2_$wv.ncu-rep --page source --csv > $O/p2_${wv}_source.csv 2>/dev/null
ncu -i $O/p2_$wv.ncu-rep --page raw --csv > $O/p2_${wv}_raw.csv 2>/dev/null
done
ls -la $O/
