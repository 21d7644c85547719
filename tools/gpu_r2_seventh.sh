O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_fused.py tests/test_abi_v2.py -m gpu -x -q 2>&1 | tail -3
for pf in 1 3 5 7 0; do
  python tools/quickbench.py --configs c2_haar,c2_db4,c3_sym8 --reps 30 --l2pf $pf | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('l2pf $pf', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'])"
done
python tools/quickbench.py --configs c5_db8 --reps 5 --mode 2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('symmetric', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'])"
