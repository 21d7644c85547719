O=gpurun_out; mkdir -p $O
python -m pytest tests/test_cpp_host.py -m gpu -x -q 2>&1 | tail -2
for th in 0 192 256; do
  python tools/quickbench.py --configs c3_sym8,c5_db8 --reps 10 --threads $th | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('threads $th', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'])"
done
echo "== zero copy"
timeout 120 tools/_build/latency 0 2000 zc
