O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/quickbench.py --configs c2_haar,c2_db4,c3_sym8,c5_db8 --reps 20 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('default', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'], d['fwd_launches'], d['inv_launches'], d['rt_err'])"
python tools/quickbench.py --configs c2_haar,c2_db4 --reps 20 --lean_small 0 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('lean_small=0', d['config'], d['fwd_ms'], d['inv_ms'], d['fwdinv_gsamples'])"
tools/_build/latency 0 2000
( time python bench.py > $O/bench_r02a.json 2> $O/bench_r02a.err ) 2>&1 | tail -3
tail -3 $O/bench_r02a.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02a.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"],"fused",d["roofline"]["frac_of_fused_compulsory_bound"],d["roofline"]["inverse_frac_of_fused_compulsory_bound"],"fp64 peak",d["roofline"]["fp64_tflops_peak_measured"])
for k,v in (d.get("extra") or {}).items():
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","failed","halo_exchange_ms_per_step","frac_of_roofline_model","h2d_gbs","d2h_gbs","unavailable","c1_1x4096_db4_J1_forward_host_buffers_us","c1_graph_replay_us","c2s_16x4096_db4_J4_forward_device_sync_us","c2s_no_sync_back_to_back_us","c2s_graph_replay_back_to_back_us")})
PY
