N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/benchN${N}_batch.json 2> gpurun_out/benchN${N}_batch.err
tail -c 600 gpurun_out/benchN${N}_batch.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --workload span > gpurun_out/benchN${N}_span.json 2> gpurun_out/benchN${N}_span.err
tail -c 600 gpurun_out/benchN${N}_span.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --workload batch1024x65536_sym8_J8 --no-cpu-baseline --no-e2e > gpurun_out/benchN${N}_c3.json 2> gpurun_out/benchN${N}_c3.err
python - <<PY
import json
for f in ("batch","span","c3"):
    try:
        d=json.loads(open(f"gpurun_out/benchN${N}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"],2), d["ms_per_step"], d["scaling"], d.get("round_trip_max_abs_err"), d["roofline"].get("frac"))
    except Exception as e: print(f, "failed", e)
PY
