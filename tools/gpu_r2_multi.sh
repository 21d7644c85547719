# multi-GPU box: the tests that need >= 2 devices, then the bench under torchrun
N=${1:-2}
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
python -m pytest tests/test_abi_v2.py -m gpu -x -q -k "real_devices or nccl or sharded" 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_r02_n$N.json 2> $O/bench_r02_n$N.err
tail -2 $O/bench_r02_n$N.err
python - $N <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/bench_r02_n{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("N",d["n_gpus"],"value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"], d["e2e"].get("numa_binding"), d["e2e"].get("pcie_gbs_per_rank_each_way_in_the_step"))
for k,v in (d.get("extra") or {}).items():
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk in ("value","ms_per_step","failed","halo_exchange_ms_per_step","frac_of_roofline_model","rows_this_rank")})
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload span --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-400
