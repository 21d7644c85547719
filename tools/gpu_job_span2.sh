cat > /tmp/span_once.py <<PY
import torch
import vectorwave_b200 as vw
from vectorwave_b200.sharded import SpanShardedMODWT
n=1<<27
sh=SpanShardedMODWT(vw.Coiflet.COIF5, 10, n, vw.BoundaryMode.PERIODIC, rank=0, world=1, engine=vw.Engine.get(0))
x=torch.randn(n,dtype=torch.float64,device="cuda")
r=sh.forward(x); y=sh.inverse(r)
r=sh.forward(x, result=r); y=sh.inverse(r)
torch.cuda.synchronize()
print(float((y-x).abs().max()))
PY
PYTHONPATH=$PWD timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 0 -c 120 --csv --log-file gpurun_out/kern_span.csv python /tmp/span_once.py > gpurun_out/span_once.log 2>&1
tail -2 gpurun_out/span_once.log
