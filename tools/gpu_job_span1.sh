python -m pytest tests -m gpu -x -q -k span 2>&1 | tail -2
python - <<PY
import torch, time, math
import vectorwave_b200 as vw
from vectorwave_b200.sharded import SpanShardedMODWT
n=1<<27
sh=SpanShardedMODWT(vw.Coiflet.COIF5, 10, n, vw.BoundaryMode.PERIODIC, rank=0, world=1, engine=vw.Engine.get(0))
x=torch.randn(n,dtype=torch.float64,device="cuda")
for _ in range(2):
    r=sh.forward(x); y=sh.inverse(r)
torch.cuda.synchronize()
for name,fn in (("fwd",lambda: sh.forward(x, result=r)),("inv",lambda: sh.inverse(r))):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record()
    for _ in range(5): out=fn()
    torch.cuda.synchronize(); t1=time.perf_counter()
    e1.record(); t1=time.perf_counter(); torch.cuda.synchronize()
    print(name, "gpu ms", e0.elapsed_time(e1)/5, "host enqueue ms", (t1-t0)*1e3/5)
PY
