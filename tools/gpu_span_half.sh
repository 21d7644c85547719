# single-rank timing of the span path at the per-rank sizes of 2- and 8-GPU jobs (2^27 and 2^25 samples)
python - <<PY
import torch, time
import vectorwave_b200 as vw
from vectorwave_b200.sharded import SpanShardedMODWT
for lg in (27, 25):
    n = 1 << lg
    sh = SpanShardedMODWT(vw.Coiflet.COIF5, 10, n, vw.BoundaryMode.PERIODIC, rank=0, world=1, engine=vw.Engine.get(0))
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    r = sh.forward(x); y = sh.inverse(r)
    r = sh.forward(x, result=r); y = sh.inverse(r)
    torch.cuda.synchronize()
    out = {}
    for name, fn in (("fwd", lambda: sh.forward(x, result=r)), ("inv", lambda: sh.inverse(r))):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 10
    print(lg, out, "GS/s per rank", n / (out["fwd"] + out["inv"]) * 1e-6, "err", float((y - x).abs().max()))
PY
