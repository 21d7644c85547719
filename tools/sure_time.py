"""Developer timing of the SURE risk scan (device-resident row, CUDA events)."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorwave_b200 as vw
eng = vw.Engine.get()
for n in (4096, 65536, 262144, 1 << 20):
    c = torch.randn((1, n), dtype=torch.float64, device="cuda")
    eng.sure_threshold(c, 1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t = eng.sure_threshold(c, 1.0); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"n": n, "ms": round(ms, 3), "pairs_per_s": round(n * n / ms * 1e3, 0), "thr": float(t[0])}), flush=True)
