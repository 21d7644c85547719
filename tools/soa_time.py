"""Developer timing: SoA batch forward (native flat-signal path) vs the AoS forward of the same batch."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorwave_b200 as vw
S = 1.0 / math.sqrt(2.0)
eng = vw.Engine.get()
for wname, b, n, levels in (("db4", 4096, 4096, 4), ("haar", 4096, 4096, 4), ("sym8", 1024, 65536, 8)):
    wv = vw.get_wavelet(wname)
    hs, gs = wv.lowPassDecomposition() * S, wv.highPassDecomposition() * S
    x = torch.randn(b * n, dtype=torch.float64, device="cuda")
    w = [torch.empty(b * n, dtype=torch.float64, device="cuda") for _ in range(levels)]
    v = torch.empty(b * n, dtype=torch.float64, device="cuda")
    def run():
        eng.forward_soa(x, b, n, hs, gs, w, v)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"soa_forward": wname, "batch": b, "n": n, "levels": levels, "ms": round(ms, 4),
                      "gbs_alg": round(24.0 * levels * b * n / ms * 1e-6, 1)}), flush=True)
