"""Developer timing of BASELINE config #1 (1 x 4096, db4, J = 1, PERIODIC, forward only): the one shape with a published
reference number (core 0.358 ms / extensions 0.117 ms, docs/BENCHMARK-RESULTS.md:26).  Launch-latency bound: reports
microseconds per call through the host-buffer API (H2D + kernel + D2H + sync) and device-resident."""
import json, math, os, statistics, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorwave_b200 as vw
S = 1.0 / math.sqrt(2.0)
eng = vw.Engine.get()
for kv in filter(None, os.environ.get("VW_OPTS", "").split(",")):   # e.g. VW_OPTS=l2pf=0,tile=2048
    k, v = kv.split("=")
    eng.set_option(k, int(v))
wv = vw.get_wavelet("db4")
hs, gs = wv.lowPassDecomposition() * S, wv.highPassDecomposition() * S
n = 4096
x = eng.pinned_empty((1, n)); x[...] = np.random.default_rng(42).standard_normal((1, n))
w = eng.pinned_empty((1, 1, n)); v = eng.pinned_empty((1, n))
def host():
    eng.forward(x, hs, gs, 1, 0, 0, w, v)
for _ in range(50):
    host()
ts = []
for _ in range(500):
    t0 = time.perf_counter_ns(); host(); ts.append(time.perf_counter_ns() - t0)
xd = torch.as_tensor(x, device="cuda"); wd = torch.empty((1, 1, n), dtype=torch.float64, device="cuda"); vd = torch.empty((1, n), dtype=torch.float64, device="cuda")
for _ in range(50):
    eng.forward(xd, hs, gs, 1, 0, 0, wd, vd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(500):
    eng.forward(xd, hs, gs, 1, 0, 0, wd, vd)
e1.record(); torch.cuda.synchronize()
t = vw.MODWTTransform(wv, vw.BoundaryMode.PERIODIC)
xs = np.asarray(x[0])
for _ in range(50):
    t.forward(xs)
tt = []
for _ in range(300):
    t0 = time.perf_counter_ns(); t.forward(xs); tt.append(time.perf_counter_ns() - t0)
print(json.dumps({"config": "c1_db4_J1_forward", "host_api_us_median": round(statistics.median(ts) / 1e3, 1),
                  "host_api_us_p10": round(sorted(ts)[50] / 1e3, 1), "device_resident_us": round(e0.elapsed_time(e1) * 2, 2),
                  "MODWTTransform_forward_us_median": round(statistics.median(tt) / 1e3, 1),
                  "reference_published_us": {"core": 358, "extensions": 117}}))
