"""Developer tool: where a tile's lifetime goes in k_fused_analysis.  Needs an instrumented build
(make -C vectorwave_b200/csrc TARGET=../libvwmodwt_dbg.so BUILD=build_dbg EXTRA=-DVW_PHASE_CLOCKS) loaded via
VW_LIB_PATH.  Runs one forward of a named config and prints, per phase, the mean / p50 / p90 duration over the CTAs, the
mean CTA lifetime, how many CTAs are alive per SM over time, and the share of CTA-time spent waiting for the input tile."""
import argparse, ctypes as C, math, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorwave_b200 as vw
from tools.quickbench import CONFIGS, S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2_db4")
ap.add_argument("--inverse", type=int, default=0, help="1: the synthesis tile kernel (last launch of one inverse)")
a = ap.parse_args()
wname, b, n, levels = CONFIGS[a.config]
eng = vw.Engine.get()
if not hasattr(eng.lib, "vw_debug_phase_log"):
    raise SystemExit("not an instrumented build: set VW_LIB_PATH to a -DVW_PHASE_CLOCKS library")
wv = vw.get_wavelet(wname)
hs, gs = wv.lowPassDecomposition() * S, wv.highPassDecomposition() * S
x = torch.randn((b, n), dtype=torch.float64, device="cuda")
w = torch.empty((levels, b, n), dtype=torch.float64, device="cuda")
v = torch.empty((b, n), dtype=torch.float64, device="cuda")
for _ in range(3):
    eng.forward(x, hs, gs, levels, 0, 0, w, v)
torch.cuda.synchronize()
if a.inverse:
    xr = torch.empty_like(x)
    for _ in range(3):
        eng.inverse(w, v, hs, gs, 0, None, 0, out=xr)
    torch.cuda.synchronize()
    groups = vw._native.plan_groups(0, hs.size, levels, n)
    print("inverse plan", groups, "(the log holds the LAST launch: the finest group)")
    ctas = 16384
    log = np.zeros((ctas, 16), dtype=np.uint64)
    assert eng.lib.vw_debug_phase_log_inv(log.ctypes.data_as(C.c_void_p), ctas) == 0
    log = log[log[:, 11] > 0].astype(np.int64)
    nlev = min(groups[0][1], 4)
    life = log[:, 11] - log[:, 0]
    span = log[:, 11].max() - log[:, 0].min()
    print(f"CTAs logged: {log.shape[0]}, kernel span {span / 1e3:.1f} us, mean CTA lifetime {life.mean() / 1e3:.2f} us")

    def row(name, d):
        print(f"  {name:16s} mean {d.mean() / 1e3:6.2f} us  p50 {np.percentile(d, 50) / 1e3:6.2f}  p90 {np.percentile(d, 90) / 1e3:6.2f}  "
              f"share of lifetime {100 * d.sum() / life.sum():5.1f} %")
    row("V tile wait", log[:, 1] - log[:, 0])
    prev = log[:, 1]
    for k in range(nlev):
        row(f"level {nlev - k} W wait", log[:, 2 + 2 * k] - prev)
        row(f"level {nlev - k} compute", log[:, 3 + 2 * k] - log[:, 2 + 2 * k])
        prev = log[:, 3 + 2 * k]
    row("store drain", log[:, 11] - prev)
    sms = len(np.unique(log[:, 10]))
    print(f"CTAs alive per SM, time-averaged: {life.sum() / span / sms:.2f} over {sms} SMs")
    sys.exit(0)
groups = vw._native.plan_groups(1, hs.size, levels, n)
print("plan", groups, "(the log holds the LAST launch of the forward)")
ctas = 16384
log = np.zeros((ctas, 8), dtype=np.uint64)
rc = eng.lib.vw_debug_phase_log(log.ctypes.data_as(C.c_void_p), ctas)
assert rc == 0, rc
live = log[:, 7] > 0
log = log[live].astype(np.int64)
print("CTAs logged:", log.shape[0])
t0 = log[:, 0].min()
names = ["input wait", "level 1", "level 2", "level 3", "level 4"]
nlev = groups[-1][1]
stamps = [log[:, 0], log[:, 1]] + [log[:, 2 + i] for i in range(min(nlev, 4))]
life = log[:, 7] - log[:, 0]
print(f"kernel span {(log[:, 7].max() - t0) / 1e3:.1f} us, mean CTA lifetime {life.mean() / 1e3:.2f} us")
for i in range(len(stamps) - 1):
    d = stamps[i + 1] - stamps[i]
    print(f"  {names[i]:11s} mean {d.mean() / 1e3:6.2f} us  p50 {np.percentile(d, 50) / 1e3:6.2f}  p90 {np.percentile(d, 90) / 1e3:6.2f}  "
          f"share of lifetime {100 * d.sum() / life.sum():5.1f} %")
d = log[:, 7] - stamps[-1]
print(f"  {'store drain':11s} mean {d.mean() / 1e3:6.2f} us  p50 {np.percentile(d, 50) / 1e3:6.2f}  p90 {np.percentile(d, 90) / 1e3:6.2f}  "
      f"share of lifetime {100 * d.sum() / life.sum():5.1f} %")
sm = log[:, 6]
span = log[:, 7].max() - t0
print(f"CTAs alive per SM, time-averaged: {life.sum() / span / len(np.unique(sm)):.2f} over {len(np.unique(sm))} SMs")
