"""Developer helper for ncu captures: one forward + one inverse of a named shape (no timing loop)."""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorwave_b200 as vw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--wavelet", default="coif5")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--log2n", type=int, default=26)
ap.add_argument("--levels", type=int, default=10)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--denoise", type=int, default=0)
a = ap.parse_args()
S = 1.0 / math.sqrt(2.0)
eng = vw.Engine.get()
wv = vw.get_wavelet(a.wavelet)
hs, gs = wv.lowPassDecomposition() * S, wv.highPassDecomposition() * S
n = 1 << a.log2n
x = torch.randn((a.batch, n), dtype=torch.float64, device="cuda")
w = torch.empty((a.levels, a.batch, n), dtype=torch.float64, device="cuda")
v = torch.empty((a.batch, n), dtype=torch.float64, device="cuda")
xr = torch.empty((a.batch, n), dtype=torch.float64, device="cuda")
from vectorwave_b200.modwt import multilevel_alignment  # noqa: E402
bm = [vw.BoundaryMode.PERIODIC, vw.BoundaryMode.ZERO_PADDING, vw.BoundaryMode.SYMMETRIC][a.mode]
align, order = multilevel_alignment(wv, bm, a.levels)
for _ in range(a.warm + 1):
    eng.forward(x, hs, gs, a.levels, a.mode, 0, w, v)
    eng.inverse(w, v, hs, gs, a.mode, align, order, out=xr)
    if a.denoise:
        eng.denoise(x, hs, gs, a.levels, a.mode, align, order, -1.0, True)
torch.cuda.synchronize()
print("rt_err", float((xr - x).abs().max()))
