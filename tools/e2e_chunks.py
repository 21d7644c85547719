"""Developer helper: e2e (host buffers, all PCIe crossings) of the bench workload as a function of the pipeline chunk cap."""
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorwave_b200 as vw  # noqa: E402

S = 1.0 / math.sqrt(2.0)
eng = vw.Engine.get()
wv = vw.get_wavelet("db4")
hs, gs = wv.lowPassDecomposition() * S, wv.highPassDecomposition() * S
b, n, levels = 4096, 4096, 4
xh = eng.pinned_empty((b, n)); xh[...] = np.random.default_rng(7).standard_normal((b, n))
wh = eng.pinned_empty((levels, b, n)); vh = eng.pinned_empty((b, n)); oh = eng.pinned_empty((b, n))
for chunks in [int(a) for a in (sys.argv[1:] or ["8", "16", "32", "64", "8"])]:
    eng.set_option("pipe_chunks", chunks)
    for _ in range(2):
        eng.forward(xh, hs, gs, levels, 0, 0, wh, vh); eng.inverse(wh, vh, hs, gs, 0, None, 0, out=oh)
    t0 = time.perf_counter()
    for _ in range(8):
        eng.forward(xh, hs, gs, levels, 0, 0, wh, vh); eng.inverse(wh, vh, hs, gs, 0, None, 0, out=oh)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 8
    print(f"pipe_chunks {chunks}: {dt * 1e3:.2f} ms/step, e2e {b * n / dt * 1e-9:.3f} GSamples/s, err {float(np.max(np.abs(oh - xh))):.2e}", flush=True)
