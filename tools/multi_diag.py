"""Developer diagnostic: where does the one-host-thread multi-GPU span path spend its time?"""
import math, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorwave_b200 as vw
from vectorwave_b200 import _native
S = 1.0 / math.sqrt(2.0)
world = torch.cuda.device_count()
print("devices", world, "peer 0->1", torch.cuda.can_device_access_peer(0, 1) if world > 1 else None)
n_total, levels = 1 << 28, 10
n_local = n_total // world
wv = vw.Coiflet.COIF5
hs, gs = wv.lowPassDecomposition() * S, wv.highPassDecomposition() * S
plan = _native.span_plan(hs.size, levels, n_local, world)
lead, lead_w, pad = int(plan.lead), int(plan.lead_w), int(plan.pad)
row = lead_w + n_local + pad
devices = list(range(world))
me = _native.MultiEngine(devices)
xext, w, v, xo = [], [], [], []
for d in devices:
    dev = torch.device("cuda", d)
    e = torch.randn(lead + n_local, dtype=torch.float64, device=dev)
    xext.append(e); w.append(torch.empty((levels, row), dtype=torch.float64, device=dev))
    v.append(torch.empty(n_local + pad, dtype=torch.float64, device=dev)); xo.append(torch.empty(n_local, dtype=torch.float64, device=dev))
for d in devices: torch.cuda.synchronize(d)
def t(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
print("sharded forward ms", t(lambda: me.forward(plan, xext, hs, gs, 0, w, v)))
print("sharded inverse ms", t(lambda: me.inverse(plan, w, v, hs, gs, 0, 0, xo)))
# the same cascades per device, no exchange, each device alone
import ctypes as C
lib = me.lib
for d in devices:
    ctx = C.c_void_p(lib.vw_multi_ctx(me.handle, d))
    def one():
        lib.vw_modwt_forward_span_all(ctx, C.c_void_p(xext[d].data_ptr()), C.byref(plan), hs.ctypes.data_as(_native._dp), gs.ctypes.data_as(_native._dp),
                                      C.c_void_p(w[d].data_ptr()), row, C.c_void_p(v[d].data_ptr()), 1)
    print("device", d, "forward_span_all alone ms", t(one))
# both devices, enqueue without sync then sync all
def both():
    for d in devices:
        ctx = C.c_void_p(lib.vw_multi_ctx(me.handle, d))
        lib.vw_modwt_forward_span_all(ctx, C.c_void_p(xext[d].data_ptr()), C.byref(plan), hs.ctypes.data_as(_native._dp), gs.ctypes.data_as(_native._dp),
                                      C.c_void_p(w[d].data_ptr()), row, C.c_void_p(v[d].data_ptr()), 1 | 16)
    me.synchronize()
print("all devices forward_span_all NO_SYNC + sync ms", t(both))
t0 = time.perf_counter()
for d in devices:
    ctx = C.c_void_p(lib.vw_multi_ctx(me.handle, d))
    lib.vw_modwt_forward_span_all(ctx, C.c_void_p(xext[d].data_ptr()), C.byref(plan), hs.ctypes.data_as(_native._dp), gs.ctypes.data_as(_native._dp),
                                  C.c_void_p(w[d].data_ptr()), row, C.c_void_p(v[d].data_ptr()), 1 | 16)
    print("  enqueue device", d, "returned after ms", (time.perf_counter() - t0) * 1e3)
me.synchronize()
print("  all done after ms", (time.perf_counter() - t0) * 1e3)
me.close()
