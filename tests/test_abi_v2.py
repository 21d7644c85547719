"""GPU tests of the second ABI generation (include/vw_modwt.h): whole-cascade span calls, several devices driven by one
host thread, device-resident results, CUDA-graph replay and the timing record -- all through the C ABI, all against
the oracle.  The NCCL test needs two GPUs and is skipped on a one-GPU box."""
import os
import socket

import numpy as np
import pytest

import vectorwave_b200 as vw
from oracle import cref, nptwin
from oracle.wavelets import filters
from vectorwave_b200 import _native

pytestmark = pytest.mark.gpu
S = nptwin.S
REL = 1e-12


def _sharded_roundtrip(devices, name, levels, n_local, mode):
    """decompose + reconstruct of one long signal over len(devices) spans through vw_modwt_*_sharded."""
    import torch
    world = len(devices)
    h, g, wid = filters(name)
    hs, gs = h * S, g * S
    x = np.random.default_rng(17 + world).standard_normal(world * n_local)
    plan = _native.span_plan(h.size, levels, n_local, world)
    lead, lead_w, pad = int(plan.lead), int(plan.lead_w), int(plan.pad)
    row = lead_w + n_local + pad
    me = _native.MultiEngine(devices)
    try:
        xext, w, v, xo = [], [], [], []
        for r, d in enumerate(devices):
            dev = torch.device("cuda", d)
            e = torch.full((lead + n_local,), float("nan"), dtype=torch.float64, device=dev)   # the engine must fill the lead
            e[lead:] = torch.as_tensor(x[r * n_local:(r + 1) * n_local], device=dev)
            xext.append(e)
            w.append(torch.full((levels, row), float("nan"), dtype=torch.float64, device=dev))
            v.append(torch.full((n_local + pad,), float("nan"), dtype=torch.float64, device=dev))
            xo.append(torch.empty(n_local, dtype=torch.float64, device=dev))
        for d in set(devices):
            torch.cuda.synchronize(d)
        ex_f = me.forward(plan, xext, hs, gs, mode, w, v, timed=True)
        order = 1 if mode == 1 else 0
        ex_i = me.inverse(plan, w, v, hs, gs, mode, order, xo, timed=True)
        me.inverse(plan, w, v, hs, gs, mode, order, xo)           # a result survives an inverse
        assert ex_f >= 0.0 and ex_i >= 0.0
        wg = np.concatenate([w[r][:, lead_w:lead_w + n_local].cpu().numpy() for r in range(world)], axis=1)
        vg = np.concatenate([v[r][:n_local].cpu().numpy() for r in range(world)])
        xg = np.concatenate([xo[r].cpu().numpy() for r in range(world)])
    finally:
        me.close()
    wo, vo = cref.decompose(x, h, g, levels, mode)
    t = REL * float(np.max(np.abs(x)))
    np.testing.assert_allclose(wg, wo, rtol=0, atol=t)
    np.testing.assert_allclose(vg, vo, rtol=0, atol=t)
    np.testing.assert_allclose(xg, cref.reconstruct(wo, vo, h, g, mode, wid), rtol=0, atol=t)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name,levels,n_local,world", [("db4", 5, 1 << 14, 4), ("coif5", 6, 1 << 15, 2), ("haar", 9, 1 << 13, 3),
                                                       ("sym8", 7, 1 << 16, 1)])
def test_sharded_abi_spans_on_one_device(name, levels, n_local, world, mode):
    """vw_init_multi accepts a device more than once: `world` spans of one signal, all on cuda:0, exchange halos through
    the same peer-copy path and must reproduce the unsharded transform (ring wrap for PERIODIC, zeros at the open ends for
    ZERO_PADDING; world = 1 wraps onto itself)."""
    _sharded_roundtrip([0] * world, name, levels, n_local, mode)


def test_sharded_abi_over_real_devices():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    _sharded_roundtrip(list(range(min(n, 8))), "coif5", 8, 1 << 18, 0)
    _sharded_roundtrip([0, 1], "db4", 6, 1 << 16, 1)


def _nccl_worker(rank, world, port, name, levels, n_local, ret):
    import torch
    import torch.distributed as dist
    from vectorwave_b200.sharded import SpanShardedMODWT
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        x = np.random.default_rng(23).standard_normal(world * n_local)
        sh = SpanShardedMODWT(vw.get_wavelet(name), levels, n_local, vw.BoundaryMode.PERIODIC, rank=rank, world=world,
                              engine=vw.Engine.get(rank))
        res = sh.forward(torch.as_tensor(x[rank * n_local:(rank + 1) * n_local], device=f"cuda:{rank}"))
        xr = sh.inverse(res)
        torch.cuda.synchronize()
        ret[rank] = (res.details().cpu().numpy(), res.approximation().cpu().numpy(), xr.cpu().numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_span_sharded_over_nccl_two_ranks_equals_unsharded():
    """The per-rank path of the bench (one process per GPU, NCCL send/recv halo exchange, PERIODIC ring wrap) on two real
    ranks: coefficients and reconstruction must equal the UNSHARDED transform of the oracle -- not just round-trip."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    name, levels, n_local, world = "coif5", 8, 1 << 17, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, port, name, levels, n_local, ret), nprocs=world, join=True)
    h, g, wid = filters(name)
    x = np.random.default_rng(23).standard_normal(world * n_local)
    wo, vo = cref.decompose(x, h, g, levels, 0)
    t = REL * float(np.max(np.abs(x)))
    np.testing.assert_allclose(np.concatenate([ret[r][0] for r in range(world)], axis=1), wo, rtol=0, atol=t)
    np.testing.assert_allclose(np.concatenate([ret[r][1] for r in range(world)]), vo, rtol=0, atol=t)
    np.testing.assert_allclose(np.concatenate([ret[r][2] for r in range(world)]), cref.reconstruct(wo, vo, h, g, 0, wid), rtol=0, atol=t)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_device_resident_result_pipeline(mode):
    """decompose -> (levels on request) -> threshold -> reconstruct with the coefficients resident in HBM: every piece
    equals the oracle, and the handle is reused on the steady state."""
    from vectorwave_b200.modwt import multilevel_alignment
    eng = vw.Engine.get()
    name, b, n, levels = "db8", 5, 6000, 4
    h, g, wid = filters(name)
    hs, gs = h * S, g * S
    bm = [vw.BoundaryMode.PERIODIC, vw.BoundaryMode.ZERO_PADDING, vw.BoundaryMode.SYMMETRIC][mode]
    align, order = multilevel_alignment(vw.get_wavelet(name), bm, levels)
    x = np.random.default_rng(3 + mode).standard_normal((b, n))
    t = REL * float(np.max(np.abs(x)))
    res = eng.decompose_resident(x, hs, gs, levels, mode)
    assert res.shape() == (b, n, levels)
    handle0 = res.handle.value
    for i in range(b):
        wo, vo = cref.decompose(x[i], h, g, levels, mode)
        for j in range(1, levels + 1):
            np.testing.assert_allclose(res.get_level(j)[i], wo[j - 1], rtol=0, atol=t)
        np.testing.assert_allclose(res.get_level(0)[i], vo, rtol=0, atol=t)
        np.testing.assert_allclose(res.energy(2)[i], float(np.sum(wo[1] * wo[1])), rtol=1e-12)
    xr = res.reconstruct(hs, gs, mode, align, order)
    for i in range(b):
        wo, vo = cref.decompose(x[i], h, g, levels, mode)
        np.testing.assert_allclose(xr[i], cref.reconstruct(wo, vo, h, g, mode, wid), rtol=0, atol=t)
    # universal threshold on the resident coefficients == VectorWaveSwtAdapter.denoise
    thr = res.universal_threshold(True)
    den = res.reconstruct(hs, gs, mode, align, order)
    for i in range(b):
        dref, tref = cref.swt_denoise(x[i], h, g, levels, mode, wid, -1.0, True)
        assert abs(thr[i] - tref) <= 1e-12 * tref
        np.testing.assert_allclose(den[i], dref, rtol=0, atol=t)
    # steady state: same shape -> same allocation; fixed hard threshold on one level; set_level round trip
    res = eng.decompose_resident(x, hs, gs, levels, mode, result=res)
    assert res.handle.value == handle0
    res.threshold(1, 0.5, False)
    w1 = res.get_level(1)
    for i in range(b):
        wo, _ = cref.decompose(x[i], h, g, levels, mode)
        np.testing.assert_allclose(w1[i], np.where(np.abs(wo[0]) <= 0.5, 0.0, wo[0]), rtol=0, atol=t)
    res.set_level(1, np.zeros((b, n)))
    assert float(np.max(np.abs(res.get_level(1)))) == 0.0
    with pytest.raises(vw.InvalidArgumentException):
        res.get_level(levels + 1)
    res.free()


@pytest.mark.parametrize("mode", [vw.BoundaryMode.PERIODIC, vw.BoundaryMode.SYMMETRIC])
def test_resident_result_through_the_facade(mode):
    """MultiLevelMODWTTransform.decomposeResident: the facade-level twin of GpuResidentResult.java -- same getters as the
    host-backed result, reconstruct / partial reconstruct / thresholds on the device."""
    name, n, levels = "db4", 3000, 5
    h, g, wid = filters(name)
    x = np.random.default_rng(8).standard_normal(n)
    t = vw.MultiLevelMODWTTransform(vw.get_wavelet(name), mode)
    host = t.decompose(x, levels)
    res = t.decomposeResident(x, levels)
    assert res.getLevels() == levels and res.getSignalLength() == n
    tol = REL * float(np.max(np.abs(x)))
    for j in range(1, levels + 1):
        np.testing.assert_allclose(res.getDetailCoeffsAtLevel(j), host.getDetailCoeffsAtLevel(j), rtol=0, atol=tol)
        assert abs(res.getDetailEnergyAtLevel(j) - host.getDetailEnergyAtLevel(j)) <= 1e-12 * host.getDetailEnergyAtLevel(j)
    np.testing.assert_allclose(res.reconstruct(), t.reconstruct(host), rtol=0, atol=tol)
    np.testing.assert_allclose(t.reconstructFromLevel(res, 3), t.reconstructFromLevel(host, 3), rtol=0, atol=tol)
    thr = res.applyUniversalThreshold(True)
    dref, tref = cref.swt_denoise(x, h, g, levels, mode.value, wid, -1.0, True)
    assert abs(thr - tref) <= 1e-12 * tref
    np.testing.assert_allclose(res.reconstruct(), dref, rtol=0, atol=tol)
    with pytest.raises(vw.IllegalArgumentException):
        res.getDetailCoeffsAtLevel(levels + 1)
    res.close()


def test_graph_replay_of_a_small_transform():
    """Config #1's shape (1 x 4096, db4, J = 1) recorded once -- H2D, kernel, two D2H copies -- and replayed as one CUDA
    graph launch on new data in the same pinned buffers."""
    eng = vw.Engine.get()
    h, g, _ = filters("db4")
    hs, gs = h * S, g * S
    n = 4096
    x = eng.pinned_empty((1, n)); w = eng.pinned_empty((1, 1, n)); v = eng.pinned_empty((1, n))
    x[...] = np.random.default_rng(1).standard_normal((1, n))
    eng.forward(x, hs, gs, 1, 0, 0, w, v)                       # un-captured first: scratch exists afterwards
    with eng.capture() as gr:
        eng.forward(x, hs, gs, 1, 0, 0, w, v)
    for seed in (2, 3):
        x[...] = np.random.default_rng(seed).standard_normal((1, n))
        w[...] = 0.0; v[...] = 0.0
        gr.launch()
        wo, vo = cref.decompose(x[0], h, g, 1, 0)
        t = REL * float(np.max(np.abs(x)))
        np.testing.assert_allclose(w[0, 0], wo[0], rtol=0, atol=t)
        np.testing.assert_allclose(v[0], vo, rtol=0, atol=t)
    with pytest.raises(Exception):                               # a call that needs an answer cannot be captured
        with eng.capture():
            eng.universal_threshold(w[0])
    gr.close()
    eng.forward(x, hs, gs, 1, 0, 0, w, v)                       # the ctx is usable after a failed capture


def test_timing_record_of_the_last_call():
    import torch
    eng = vw.Engine.get()
    h, g, _ = filters("db4")
    x = torch.randn((64, 8192), dtype=torch.float64, device="cuda")
    eng.set_option("timing", 1)
    try:
        eng.forward(x, h * S, g * S, 4, 0)
        dev_ms, host_ms, launches = eng.last_timing()
        assert launches >= 1 and dev_ms > 0.0 and host_ms > 0.0
    finally:
        eng.set_option("timing", 0)


def test_two_streams_share_one_ctx_scratch_safely():
    """device-pointer calls return without synchronising and the multi-level paths ping-pong through per-ctx
    scratch; two torch streams issuing back to back must not corrupt each other's intermediate approximations."""
    import torch
    eng = vw.Engine.get()
    h, g, _ = filters("sym8")
    hs, gs = h * S, g * S
    xs = [torch.randn((8, 1 << 16), dtype=torch.float64, device="cuda") for _ in range(2)]
    ref = [eng.forward(x, hs, gs, 8, 0) for x in xs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(5):
        outs = []
        for x, st in zip(xs, streams):
            with torch.cuda.stream(st):
                outs.append(eng.forward(x, hs, gs, 8, 0))
        torch.cuda.synchronize()
        for (w, v), (w0, v0) in zip(outs, ref):
            assert torch.equal(w, w0) and torch.equal(v, v0)


def test_threads_mixing_numpy_and_torch_callers():
    """stream binding and launch are one critical section in the Python binding."""
    import threading

    import torch
    eng = vw.Engine.get()
    h, g, _ = filters("db4")
    hs, gs = h * S, g * S
    xh = np.random.default_rng(5).standard_normal((4, 4096))
    xd = torch.randn((4, 4096), dtype=torch.float64, device="cuda")
    wh0, vh0 = eng.forward(xh, hs, gs, 4, 0)
    wd0, vd0 = eng.forward(xd, hs, gs, 4, 0)
    torch.cuda.synchronize()
    errs = []

    def host():
        for _ in range(50):
            w, v = eng.forward(xh, hs, gs, 4, 0)
            if not (np.array_equal(w, wh0) and np.array_equal(v, vh0)):
                errs.append("host")

    def device():
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(50):
                w, v = eng.forward(xd, hs, gs, 4, 0)
                st.synchronize()
                if not (torch.equal(w, wd0) and torch.equal(v, vd0)):
                    errs.append("device")

    ts = [threading.Thread(target=host), threading.Thread(target=device)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs
