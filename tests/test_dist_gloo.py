"""world_size-2 gloo tests (CPU) of the multi-process host logic: batch sharding by signal and the span-sharded halo
exchange.  The CUDA engine cannot run here, so the span compute is a numpy stand-in with the engine's span-call
contract (vw_modwt_forward_span / vw_modwt_inverse_span: linear filtering of a [halo|span] / [span|halo] buffer);
what is under test is the exchange pattern, the buffer layout and that the result equals the UNSHARDED transform."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cref, nptwin
from oracle.wavelets import filters
from vectorwave_b200 import BoundaryMode, get_wavelet
from vectorwave_b200.sharded import SpanShardedMODWT, shard_batch


class NumpySpanEngine:
    """Stand-in with the Engine.forward_span / inverse_span contract (test double; CPU tensors)."""

    def span_halo(self, l, first, nlev):
        return (l - 1) * (1 << (first - 1)) * ((1 << nlev) - 1)

    def forward_span(self, vin_ext, halo, hs, gs, first, nlev, flags=0, w_out=None, v_out=None):
        cur = vin_ext.numpy().copy()
        n = cur.size - halo
        for i in range(nlev):
            d = 1 << (first - 1 + i)
            w = nptwin.conv(cur, gs, nptwin.ZERO_PADDING, d)
            cur = nptwin.conv(cur, hs, nptwin.ZERO_PADDING, d)
            w_out[i, :n] = torch.from_numpy(w[halo:])
        v_out[:n] = torch.from_numpy(cur[halo:])
        return w_out, v_out

    def inverse_span(self, vin_ext, w_ext, halo, hs, gs, first, nlev, order=0, flags=0, out=None):
        cur = vin_ext.numpy().copy()
        n = cur.size - halo
        t = np.arange(cur.size)
        for i in range(nlev - 1, -1, -1):
            d = 1 << (first - 1 + i)
            w = w_ext[i].numpy()
            acc = np.zeros_like(cur)
            for k in range(len(hs)):
                idx = t + k * d
                ok = idx < cur.size
                acc = np.where(ok, acc + hs[k] * cur[np.minimum(idx, cur.size - 1)], acc)
            for k in range(len(gs)):
                idx = t + k * d
                ok = idx < cur.size
                acc = np.where(ok, acc + gs[k] * w[np.minimum(idx, cur.size - 1)], acc)
            cur = acc
        out[:n] = torch.from_numpy(cur[:n])
        return out


    # -- the whole-cascade calls of the default path, restated from the contract in include/vw_modwt.h (vw_span_plan,
    #    vw_modwt_forward_span_all, vw_span_pack_inverse / unpack, vw_modwt_inverse_span_all) on top of the calls above
    def forward_span_all(self, xext, plan, hs, gs, w, v, flags=0):
        n, lead, lead_w = int(plan.n_local), int(plan.lead), int(plan.lead_w)
        ng = plan.ngroups_f
        rf = [sum(plan.halo_f[g + 1:ng]) for g in range(ng)]
        cur, have = xext, lead                                   # cur[0] is position -have
        for g in range(ng):
            keep, first, nlev = rf[g], plan.first_f[g], plan.nlev_f[g]
            vout = v if g + 1 == ng else torch.empty(keep + n, dtype=torch.float64)
            self.forward_span(cur, have - keep, hs, gs, first, nlev, w_out=w[first - 1:first - 1 + nlev, lead_w - keep:],
                              v_out=vout)
            cur, have = vout, keep

    @staticmethod
    def _msg_layout(plan):
        ng = plan.ngroups_i
        si_in, acc = [], 0
        for g in range(ng):
            acc += plan.halo_i[g]
            si_in.append(acc)
        pieces = [(None, si_in[ng - 1])]                         # (row or None for V_J, count)
        for g in range(ng):
            pieces += [(plan.first_i[g] - 1 + i, si_in[g]) for i in range(plan.nlev_i[g])]
        return si_in, pieces

    def span_pack_inverse(self, plan, w, v, msg):
        _, pieces = self._msg_layout(plan)
        at, lw = 0, int(plan.lead_w)
        for row, cnt in pieces:
            msg[at:at + cnt] = v[:cnt] if row is None else w[row, lw:lw + cnt]
            at += cnt
        assert at == int(plan.inverse_msg)

    def span_unpack_inverse(self, plan, msg, w, v):
        _, pieces = self._msg_layout(plan)
        at, lw, n = 0, int(plan.lead_w), int(plan.n_local)
        for row, cnt in pieces:
            src = msg[at:at + cnt] if msg is not None else torch.zeros(cnt, dtype=torch.float64)
            if row is None:
                v[n:n + cnt] = src
            else:
                w[row, lw + n:lw + n + cnt] = src
            at += cnt

    def inverse_span_all(self, plan, w, v, hs, gs, order, out, flags=0):
        si_in, _ = self._msg_layout(plan)
        n, lw, ng = int(plan.n_local), int(plan.lead_w), plan.ngroups_i
        vext = v
        for g in range(ng - 1, -1, -1):
            s_in, s_out = si_in[g], si_in[g] - plan.halo_i[g]
            first, nlev = plan.first_i[g], plan.nlev_i[g]
            dst = out if g == 0 else torch.empty(n + s_out, dtype=torch.float64)
            self.inverse_span(vext[:n + s_in], w[first - 1:first - 1 + nlev, lw:lw + n + s_in], s_in - s_out, hs, gs, first, nlev,
                              order, out=dst)
            vext = dst


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, n_total, levels, mode_value, groups_f, groups_i, upfront, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = np.random.default_rng(5).standard_normal(n_total)
        nl = n_total // world
        mode = BoundaryMode(mode_value)
        sh = SpanShardedMODWT(get_wavelet(name), levels, nl, mode, engine=NumpySpanEngine(), groups_forward=groups_f,
                              groups_inverse=groups_i, upfront=upfront)
        res = sh.forward(torch.from_numpy(x[rank * nl:(rank + 1) * nl].copy()))
        xr = sh.inverse(res, order=1 if mode == BoundaryMode.ZERO_PADDING else 0)
        # second inverse must give the same answer (the result object survives an inverse)
        xr2 = sh.inverse(res, order=1 if mode == BoundaryMode.ZERO_PADDING else 0)
        ret[rank] = (res.details().numpy().copy(), res.approximation().numpy().copy(), xr.numpy().copy(),
                     xr2.numpy().copy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("upfront", [True, False])     # one exchange per direction vs one per launch group
@pytest.mark.parametrize("mode", [BoundaryMode.PERIODIC, BoundaryMode.ZERO_PADDING])
@pytest.mark.parametrize("name,n_total,levels,gf,gi", [
    ("db4", 2048, 5, [(1, 3), (4, 2)], [(1, 2), (3, 2), (5, 1)]),
    ("haar", 1024, 6, [(1, 4), (5, 1), (6, 1)], [(1, 4), (5, 2)]),
    ("sym8", 4096, 4, None, None),     # groups from the engine's planner (vw_plan_query)
])
def test_span_sharded_two_ranks_equals_unsharded(mode, name, n_total, levels, gf, gi, upfront):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, name, n_total, levels, mode.value, gf, gi, upfront, ret), nprocs=world, join=True)
    h, g, wid = filters(name)
    x = np.random.default_rng(5).standard_normal(n_total)
    wo, vo = cref.decompose(x, h, g, levels, mode.value)
    ref = cref.reconstruct(wo, vo, h, g, mode.value, wid)
    w = np.concatenate([ret[r][0] for r in range(world)], axis=1)
    v = np.concatenate([ret[r][1] for r in range(world)])
    xr = np.concatenate([ret[r][2] for r in range(world)])
    xr2 = np.concatenate([ret[r][3] for r in range(world)])
    np.testing.assert_allclose(w, wo, rtol=0, atol=1e-12)
    np.testing.assert_allclose(v, vo, rtol=0, atol=1e-12)
    np.testing.assert_allclose(xr, ref, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(xr, xr2)


def test_span_sharded_single_rank_self_wrap():
    h, g, wid = filters("db4")
    x = np.random.default_rng(9).standard_normal(1024)
    sh = SpanShardedMODWT(get_wavelet("db4"), 4, 1024, BoundaryMode.PERIODIC, engine=NumpySpanEngine(), rank=0, world=1)
    res = sh.forward(torch.from_numpy(x.copy()))
    wo, vo = cref.decompose(x, h, g, 4, 0)
    np.testing.assert_allclose(res.details().numpy(), wo, rtol=0, atol=1e-12)
    np.testing.assert_allclose(sh.inverse(res).numpy(), cref.reconstruct(wo, vo, h, g, 0, wid), rtol=0, atol=1e-12)


def test_batch_sharding_by_signal_covers_every_signal_once():
    for batch in (1, 7, 8, 1024, 4097):
        for world in (1, 2, 4, 8):
            spans = [shard_batch(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_span_rejects_halo_larger_than_span():
    with pytest.raises(Exception):
        SpanShardedMODWT(get_wavelet("coif5"), 8, 256, BoundaryMode.PERIODIC, engine=NumpySpanEngine(), rank=0, world=4)
