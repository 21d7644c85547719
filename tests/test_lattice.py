"""The lattice form of the long-filter column kernels (csrc/vw_lattice.cu, csrc/vw_column.cu).

CPU part: the host-side factorisation -- which reference tables fit a paraunitary lattice to rounding (coif5 does,
sym8 / db8 / db4 do not and keep the direct form), and a numpy evaluation of the cascade with the returned coefficients
against the oracle's direct sums (analysis and synthesis), so the algebra the kernels implement is pinned without a GPU.
GPU part: the coif5 column levels in lattice form against the oracle in every boundary mode, on ragged lengths, batches,
span calls and the threshold-on-load synthesis, and against the direct-form kernels (option "lattice" = 0)."""
import numpy as np
import pytest

from oracle import cref, nptwin
from oracle.wavelets import TABLES, filters
from vectorwave_b200 import _native

S = nptwin.S
REL = 1e-12


def _lattice(name):
    h, g, _ = filters(name)
    return _native.lattice_query(h * S, g * S)


def test_coif5_fits_a_lattice_to_rounding_and_short_or_inexact_tables_do_not():
    coef, err = _lattice("coif5")
    assert coef is not None and coef.size == 15 + 3
    assert err <= 2e-16                      # every tap of h and g reproduced to half an ulp of the largest tap
    for name in ("db4", "db8", "sym8", "coif2", "coif3", "db10"):
        coef, err = _lattice(name)
        assert coef is None and err > 2e-16, name      # decimal tables that are not orthonormal to rounding


def test_lattice_query_rejects_pairs_that_are_not_quadrature_mirrors():
    h, g, _ = filters("coif5")
    g2 = g.copy()
    g2[3] *= 1.0 + 1e-15
    coef, err = _native.lattice_query(h * S, g2 * S)
    assert coef is None


def _expand(coef):
    """taps of E(z) = S_{K-1} Lam ... S_1 Lam B from the coefficients the kernels get (float64 arithmetic is enough
    to see 1e-15)"""
    k = coef.size - 3
    t, b = coef[:k - 1], coef[k - 1:].reshape(2, 2)
    e = [b.astype(np.longdouble)]
    for tk in t.astype(np.longdouble):
        f = [np.zeros((2, 2), dtype=np.longdouble) for _ in range(len(e) + 1)]
        for n, en in enumerate(e):
            f[n][0, :] = en[0, :]
            f[n + 1][1, :] = en[1, :]
        s = np.array([[1, tk], [-tk, 1]], dtype=np.longdouble)
        e = [s @ fn for fn in f]
    h = np.array([x for en in e for x in (en[0, 0], en[0, 1])])
    g = np.array([x for en in e for x in (en[1, 0], en[1, 1])])
    return h, g


def test_coefficients_expand_back_to_the_reference_table():
    h, g, _ = filters("coif5")
    coef, _ = _lattice("coif5")
    hh, gg = _expand(coef)
    assert float(np.max(np.abs(hh - (h * S).astype(np.longdouble)))) <= 2e-16
    assert float(np.max(np.abs(gg - (g * S).astype(np.longdouble)))) <= 2e-16


def _analysis_rows(u, coef):
    """the per-row recurrence of k_column_analysis_lat on one periodic column"""
    k = coef.size - 3
    t, b = coef[:k - 1], coef[k - 1:]
    n = u.size
    v, w = np.zeros(n), np.zeros(n)
    dl = np.zeros((k - 1, 2))
    uprev = 0.0
    for q in range(-30, n):
        x = u[q % n]
        aa = b[0] * x + b[1] * uprev
        bb = b[2] * x + b[3] * uprev
        uprev = x
        for s in range(k - 1):
            bd = dl[s, q & 1]
            dl[s, q & 1] = bb
            aa, bb = aa + t[s] * bd, bd - t[s] * aa
        if q >= 0:
            v[q], w[q] = aa, bb
    return v, w


def _synthesis_rows(v, w, coef):
    """the per-row recurrence of k_column_synthesis_lat on one periodic column"""
    k = coef.size - 3
    t, b = coef[:k - 1], coef[k - 1:]
    n, l = v.size, 2 * k
    out = np.zeros(n)
    dl = np.zeros((k - 1, 2))
    y0prev = 0.0
    for m in range(n + l - 1):
        aa, bb = v[m % n], w[m % n]
        for s in range(k - 2, -1, -1):
            an, bb = aa - t[s] * bb, bb + t[s] * aa
            aa = dl[s, m & 1]
            dl[s, m & 1] = an
        y0 = b[0] * aa + b[2] * bb
        y1 = b[1] * aa + b[3] * bb
        if m - (l - 1) >= 0:
            out[m - (l - 1)] = y0prev + y1
        y0prev = y0
    return out


def test_numpy_cascade_equals_the_oracle_direct_sums():
    h, g, wid = filters("coif5")
    coef, _ = _lattice("coif5")
    x = np.random.default_rng(5).standard_normal(1536)
    wo, vo = cref.decompose(x, h, g, 1, 0)
    v, w = _analysis_rows(x, coef)
    t = 1e-14 * float(np.max(np.abs(x)))
    assert float(np.max(np.abs(v - vo))) <= t and float(np.max(np.abs(w - wo[0]))) <= t
    ref = cref.reconstruct(wo, vo, h, g, 0, wid)
    assert float(np.max(np.abs(_synthesis_rows(vo, wo[0], coef) - ref))) <= t


# ---- GPU -------------------------------------------------------------------------------------------------------------

@pytest.fixture()
def eng():
    import vectorwave_b200 as vw
    e = vw.Engine.get()
    yield e
    for k, v in (("lattice", 15), ("colmin", 0), ("tile", 0), ("fuse", 0)):
        e.set_option(k, v)


def _both(eng, fn, lattice=3):
    """fn() with the lattice kernels (3: two levels per pass where the plan pairs them, 1: one level per pass) and with
    the direct-form kernels"""
    eng.set_option("lattice", lattice)
    a = fn()
    eng.set_option("lattice", 0)
    b = fn()
    eng.set_option("lattice", 15)
    return a, b


@pytest.mark.gpu
@pytest.mark.parametrize("lattice", [3, 1])
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("b,n,levels", [(1, 65536, 8), (3, 40001, 7), (2, 9000, 5), (1, 1 << 20, 10), (2, 3001, 5), (1, 12346, 6)])
def test_coif5_column_levels_in_lattice_form_against_the_oracle(eng, mode, b, n, levels, lattice):
    import vectorwave_b200 as vw
    from vectorwave_b200.modwt import multilevel_alignment
    h, g, wid = filters("coif5")
    hs, gs = h * S, g * S
    x = np.random.default_rng(n + mode).standard_normal((b, n))
    bm = [vw.BoundaryMode.PERIODIC, vw.BoundaryMode.ZERO_PADDING, vw.BoundaryMode.SYMMETRIC][mode]
    align, order = multilevel_alignment(vw.get_wavelet("coif5"), bm, levels)
    l0 = eng.launch_count()
    (w, v), (wd, vd) = _both(eng, lambda: eng.forward(x, hs, gs, levels, mode), lattice)
    assert eng.launch_count() > l0
    w, v, wd, vd = (np.asarray(a) for a in (w, v, wd, vd))
    t = REL * float(np.max(np.abs(x)))
    rows = range(b) if n <= 70000 else [b - 1]
    for i in rows:
        wo, vo = cref.decompose(x[i], h, g, levels, mode)
        assert float(np.max(np.abs(w[:, i, :] - wo))) <= t
        assert float(np.max(np.abs(v[i] - vo))) <= t
        ref = cref.reconstruct(wo, vo, h, g, mode, wid)
        xr, xrd = _both(eng, lambda: np.asarray(eng.inverse(wo[:, None, :].copy(), vo[None, :].copy(), hs, gs, mode, align, order)), lattice)
        tr = REL * max(float(np.max(np.abs(ref))), float(np.max(np.abs(x))))
        assert float(np.max(np.abs(xr[0] - ref))) <= tr
        assert float(np.max(np.abs(xrd[0] - ref))) <= tr
    # the two forms agree far inside the parity bar (they differ by rounding only)
    assert float(np.max(np.abs(w - wd))) <= 0.05 * t and float(np.max(np.abs(v - vd))) <= 0.05 * t


@pytest.mark.gpu
def test_lattice_is_what_runs_for_coif5_and_only_for_it(eng):
    """A filter whose table fits no lattice must give bit-identical results whether the option is on or off (nothing of
    the lattice path may touch it); coif5 must differ in the last bits (the lattice really ran) yet stay within rounding."""
    x = np.random.default_rng(3).standard_normal((2, 1 << 17))
    for name, same in (("db10", True), ("sym8", True), ("coif5", False)):
        h, g, _ = filters(name)
        (w, v), (wd, vd) = _both(eng, lambda: eng.forward(x, h * S, g * S, 8, 0))
        w, wd = np.asarray(w), np.asarray(wd)
        if same:
            assert np.array_equal(w, wd), name
        else:
            assert not np.array_equal(w, wd)
            assert float(np.max(np.abs(w - wd))) <= 1e-14 * float(np.max(np.abs(x)))


@pytest.mark.gpu
def test_pairs_save_launches_and_singles_take_over_where_pairs_do_not_apply(eng):
    """PERIODIC: levels 3-8 of coif5 run as three pair launches; SYMMETRIC keeps one level per launch (V_j is re-mirrored
    per level, which the in-register cascade cannot do); lattice = 1 switches the pairs off."""
    h, g, _ = filters("coif5")
    x = np.random.default_rng(8).standard_normal((1, 1 << 16))
    counts = {}
    for key, lat, mode in (("pairs", 3, 0), ("singles", 1, 0), ("symmetric", 3, 2)):
        eng.set_option("lattice", lat)
        eng.forward(x, h * S, g * S, 8, mode)          # plan + lattice fit cached
        l0 = eng.launch_count()
        eng.forward(x, h * S, g * S, 8, mode)
        counts[key] = eng.launch_count() - l0
    eng.set_option("lattice", 15)
    assert counts["singles"] - counts["pairs"] == 3 and counts["symmetric"] == counts["singles"]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_coif5_swt_denoise_thresholds_on_load_in_the_lattice_synthesis(eng, mode):
    import vectorwave_b200 as vw
    from vectorwave_b200.modwt import multilevel_alignment
    h, g, wid = filters("coif5")
    n, levels = 1 << 16, 6
    x = np.random.default_rng(11).standard_normal((2, n))
    bm = [vw.BoundaryMode.PERIODIC, vw.BoundaryMode.ZERO_PADDING, vw.BoundaryMode.SYMMETRIC][mode]
    align, order = multilevel_alignment(vw.get_wavelet("coif5"), bm, levels)
    den, thr = eng.denoise(x, h * S, g * S, levels, mode, align, order, -1.0, True)
    den = np.asarray(den)
    for i in range(2):
        dref, tref = cref.swt_denoise(x[i], h, g, levels, mode, wid, -1.0, True)
        assert abs(float(np.asarray(thr)[i]) - tref) <= 1e-12 * tref
        assert float(np.max(np.abs(den[i] - dref))) <= REL * float(np.max(np.abs(x)))


@pytest.mark.gpu
def test_coif5_span_calls_in_lattice_form_equal_the_unsharded_transform(eng):
    """rank-local span call (input = [halo | owned], VW_MODE_LINEAR inside) through the lattice column kernels"""
    import torch
    h, g, _ = filters("coif5")
    hs, gs = h * S, g * S
    n, level = 1 << 17, 6
    x = np.random.default_rng(2).standard_normal(n)
    # V_5 of the whole signal, then level 6 computed on the second half from [halo | owned] of V_5
    w, v = eng.forward(x, hs, gs, level - 1, 0)
    v5 = np.asarray(v).reshape(-1)
    wf, vf = eng.forward(x, hs, gs, level, 0)
    halo = eng.span_halo(30, level, 1)
    lo = n // 2
    ext = torch.as_tensor(np.ascontiguousarray(v5[lo - halo:]), device="cuda")
    ws, vs = eng.forward_span(ext, halo, hs, gs, level, 1)
    torch.cuda.synchronize()
    t = REL * float(np.max(np.abs(x)))
    assert float(np.max(np.abs(ws.cpu().numpy().reshape(-1) - np.asarray(wf)[level - 1].reshape(-1)[lo:]))) <= t
    assert float(np.max(np.abs(vs.cpu().numpy().reshape(-1) - np.asarray(vf).reshape(-1)[lo:]))) <= t


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("name,b,n,levels", [("sym8", 1, 65536, 8), ("db8", 2, 40001, 7), ("coif3", 3, 9000, 6), ("db10", 2, 30002, 7),
                                             ("sym8", 2, 4096, 5)])
def test_direct_form_synthesis_pairs_of_16_to_20_tap_filters_against_the_oracle(eng, mode, name, b, n, levels):
    """16-20-tap quadrature-mirror pairs fit no lattice (decimal tables), so their deep synthesis levels run two per pass in
    direct form (DirSynCore in csrc/vw_column.cu): against the oracle, and against the one-level-per-pass kernels (bit 2 of
    the option off).  SYMMETRIC never pairs (per-level re-mirroring), so there both settings must agree bit for bit."""
    import vectorwave_b200 as vw
    from vectorwave_b200.modwt import multilevel_alignment
    h, g, wid = filters(name)
    hs, gs = h * S, g * S
    x = np.random.default_rng(n + mode + levels).standard_normal((b, n))
    bm = [vw.BoundaryMode.PERIODIC, vw.BoundaryMode.ZERO_PADDING, vw.BoundaryMode.SYMMETRIC][mode]
    align, order = multilevel_alignment(vw.get_wavelet(name), bm, levels)
    w = np.empty((levels, b, n))
    v = np.empty((b, n))
    refs = []
    for i in range(b):
        wo, vo = cref.decompose(x[i], h, g, levels, mode)
        w[:, i, :], v[i] = wo, vo
        refs.append(cref.reconstruct(wo, vo, h, g, mode, wid))
    out = {}
    fwd = {}
    for lat in (15, 3):
        eng.set_option("lattice", lat)
        eng.inverse(w, v, hs, gs, mode, align, order)           # plan cached
        l0 = eng.launch_count()
        out[lat] = (np.asarray(eng.inverse(w, v, hs, gs, mode, align, order)), eng.launch_count() - l0)
        fw, fv = eng.forward(x, hs, gs, levels, mode)
        fwd[lat] = (np.asarray(fw), np.asarray(fv))
    eng.set_option("lattice", 15)
    out[7] = out[15]
    tx = REL * float(np.max(np.abs(x)))
    for lat in (15, 3):                                            # the analysis pairs (bit 3) against the oracle too
        assert float(np.max(np.abs(fwd[lat][0] - w))) <= tx and float(np.max(np.abs(fwd[lat][1] - v))) <= tx
    if mode == 2:
        assert np.array_equal(fwd[15][0], fwd[3][0])
    for i in range(b):
        tr = REL * max(float(np.max(np.abs(refs[i]))), float(np.max(np.abs(x))))
        assert float(np.max(np.abs(out[7][0][i] - refs[i]))) <= tr
        assert float(np.max(np.abs(out[3][0][i] - refs[i]))) <= tr
    if mode == 2:
        assert np.array_equal(out[7][0], out[3][0])
    elif (name, n, levels) == ("sym8", 65536, 8):
        assert out[7][1] < out[3][1]             # config #3's plan: a pair replaces two single-level column launches


@pytest.mark.gpu
def test_direct_form_pairs_threshold_on_load_equals_the_oracle_denoise(eng):
    h, g, wid = filters("db8")
    n, levels = 1 << 16, 7
    x = np.random.default_rng(13).standard_normal((2, n))
    for mode in (0, 1):
        den, thr = eng.denoise(x, h * S, g * S, levels, mode, threshold=-1.0, soft=True)
        den = np.asarray(den)
        for i in range(2):
            dref, tref = cref.swt_denoise(x[i], h, g, levels, mode, wid, -1.0, True)
            assert abs(float(np.asarray(thr)[i]) - tref) <= 1e-12 * tref
            assert float(np.max(np.abs(den[i] - dref))) <= REL * float(np.max(np.abs(x)))


# ---- CPU: the lane-pair recurrences of the two-levels-per-pass kernels, lane by lane, against the oracle ----------------

def _pair_analysis_lattice(x, coef):
    """k_column_analysis_lat2 on the dilation-1 column of a periodic signal: the even lane owns the even rows, the odd
    lane the odd rows; both walk in step, the neighbouring input row crosses by 'shuffle'.  Returns W_1, W_2, V_2."""
    k = coef.size - 3
    t, b = coef[:k - 1], coef[k - 1:]
    n = x.size
    n2 = n // 2
    lead = 48
    w1, w2, v2 = np.zeros(n), np.zeros(n), np.zeros(n)
    dl1 = np.zeros((2, k - 1))
    dl2 = np.zeros((2, k - 1, 2))
    xprev, vprev = [0.0, 0.0], [0.0, 0.0]
    for s in range(-lead, n2):
        xv = [x[(2 * s) % n], x[(2 * s + 1) % n]]
        send = [xv[0], xprev[1]]                  # the even lane sends its current row, the odd lane its previous one
        un = [send[1], send[0]]
        for rho in (0, 1):
            xprev[rho] = xv[rho]
            aa = b[0] * xv[rho] + b[1] * un[rho]
            bb = b[2] * xv[rho] + b[3] * un[rho]
            for i in range(k - 1):
                bd = dl1[rho, i]
                dl1[rho, i] = bb
                aa, bb = aa + t[i] * bd, bd - t[i] * aa
            o1 = bb
            a2 = b[0] * aa + b[1] * vprev[rho]
            b2 = b[2] * aa + b[3] * vprev[rho]
            vprev[rho] = aa
            for i in range(k - 1):
                bd = dl2[rho, i, s & 1]
                dl2[rho, i, s & 1] = b2
                a2, b2 = a2 + t[i] * bd, bd - t[i] * a2
            if s >= 0:
                q = 2 * s + rho
                w1[q], v2[q], w2[q] = o1, a2, b2
    return w1, w2, v2


def _pair_analysis_direct(x, h, g):
    """DirAnaCore of k_column_analysis_pair on the dilation-1 column of a periodic signal: each lane keeps the column around its
    own rows (own rows + the rows received from the partner) and its own V_1 rows.  Returns W_1, W_2, V_2."""
    ln = h.size
    n = x.size
    n2 = n // 2
    lead = 24
    w1, w2, v2 = np.zeros(n), np.zeros(n), np.zeros(n)
    col = [dict(), dict()]        # per lane: index 2s = own row of step s, 2s-1 = the row received at step s
    v1 = [dict(), dict()]
    xprev = [0.0, 0.0]
    for s in range(-lead, n2):
        xv = [x[(2 * s) % n], x[(2 * s + 1) % n]]
        send = [xv[0], xprev[1]]
        un = [send[1], send[0]]
        for rho in (0, 1):
            xprev[rho] = xv[rho]
            col[rho][2 * s - 1], col[rho][2 * s] = un[rho], xv[rho]
            ah = ag = 0.0
            for k in range(ln):
                c = col[rho].get(2 * s - k, 0.0)
                ah += h[k] * c
                ag += g[k] * c
            v1[rho][s] = ah
            bh = bg = 0.0
            for k in range(ln):
                c = v1[rho].get(s - k, 0.0)
                bh += h[k] * c
                bg += g[k] * c
            if s >= 0:
                q = 2 * s + rho
                w1[q], v2[q], w2[q] = ag, bh, bg
    return w1, w2, v2


def _pair_synthesis(v2, w2, w1, row_fn, lag1, lag0):
    """the skeleton of k_column_synthesis_pair on the dilation-1 column of a periodic signal: both lanes consume own row s of
    (V_2, W_2) and own row s - lag1 of W_1 in step; row_fn(rho, s, cv, cw2, cw1, exchange) -> output; lags as in the kernel"""
    n = v2.size
    n2 = n // 2
    out = np.zeros(n)
    steps = n2 + lag1 + lag0 + 2
    for s in range(steps):
        cv = [v2[(2 * s + rho) % n] for rho in (0, 1)]
        cw2 = [w2[(2 * s + rho) % n] for rho in (0, 1)]
        cw1 = [w1[(2 * (s - lag1) + rho) % n] if s >= lag1 else 0.0 for rho in (0, 1)]
        res = row_fn(s, cv, cw2, cw1)
        for rho in (0, 1):
            o = s - (lag1 + lag0 + rho)
            if 0 <= o < n2:
                out[2 * o + rho] = res[rho]
    return out


def _lattice_syn_rows(coef):
    k = coef.size - 3
    t, b = coef[:k - 1], coef[k - 1:]
    dl1 = np.zeros((2, k - 1))
    dl2 = np.zeros((2, k - 1, 2))
    y0prev, z0prev = [0.0, 0.0], [0.0, 0.0]

    def row(s, cv, cw2, cw1):
        z0, z1 = [0.0, 0.0], [0.0, 0.0]
        for rho in (0, 1):
            aa, bb = cv[rho], cw2[rho]
            for i in range(k - 2, -1, -1):
                an, bb = aa - t[i] * bb, bb + t[i] * aa
                aa = dl2[rho, i, s & 1]
                dl2[rho, i, s & 1] = an
            y0 = b[0] * aa + b[2] * bb
            y1 = b[1] * aa + b[3] * bb
            aa = y0prev[rho] + y1
            y0prev[rho] = y0
            bb = cw1[rho]
            for i in range(k - 2, -1, -1):
                an, bb = aa - t[i] * bb, bb + t[i] * aa
                aa = dl1[rho, i]
                dl1[rho, i] = an
            z0[rho] = b[0] * aa + b[2] * bb
            z1[rho] = b[1] * aa + b[3] * bb
        res = [z0[0] + z1[1], z0prev[1] + z1[0]]      # each lane adds the partner's second channel ('shuffle')
        z0prev[0], z0prev[1] = z0[0], z0[1]
        return res
    return row


def _direct_syn_rows(h, g):
    ln = h.size
    hh = ln // 2
    acc2 = [dict(), dict()]
    acc1 = [dict(), dict()]
    v1prev, w1prev = [0.0, 0.0], [0.0, 0.0]

    def row(s, cv, cw2, cw1):
        v1 = [0.0, 0.0]
        for rho in (0, 1):
            for k in range(ln):
                acc2[rho][s - k] = acc2[rho].get(s - k, 0.0) + h[k] * cv[rho] + g[k] * cw2[rho]
            v1[rho] = acc2[rho].pop(s - (ln - 1), 0.0)
        res = [0.0, 0.0]
        for rho in (0, 1):
            ov, ow = (v1prev[1], w1prev[1]) if rho else (v1[0], cw1[0])      # the odd lane applies its own row one step late
            pv, pw = v1[1 - rho], cw1[1 - rho]
            for kk in range(hh):
                j = s - kk
                acc1[rho][j] = acc1[rho].get(j, 0.0) + h[2 * kk] * ov + g[2 * kk] * ow + h[2 * kk + 1] * pv + g[2 * kk + 1] * pw
            res[rho] = acc1[rho].pop(s - (hh - 1), 0.0)
        v1prev[1], w1prev[1] = v1[1], cw1[1]
        return res
    return row


def test_numpy_lane_pair_recurrences_equal_two_oracle_levels():
    """What the pair kernels compute, lane by lane with their lags and exchanges, on a periodic signal whose levels 1-2 are the
    pair (dilation 1): the lattice analysis and synthesis (coif5) and the direct-form synthesis (sym8) against the oracle."""
    n = 1024
    x = np.random.default_rng(17).standard_normal(n)
    t14 = 1e-14 * float(np.max(np.abs(x)))
    h, g, wid = filters("coif5")
    coef, _ = _lattice("coif5")
    wo, vo = cref.decompose(x, h, g, 2, 0)
    w1, w2, v2 = _pair_analysis_lattice(x, coef)
    assert max(float(np.max(np.abs(w1 - wo[0]))), float(np.max(np.abs(w2 - wo[1]))), float(np.max(np.abs(v2 - vo)))) <= t14
    ref = cref.reconstruct(wo, vo, h, g, 0, wid)
    rec = _pair_synthesis(vo, wo[1], wo[0], _lattice_syn_rows(coef), 29, 14)
    assert float(np.max(np.abs(rec - ref))) <= t14
    h, g, wid = filters("sym8")
    wo, vo = cref.decompose(x, h, g, 2, 0)
    ref = cref.reconstruct(wo, vo, h, g, 0, wid)
    rec = _pair_synthesis(vo, wo[1], wo[0], _direct_syn_rows(h * S, g * S), 15, 7)
    assert float(np.max(np.abs(rec - ref))) <= t14
    w1, w2, v2 = _pair_analysis_direct(x, h * S, g * S)
    assert max(float(np.max(np.abs(w1 - wo[0]))), float(np.max(np.abs(w2 - wo[1]))), float(np.max(np.abs(v2 - vo)))) <= t14
