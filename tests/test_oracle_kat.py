"""Pins the CPU oracle against the reference's own known answers, formulas and properties
(SURVEY.md 8c).  CPU only.  CTEST/ETEST = the reference's core / extensions test trees."""
import math

import numpy as np
import pytest

from oracle import cref, nptwin
from oracle.javarandom import JavaRandom, composite_sin, uniform_pm1
from oracle.wavelets import TABLES, filters

P, Z, S = cref.PERIODIC, cref.ZERO_PADDING, cref.SYMMETRIC
SQ = 1.0 / math.sqrt(2.0)


def test_java_random_known_values():
    # published java.util.Random outputs
    assert JavaRandom(42)._next(32) == -1170105035
    assert JavaRandom(42).next_double() == 0.7275636800328681
    assert JavaRandom(0).next_double() == 0.730967787376657
    assert abs(JavaRandom(42).next_gaussian() - 1.1419053154730547) < 1e-15


def test_haar_1234_percival_walden():
    # CTEST/modwt/MODWTPercivalWaldenValidationTest.java:40-73
    h, g, _ = filters("haar")
    out = cref.conv([1.0, 2.0, 3.0, 4.0], h * SQ, P)
    np.testing.assert_allclose(out, [2.5, 1.5, 2.5, 3.5], atol=1e-12, rtol=0)
    v, w = cref.forward_single([1.0, 2.0, 3.0, 4.0], h, g, P)
    np.testing.assert_allclose(v, [2.5, 1.5, 2.5, 3.5], atol=1e-12, rtol=0)
    np.testing.assert_allclose(w, [-1.5, 0.5, 0.5, 0.5], atol=1e-12, rtol=0)


def test_time_reversed_filter_kat():
    # ETEST/modwt/TimeReversedFilterTest.java:19-51
    f = [0.7071067811865475, 0.7071067811865475]
    out = cref.conv(np.arange(1.0, 9.0), f, P)
    assert abs(out[0] - (f[0] * 1 + f[1] * 8)) < 1e-10
    assert abs(out[1] - (f[0] * 2 + f[1] * 1)) < 1e-10


def test_half_minus_half_kat():
    # CTEST/modwt/MODWTMathematicalValidationTest.java:296-315
    out = cref.conv([1, 2, 3, 4], [0.5, -0.5], P)
    np.testing.assert_array_equal(out, [-1.5, 0.5, 0.5, 0.5])


def test_level_cap_is_nine():
    # CTEST/modwt/MultiLevelMODWTTransformTest.java:271-305
    assert cref.max_levels(10000, 2) == 9
    assert cref.max_levels(10000, 2, cap=0) == 14      # 2^(j-1)+1 <= 10000 without the cap
    assert cref.max_levels(8, 8) == 0                   # MultiLevelMODWTFilterTruncationTest.java:26-33
    assert cref.max_levels(4096, 8) == 9
    assert cref.max_levels(2 ** 28, 30, cap=0) >= 10    # config #4 needs the uncapped form
    with pytest.raises(ValueError):
        cref.decompose(np.ones(7), *filters("db4")[:2], 1, P)   # L_1=8 > N=7 (VAL_TOO_LARGE, :717-729)


def test_upsample_scale_layout():
    # CORE/internal/ScalarOps.java:909-916
    h = np.array(TABLES["db2"])
    f3 = cref.upsample_scale(h, 3)
    assert f3.size == 3 * 4 + 1
    np.testing.assert_array_equal(f3[::4], h * SQ)
    assert np.count_nonzero(f3) == 4


@pytest.mark.parametrize("mode", [P, Z, S])
@pytest.mark.parametrize("n", [1, 2, 5, 7, 67, 128])
def test_c_matches_numpy_twin_single_level(mode, n):
    rng = np.random.default_rng(n * 3 + mode)
    x = rng.standard_normal(n)
    for name in ("haar", "db4", "sym8", "coif5"):
        h, g, _ = filters(name)
        v, w = cref.forward_single(x, h, g, mode)
        v2, w2 = nptwin.forward_single(x, h, g, mode)
        np.testing.assert_array_equal(v, v2)
        np.testing.assert_array_equal(w, w2)
        for bv in (False, True):
            xr = cref.inverse_single(v, w, h, g, mode, bv)
            xr2 = nptwin.inverse_single(v, w, h, g, mode, bv)
            np.testing.assert_array_equal(xr, xr2)


@pytest.mark.parametrize("mode", [P, Z, S])
@pytest.mark.parametrize("name,n,levels", [("haar", 129, 6), ("db4", 257, 5), ("db8", 500, 5), ("sym8", 512, 5),
                                           ("sym4", 129, 4), ("coif2", 257, 4), ("coif3", 300, 4),
                                           ("db6", 256, 4), ("coif5", 1024, 5), ("db10", 400, 4)])
def test_c_dense_sparse_and_twin_agree_multilevel(mode, name, n, levels):
    h, g, wid = filters(name)
    x = uniform_pm1(n, 123)
    wd, vd = cref.decompose(x, h, g, levels, mode, dense=True)
    ws, vs = cref.decompose(x, h, g, levels, mode, dense=False)
    wt, vt = nptwin.decompose(x, h, g, levels, mode)
    np.testing.assert_array_equal(wd, ws)
    np.testing.assert_array_equal(vd, vs)
    np.testing.assert_array_equal(wd, wt)
    np.testing.assert_array_equal(vd, vt)
    rd = cref.reconstruct(wd, vd, h, g, mode, wid, dense=True)
    rs = cref.reconstruct(wd, vd, h, g, mode, wid, dense=False)
    rt = nptwin.reconstruct(wd, vd, h, g, mode, wid)
    np.testing.assert_array_equal(rd, rs)
    np.testing.assert_array_equal(rd, rt)


@pytest.mark.parametrize("n", [128, 129, 256])
@pytest.mark.parametrize("name", ["haar", "db4"])
def test_periodic_round_trip_single_level(n, name):
    # CTEST/modwt/ModwtPeriodicRoundTripTest.java:25-42 (1e-9, compositeSin(n,42,0.05))
    h, g, _ = filters(name)
    x = composite_sin(n, 42, 0.05)
    v, w = cref.forward_single(x, h, g, P)
    xr = cref.inverse_single(v, w, h, g, P)
    assert np.max(np.abs(xr - x)) < 1e-9


@pytest.mark.parametrize("name,tol", [("haar", 1e-9), ("db4", 1e-9), ("sym4", 1e-6), ("coif2", 5e-4), ("db8", 1e-8),
                                      ("coif5", 1e-8), ("sym8", 1e-6)])
def test_multilevel_pr_and_energy(name, tol):
    # CTEST/modwt/MultiLevelModwtCorrectnessTest.java:27-72 (N=512, J=min(5,max), per-wavelet tolerances)
    h, g, wid = filters(name)
    x = composite_sin(512, 7, 0.1)
    levels = min(5, cref.max_levels(512, len(h)))
    w, v = cref.decompose(x, h, g, levels, P)
    xr = cref.reconstruct(w, v, h, g, P, wid)
    assert np.max(np.abs(xr - x)) < tol
    e_in = float(np.sum(x * x))
    e_out = float(np.sum(w * w) + np.sum(v * v))
    assert abs(e_in - e_out) / e_in < max(tol, 1e-9) * 10


def test_haar_million_sample_round_trip():
    # CTEST/modwt/MultiLevelMODWTOverflowTest.java:99-125 (N=1e6 Haar, PR < 1e-10)
    h, g, _ = filters("haar")
    x = np.random.default_rng(5).standard_normal(1_000_000)
    w, v = cref.decompose(x, h, g, 9, P)
    xr = cref.reconstruct(w, v, h, g, P)
    assert np.max(np.abs(xr - x)) < 1e-10


def test_linearity_and_shift_equivariance():
    # CTEST/modwt/MODWTPercivalWaldenValidationTest.java:312-345; ETEST/modwt/TimeReversedFilterTest.java:76-89
    h, g, _ = filters("db4")
    rng = np.random.default_rng(11)
    a, b = rng.standard_normal(200), rng.standard_normal(200)
    wa, va = cref.decompose(a, h, g, 3, P)
    wb, vb = cref.decompose(b, h, g, 3, P)
    wc, vc = cref.decompose(2.0 * a + 3.0 * b, h, g, 3, P)
    np.testing.assert_allclose(wc, 2 * wa + 3 * wb, atol=1e-12)
    np.testing.assert_allclose(vc, 2 * va + 3 * vb, atol=1e-12)
    ws, vs = cref.decompose(np.roll(a, 5), h, g, 3, P)
    np.testing.assert_allclose(ws, np.roll(wa, 5, axis=1), atol=1e-13)
    np.testing.assert_allclose(vs, np.roll(va, 5), atol=1e-13)


def test_swt_single_element_and_multiwrap():
    # CTEST/modwt/MODWTTransformTest.java:112-122 -- N=1 Haar must round trip (true multi-wrap modulo)
    h, g, _ = filters("haar")
    v, w = cref.forward_single([5.0], h, g, P)
    assert cref.inverse_single(v, w, h, g, P)[0] == pytest.approx(5.0, abs=1e-12)
    h8, g8, _ = filters("db4")
    x = np.array([1.0, -2.0, 3.0])
    v, w = cref.forward_single(x, h8, g8, P)       # L=8 > N=3 wraps more than once
    exp = np.array([sum(h8[k] * SQ * x[(t - k) % 3] for k in range(8)) for t in range(3)])
    np.testing.assert_allclose(v, exp, atol=1e-15)


def test_symmetric_nrmse_within_reference_budget():
    # CTEST/modwt/SymmetricNRMSEBaselineGuardTest.java + test/resources/baselines/symmetric_nrmse_baseline.properties
    budget = {("haar", 129, 6): 1.15, ("haar", 257, 6): 0.90, ("db4", 129, 5): 1.45, ("db4", 257, 5): 1.20,
              ("sym4", 129, 5): 1.60, ("sym4", 257, 5): 1.65, ("coif2", 129, 5): 1.70, ("coif2", 257, 5): 1.65}
    for (name, n, level), base in budget.items():
        h, g, wid = filters(name)
        x = uniform_pm1(n, 123)
        levels = max(1, min(level, cref.max_levels(n, len(h))))
        w, v = cref.decompose(x, h, g, levels, S)
        y = cref.reconstruct(w, v, h, g, S, wid)
        lups = (len(h) - 1) * (1 << max(0, levels - 1)) + 1
        margin = min(n // 4, max(1, lups // 2))
        a, b = x[margin:n - margin], y[margin:n - margin]
        nrmse = math.sqrt(float(np.sum((a - b) ** 2) / np.sum(a * a)))
        assert nrmse <= base * 1.10, (name, n, levels, nrmse)


def test_alignment_table_matches_twin():
    for wid in range(7):
        for l0 in (2, 8, 12, 16, 18, 30):
            for level in range(1, 8):
                c = cref.alignment(wid, l0, level)
                t = nptwin.alignment(wid, l0, level)
                assert (bool(c[0]), c[1], bool(c[2]), c[3]) == t


def test_threshold_rules_and_universal():
    # CORE/modwt/MutableMultiLevelMODWTResult.java:97-118; CORE/swt/VectorWaveSwtAdapter.java:505-520,627-645
    c = np.array([-3.0, -1.0, -0.5, 0.0, 0.5, 1.0, 3.0])
    np.testing.assert_array_equal(cref.threshold(c, 1.0, True), [-2.0, 0.0, 0.0, 0.0, 0.0, 0.0, 2.0])
    np.testing.assert_array_equal(cref.threshold(c, 1.0, False), [-3.0, 0.0, 0.0, 0.0, 0.0, 0.0, 3.0])
    np.testing.assert_array_equal(nptwin.threshold(c, 1.0, True), cref.threshold(c, 1.0, True))
    w1 = np.random.default_rng(3).standard_normal(1000)
    med = np.median(np.abs(w1))
    expect = med / 0.6745 * math.sqrt(2 * math.log(1000))
    assert cref.universal_threshold(w1) == pytest.approx(expect, rel=1e-15)
    assert nptwin.universal_threshold(w1[:999]) == pytest.approx(cref.universal_threshold(w1[:999]), rel=1e-15)


def test_partial_reconstruction_masks():
    # CORE/modwt/MultiLevelMODWTTransform.java:361-446 reconstructFromLevel / reconstructLevels
    h, g, wid = filters("db4")
    x = uniform_pm1(256, 9)
    w, v = cref.decompose(x, h, g, 4, P)
    full = cref.reconstruct(w, v, h, g, P, wid)
    parts = cref.reconstruct(w, v, h, g, P, wid, detail_mask=0, use_approx=True)
    for j in range(4):
        parts = parts + cref.reconstruct(w, v, h, g, P, wid, detail_mask=1 << j, use_approx=False)
    np.testing.assert_allclose(parts, full, atol=1e-12)     # additive MRA


def test_batch_soa_matches_scalar_core():
    # ETEST/modwt/BatchMODWTMultiLevelParityTest.java:24-46 (B=3/4, N=128, J=3, 1e-10) -- bit-equal here
    h, g, _ = filters("db4")
    for b in (3, 4):
        rng = JavaRandom(17)
        x = np.array([[rng.next_double() * 2 - 1 for _ in range(128)] for _ in range(b)])
        soa = np.ascontiguousarray(x.T).ravel()
        w_soa, v_soa = cref.batch_soa_decompose(soa, b, 128, h, g, 3)
        for i in range(b):
            w, v = cref.decompose(x[i], h, g, 3, P, dense=True)
            np.testing.assert_array_equal(w_soa.reshape(3, 128, b)[:, :, i], w)
            np.testing.assert_array_equal(v_soa.reshape(128, b)[:, i], v)
    hh, gh, _ = filters("haar")
    x = uniform_pm1(4 * 64, 3).reshape(4, 64)
    v_soa, w_soa = cref.batch_soa_haar_single(np.ascontiguousarray(x.T).ravel(), 4, 64)
    for i in range(4):
        v, w = cref.forward_single(x[i], hh, gh, P)
        np.testing.assert_allclose(v_soa.reshape(64, 4)[:, i], v, atol=1e-15)
        np.testing.assert_allclose(w_soa.reshape(64, 4)[:, i], w, atol=1e-15)


def test_threaded_batch_equals_serial():
    h, g, wid = filters("sym8")
    x = np.random.default_rng(2).standard_normal((6, 300))
    w1, v1, r1 = cref.batch_fwd_inv(x, h, g, 3, P, wid, dense=True, threads=1)
    w4, v4, r4 = cref.batch_fwd_inv(x, h, g, 3, P, wid, dense=False, threads=4)
    np.testing.assert_array_equal(w1, w4)
    np.testing.assert_array_equal(v1, v4)
    np.testing.assert_array_equal(r1, r4)


def test_denoise_pipeline():
    # CORE/swt/VectorWaveSwtAdapter.java:546-562
    h, g, wid = filters("db8")
    x = composite_sin(1024, 42, 0.2)
    for mode in (P, Z, S):
        out, thr = cref.swt_denoise(x, h, g, 4, mode, wid, -1.0, True)
        w, v = cref.decompose(x, h, g, 4, mode)
        assert thr == pytest.approx(nptwin.universal_threshold(w[0]), rel=1e-15)
        wt = np.stack([nptwin.threshold(w[j], thr, True) for j in range(4)])
        np.testing.assert_array_equal(out, nptwin.reconstruct(wt, v, h, g, mode, wid))
