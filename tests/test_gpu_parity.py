"""GPU parity tests: the CUDA engine, called through the C ABI (ctypes mirror of the Java classes), against the
CPU oracle on the same seeded inputs.  Integer / index behaviour must be exact; fp64 coefficients must be within
1e-12 * max|x| per level (north_star); with VW_FLAG_BITEXACT they must be bit-identical to the oracle."""
import math

import numpy as np
import pytest

import vectorwave_b200 as vw
from oracle import cref, nptwin
from oracle.javarandom import composite_sin, uniform_pm1
from oracle.wavelets import WAVELET_ID, filters
from vectorwave_b200 import _native

pytestmark = pytest.mark.gpu

BM = vw.BoundaryMode
MODES = [BM.PERIODIC, BM.ZERO_PADDING, BM.SYMMETRIC]
WAVELETS = ["haar", "db2", "db4", "db6", "db8", "db10", "sym4", "sym8", "coif2", "coif3", "coif5"]
REL = 1e-12  # tolerance factor on max|x|, stated by BASELINE.json north_star


def tol(x):
    return REL * max(1.0, float(np.max(np.abs(x))))


def close(a, b, x):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=0, atol=tol(x))


# ---- known answers from the reference's tests -------------------------------------------------------------
def test_haar_1234_known_answer():
    # CTEST/modwt/MODWTPercivalWaldenValidationTest.java:40-73
    r = vw.MODWTTransform(vw.Haar(), BM.PERIODIC).forward([1.0, 2.0, 3.0, 4.0])
    np.testing.assert_allclose(r.approximationCoeffs(), [2.5, 1.5, 2.5, 3.5], atol=1e-12)
    np.testing.assert_allclose(r.detailCoeffs(), [-1.5, 0.5, 0.5, 0.5], atol=1e-12)
    assert r.getSignalLength() == 4


def test_wavelet_operations_primitives():
    # ETEST/modwt/TimeReversedFilterTest.java:19-51; CTEST/modwt/MODWTMathematicalValidationTest.java:296-315
    f = [0.7071067811865475, 0.7071067811865475]
    out = np.empty(8)
    vw.WaveletOperations.circularConvolveMODWT(np.arange(1.0, 9.0), f, out)
    assert abs(out[0] - (f[0] * 1 + f[1] * 8)) < 1e-10 and abs(out[1] - (f[0] * 2 + f[1] * 1)) < 1e-10
    out4 = np.empty(4)
    vw.WaveletOperations.circularConvolveMODWT([1, 2, 3, 4], [0.5, -0.5], out4)
    np.testing.assert_array_equal(out4, [-1.5, 0.5, 0.5, 0.5])
    x = uniform_pm1(300, 4)
    dense = cref.upsample_scale(filters("db4")[0], 4)       # 57 dense taps, as the reference's callers pass it
    for op, mode in ((vw.WaveletOperations.circularConvolveMODWT, 0), (vw.WaveletOperations.zeroPaddingConvolveMODWT, 1),
                     (vw.WaveletOperations.symmetricConvolveMODWT, 2)):
        o = np.empty(300)
        op(x, dense, o)
        close(o, cref.conv(x, dense, mode), x)


# ---- single level ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 64, 67, 129, 500, 4096])
def test_single_level_forward_inverse(mode, n):
    x = np.random.default_rng(n).standard_normal(n)
    for name in ("haar", "db4", "sym8", "coif5"):
        h, g, _ = filters(name)
        t = vw.MODWTTransform(vw.get_wavelet(name), mode)
        r = t.forward(x)
        v, w = cref.forward_single(x, h, g, mode.value)
        close(r.approximationCoeffs(), v, x)
        close(r.detailCoeffs(), w, x)
        xr = t.inverse(r)
        close(xr, cref.inverse_single(r.approximationCoeffs(), r.detailCoeffs(), h, g, mode.value), x)


@pytest.mark.parametrize("mode", MODES)
def test_single_level_bitexact_mode(mode):
    for name, n in (("haar", 5), ("db4", 3), ("db4", 67), ("coif5", 129), ("sym8", 1000)):
        x = uniform_pm1(n, 31)
        h, g, _ = filters(name)
        t = vw.MODWTTransform(vw.get_wavelet(name), mode, flags=_native.FLAG_BITEXACT)
        r = t.forward(x)
        v, w = cref.forward_single(x, h, g, mode.value)
        np.testing.assert_array_equal(r.approximationCoeffs(), v)
        np.testing.assert_array_equal(r.detailCoeffs(), w)
        np.testing.assert_array_equal(t.inverse(r), cref.inverse_single(v, w, h, g, mode.value))


@pytest.mark.parametrize("mode", MODES)
def test_batch_single_level_and_symmetric_dispatch(mode):
    # CORE/modwt/MODWTTransform.java:486-559; B>=4 and n>=64 takes the optimized body whose SYMMETRIC rule is t+l
    h, g, _ = filters("db4")
    t = vw.MODWTTransform(vw.Daubechies.DB4, mode)
    for b, n in ((3, 128), (4, 63), (4, 64), (6, 200)):
        x = np.random.default_rng(b * n).standard_normal((b, n))
        res = t.forwardBatch(list(x))
        assert len(res) == b
        back = t.inverseBatch(res)
        optimized = b >= 4 and n >= 64
        for i in range(b):
            v, w = cref.forward_single(x[i], h, g, mode.value)
            close(res[i].approximationCoeffs(), v, x)
            close(res[i].detailCoeffs(), w, x)
            close(back[i], cref.inverse_single(v, w, h, g, mode.value, batch_variant=optimized), x)
    assert t.forwardBatch([]) == [] and t.inverseBatch([]) == []
    mixed = t.forwardBatch([np.ones(5), np.ones(9)])
    assert [r.getSignalLength() for r in mixed] == [5, 9]


# ---- multi level ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", WAVELETS)
def test_multilevel_decompose_reconstruct(mode, name):
    h, g, wid = filters(name)
    for n in (129, 500, 512, 2048):
        x = uniform_pm1(n, 123)
        t = vw.MultiLevelMODWTTransform(vw.get_wavelet(name), mode)
        levels = min(5, t.getMaximumLevels(n))
        if levels < 1:
            continue
        r = t.decompose(x, levels)
        w, v = cref.decompose(x, h, g, levels, mode.value)
        assert r.getLevels() == levels and r.getSignalLength() == n
        for j in range(1, levels + 1):
            close(r.getDetailCoeffsAtLevel(j), w[j - 1], x)
        close(r.getApproximationCoeffs(), v, x)
        xr = t.reconstruct(r)
        close(xr, cref.reconstruct(w, v, h, g, mode.value, wid), x)


@pytest.mark.parametrize("mode", MODES)
def test_multilevel_bitexact_mode(mode):
    for name, n, levels in (("haar", 129, 6), ("db4", 257, 5), ("db8", 500, 5), ("sym8", 777, 5), ("coif2", 300, 4),
                            ("sym4", 129, 4), ("coif3", 400, 4), ("db6", 256, 4), ("coif5", 1024, 5)):
        h, g, wid = filters(name)
        x = uniform_pm1(n, 77)
        t = vw.MultiLevelMODWTTransform(vw.get_wavelet(name), mode, flags=_native.FLAG_BITEXACT)
        r = t.decompose(x, levels)
        w, v = cref.decompose(x, h, g, levels, mode.value)
        np.testing.assert_array_equal(r._w, w)
        np.testing.assert_array_equal(r._v, v)
        np.testing.assert_array_equal(t.reconstruct(r), cref.reconstruct(w, v, h, g, mode.value, wid))


def test_periodic_perfect_reconstruction_where_the_table_permits():
    # CTEST/modwt/MultiLevelModwtCorrectnessTest.java:27-72; 1e-10 only for tables that are accurate enough
    for name in ("haar", "db2", "db6", "db8", "coif5"):
        t = vw.MultiLevelMODWTTransform(vw.get_wavelet(name), BM.PERIODIC)
        x = composite_sin(512, 7, 0.1)
        r = t.decompose(x, min(5, t.getMaximumLevels(512)))
        assert np.max(np.abs(t.reconstruct(r) - x)) < 1e-10
        e = r.getTotalEnergy()
        assert abs(e - float(np.sum(x * x))) / e < 1e-9
        assert abs(np.sum(r.getRelativeEnergyDistribution()) - 1.0) < 1e-12


def test_partial_reconstruction_and_extract_level():
    # CORE/modwt/MultiLevelMODWTTransform.java:361-446; CORE/swt/VectorWaveSwtAdapter.java:576-598
    h, g, wid = filters("db4")
    x = uniform_pm1(600, 9)
    for mode in MODES:
        t = vw.MultiLevelMODWTTransform(vw.Daubechies.DB4, mode)
        r = t.decompose(x, 4)
        w, v = cref.decompose(x, h, g, 4, mode.value)
        close(t.reconstructFromLevel(r, 3), cref.reconstruct(w, v, h, g, mode.value, wid, detail_mask=0b1100), x)
        close(t.reconstructLevels(r, 2, 3), cref.reconstruct(w, v, h, g, mode.value, wid, detail_mask=0b0110,
                                                              use_approx=False), x)
        close(t.reconstructLevels(r, 3, 4), cref.reconstruct(w, v, h, g, mode.value, wid, detail_mask=0b1100), x)
        swt = vw.VectorWaveSwtAdapter(vw.Daubechies.DB4, mode)
        close(swt.extractLevel(x, 4, 2), cref.reconstruct(w, v, h, g, mode.value, wid, detail_mask=0b0010,
                                                          use_approx=False), x)
        close(swt.extractLevel(x, 4, 0), cref.reconstruct(w, v, h, g, mode.value, wid, detail_mask=0), x)
    with pytest.raises(vw.InvalidArgumentException):
        t.reconstructFromLevel(r, 5)
    with pytest.raises(vw.InvalidArgumentException):
        t.reconstructLevels(r, 3, 2)


# ---- batch facade ------------------------------------------------------------------------------------------------
def test_batch_facade_matches_core():
    # ETEST/modwt/BatchMODWTApiTest.java:17-46, BatchMODWTMultiLevelParityTest.java:24-46 (B=3/4, N=128, J=3, 1e-10)
    for name in ("haar", "db2", "db4"):
        wv = vw.get_wavelet(name)
        h, g, _ = filters(name)
        for b in (3, 4):
            x = uniform_pm1(b * 128, 5 + b).reshape(b, 128)
            sl = vw.BatchMODWT.singleLevelAoS(wv, x)
            ml = vw.BatchMODWT.multiLevelAoS(wv, x, 3)
            for i in range(b):
                v, w = cref.forward_single(x[i], h, g, 0)
                np.testing.assert_allclose(sl.approx()[i], v, atol=1e-10)
                np.testing.assert_allclose(sl.detail()[i], w, atol=1e-10)
                wm, vm = cref.decompose(x[i], h, g, 3, 0)
                close(ml.detailPerLevel()[:, i, :], wm, x)
                close(ml.finalApprox()[i], vm, x)
            xr = vw.BatchMODWT.inverseMultiLevelAoS(wv, ml.detailPerLevel(), ml.finalApprox())
            xs = vw.BatchMODWT.inverseSingleLevelAoS(wv, sl.approx(), sl.detail())
            for i in range(b):
                close(xr[i], cref.reconstruct(ml.detailPerLevel()[:, i, :], ml.finalApprox()[i], h, g, 0), x)
                close(xs[i], cref.inverse_single(sl.approx()[i], sl.detail()[i], h, g, 0), x)
    # SoA statics
    x = uniform_pm1(4 * 64, 2).reshape(4, 64)
    soa = vw.BatchSIMDMODWT.convertToSoA(x)
    d = [np.empty(256) for _ in range(2)]
    a = np.empty(256)
    vw.BatchSIMDMODWT.batchMultiLevelMODWTSoA(soa, d, a, vw.Daubechies.DB2, 4, 64, 2)
    w_soa, v_soa = cref.batch_soa_decompose(soa, 4, 64, *filters("db2")[:2], 2)
    close(np.stack(d), w_soa, x)
    close(a, v_soa, x)
    # the same statics on device-resident SoA buffers (transposes on the device, nothing crosses PCIe)
    import torch
    soa_d = torch.as_tensor(soa, device="cuda")
    d_d = [torch.empty(256, dtype=torch.float64, device="cuda") for _ in range(2)]
    a_d = torch.empty(256, dtype=torch.float64, device="cuda")
    vw.BatchSIMDMODWT.batchMultiLevelMODWTSoA(soa_d, d_d, a_d, vw.Daubechies.DB2, 4, 64, 2)
    close(np.stack([t.cpu().numpy() for t in d_d]), w_soa, x)
    close(a_d.cpu().numpy(), v_soa, x)
    s_a, s_d = torch.empty_like(a_d), torch.empty_like(a_d)
    vw.BatchSIMDMODWT.batchMODWTSoA(soa_d, s_a, s_d, vw.Haar.INSTANCE if hasattr(vw.Haar, "INSTANCE") else vw.get_wavelet("haar"), 4, 64)
    close(s_a.cpu().numpy().reshape(64, 4).T, np.stack([cref.forward_single(x[i], np.array([0.5, 0.5]) / nptwin.S, np.array([0.5, -0.5]) / nptwin.S, 0)[0] for i in range(4)]), x)
    with pytest.raises(vw.IllegalArgumentException):
        vw.BatchMODWT.multiLevelAoS(vw.Daubechies.DB4, x, 0)
    with pytest.raises(vw.IllegalArgumentException):
        vw.BatchMODWT.singleLevelAoS(vw.Daubechies.DB4, [])
    with pytest.raises(vw.IllegalArgumentException):
        vw.BatchMODWT.singleLevelAoS(vw.Daubechies.DB4, [np.ones(4), np.ones(5)])


# ---- SWT adapter -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", MODES)
def test_swt_adapter_equals_modwt_and_denoise(mode):
    # CTEST/swt/SwtAdapterParityTest.java:28-56 (1e-10); CORE/swt/VectorWaveSwtAdapter.java:505-562
    h, g, wid = filters("db8")
    x = composite_sin(4096, 42, 0.2)
    with vw.VectorWaveSwtAdapter(vw.Daubechies.DB8, mode) as swt:
        r = swt.forward(x, 5)
        w, v = cref.decompose(x, h, g, 5, mode.value)
        close(r._w, w, x)
        close(swt.inverse(r), cref.reconstruct(w, v, h, g, mode.value, wid), x)
        for soft in (True, False):
            r = swt.forward(x, 5)
            thr = swt.applyUniversalThreshold(r, soft)
            assert thr == pytest.approx(cref.universal_threshold(w[0]), rel=1e-12)
            for j in range(5):
                close(r.getMutableDetailCoeffs(j + 1), cref.threshold(w[j], thr, soft), x)
            out, used = cref.swt_denoise(x, h, g, 5, mode.value, wid, -1.0, soft)
            close(swt.denoise(x, 5, -1.0, soft), out, x)
            out2, _ = cref.swt_denoise(x, h, g, 5, mode.value, wid, 0.3, soft)
            close(swt.denoise(x, 5, 0.3, soft), out2, x)
        r = swt.forward(x, 3)
        swt.applyThreshold(r, 2, 0.25, True)
        close(r.getMutableDetailCoeffs(2), cref.threshold(w[1], 0.25, True), x)
        swt.applyThreshold(r, 0, 0.1, False)
        close(r.getMutableApproximationCoeffs(), cref.threshold(cref.decompose(x, h, g, 3, mode.value)[1], 0.1, False), x)


def test_exact_median_selection_edge_cases():
    eng = vw.Engine.get()
    rng = np.random.default_rng(8)
    for n in (1, 2, 3, 4, 5, 100, 101, 4096, 100001):
        w1 = rng.standard_normal(n)
        assert eng.universal_threshold(w1) == pytest.approx(cref.universal_threshold(w1), rel=1e-14, abs=0)
    w1 = np.array([0.0, -0.0, 3.0, -3.0, 3.0, 1e-310, -1e-310, 2.5])      # ties, signed zeros, denormals
    assert eng.universal_threshold(w1) == pytest.approx(cref.universal_threshold(w1), rel=1e-14)
    rows = rng.standard_normal((7, 1000))
    got = eng.universal_threshold(rows)
    for i in range(7):
        assert got[i] == pytest.approx(cref.universal_threshold(rows[i]), rel=1e-14)
    # candidate-buffer overflow (every key in one 24-bit bucket) falls back to full-row passes: constant rows, heavy ties
    for w1 in (np.full(200000, 1.2345), np.repeat(rng.standard_normal(7), 30000),
               np.concatenate([np.full(100000, 1.0), np.full(100000, 2.0)]),            # the two middle ranks in different buckets
               np.concatenate([np.full(70000, 0.5), 0.5 + 1e-13 * rng.random(60001), np.full(70000, 3.0)]),
               rng.standard_normal(1 << 22) * 1e-3, np.abs(rng.standard_cauchy(300001))):
        assert eng.universal_threshold(w1) == pytest.approx(cref.universal_threshold(w1), rel=1e-14, abs=0)
    rows = np.stack([rng.standard_normal(65536), np.full(65536, -7.0), rng.standard_normal(65536) * 1e200,
                     np.where(rng.random(65536) < 0.5, 1.0, 1.0 + 2.0 ** -30)])
    got = eng.universal_threshold(rows)
    for i in range(rows.shape[0]):
        assert got[i] == pytest.approx(cref.universal_threshold(rows[i]), rel=1e-14, abs=0)


# ---- device-resident path ------------------------------------------------------------------------------------------
def test_device_tensor_path_matches_host_path():
    import torch
    x = np.random.default_rng(3).standard_normal((5, 3000))
    xd = torch.as_tensor(x, device="cuda")
    h, g, wid = filters("sym8")
    eng = vw.Engine.get()
    hs, gs = h * nptwin.S, g * nptwin.S
    for mode in MODES:
        wd, vd = eng.forward(xd, hs, gs, 4, mode.value)
        assert wd.is_cuda and tuple(wd.shape) == (4, 5, 3000)
        for i in range(5):
            w, v = cref.decompose(x[i], h, g, 4, mode.value)
            close(wd[:, i, :].cpu().numpy(), w, x)
            close(vd[i].cpu().numpy(), v, x)
    t = vw.MultiLevelMODWTTransform(vw.Symlet.SYM8, BM.PERIODIC)
    r = t.decompose(xd[0], 4)
    xr = t.reconstruct(r)
    assert xr.is_cuda
    w, v = cref.decompose(x[0], h, g, 4, 0)
    close(xr.cpu().numpy(), cref.reconstruct(w, v, h, g, 0, wid), x)
    assert r.getDetailEnergyAtLevel(2) == pytest.approx(float(np.sum(w[1] ** 2)), rel=1e-12)


# ---- error behaviour (same exceptions as the reference) ---------------------------------------------------------------
def test_error_behaviour():
    t = vw.MODWTTransform(vw.Haar(), BM.PERIODIC)
    with pytest.raises(vw.NullPointerException):
        t.forward(None)
    with pytest.raises(vw.InvalidSignalException) as e:
        t.forward([])
    assert e.value.getErrorCode() == vw.ErrorCode.VAL_EMPTY
    for bad in (float("nan"), float("inf")):
        with pytest.raises(vw.InvalidSignalException) as e:
            t.forward([1.0, bad, 3.0])
        assert e.value.getErrorCode() == vw.ErrorCode.VAL_NON_FINITE_VALUES
    with pytest.raises(vw.InvalidArgumentException) as e:
        vw.MODWTTransform(vw.Haar(), BM.CONSTANT)
    assert e.value.getErrorCode() == vw.ErrorCode.CFG_UNSUPPORTED_BOUNDARY_MODE
    with pytest.raises(vw.NullPointerException):
        vw.MODWTTransform(None, BM.PERIODIC)
    with pytest.raises(vw.NullPointerException):
        t.inverse(None)
    with pytest.raises(vw.IllegalArgumentException):
        vw.MODWTResult.create([1.0, 2.0], [1.0])
    ml = vw.MultiLevelMODWTTransform(vw.Haar(), BM.PERIODIC)
    # CTEST/modwt/MultiLevelMODWTTransformTest.java:268-305: cap 9, max+1 throws
    assert ml.getMaximumLevels(10000) == 9
    with pytest.raises(vw.InvalidArgumentException) as e:
        ml.decompose(np.ones(10000), 10)
    assert e.value.getErrorCode() == vw.ErrorCode.CFG_INVALID_DECOMPOSITION_LEVEL
    with pytest.raises(vw.InvalidArgumentException):
        ml.decompose(np.ones(100), 0)
    with pytest.raises(vw.InvalidSignalException):
        ml.decompose(np.array([]), 1)
    with pytest.raises(vw.InvalidSignalException):
        ml.decompose(np.array([1.0, float("nan"), 2.0, 3.0]), 1)
    # CTEST/modwt/MultiLevelMODWTFilterTruncationTest.java:26-33: DB4 at N=8 has no admissible level
    with pytest.raises(vw.InvalidArgumentException):
        vw.MultiLevelMODWTTransform(vw.Daubechies.DB4, BM.PERIODIC).decompose(np.ones(8), 1)
    # opt-out of the cap: J=10 Haar on 10000 samples matches the uncapped oracle
    free = vw.MultiLevelMODWTTransform(vw.Haar(), BM.PERIODIC, enforce_level_cap=False)
    x = uniform_pm1(10000, 1)
    r = free.decompose(x, 10)
    w, v = cref.decompose(x, *filters("haar")[:2], 10, 0)
    close(r._w, w, x)
    # native status for L_j > n
    eng = vw.Engine.get()
    with pytest.raises(vw.InvalidArgumentException) as e:
        eng.forward(np.ones(20), np.ones(8), np.ones(8), 3, 0)
    assert e.value.getErrorCode() == vw.ErrorCode.VAL_TOO_LARGE


def test_parallel_multilevel_variant_and_factory():
    # CORE/modwt/ParallelMultiLevelMODWT.java:84-176 == the sequential cascade (ParallelVsSequentialEquivalenceTest.java
    # :18-49, 1e-12, N in {256, 500}, PERIODIC + ZERO_PADDING); SYMMETRIC is computed as ZERO_PADDING there (D6)
    h, g, _ = filters("db4")
    for n in (256, 500):
        x = composite_sin(n, 11, 0.2)
        par = vw.ParallelMultiLevelMODWT()
        for mode in MODES:
            r = par.decompose(x, vw.Daubechies.DB4, mode, 3)
            eff = 0 if mode == BM.PERIODIC else 1
            w, v = cref.decompose(x, h, g, 3, eff)
            close(r._w, w, x)
            close(r._v, v, x)
        with pytest.raises(vw.InvalidArgumentException):
            par.decompose(x, vw.Daubechies.DB4, BM.PERIODIC, 0)
    t = vw.MODWTTransformFactory.createMultiLevel(vw.Daubechies.DB4)
    assert t.getBoundaryMode() == BM.PERIODIC and isinstance(vw.MODWTTransformFactory.create(vw.Haar.INSTANCE if hasattr(vw.Haar, "INSTANCE") else vw.get_wavelet("haar")), vw.MODWTTransform)


@pytest.mark.parametrize("name,b,n,levels", [("db4", 64, 512, 4), ("haar", 100, 300, 3), ("sym8", 48, 1000, 5),
                                             ("coif5", 33, 2048, 4), ("db2", 3, 128, 3), ("db4", 8, 256, 3)])
def test_soa_layout_runs_natively(name, b, n, levels):
    """BatchSIMDMODWT.batchMultiLevelMODWTSoA on the caller's [t*B+b] layout (no transposes): wide batches take the
    column kernels at dilation 2^(j-1)*B (any integer B), narrow ones the per-level kernels; host and device buffers."""
    import torch
    h, g, _ = filters(name)
    x = np.random.default_rng(b * 1000 + n).standard_normal((b, n))
    soa = vw.BatchSIMDMODWT.convertToSoA(x)
    w_ref, v_ref = nptwin.decompose(x, h, g, levels, 0)                       # [J][B][N], [B][N]
    w_ref_soa = np.ascontiguousarray(np.transpose(w_ref, (0, 2, 1))).reshape(levels, -1)
    v_ref_soa = np.ascontiguousarray(v_ref.T).ravel()
    wv = vw.get_wavelet(name)
    d = [np.empty(b * n) for _ in range(levels)]
    a = np.empty(b * n)
    vw.BatchSIMDMODWT.batchMultiLevelMODWTSoA(soa, d, a, wv, b, n, levels)
    close(np.stack(d), w_ref_soa, x)
    close(a, v_ref_soa, x)
    eng = vw.Engine.get()
    soa_d = torch.as_tensor(soa, device="cuda")
    d_d = [torch.full((b * n,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(levels)]
    a_d = torch.full((b * n,), float("nan"), dtype=torch.float64, device="cuda")
    l0 = eng.launch_count()
    vw.BatchSIMDMODWT.batchMultiLevelMODWTSoA(soa_d, d_d, a_d, wv, b, n, levels)
    assert eng.launch_count() - l0 == levels                                # one kernel per level, nothing else
    close(np.stack([t.cpu().numpy() for t in d_d]), w_ref_soa, x)
    close(a_d.cpu().numpy(), v_ref_soa, x)
    assert torch.equal(soa_d.cpu(), torch.as_tensor(soa))                    # input untouched
    # bit-exact order on the SoA layout == the reference's SoA loop (same ascending-tap sums per lane)
    S = 1.0 / math.sqrt(2.0)
    eng.forward_soa(soa, b, n, np.asarray(h) * S, np.asarray(g) * S, d, a, flags=_native.FLAG_BITEXACT)
    w_c, v_c = cref.batch_soa_decompose(soa, b, n, h, g, levels)
    np.testing.assert_array_equal(np.stack(d), w_c)
    np.testing.assert_array_equal(a, v_c)
    with pytest.raises(vw.IllegalArgumentException):
        vw.BatchSIMDMODWT.batchMultiLevelMODWTSoA(soa, d[:-1], a, wv, b, n, levels)
    with pytest.raises(vw.IllegalArgumentException):
        vw.BatchSIMDMODWT.batchMultiLevelMODWTSoA(soa, d, np.empty(b * n + 1), wv, b, n, levels)


def test_one_context_shared_by_host_threads():
    """The reference's transform objects are shared between threads (ConcurrentExecutionIntegrationTest): calls on one
    ctx from several host threads serialise on the ctx mutex.  Raw C-ABI calls here -- the Python mirror's own call lock
    is bypassed on purpose, and ctypes drops the GIL for the duration of each call."""
    import ctypes as C
    import threading
    eng = vw.Engine.get()
    S = 1.0 / math.sqrt(2.0)
    dp = C.POINTER(C.c_double)
    eng.lib.vw_reset_stream(eng.ctx)
    jobs = []
    for i, (name, b, n, levels) in enumerate([("db4", 5, 1024, 4), ("haar", 3, 4096, 5), ("sym8", 7, 2048, 3), ("coif5", 2, 8192, 4)]):
        h, g, _ = filters(name)
        x = np.random.default_rng(100 + i).standard_normal((b, n))
        w_ref, v_ref = nptwin.decompose(x, h, g, levels, 0)
        jobs.append((x, np.asarray(h) * S, np.asarray(g) * S, levels, w_ref, v_ref))
    errors = []

    def worker(job):
        x, hs, gs, levels, w_ref, v_ref = job
        b, n = x.shape
        for _ in range(25):
            w = np.full((levels, b, n), np.nan)
            v = np.full((b, n), np.nan)
            rc = eng.lib.vw_modwt_forward(eng.ctx, x.ctypes.data_as(C.c_void_p), b, n, n, hs.ctypes.data_as(dp),
                                          gs.ctypes.data_as(dp), hs.size, levels, 0, w.ctypes.data_as(C.c_void_p), n, b * n,
                                          v.ctypes.data_as(C.c_void_p), n, 0)
            if rc != 0 or not (np.max(np.abs(w - w_ref)) <= tol(x) and np.max(np.abs(v - v_ref)) <= tol(x)):
                errors.append((rc, x.shape))
                return

    threads = [threading.Thread(target=worker, args=(j,)) for j in jobs for _ in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
