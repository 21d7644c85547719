"""WaveletDenoiser (CORE/denoising/WaveletDenoiser.java) -- SURVEY.md 8f row 4: device reductions + host selectors
against the numpy restatement of the reference."""
import numpy as np
import pytest

from oracle import cref, nptwin
from oracle.javarandom import composite_sin
from oracle.wavelets import filters

REL = 1e-12


def test_oracle_denoiser_reduces_noise_on_a_smooth_signal():
    h, g, wid = filters("db4")
    clean = composite_sin(2048, 3, 0.0)
    noisy = clean + 0.2 * np.random.default_rng(1).standard_normal(2048)
    for method in ("UNIVERSAL", "MINIMAX", "BAYES", "SURE"):
        den, thrs = nptwin.denoiser_multilevel(noisy, h, g, 4, 0, wid, method, True)
        assert np.sqrt(np.mean((den - clean) ** 2)) < np.sqrt(np.mean((noisy - clean) ** 2))
        assert all(t >= 0 for t in thrs)
    assert nptwin.denoiser_threshold(np.ones(32), 1.0, "MINIMAX") == 0.0      # n <= 32 (:500-501)


def _sure_scalar(coeffs, sigma):
    """calculateSUREThreshold (:441-472) + calculateSURERisk (:477-492) as the plain double loop, scalar Python floats"""
    import math
    n = len(coeffs)
    sorted_abs = sorted(abs(float(c)) for c in coeffs)
    min_risk, best = math.inf, 0.0
    sigma2 = sigma * sigma
    for t in sorted_abs:
        risk = -n * sigma2
        for c in coeffs:
            c = float(c)
            a = abs(c)
            if a <= t:
                risk += c * c
            else:
                risk += sigma2 + (a - t) * (a - t)
        risk /= n
        if risk < min_risk:
            min_risk, best = risk, t
    universal = sigma * math.sqrt(2.0 * math.log(n))
    return (universal if best > universal else best), min_risk


def test_oracle_sure_is_the_reference_double_loop():
    rng = np.random.default_rng(11)
    cases = [(rng.standard_normal(150), 1.0), (rng.standard_normal(97) * 3.0, 0.5),
             (np.concatenate([rng.standard_normal(60) * 0.1, rng.standard_normal(6) * 8.0]), 0.1),
             (np.array([1.0, -1.0, 1.0, 2.0, -2.0, 0.0, 0.0]), 0.7), (np.array([3.0]), 1.0), (np.zeros(5), 0.0)]
    for c, sigma in cases:
        assert nptwin.sure_threshold(c, sigma) == _sure_scalar(c, sigma)
        assert cref.sure_threshold(c, sigma) == _sure_scalar(c, sigma)           # the C restatement, same loops
    big = rng.standard_normal(3000) * np.where(rng.random(3000) < 0.05, 6.0, 1.0)
    assert cref.sure_threshold(big, 1.0) == nptwin.sure_threshold(big, 1.0)


@pytest.mark.gpu
def test_sure_threshold_kernel_is_bit_identical_to_the_double_loop():
    import vectorwave_b200 as vw
    eng = vw.Engine.get()
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 255, 1024, 1025, 4097, 9000):
        rows = rng.standard_normal((3, n)) * np.array([[1.0], [1e-4], [50.0]])
        rows[2, :: 3] = np.round(rows[2, :: 3])                       # ties between candidates
        sig = np.array([1.0, 1.1e-4, 20.0])
        thr, risk = eng.sure_threshold(rows, sig, with_risk=True)
        for b in range(3):
            t_ref, r_ref = nptwin.sure_threshold(rows[b], float(sig[b]))
            assert thr[b] == t_ref and risk[b] == r_ref, (n, b)
    # a sparse signal in noise: the minimum sits well below the universal cap
    c = np.concatenate([rng.standard_normal(4000), rng.standard_normal(96) * 5.0])
    t_ref, r_ref = nptwin.sure_threshold(c, 1.0)
    assert eng.sure_threshold(c, 1.0, with_risk=True) == (t_ref, r_ref)
    assert 0 < t_ref < np.sqrt(2 * np.log(c.size))
    # tiny sigma: the arg-min exceeds the universal threshold and is capped (:465-469)
    assert eng.sure_threshold(c, 0.05) == nptwin.sure_threshold(c, 0.05)[0] == 0.05 * np.sqrt(2.0 * np.log(4096.0))
    # device-resident rows with a leading dimension, scalar sigma broadcast
    import torch
    big = torch.from_numpy(rng.standard_normal((2, 3000))).cuda()
    view = big[:, 100:2100]
    got = eng.sure_threshold(view, 0.9)
    for b in range(2):
        assert got[b] == nptwin.sure_threshold(view[b].cpu().numpy(), 0.9)[0]
    with pytest.raises(vw.NativeEngineError):                           # batch * n^2 > 2^44
        eng.sure_threshold(torch.zeros((1, 1 << 23), dtype=torch.float64, device="cuda"), 1.0)


@pytest.mark.gpu
def test_denoiser_sure_selector():
    """SURE end to end: the thresholds are the oracle's selector applied to the engine's own coefficients bit for bit;
    the denoised signal matches the oracle pipeline run with those thresholds."""
    import vectorwave_b200 as vw
    rng = np.random.default_rng(21)
    for name, n, levels, mode in (("db4", 4096, 4, 0), ("sym8", 2049, 3, 2), ("haar", 1000, 3, 1)):
        bm = [vw.BoundaryMode.PERIODIC, vw.BoundaryMode.ZERO_PADDING, vw.BoundaryMode.SYMMETRIC][mode]
        h, g, wid = filters(name)
        x = composite_sin(n, 42, 0.0) + 0.3 * rng.standard_normal(n)
        tol = REL * float(np.max(np.abs(x)))
        den = vw.WaveletDenoiser(vw.get_wavelet(name), bm)
        got = np.asarray(den.denoiseMultiLevel(x, levels, vw.ThresholdMethod.SURE, vw.ThresholdType.SOFT))
        res = vw.MultiLevelMODWTTransform(vw.get_wavelet(name), bm).decompose(x, levels)
        sigma = nptwin.denoiser_sigma(np.asarray(res.getDetailCoeffsAtLevel(1)))
        for j in range(1, levels + 1):
            wj = np.asarray(res.getDetailCoeffsAtLevel(j))
            assert den.lastThresholds[j - 1] == nptwin.sure_threshold(wj, sigma / np.sqrt(float(1 << j)))[0], (name, j)
        w, v = nptwin.decompose(x, h, g, levels, mode)
        w = np.stack([nptwin.threshold(w[j], den.lastThresholds[j], True) for j in range(levels)])
        ref = nptwin.reconstruct(w, v, h, g, mode, wid)
        assert float(np.max(np.abs(got - ref))) <= tol, name
        got1 = np.asarray(den.denoise(x, vw.ThresholdMethod.SURE, vw.ThresholdType.HARD))
        assert np.all(np.isfinite(got1)) and got1.shape == x.shape


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_denoiser_matches_the_reference_restatement(mode):
    import vectorwave_b200 as vw
    bm = [vw.BoundaryMode.PERIODIC, vw.BoundaryMode.ZERO_PADDING, vw.BoundaryMode.SYMMETRIC][mode]
    rng = np.random.default_rng(5 + mode)
    for name, n, levels in (("db4", 4096, 4), ("haar", 1000, 5), ("sym8", 2049, 3)):
        h, g, wid = filters(name)
        x = composite_sin(n, 42, 0.0) + 0.3 * rng.standard_normal(n)
        tol = REL * float(np.max(np.abs(x)))
        den = vw.WaveletDenoiser(vw.get_wavelet(name), bm)
        for method in (vw.ThresholdMethod.UNIVERSAL, vw.ThresholdMethod.MINIMAX, vw.ThresholdMethod.BAYES):
            for ttype in (vw.ThresholdType.SOFT, vw.ThresholdType.HARD):
                soft = ttype == vw.ThresholdType.SOFT
                ref, thrs = nptwin.denoiser_multilevel(x, h, g, levels, mode, wid, method.value, soft)
                got = np.asarray(den.denoiseMultiLevel(x, levels, method, ttype))
                assert np.allclose(den.lastThresholds, thrs, rtol=1e-12, atol=0)
                assert float(np.max(np.abs(got - ref))) <= tol, (name, method, ttype)
                ref1, _ = nptwin.denoiser_single(x, h, g, mode, method.value, soft)
                assert float(np.max(np.abs(np.asarray(den.denoise(x, method, ttype)) - ref1))) <= tol
        ref_f, _ = nptwin.denoiser_single(x, h, g, mode, None, False, fixed=0.25)
        assert float(np.max(np.abs(np.asarray(den.denoiseFixed(x, 0.25, vw.ThresholdType.HARD)) - ref_f))) <= tol


@pytest.mark.gpu
def test_denoiser_reductions_and_errors():
    import vectorwave_b200 as vw
    eng = vw.Engine.get()
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((5, 10001)) * np.array([[1.0], [1e-3], [1e3], [5.0], [0.1]]) + np.array([[0.0], [1.0], [-7.0], [0.0], [2.0]])
    med = eng.median_abs(rows)
    m, v = eng.mean_variance(rows)
    for i in range(5):
        assert med[i] == np.median(np.abs(rows[i]))
        assert m[i] == pytest.approx(rows[i].mean(), rel=1e-12, abs=1e-15)
        assert v[i] == pytest.approx(rows[i].var(), rel=1e-12)
    d = vw.WaveletDenoiser.forFinancialData()
    assert d.wavelet is vw.Daubechies.DB4 and d.boundaryMode == vw.BoundaryMode.PERIODIC
    with pytest.raises(vw.InvalidArgumentException):
        d.denoise(rows[0], vw.ThresholdMethod.FIXED)
    with pytest.raises(vw.InvalidArgumentException):
        vw.WaveletDenoiser(None, vw.BoundaryMode.PERIODIC)
