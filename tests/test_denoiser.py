"""WaveletDenoiser (CORE/denoising/WaveletDenoiser.java) -- SURVEY.md 8f row 4: device reductions + host selectors
against the numpy restatement of the reference."""
import numpy as np
import pytest

from oracle import nptwin
from oracle.javarandom import composite_sin
from oracle.wavelets import filters

REL = 1e-12


def test_oracle_denoiser_reduces_noise_on_a_smooth_signal():
    h, g, wid = filters("db4")
    clean = composite_sin(2048, 3, 0.0)
    noisy = clean + 0.2 * np.random.default_rng(1).standard_normal(2048)
    for method in ("UNIVERSAL", "MINIMAX", "BAYES"):
        den, thrs = nptwin.denoiser_multilevel(noisy, h, g, 4, 0, wid, method, True)
        assert np.sqrt(np.mean((den - clean) ** 2)) < np.sqrt(np.mean((noisy - clean) ** 2))
        assert all(t >= 0 for t in thrs)
    assert nptwin.denoiser_threshold(np.ones(32), 1.0, "MINIMAX") == 0.0      # n <= 32 (:500-501)


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
        d.denoise(rows[0], vw.ThresholdMethod.SURE)
    with pytest.raises(vw.InvalidArgumentException):
        vw.WaveletDenoiser(None, vw.BoundaryMode.PERIODIC)
