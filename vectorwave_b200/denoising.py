"""WaveletDenoiser over the native engine (CORE/denoising/WaveletDenoiser.java) -- SURVEY.md 8f row 4.

The reference's denoiser is a caller of the MODWT path: decompose, estimate sigma = median(|W_1|) / 0.6745, pick a
threshold per level from (W_j, sigma / sqrt(2^j)) with one of the selectors, threshold the details, reconstruct.  Here
every pass over the coefficients runs on the device (exact median select, mean / variance reductions, the SURE risk
scan, thresholding, the transforms); only the scalar selector formulas are evaluated on the host, exactly as written in
the reference.

SURE: the reference evaluates the risk of every candidate with an O(n^2) double loop (:441-492).  A sort + prefix-sum
evaluation would change the summation order enough to flip the arg-min between neighbouring candidates, so the device
kernel keeps the double loop -- one thread per candidate, the reference's coefficient order and roundings -- and returns
the JVM's threshold bit for bit (n = 65536 costs about a millisecond; the engine refuses batch * n^2 > 2^44).
"""
import enum
import math

import numpy as np

from . import _native
from ._native import ORDER_PAIR
from .errors import ErrorCode, InvalidArgumentException, NativeEngineError
from .modwt import MODWTTransform, MultiLevelMODWTTransform, _as_signal
from .wavelets import BoundaryMode, Daubechies


class ThresholdMethod(enum.Enum):   # :602-632
    UNIVERSAL = "UNIVERSAL"
    SURE = "SURE"
    MINIMAX = "MINIMAX"
    BAYES = "BAYES"
    FIXED = "FIXED"


class ThresholdType(enum.Enum):     # :636-650
    SOFT = "SOFT"
    HARD = "HARD"


MAX_SAFE_LEVEL_FOR_SCALING = 31     # :57
BAYES_EPSILON = 1e-10               # :67


class WaveletDenoiser:
    ThresholdMethod = ThresholdMethod
    ThresholdType = ThresholdType

    def __init__(self, wavelet, boundaryMode, engine=None):
        if wavelet is None:
            raise InvalidArgumentException("wavelet cannot be null", ErrorCode.VAL_NULL_ARGUMENT)
        if boundaryMode is None:
            raise InvalidArgumentException("boundaryMode cannot be null", ErrorCode.VAL_NULL_ARGUMENT)
        self.wavelet, self.boundaryMode, self._engine = wavelet, boundaryMode, engine

    @staticmethod
    def forFinancialData():
        """:99-101"""
        return WaveletDenoiser(Daubechies.DB4, BoundaryMode.PERIODIC)

    # ---- selectors (:391-552), scalar formulas on the host, reductions on the device ---------------------------
    def _threshold(self, eng, coeffs, n, sigma, method):
        if method == ThresholdMethod.UNIVERSAL:
            return sigma * math.sqrt(2.0 * math.log(n))
        if method == ThresholdMethod.MINIMAX:                       # :497-509
            log_n = math.log(n)
            if n <= 32:
                return 0.0
            if n <= 64:
                return sigma * 0.3936 + 0.1829 * sigma * log_n
            return sigma * (0.4745 + 0.1148 * log_n)
        if method == ThresholdMethod.BAYES:                         # :521-552
            sigma2 = sigma * sigma
            _, variance = eng.mean_variance(coeffs)
            sigma_x = math.sqrt(max(0.0, variance - sigma2) + BAYES_EPSILON)
            return sigma2 / sigma_x
        if method == ThresholdMethod.SURE:                          # :441-472
            try:
                return eng.sure_threshold(coeffs, sigma)
            except NativeEngineError as e:                          # n^2 beyond the engine's bound
                raise InvalidArgumentException(str(e), ErrorCode.CFG_UNSUPPORTED_OPERATION) from e
        if method == ThresholdMethod.FIXED:                         # :414-425
            raise InvalidArgumentException("Fixed threshold method requires explicit threshold value",
                                           ErrorCode.CFG_UNSUPPORTED_OPERATION)
        raise InvalidArgumentException("Unknown threshold selection method", ErrorCode.CFG_UNSUPPORTED_OPERATION)

    def _sigma(self, eng, detail):
        """estimateNoiseSigma (:376-387): exact median of |detail| on the device, / 0.6745"""
        return eng.median_abs(detail) / 0.6745

    # ---- API ---------------------------------------------------------------------------------------------
    def denoise(self, signal, method, type=ThresholdType.SOFT):
        """:111-145 single level."""
        t = MODWTTransform(self.wavelet, self.boundaryMode, self._engine)
        eng = t._eng()
        x = _as_signal(signal)
        w, v = eng.forward(x, t._hs, t._gs, 1, self.boundaryMode.value, _native.FLAG_CHECK_FINITE)
        sigma = self._sigma(eng, w[0])
        thr = self._threshold(eng, w[0], w[0].shape[-1], sigma, method)
        return self._finish_single(t, eng, w, v, thr, type)

    def denoiseFixed(self, signal, threshold, type):
        """:354-366"""
        t = MODWTTransform(self.wavelet, self.boundaryMode, self._engine)
        eng = t._eng()
        w, v = eng.forward(_as_signal(signal), t._hs, t._gs, 1, self.boundaryMode.value, _native.FLAG_CHECK_FINITE)
        return self._finish_single(t, eng, w, v, float(threshold), type)

    def _finish_single(self, t, eng, w, v, thr, type):
        eng.threshold(w[0], thr, type == ThresholdType.SOFT)
        align = [(-1, 0, -1, 0)] if self.boundaryMode == BoundaryMode.SYMMETRIC else None   # MODWTTransform.inverse :277-295
        return eng.inverse(w, v, t._hrs, t._grs, self.boundaryMode.value, align, ORDER_PAIR)

    def denoiseMultiLevel(self, signal, levels, method, type):
        """:155-171 + DenoisedMultiLevelResult (:183-231): sigma from level 1, per-level sigma / sqrt(2^j)."""
        t = MultiLevelMODWTTransform(self.wavelet, self.boundaryMode, self._engine)
        res = t.decomposeMutable(signal, levels)
        eng = t._eng()
        if res.getLevels() > MAX_SAFE_LEVEL_FOR_SCALING:
            raise InvalidArgumentException("Decomposition level exceeds safe limit for scale-dependent thresholds",
                                           ErrorCode.VAL_TOO_LARGE)
        sigma = self._sigma(eng, res.getMutableDetailCoeffs(1))
        n = res.getSignalLength()
        self.lastThresholds = []
        for level in range(1, res.getLevels() + 1):
            d = res.getMutableDetailCoeffs(level)
            level_scale = math.sqrt(1 << level)                      # :221
            thr = self._threshold(eng, d, n, sigma / level_scale, method)
            self.lastThresholds.append(thr)
            eng.threshold(d, thr, type == ThresholdType.SOFT)
        return t.reconstruct(res)
