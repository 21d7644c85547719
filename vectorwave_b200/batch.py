"""BatchMODWT / BatchSIMDMODWT facades over the native engine
(EXT/extensions/modwt/BatchMODWT.java:62-178, BatchSIMDMODWT.java:64-81,282-308,343-381).

AoS = [batch][n] rows (the engine's native layout, no transposes needed); the SoA statics keep the
reference's flat `t*batch + b` indexing for drop-in use.  PERIODIC only, like the reference.
"""
import numpy as np

from ._native import Engine, ORDER_PAIR, ORDER_SPLIT
from .errors import ErrorCode, IllegalArgumentException, InvalidArgumentException, NullPointerException
from .modwt import SCALE, _is_torch
from .wavelets import BoundaryMode, Haar

_P = BoundaryMode.PERIODIC.value


class SingleLevelResult:
    """record SingleLevelResult(double[][] approx, double[][] detail) -- BatchMODWT.java:33"""

    def __init__(self, approx, detail):
        self._approx, self._detail = approx, detail

    def approx(self):
        return self._approx

    def detail(self):
        return self._detail


class MultiLevelResult:
    """record MultiLevelResult(double[][][] detailPerLevel, double[][] finalApprox) -- BatchMODWT.java:41"""

    def __init__(self, detailPerLevel, finalApprox):
        self._d, self._a = detailPerLevel, finalApprox

    def detailPerLevel(self):
        return self._d

    def finalApprox(self):
        return self._a


def _validate_aos(signals):
    """BatchMODWT.java:201-212"""
    if signals is None or len(signals) == 0:
        raise IllegalArgumentException("signals must be non-null and non-empty")
    if _is_torch(signals):
        if signals.dim() != 2 or signals.shape[1] == 0:
            raise IllegalArgumentException("signal length must be > 0")
        return signals
    if any(s is None for s in signals):
        raise IllegalArgumentException("all signals must be non-null and same length")
    n = len(signals[0])
    if n == 0:
        raise IllegalArgumentException("signal length must be > 0")
    if any(len(s) != n for s in signals):
        raise IllegalArgumentException("all signals must be non-null and same length")
    return np.asarray(signals, dtype=np.float64)


def _scaled(wavelet, single_level):
    h, g = wavelet.lowPassDecomposition(), wavelet.highPassDecomposition()
    if single_level and isinstance(wavelet, Haar):
        # literal +-0.5 taps of haarBatchMODWTSoA (BatchSIMDMODWT.java:90-93)
        return np.array([0.5, 0.5]), np.array([0.5, -0.5])
    return h * SCALE, g * SCALE


def _check_levels(x, l, levels):
    # the reference has no L_j <= N validation here and its index goes negative; reject like core
    n = x.shape[-1]
    if (l - 1) * (1 << (levels - 1)) + 1 > n:
        raise InvalidArgumentException("Upsampled analysis filter length exceeds signal length", ErrorCode.VAL_TOO_LARGE)


class BatchMODWT:
    @staticmethod
    def singleLevelAoS(wavelet, signals, engine=None):
        x = _validate_aos(signals)
        hs, gs = _scaled(wavelet, True)
        w, v = (engine or Engine.get()).forward(x, hs, gs, 1, _P)
        return SingleLevelResult(v, w[0])

    @staticmethod
    def multiLevelAoS(wavelet, signals, levels, engine=None):
        if levels < 1:
            raise IllegalArgumentException("levels must be >= 1")
        x = _validate_aos(signals)
        hs, gs = _scaled(wavelet, False)
        _check_levels(x, hs.size, levels)
        w, v = (engine or Engine.get()).forward(x, hs, gs, levels, _P)
        return MultiLevelResult(w, v)

    @staticmethod
    def inverseSingleLevelAoS(wavelet, approx, detail, engine=None):
        a, d = _validate_aos(approx), _validate_aos(detail)
        if tuple(a.shape) != tuple(d.shape):
            raise IllegalArgumentException("approx/detail shapes must match")
        hs = wavelet.lowPassReconstruction() * SCALE
        gs = wavelet.highPassReconstruction() * SCALE
        # loops core MODWTTransform.inverse per signal (BatchMODWT.java:122-139): pair-added order
        return (engine or Engine.get()).inverse(d.reshape(1, *d.shape), a, hs, gs, _P, None, ORDER_PAIR)

    @staticmethod
    def inverseMultiLevelAoS(wavelet, detailPerLevel, finalApprox, engine=None):
        if detailPerLevel is None or len(detailPerLevel) == 0:
            raise IllegalArgumentException("levels must be > 0")
        a = _validate_aos(finalApprox)
        d = detailPerLevel if _is_torch(detailPerLevel) else np.asarray(detailPerLevel, dtype=np.float64)
        if d.ndim != 3 or tuple(d.shape[1:]) != tuple(a.shape):
            raise IllegalArgumentException("all detail rows must have consistent length")
        hs = wavelet.lowPassReconstruction() * SCALE
        gs = wavelet.highPassReconstruction() * SCALE
        # loops core MultiLevelMODWTTransform.reconstruct PERIODIC per signal (BatchMODWT.java:151-178): split order
        return (engine or Engine.get()).inverse(d, a, hs, gs, _P, None, ORDER_SPLIT)


def _soa_flat(a, tot, what, writable=False):
    """the caller's flat SoA array as the engine wants it: contiguous float64 of batchSize * signalLength, numpy or CUDA
    tensor; outputs must already be that (they are written in place), inputs are converted if needed"""
    if a is None:
        raise NullPointerException(f"{what} cannot be null")
    if _is_torch(a):
        if a.numel() != tot:
            raise IllegalArgumentException(f"{what} length must be batchSize * signalLength")
        if writable:
            if not a.is_contiguous():
                raise IllegalArgumentException(f"{what} must be contiguous")
            return a.view(-1)
        return a.contiguous().view(-1)
    if writable:
        if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.size == tot):
            raise IllegalArgumentException(f"{what} must be a contiguous float64 array of batchSize * signalLength")
        return a.reshape(-1)
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if a.size != tot:
        raise IllegalArgumentException(f"{what} length must be batchSize * signalLength")
    return a


class BatchSIMDMODWT:
    """SoA statics: flat arrays indexed t*batchSize + b (BatchSIMDMODWT.java:282-308), numpy arrays or CUDA tensors.
    The engine runs the cascade on that layout directly (`vw_modwt_forward_soa`: the flat array is one periodic signal of
    n*B samples at dilation 2^(j-1)*B) -- no AoS <-> SoA transposes, and with CUDA tensors nothing crosses PCIe
    (SURVEY 8f row 2)."""

    @staticmethod
    def convertToSoA(signals, soaOutput=None):
        x = np.asarray(signals, dtype=np.float64)
        soa = np.ascontiguousarray(x.T).ravel()
        if soaOutput is not None:
            soaOutput[...] = soa
            return soaOutput
        return soa

    @staticmethod
    def convertFromSoA(soaData, output):
        b, n = output.shape
        output[...] = np.asarray(soaData).reshape(n, b).T
        return output

    @staticmethod
    def batchMODWTSoA(soaSignals, soaApprox, soaDetail, wavelet, batchSize, signalLength, engine=None):
        """:64-81; outputs written into the caller's SoA arrays."""
        tot = int(batchSize) * int(signalLength)
        hs, gs = _scaled(wavelet, True)
        (engine or Engine.get()).forward_soa(_soa_flat(soaSignals, tot, "soaSignals"), batchSize, signalLength, hs, gs,
                                             [_soa_flat(soaDetail, tot, "soaDetail", True)],
                                             _soa_flat(soaApprox, tot, "soaApprox", True))

    @staticmethod
    def batchMultiLevelMODWTSoA(soaSignals, soaDetailPerLevel, soaApproxOut, wavelet, batchSize, signalLength,
                                levels, engine=None):
        """:343-381"""
        if len(soaDetailPerLevel) != levels:
            raise IllegalArgumentException("soaDetailPerLevel length must equal levels")
        tot = int(batchSize) * int(signalLength)
        hs, gs = _scaled(wavelet, False)
        # the reference has no L_j <= N validation here and its index goes negative; reject like core
        if (hs.size - 1) * (1 << (levels - 1)) + 1 > signalLength:
            raise InvalidArgumentException("Upsampled analysis filter length exceeds signal length", ErrorCode.VAL_TOO_LARGE)
        (engine or Engine.get()).forward_soa(_soa_flat(soaSignals, tot, "soaSignals"), batchSize, signalLength, hs, gs,
                                             [_soa_flat(d, tot, "soaDetailPerLevel", True) for d in soaDetailPerLevel],
                                             _soa_flat(soaApproxOut, tot, "soaApproxOut", True))
