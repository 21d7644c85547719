"""VectorWaveSwtAdapter over the native engine (CORE/swt/VectorWaveSwtAdapter.java).

The reference's SWT is its MODWT cascade plus thresholding: forward == MultiLevel decompose (never the
FFT path, :337-394), PERIODIC inverse == the split-order cascade (:444-474), other modes delegate to
MultiLevelMODWTTransform.reconstruct (:435-442).  The <=4-thread executor of the reference has no
equivalent here (the GPU is the parallelism); enableParallel / parallelThreshold are accepted and ignored.
"""
from . import _native
from ._native import Engine
from .errors import ErrorCode, InvalidArgumentException, InvalidSignalException, NullPointerException
from .modwt import (MultiLevelMODWTTransform, MutableMultiLevelMODWTResult, SCALE, _as_signal, _is_torch, _length,
                    multilevel_alignment)
from .wavelets import BoundaryMode


class VectorWaveSwtAdapter:
    DEFAULT_PARALLEL_THRESHOLD = 4096  # :98

    def __init__(self, wavelet, boundaryMode=BoundaryMode.PERIODIC, enableParallel=True,
                 parallelThreshold=DEFAULT_PARALLEL_THRESHOLD, engine=None, flags=0, enforce_level_cap=True):
        if wavelet is None:
            raise NullPointerException("Wavelet cannot be null")
        if boundaryMode is None:
            raise NullPointerException("Boundary mode cannot be null")
        self.wavelet = wavelet
        self.boundaryMode = boundaryMode
        self.enableParallel = enableParallel
        self.parallelThreshold = parallelThreshold
        self._flags = flags
        self._modwt = MultiLevelMODWTTransform(wavelet, boundaryMode, engine, flags, enforce_level_cap)
        self._closed = False

    def _eng(self):
        return self._modwt._eng()

    def getWavelet(self):
        return self.wavelet

    def getBoundaryMode(self):
        return self.boundaryMode

    def forward(self, signal, levels=None):
        """:184-204 -> MutableMultiLevelMODWTResult."""
        if signal is None:
            raise NullPointerException("signal cannot be null")
        if levels is None:
            levels = self._modwt.getMaximumLevels(_length(_as_signal(signal)))
        return self._modwt.decomposeMutable(signal, levels)

    def inverse(self, result):
        """:435-474."""
        if result is None:
            raise NullPointerException("Result cannot be null")
        return self._modwt.reconstruct(result)

    def applyThreshold(self, result, level, threshold, soft):
        """:489-493."""
        if result is None:
            raise NullPointerException("Result cannot be null")
        result.applyThreshold(level, threshold, soft)

    def applyUniversalThreshold(self, result, soft):
        """:505-520: sigma = median|W_1| / 0.6745 (exact selection on the device), thr = sigma*sqrt(2 ln N),
        applied to every detail level; the approximation is untouched."""
        if result is None:
            raise NullPointerException("Result cannot be null")
        thr = self._eng().universal_threshold(result.getMutableDetailCoeffs(1))
        self._eng().threshold(result._w, thr, soft)   # all J rows in one launch
        return thr

    def denoise(self, signal, levels, threshold=-1.0, soft=True):
        """:532-562 as ONE native call (decompose, threshold, reconstruct stay on the device)."""
        if signal is None:
            raise NullPointerException("signal cannot be null")
        x = _as_signal(signal)
        n = _length(x)
        if n == 0:
            raise InvalidSignalException("Signal cannot be empty for SWT", ErrorCode.VAL_EMPTY)
        max_levels = self._modwt.getMaximumLevels(n)
        if levels < 1 or levels > max_levels:
            raise InvalidArgumentException(f"Invalid SWT decomposition levels: {levels} (maximum {max_levels})",
                                           ErrorCode.CFG_INVALID_DECOMPOSITION_LEVEL)
        align, order = multilevel_alignment(self.wavelet, self.boundaryMode, levels)
        out, _ = self._eng().denoise(x, self._modwt._hs, self._modwt._gs, levels, self.boundaryMode.value, align, order,
                                     float(threshold), soft, _native.FLAG_CHECK_FINITE | self._flags)
        return out

    def extractLevel(self, signal, levels, targetLevel):
        """:576-598: keep one detail level (or the approximation when targetLevel == 0), reconstruct."""
        result = self.forward(signal, levels)
        mask = 0 if targetLevel == 0 else (1 << (targetLevel - 1)) if 1 <= targetLevel <= levels else 0
        return self._modwt._reconstruct(result, mask, targetLevel == 0)

    def getCacheStatistics(self):
        return {"filterCacheSize": 0, "parallelExecutorActive": False, "parallelThreshold": self.parallelThreshold}

    def cleanup(self):
        self._closed = True

    def close(self):
        self.cleanup()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
