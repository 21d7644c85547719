"""WaveletOperations facade: the three public MODWT convolution primitives and the threshold helpers
(CORE/WaveletOperations.java:29-39,48,59 and softThreshold/hardThreshold; kernels
CORE/internal/ScalarOps.java:700-723,790-808,818-835,1013-1044).  `filter` is the already scaled,
already upsampled dense filter, exactly as the reference's callers pass it; `output` is caller allocated.
The reference's FFT heuristic (N >= 1024 and L_j > N/8) is not reproduced: the engine always does the true
circular convolution."""
import numpy as np

from ._native import Engine
from .errors import NullPointerException
from .wavelets import BoundaryMode


def _conv(signal, filt, output, mode, engine):
    if signal is None or filt is None or output is None:
        raise NullPointerException("signal, filter and output cannot be null")
    (engine or Engine.get()).conv(signal, filt, mode.value, out=output)
    return output


class WaveletOperations:
    @staticmethod
    def circularConvolveMODWT(signal, filter, output, engine=None):
        return _conv(signal, filter, output, BoundaryMode.PERIODIC, engine)

    @staticmethod
    def zeroPaddingConvolveMODWT(signal, filter, output, engine=None):
        return _conv(signal, filter, output, BoundaryMode.ZERO_PADDING, engine)

    @staticmethod
    def symmetricConvolveMODWT(signal, filter, output, engine=None):
        return _conv(signal, filter, output, BoundaryMode.SYMMETRIC, engine)

    @staticmethod
    def softThreshold(coefficients, threshold, engine=None):
        out = np.array(coefficients, dtype=np.float64, copy=True)
        return (engine or Engine.get()).threshold(out, float(threshold), True)

    @staticmethod
    def hardThreshold(coefficients, threshold, engine=None):
        out = np.array(coefficients, dtype=np.float64, copy=True)
        return (engine or Engine.get()).threshold(out, float(threshold), False)
