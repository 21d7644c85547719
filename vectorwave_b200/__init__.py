"""vectorwave_b200 -- B200-native MODWT / SWT engine behind VectorWave's public API for that path.

The product is libvwmodwt.so (hand-written sm_100a CUDA behind the C ABI of include/vw_modwt.h);
this package is the host-side mirror of the reference's Java classes for the path, used by the
tests and the benchmark, and the ctypes twin of the Java FFM binding in java/.
"""
from ._native import Engine, load_library, LIB_PATH  # noqa: F401
from .batch import BatchMODWT, BatchSIMDMODWT, MultiLevelResult, SingleLevelResult  # noqa: F401
from .denoising import ThresholdMethod, ThresholdType, WaveletDenoiser  # noqa: F401
from .errors import (ErrorCode, IllegalArgumentException, InvalidArgumentException, InvalidSignalException,  # noqa: F401
                     NativeEngineError, NullPointerException, WaveletTransformException)
from .modwt import (MODWTResult, MODWTTransform, MODWTTransformFactory, MultiLevelMODWTResult,  # noqa: F401
                    MultiLevelMODWTTransform, MutableMultiLevelMODWTResult, ParallelMultiLevelMODWT,
                    ResidentMultiLevelMODWTResult, SymmetricAlignmentStrategy)
from .ops import WaveletOperations  # noqa: F401
from .streaming import BatchStreamingMODWT  # noqa: F401
from .streaming_core import InvalidStateException, MODWTStreamingTransform  # noqa: F401
from .swt import VectorWaveSwtAdapter  # noqa: F401
from .wavelets import BoundaryMode, Coiflet, Daubechies, Haar, Symlet, Wavelet, get_wavelet  # noqa: F401
