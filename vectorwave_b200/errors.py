"""Exception vocabulary of the reference (CORE/exception/*.java), raised by the host mirror.

ErrorCode values follow CORE/exception/ErrorCode.java:24-118; the native status integers
(include/vw_modwt.h vw_status) map onto them 1:1.
"""
import enum


class ErrorCode(enum.Enum):
    VAL_NULL_ARGUMENT = "VAL_001"
    VAL_NON_FINITE_VALUES = "VAL_003"
    VAL_TOO_LARGE = "VAL_005"
    VAL_EMPTY = "VAL_006"
    VAL_LENGTH_MISMATCH = "VAL_007"
    CFG_UNSUPPORTED_OPERATION = "CFG_001"
    CFG_UNSUPPORTED_BOUNDARY_MODE = "CFG_003"
    CFG_INVALID_DECOMPOSITION_LEVEL = "CFG_004"
    STATE_CLOSED = "STATE_001"
    STATE_INVALID = "STATE_002"


class WaveletTransformException(RuntimeError):
    """CORE/exception/WaveletTransformException.java -- unchecked base class."""

    def __init__(self, message, error_code=None):
        super().__init__(message)
        self.error_code = error_code

    def getErrorCode(self):
        return self.error_code


class InvalidSignalException(WaveletTransformException):
    """CORE/exception/InvalidSignalException.java (VAL_EMPTY, VAL_NON_FINITE_VALUES)."""


class InvalidArgumentException(WaveletTransformException):
    """CORE/exception/InvalidArgumentException.java (CFG_*, VAL_TOO_LARGE)."""


class NullPointerException(TypeError):
    """java.lang.NullPointerException analogue for None arguments (Objects.requireNonNull)."""


class IllegalArgumentException(ValueError):
    """java.lang.IllegalArgumentException analogue (shape errors, EXT/extensions/modwt/BatchMODWT.java:201-212)."""


class NativeEngineError(RuntimeError):
    """CUDA / allocation failures reported by libvwmodwt.so (VW_ECUDA, VW_ENOMEM, VW_EUNSUPPORTED)."""
