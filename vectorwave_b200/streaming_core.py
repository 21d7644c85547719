"""MODWTStreamingTransform over the native engine (CORE/modwt/streaming/MODWTStreamingTransform.java,
MODWTStreamingTransformImpl.java, MultiLevelMODWTStreamingTransform.java) -- SURVEY.md 8f row 1.

The reference pushes samples one at a time into a ring and transforms a window whenever `bufferSize` samples are waiting:
single level = sliding windows that overlap by L-1 samples (Impl :186-229), multi level = back-to-back windows, one
MODWTResult published per level with the approximation only on the last (MultiLevel :163-190).  Every window is an
ordinary `forward` / `decompose` with the transform's boundary mode applied to the window itself.

Here `process(data)` finds every window the new samples complete and runs them as ONE batched engine call (rows =
windows), then publishes the results in window order -- same windows, same coefficients, one launch instead of one JVM
transform per window.  Subscribers are called synchronously (the reference's SubmissionPublisher delivers in the same
order, asynchronously).
"""
import time

import numpy as np

from ._native import FLAG_CHECK_FINITE
from .errors import ErrorCode, InvalidArgumentException, InvalidSignalException, WaveletTransformException
from .modwt import MODWTTransform, MultiLevelMODWTTransform


class InvalidStateException(WaveletTransformException):
    """CORE/exception/InvalidStateException.java"""

    @staticmethod
    def closed(what):
        return InvalidStateException(f"{what} is closed", ErrorCode.STATE_CLOSED)


class StreamingResult:
    """What subscribers receive: MODWTResult's accessors over the window's coefficients (numpy, defensive copies).
    Multi-level windows publish one per level with an EMPTY approximation except on the last level
    (MultiLevelMODWTStreamingTransform.java:281-283)."""

    def __init__(self, approx, details):
        self._a, self._d = approx, details

    def approximationCoeffs(self):
        return self._a.copy()

    def detailCoeffs(self):
        return self._d.copy()

    def getSignalLength(self):
        return int(self._a.size)      # MODWTResultWrapper.getSignalLength = approx.length (:305-307)

    def isValid(self):
        return bool(np.all(np.isfinite(self._a)) and np.all(np.isfinite(self._d)))


class StreamingStatistics:
    """MODWTStreamingTransform.StreamingStatistics (:153-189)"""

    def __init__(self):
        self.reset()

    def reset(self):
        self._samples = self._blocks = self._total_ns = self._max_ns = 0
        self._min_ns = None
        self._start = time.perf_counter_ns()

    def _add_samples(self, count):
        self._samples += count

    def _record_blocks(self, count, total_ns):
        if count <= 0:
            return
        per = total_ns // count
        self._blocks += count
        self._total_ns += total_ns
        self._max_ns = max(self._max_ns, per)
        self._min_ns = per if self._min_ns is None else min(self._min_ns, per)

    def getSamplesProcessed(self):
        return self._samples

    def getBlocksProcessed(self):
        return self._blocks

    def getAverageProcessingTimeNanos(self):
        return self._total_ns // self._blocks if self._blocks else 0

    def getMaxProcessingTimeNanos(self):
        return self._max_ns

    def getMinProcessingTimeNanos(self):
        return 0 if self._min_ns is None else self._min_ns

    def getThroughputSamplesPerSecond(self):
        dt = (time.perf_counter_ns() - self._start) * 1e-9
        return self._samples / dt if dt > 0 else 0.0


class MODWTStreamingTransform:
    """Factory + the Flow.Publisher surface shared by both implementations (MODWTStreamingTransform.java:67-96)."""

    @staticmethod
    def create(wavelet, boundaryMode, bufferSize=256, engine=None):
        return _SingleLevelStreaming(wavelet, boundaryMode, bufferSize, engine)

    @staticmethod
    def createMultiLevel(wavelet, boundaryMode, bufferSize, levels, engine=None):
        return _MultiLevelStreaming(wavelet, boundaryMode, bufferSize, levels, engine)


class _StreamingBase:
    def __init__(self, wavelet, boundaryMode, bufferSize):
        if wavelet is None:
            raise InvalidArgumentException("Wavelet cannot be null")
        if boundaryMode is None:
            raise InvalidArgumentException("Boundary mode cannot be null")
        if bufferSize <= 0:
            raise InvalidArgumentException(f"Buffer size must be positive, got: {bufferSize}")
        self.wavelet, self.boundaryMode, self.bufferSize = wavelet, boundaryMode, int(bufferSize)
        self._tail = np.empty(0)              # the samples waiting in the ring, oldest first
        self._closed = False
        self._stats = StreamingStatistics()
        self._subscribers = []

    # ---- Flow.Publisher -------------------------------------------------------------------------------------
    def subscribe(self, subscriber):
        """subscriber: a callable(result) or an object with onNext (and optionally onSubscribe / onComplete / onError)"""
        self._subscribers.append(subscriber)
        if hasattr(subscriber, "onSubscribe"):
            subscriber.onSubscribe(self)

    def _submit(self, result):
        for s in self._subscribers:
            (s.onNext if hasattr(s, "onNext") else s)(result)

    def _complete(self):
        for s in self._subscribers:
            if hasattr(s, "onComplete"):
                s.onComplete()

    # ---- MODWTStreamingTransform ----------------------------------------------------------------------------
    def process(self, data):
        if self._closed:
            raise InvalidStateException.closed("Transform")
        if data is None or len(data) == 0:
            raise InvalidSignalException("Data cannot be null or empty")
        self._push(np.ascontiguousarray(data, dtype=np.float64).reshape(-1))

    def processSample(self, sample):
        if self._closed:
            raise InvalidStateException.closed("Transform")
        self._push(np.array([float(sample)]))

    def getStatistics(self):
        return self._stats

    def getBufferLevel(self):
        return int(self._tail.size)

    def isClosed(self):
        return self._closed

    def reset(self):
        if self._closed:
            raise InvalidStateException.closed("Transform")
        self._tail = np.empty(0)
        self._stats.reset()

    def _push(self, data):
        buf = np.concatenate([self._tail, data]) if self._tail.size else data
        bs, hop = self.bufferSize, self._hop
        nw = 0 if buf.size < bs else (buf.size - bs) // hop + 1
        # The reference has already stored the samples in its ring when a window fails (non-finite values, level
        # errors): commit the tail -- with the attempted windows consumed -- whether or not _run raises
        self._tail = buf[nw * hop:].copy()
        self._stats._add_samples(int(data.size))
        if nw:
            t0 = time.perf_counter_ns()
            windows = np.lib.stride_tricks.as_strided(buf, shape=(nw, bs), strides=(hop * 8, 8), writeable=False)
            self._run(np.ascontiguousarray(windows))
            self._stats._record_blocks(nw, time.perf_counter_ns() - t0)

    def _flush_window(self):
        final = np.zeros((1, self.bufferSize))
        final[0, :self._tail.size] = self._tail
        self._run(final)
        self._tail = np.empty(0)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class _SingleLevelStreaming(_StreamingBase):
    """MODWTStreamingTransformImpl: windows of bufferSize samples sliding by bufferSize - (L-1)."""

    def __init__(self, wavelet, boundaryMode, bufferSize, engine=None):
        super().__init__(wavelet, boundaryMode, bufferSize)
        self.filterLength = len(wavelet.lowPassDecomposition())
        self.overlapSize = self.filterLength - 1
        if (self.bufferSize + self.overlapSize) * 8 > 100 * 1024 * 1024:                      # :104-113
            raise InvalidArgumentException("Buffer size too large, would require "
                                           f"{(self.bufferSize + self.overlapSize) * 8 // (1024 * 1024)}MB. "
                                           "Maximum allowed is 100MB")
        if self.bufferSize < self.filterLength:                                               # :116-120
            raise InvalidArgumentException("Buffer size must be at least as large as filter length: "
                                           f"bufferSize={self.bufferSize}, filterLength={self.filterLength}")
        self._hop = self.bufferSize - self.overlapSize
        self._t = MODWTTransform(wavelet, boundaryMode, engine)

    def _run(self, windows):
        eng = self._t._eng()
        w, v = eng.forward(windows, self._t._hs, self._t._gs, 1, self.boundaryMode.value, FLAG_CHECK_FINITE)
        for k in range(windows.shape[0]):
            self._submit(StreamingResult(v[k], w[0][k]))

    def flush(self):
        """:232-256: whatever waits in the ring (after a window that is its last L-1 samples), zero padded"""
        if self._closed:
            raise InvalidStateException.closed("Transform")
        if self._tail.size > 0:
            self._flush_window()

    def close(self):
        if not self._closed:
            self._closed = True
            if self._tail.size > 0:
                self._flush_window()
            self._complete()


class _MultiLevelStreaming(_StreamingBase):
    """MultiLevelMODWTStreamingTransform: back-to-back windows, `levels` results per window."""

    def __init__(self, wavelet, boundaryMode, bufferSize, levels, engine=None):
        super().__init__(wavelet, boundaryMode, bufferSize)
        if levels < 1:
            raise InvalidArgumentException(f"Levels must be at least 1, got: {levels}")
        self.levels = int(levels)
        self._hop = self.bufferSize
        self._t = MultiLevelMODWTTransform(wavelet, boundaryMode, engine)

    def _run(self, windows):
        max_levels = self._t._calculate_max_levels(self.bufferSize)                              # MultiLevelMODWTTransform :226-239
        if self.levels > max_levels:
            raise InvalidArgumentException(
                f"Invalid number of decomposition levels: {self.levels} (maximum {max_levels} for signal length "
                f"{self.bufferSize})", ErrorCode.CFG_INVALID_DECOMPOSITION_LEVEL)
        eng = self._t._eng()
        w, v = eng.forward(windows, self._t._hs, self._t._gs, self.levels, self.boundaryMode.value, FLAG_CHECK_FINITE)
        empty = np.empty(0)
        for k in range(windows.shape[0]):
            for level in range(1, self.levels + 1):
                self._submit(StreamingResult(v[k] if level == self.levels else empty, w[level - 1][k]))

    def flush(self):
        """:193-218 (no closed check in the reference)"""
        if self._tail.size > 0:
            self._flush_window()

    def close(self):
        if not self._closed:
            self._closed = True
            self.flush()
            self._complete()
