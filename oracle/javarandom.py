"""java.util.Random restatement (48-bit LCG) so the reference's seeded test fixtures can
be regenerated without a JVM.  TEST INFRASTRUCTURE ONLY.

nextDouble() is exact.  nextGaussian() follows the Marsaglia polar method of the JDK;
the JDK uses StrictMath.log/sqrt (fdlibm) where this uses libm, so a last-ulp
difference is possible -- parity never depends on it (oracle and engine consume the
same generated input).
"""
import math

import numpy as np

_MASK = (1 << 48) - 1
_MULT = 0x5DEECE66D


class JavaRandom:
    def __init__(self, seed):
        self.state = (seed ^ _MULT) & _MASK
        self._have_gauss = False
        self._next_gauss = 0.0

    def _next(self, bits):
        self.state = (self.state * _MULT + 0xB) & _MASK
        r = self.state >> (48 - bits)
        if r >= 1 << (bits - 1):  # signed (int) cast
            r -= 1 << bits
        return r

    def next_double(self):
        hi = self._next(26) & ((1 << 26) - 1)
        lo = self._next(27) & ((1 << 27) - 1)
        return ((hi << 27) + lo) * (1.0 / (1 << 53))

    def next_gaussian(self):
        if self._have_gauss:
            self._have_gauss = False
            return self._next_gauss
        while True:
            v1 = 2 * self.next_double() - 1
            v2 = 2 * self.next_double() - 1
            s = v1 * v1 + v2 * v2
            if 0 < s < 1:
                break
        m = math.sqrt(-2 * math.log(s) / s)
        self._next_gauss = v2 * m
        self._have_gauss = True
        return v1 * m


def uniform_pm1(n, seed):
    """`new Random(seed)`; x[i] = nextDouble()*2-1 (ETEST/modwt/BatchMODWTApiTest.java:67-74 randomAoS,
    CTEST/modwt/SymmetricNRMSEBaselineGuardTest.java randomSignal)."""
    r = JavaRandom(seed)
    return np.array([r.next_double() * 2 - 1 for _ in range(n)])


def composite_sin(n, seed, noise_std):
    """CTEST/testing/TestSignals.java:18-30 compositeSin."""
    r = JavaRandom(seed)
    out = np.empty(n)
    for i in range(n):
        t = i / float(n)
        v = (math.sin(2 * math.pi * 2 * t) + 0.5 * math.sin(2 * math.pi * 7 * t)
             + 0.25 * math.cos(2 * math.pi * 13 * t))
        if noise_std > 0:
            v += noise_std * r.next_gaussian()
        out[i] = v
    return out
