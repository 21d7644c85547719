"""ctypes front-end for the C oracle (oracle/modwt_oracle.c).  TEST INFRASTRUCTURE ONLY.

`build()` compiles oracle/_build/libvw_oracle.so with gcc via oracle/Makefile; `lib()`
loads it (building on demand).  Wrappers take / return numpy float64 arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "_build", "libvw_oracle.so")
_lib = None

PERIODIC, ZERO_PADDING, SYMMETRIC = 0, 1, 2
MODE = {"PERIODIC": 0, "ZERO_PADDING": 1, "SYMMETRIC": 2}

_dp = C.POINTER(C.c_double)
_i64 = C.c_int64


def build(force=False):
    src = os.path.join(_DIR, "modwt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.vwo_conv.argtypes = [_dp, _i64, _dp, _i64, C.c_int, _dp]
        L.vwo_upsampled_len.argtypes = [_i64, C.c_int]
        L.vwo_upsampled_len.restype = _i64
        L.vwo_upsample_scale.argtypes = [_dp, _i64, C.c_int, _dp]
        L.vwo_max_levels.argtypes = [_i64, _i64, C.c_int]
        L.vwo_max_levels.restype = C.c_int
        L.vwo_forward_single.argtypes = [_dp, _i64, _dp, _dp, _i64, C.c_int, _dp, _dp]
        L.vwo_inverse_single.argtypes = [_dp, _dp, _i64, _dp, _dp, _i64, C.c_int, C.c_int, _dp]
        L.vwo_decompose.argtypes = [_dp, _i64, _dp, _dp, _i64, C.c_int, C.c_int, C.c_int, _dp, _dp]
        L.vwo_decompose.restype = C.c_int
        L.vwo_alignment.argtypes = [C.c_int, _i64, C.c_int, C.POINTER(C.c_int)]
        L.vwo_reconstruct.argtypes = [_dp, _dp, _i64, _dp, _dp, _i64, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_uint64, C.c_int, _dp]
        L.vwo_threshold.argtypes = [_dp, _i64, C.c_double, C.c_int]
        L.vwo_universal_threshold.argtypes = [_dp, _i64]
        L.vwo_universal_threshold.restype = C.c_double
        L.vwo_swt_denoise.argtypes = [_dp, _i64, _dp, _dp, _i64, C.c_int, C.c_int, C.c_int, C.c_double,
                                      C.c_int, C.c_int, _dp]
        L.vwo_swt_denoise.restype = C.c_double
        L.vwo_sure_threshold.argtypes = [_dp, _i64, C.c_double, _dp]
        L.vwo_sure_threshold.restype = C.c_double
        L.vwo_batch_fwd_inv.argtypes = [_dp, _i64, _i64, _dp, _dp, _i64, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, _dp, _dp, _dp]
        L.vwo_batch_soa_decompose.argtypes = [_dp, _i64, _i64, _dp, _dp, _i64, C.c_int, _dp, _dp]
        L.vwo_batch_soa_haar_single.argtypes = [_dp, _i64, _i64, _dp, _dp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def conv(x, f, mode):
    x, f = _c(x), _c(f)
    out = np.empty_like(x)
    lib().vwo_conv(_p(x), x.size, _p(f), f.size, mode, _p(out))
    return out


def upsample_scale(base, level):
    base = _c(base)
    out = np.empty(lib().vwo_upsampled_len(base.size, level))
    lib().vwo_upsample_scale(_p(base), base.size, level, _p(out))
    return out


def max_levels(n, l, cap=10):
    """cap=10 is the reference's MAX_DECOMPOSITION_LEVELS (=> at most 9); cap=0 is uncapped."""
    return lib().vwo_max_levels(n, l, cap)


def forward_single(x, h, g, mode):
    x, h, g = _c(x), _c(h), _c(g)
    v, w = np.empty_like(x), np.empty_like(x)
    lib().vwo_forward_single(_p(x), x.size, _p(h), _p(g), h.size, mode, _p(v), _p(w))
    return v, w


def inverse_single(v, w, hr, gr, mode, batch_variant=False):
    v, w, hr, gr = _c(v), _c(w), _c(hr), _c(gr)
    out = np.empty_like(v)
    lib().vwo_inverse_single(_p(v), _p(w), v.size, _p(hr), _p(gr), hr.size, mode, int(batch_variant), _p(out))
    return out


def decompose(x, h, g, levels, mode, dense=False):
    """Returns (W[J][N], V_J[N]); raises ValueError when L_j > N (VAL_TOO_LARGE)."""
    x, h, g = _c(x), _c(h), _c(g)
    w = np.empty((levels, x.size))
    v = np.empty(x.size)
    rc = lib().vwo_decompose(_p(x), x.size, _p(h), _p(g), h.size, levels, mode, int(dense), _p(w), _p(v))
    if rc != 0:
        raise ValueError("VAL_TOO_LARGE: upsampled filter longer than signal")
    return w, v


def alignment(wavelet_id, l0, level):
    out = (C.c_int * 4)()
    lib().vwo_alignment(wavelet_id, l0, level, out)
    return tuple(out)


def reconstruct(w, v, hr, gr, mode, wavelet_id=0, dense=False, detail_mask=None, use_approx=True):
    w, v, hr, gr = _c(w), _c(v), _c(hr), _c(gr)
    levels = w.shape[0]
    if detail_mask is None:
        detail_mask = (1 << levels) - 1
    out = np.empty_like(v)
    lib().vwo_reconstruct(_p(w), _p(v), v.size, _p(hr), _p(gr), hr.size, levels, mode, wavelet_id, int(dense),
                          detail_mask, int(use_approx), _p(out))
    return out


def threshold(c, thr, soft):
    c = _c(c).copy()
    lib().vwo_threshold(_p(c), c.size, thr, int(soft))
    return c


def universal_threshold(w1):
    w1 = _c(w1)
    return lib().vwo_universal_threshold(_p(w1), w1.size)


def sure_threshold(c, sigma):
    """WaveletDenoiser.calculateSUREThreshold -> (threshold, minimal risk)"""
    c = _c(c)
    risk = np.empty(1)
    thr = lib().vwo_sure_threshold(_p(c), c.size, float(sigma), _p(risk))
    return thr, float(risk[0])


def swt_denoise(x, h, g, levels, mode, wavelet_id=0, thr=-1.0, soft=True, dense=False):
    x, h, g = _c(x), _c(h), _c(g)
    out = np.empty_like(x)
    used = lib().vwo_swt_denoise(_p(x), x.size, _p(h), _p(g), h.size, levels, mode, wavelet_id, thr, int(soft),
                                 int(dense), _p(out))
    return out, used


def batch_fwd_inv(x, h, g, levels, mode, wavelet_id=0, dense=True, threads=1, inverse=True):
    """x [B][N] -> (W [B][J][N], V [B][N], xr [B][N] or None).  dense=True is the reference's cost."""
    x, h, g = _c(x), _c(h), _c(g)
    b, n = x.shape
    w = np.empty((b, levels, n))
    v = np.empty((b, n))
    xr = np.empty((b, n)) if inverse else None
    lib().vwo_batch_fwd_inv(_p(x), b, n, _p(h), _p(g), h.size, levels, mode, wavelet_id, int(dense), threads,
                            _p(w), _p(v), _p(xr) if inverse else None)
    return w, v, xr


def batch_soa_decompose(soa_x, b, n, h, g, levels):
    soa_x, h, g = _c(soa_x), _c(h), _c(g)
    w = np.empty((levels, b * n))
    v = np.empty(b * n)
    lib().vwo_batch_soa_decompose(_p(soa_x), b, n, _p(h), _p(g), h.size, levels, _p(w), _p(v))
    return w, v


def batch_soa_haar_single(soa_x, b, n):
    soa_x = _c(soa_x)
    v, w = np.empty(b * n), np.empty(b * n)
    lib().vwo_batch_soa_haar_single(_p(soa_x), b, n, _p(v), _p(w))
    return v, w
