"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol include/vw_modwt.h
declares, the ctypes table covers the same set, and the product never reaches for the oracle."""
import os
import re

import pytest

import vectorwave_b200 as vw
from vectorwave_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vw_modwt.h")).read()
    return sorted(set(re.findall(r"VW_API[^;(]*?\b(vw_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("vw_init", "vw_modwt_forward", "vw_modwt_inverse", "vw_swt_denoise", "vw_conv_modwt",
                 "vw_modwt_forward_span", "vw_modwt_inverse_span", "vw_modwt_stream_level", "vw_median_abs", "vw_mean_variance", "vw_universal_threshold", "vw_threshold"):
        assert must in names
    assert len(names) >= 25


def test_library_exports_every_declared_symbol():
    lib = vw.load_library()
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in vw_modwt.h but not exported by libvwmodwt.so"


def test_ctypes_table_matches_header():
    assert sorted(_native.SIGNATURES) == _declared()


def test_pure_host_entry_points_work_without_a_gpu():
    lib = vw.load_library()
    assert lib.vw_abi_version() == 2
    assert lib.vw_max_levels(10000, 2, 10) == 9          # CTEST/modwt/MultiLevelMODWTTransformTest.java:275
    assert lib.vw_max_levels(10000, 2, 0) == 14
    assert lib.vw_max_levels(8, 8, 10) == 0
    assert lib.vw_span_halo(30, 1, 3) == 29 * 7
    assert lib.vw_span_halo(16, 5, 2) == 15 * 16 * 3
    assert lib.vw_status_name(104) == b"CFG_INVALID_DECOMPOSITION_LEVEL"


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(vw.NativeEngineError):
        vw.Engine.get()
    with pytest.raises(vw.NativeEngineError):
        vw.MODWTTransform(vw.Haar(), vw.BoundaryMode.PERIODIC).forward([1.0, 2.0, 3.0, 4.0])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vectorwave_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_product_wavelet_tables_equal_oracle_tables():
    from oracle.wavelets import TABLES, highpass
    import numpy as np
    for name, table in TABLES.items():
        w = vw.get_wavelet(name)
        np.testing.assert_array_equal(w.lowPassDecomposition(), np.array(table))
        np.testing.assert_array_equal(w.highPassDecomposition(), highpass(name))


def test_host_policy_level_cap_and_alignment_match_oracle():
    from oracle import cref
    from oracle.wavelets import WAVELET_ID
    from vectorwave_b200.modwt import compute_tau_j
    for name in ("haar", "db4", "db6", "db8", "sym4", "sym8", "coif2", "coif3", "coif5", "db10"):
        w = vw.get_wavelet(name)
        l = w.lowPassDecomposition().size
        t = vw.MultiLevelMODWTTransform.__new__(vw.MultiLevelMODWTTransform)
        t._hs = w.lowPassDecomposition(); t._enforce_cap = True
        for n in (7, 8, 9, 64, 100, 129, 4096, 10000, 1 << 20):
            assert t._calculate_max_levels(n) == cref.max_levels(n, l)
        for level in range(1, 9):
            ap, dh, dp, dg = vw.SymmetricAlignmentStrategy.decide(w, level)
            c = cref.alignment(WAVELET_ID.get(name, 0), l, level)
            assert (ap, dh, dp, dg) == (bool(c[0]), c[1], bool(c[2]), c[3])
            assert compute_tau_j(l, level) == ((l - 1) * (1 << (level - 1))) // 2
