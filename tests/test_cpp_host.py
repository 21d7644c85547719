"""The C++ host mirror (include/vw_modwt.hpp): compiles against the C ABI everywhere; on a GPU box the parity program
(tests/cpp/host_parity.cpp) is run against the CPU oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_parity.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "host_parity")


def _build():
    sys.path.insert(0, ROOT)
    from oracle import cref
    cref.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib = os.path.join(ROOT, "vectorwave_b200")
    ora = os.path.join(ROOT, "oracle", "_build")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
                           "-L", lib, "-lvwmodwt", "-L", ora, "-lvw_oracle",
                           f"-Wl,-rpath,{lib}", f"-Wl,-rpath,{ora}", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return EXE


def test_cpp_host_mirror_compiles_and_fails_loudly_without_a_gpu():
    exe = _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present: covered by the parity run")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode != 0                     # vectorwave::NativeEngineError: no CPU path
    assert "vw_init failed" in (p.stderr + p.stdout)


@pytest.mark.gpu
def test_cpp_host_mirror_matches_the_oracle():
    p = subprocess.run([_build()], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.strip().endswith("ok"), p.stdout[-2000:] + p.stderr[-2000:]
