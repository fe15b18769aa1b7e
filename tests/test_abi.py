"""CPU tests: the C-ABI library loads, exports every symbol include/cvmx.h declares, and refuses to run
without a CUDA device (no CPU fallback).  No compute call is made here."""

import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cvmx.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cvmx_[a-z_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from cvmatrix_b200 import _lib, build

    build.build_lib()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 14
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (cvmx_[a-z_]+)", nm))
    assert set(names) <= exported, set(names) - exported
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.cvmx_version() == 100


def test_library_is_sm100a_only_and_self_contained():
    from cvmatrix_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcudart" not in ldd and "libtorch" not in ldd  # static CUDA runtime, no torch in the ABI


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from cvmatrix_b200 import CVMatrix, _lib

    h = C.c_void_p()
    rc = _lib.load().cvmx_create(0, _lib.F64, 15, 1, 1e-14, C.byref(h))
    assert rc == _lib.ERR_CUDA and not h.value
    assert b"no CPU fallback" in _lib.load().cvmx_last_error(None)
    with pytest.raises(_lib.CvmxError):
        CVMatrix()


def test_constructor_argument_checks():
    from cvmatrix_b200 import CVMatrix

    with pytest.raises(ImportError, match=r"cvmatrix\[jax\]"):   # the reference's exception without JAX (cvmatrix.py:85-90)
        CVMatrix(backend="jax")
    with pytest.raises(ValueError, match="Invalid backend"):
        CVMatrix(backend="torch")
    with pytest.raises(TypeError):
        CVMatrix(dtype=np.float16)


def test_product_does_not_import_oracle():
    """The shipped package must never route through oracle/ (or any CPU implementation)."""
    pkg = os.path.join(ROOT, "cvmatrix_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "build.py", (f,)
