"""CPU-side checks of the boundary: the shared library loads, exports every symbol the header
declares, sizes theta like the reference's init.m, and refuses to compute without a GPU."""
import os
import re

import numpy as np
import pytest

from gpz_b200 import _lib as L
from gpz_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "gpz_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gpz_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(L.EXPORTS) == syms


@pytest.mark.parametrize("method", synth.METHODS)
def test_theta_len_matches_reference_layout(method):
    lib = L.load()
    for (d, k, m, het) in [(1, 1, 25, True), (5, 1, 100, True), (10, 2, 7, False)]:
        mm = L.make_model(d, k, m, method, het)
        import ctypes as C
        assert lib.gpz_theta_len(C.byref(mm)) == synth.theta_len(method, m, d, k, het)
        assert lib.gpz_g_dim(C.byref(mm)) == synth.g_dim_of(method, m, d)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    X, Y = synth.make_data(16, 2)
    with pytest.raises(L.GpzError, match="no CUDA device|CUDA"):
        L.Context(L.make_model(2, 1, 4, "VD", True), X, Y)
    with pytest.raises(L.GpzError):
        L.inv_logdet(np.eye(3))
    with pytest.raises(L.GpzError):
        L.dxy(np.zeros((3, 2)), np.zeros((2, 2)))


def test_mex_gateway_source_compiles_against_stub():
    """matlab/gpz_b200_mex.cpp cannot be built without MATLAB; at least keep it syntactically valid and in
    step with include/gpz_b200.h (g++ -fsyntax-only against a stub mex.h)."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "tests", "mex_stub"), os.path.join(ROOT, "matlab", "gpz_b200_mex.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
