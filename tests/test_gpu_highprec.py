"""The CUDA path against the 40-digit objective of tests/test_oracle_highprec.py directly (no oracle in between): nlogML to 1e-12
and the gradient against central differences of the 40-digit objective, on both GEMM engines."""
import numpy as np
import pytest

from gpz_b200 import _lib as L
from gpz_b200 import synth

mp = pytest.importorskip("mpmath")
from test_oracle_highprec import _mp_nlogml  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("meth,psi,nan", [("VC", False, False), ("VD", True, False), ("GC", True, True), ("VL", False, True)])
@pytest.mark.parametrize("engine", ["fp64", "int8"])
def test_cuda_objective_and_gradient_against_40_digit_differences(meth, psi, nan, engine):
    n, d, m = 36, 2, 5
    X, Y = synth.make_data(n, d, seed=11)
    X, Y = np.array(X), np.asarray(Y)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, meth, m, het=True, seed=12), 0.2, 13)
    omega = 0.5 + np.random.default_rng(14).random((n, 1))
    if nan:
        X[3:12, 0] = np.nan
        X[20:27, 1] = np.nan
    Psi = synth.make_psi(n, d, meth, seed=15) if psi else None
    ctx = L.Context(L.make_model(d, 1, m, meth, True), X, Y, Psi, omega, np.ones(n, dtype=bool), None)
    ctx.set_option("ozaki_slices", 7 if engine == "int8" else 0)
    f, g, _ = ctx.eval(theta)
    ctx.close()
    mp.mp.dps = 40
    th = [mp.mpf(float(t)) for t in theta]
    f0 = _mp_nlogml(th, meth, X, Y[:, 0], omega[:, 0], m, d, Psi)
    assert abs(float((mp.mpf(float(f)) - f0) / f0)) <= 1e-12
    h = mp.mpf(10) ** -15
    gx = np.zeros(len(theta))
    for q in range(len(theta)):
        tp, tm = list(th), list(th)
        tp[q] += h
        tm[q] -= h
        gx[q] = float((_mp_nlogml(tp, meth, X, Y[:, 0], omega[:, 0], m, d, Psi) - _mp_nlogml(tm, meth, X, Y[:, 0], omega[:, 0], m, d, Psi)) / (2 * h))
    err = np.max(np.abs(g - gx)) / np.max(np.abs(gx))
    assert err <= 1e-10, err
