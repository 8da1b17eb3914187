"""Host-side mirror of the reference API (gpz_b200/api.py): the host-only helpers on CPU, and the
init -> train -> predict flow of demo_sinc.m on the GPU."""
import numpy as np
import pytest

from gpz_b200 import api
from oracle import gpz_oracle as O


def test_fixpsi_matches_reference_layouts():
    rng = np.random.default_rng(0)
    n, d = 13, 3
    sd = np.array([2.0, 0.5, 1.5])
    for meth in ("VD", "GL", "VC", "GC"):
        for Psi in (rng.random((n, 1)), rng.random((n, d)), rng.random((d, d, n))):
            a, b = api.fixPsi(Psi, n, sd, meth), O.fixPsi(Psi, n, sd, meth)
            assert a.shape == b.shape and np.allclose(a, b, rtol=1e-15, atol=0)
    assert api.fixPsi(None, n, sd, "VD") is None


def test_pca_and_fill_linear_host_helpers():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((400, 3)) @ np.array([[1.0, 0.3, 0.0], [0.0, 1.0, 0.5], [0.0, 0.0, 0.7]])
    mu, sig, Ti = api._pca(X)
    assert np.allclose(mu, X.mean(axis=0)) and np.allclose(Ti.T @ Ti, np.cov(X.T), rtol=1e-10)
    Xn = X.copy()
    Xn[::7, 1] = np.nan
    F = api._fill_linear(Xn, mu, sig * 400)
    assert not np.isnan(F).any() and np.allclose(F[1], X[1])


@pytest.mark.gpu
def test_sinc_demo_flow():
    """demo_sinc.m in miniature: heteroscedastic 1-D regression, d=1 forces method ?L (init.m:12-14)."""
    rng = np.random.default_rng(0)
    n = 1200
    X = rng.uniform(-10, 10, (n, 1))
    noise = 0.05 + 0.2 * (1 + np.sin(X[:, 0] / 3)) / 2
    Y = (np.sinc(X[:, 0] / np.pi) + noise * rng.standard_normal(n)).reshape(n, 1)
    tr = np.arange(n) % 5 < 3
    va = np.arange(n) % 5 == 3
    te = np.arange(n) % 5 == 4
    model = api.init(X, Y, "VD", 25, training=tr, seed=1)
    assert model["method"] == "VL" and model["g_dim"] == 25
    f0, _, _ = api.GPz(model["last"]["theta"], model, (X - model["muX"]) / model["sdX"], Y - model["muY"], training=tr)
    model = api.train(model, X, Y, maxIter=60, training=tr, validation=va, display=False)
    f1, g1, st = api.GPz(model["best"]["theta"], model, (X - model["muX"]) / model["sdX"], Y - model["muY"], training=tr)
    assert f1 < f0 - 0.1
    mu, sigma, nu, beta_i, gamma, PHI, w, iS = api.predict(X, model, selection=te)
    rmse = float(np.sqrt(np.mean((mu[:, 0] - Y[te, 0]) ** 2)))
    assert rmse < 0.2 and np.all(sigma > 0) and PHI.shape == (int(te.sum()), 25)
    # the trained model's predictions agree with the oracle's predict() on the same model
    om = O.Model(d=1, k=1, m=25, method="VL", heteroscedastic=True, muX=model["muX"], sdX=model["sdX"], muY=model["muY"])
    om.best = dict(model["best"])
    mu2, sigma2, *_ = O.predict(X, om, selection=te)
    assert np.max(np.abs(mu - mu2)) <= 1e-9 * max(1.0, np.max(np.abs(mu2))) and np.max(np.abs(sigma - sigma2)) <= 1e-8 * np.max(sigma2)
    # getPrior through its reference-named wrapper = the priors train() stored (train.m:74)
    pr = api.getPrior((X - model["muX"]) / model["sdX"], None, model["best"]["theta"], model, tr)
    assert np.allclose(pr, model["best"]["priors"], rtol=1e-9, atol=1e-12) and abs(pr.sum() - 1.0) < 1e-12
    assert model["train_info"]["fun_evals"] >= model["train_info"]["iterations"] >= 1
    # noisy-input prediction runs and adds variance
    mu3, sigma3, _, _, gamma3, *_ = api.predict(X, model, selection=te, Psi=np.full((n, 1), 0.05))
    assert np.all(np.isfinite(sigma3)) and float(np.mean(gamma3)) > 0
