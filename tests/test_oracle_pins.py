"""Pins for the CPU oracle.  The reference ships no tests or golden
vectors for this path ("parity unpinned"), so the restatement is pinned by identities that any
correct implementation of GPz.m / getPHI.m / predict*.m must satisfy."""
import itertools

import numpy as np
import pytest

from gpz_b200 import synth
from oracle import gpz_oracle as O


def problem(method, het, psi, nan, n=60, d=3, m=5, k=1, seed=0):
    X, Y = synth.make_data(n, d, seed=seed, k=k)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, method, m, het=het, seed=seed + 1), 0.1, seed + 2)
    model = O.Model(d=d, k=k, m=m, method=method, heteroscedastic=het)
    Psi = synth.make_psi(n, d, method, seed=seed + 3) if psi else None
    X = np.array(X)
    if nan:
        rng = np.random.default_rng(seed + 4)
        X[rng.random(n) < 0.25, 0] = np.nan
        X[rng.random(n) < 0.15, d - 1] = np.nan
    rng = np.random.default_rng(seed + 5)
    omega = 0.5 + rng.random((n, 1))
    return model, theta, X, np.array(Y), Psi, omega


COMBOS = list(itertools.product(synth.METHODS, (False, True), (False, True), (False, True)))


@pytest.mark.parametrize("method,het,psi,nan", COMBOS)
def test_T1_gradient_matches_finite_differences(method, het, psi, nan):
    model, theta, X, Y, Psi, omega = problem(method, het, psi, nan)
    r = O.GPz(theta, model, X, Y, Psi, omega)
    rng = np.random.default_rng(11)
    h = 1e-6
    for _ in range(4):
        u = rng.standard_normal(theta.size)
        u /= np.linalg.norm(u)
        fp = O.GPz(theta + h * u, model, X, Y, Psi, omega).nlogML
        fm = O.GPz(theta - h * u, model, X, Y, Psi, omega).nlogML
        fd = (fp - fm) / (2 * h)
        an = float(r.grad @ u)
        assert abs(fd - an) <= 2e-6 * max(1.0, abs(an), np.linalg.norm(r.grad)), (fd, an)


def test_T1_blockwise_fd_k2():
    model, theta, X, Y, Psi, omega = problem("VD", True, False, False, k=2)
    r = O.GPz(theta, model, X, Y, Psi, omega)
    h = 1e-6
    for idx in np.random.default_rng(5).choice(theta.size, 12, replace=False):
        e = np.zeros(theta.size)
        e[idx] = h
        fd = (O.GPz(theta + e, model, X, Y, Psi, omega).nlogML - O.GPz(theta - e, model, X, Y, Psi, omega).nlogML) / (2 * h)
        assert abs(fd - r.grad[idx]) <= 1e-6 * max(1.0, abs(r.grad[idx]))


@pytest.mark.parametrize("method", synth.METHODS)
def test_T2_dense_gp_evidence(method):
    """-n*k*nlogML equals ln N(y; 0, diag(1/beta) + PHI diag(1/alpha) PHI') plus the
    heteroscedastic prior terms, computed by an independent n x n Cholesky (omega = 1)."""
    n, d, m = 80, 3, 6
    model, theta, X, Y, _, _ = problem(method, True, False, False, n=n, d=d, m=m)
    r = O.GPz(theta, model, X, Y, None, None)
    PHI, _, lnBeta_i, _ = O.getPHI(X, None, theta, model, None)
    o1 = m * d + model.g_dim
    alpha = np.exp(theta[o1:o1 + m])
    v = theta[o1 + m + 1:o1 + 2 * m + 1]
    lnTau = theta[o1 + 2 * m + 1:o1 + 3 * m + 1]
    K = np.diag(np.exp(lnBeta_i[:, 0])) + (PHI / alpha[None, :]) @ PHI.T
    L = np.linalg.cholesky(K)
    z = np.linalg.solve(L, Y[:, 0])
    ll = -0.5 * z @ z - np.sum(np.log(np.diag(L))) - 0.5 * n * O.LN2PI
    ll += -0.5 * np.sum(v ** 2 * np.exp(lnTau)) + 0.5 * np.sum(lnTau) - 0.5 * m * O.LN2PI
    assert abs(-n * r.nlogML - ll) <= 1e-9 * abs(ll)


def test_T3_mode_equivalence_at_isotropic_gamma():
    n, d, m = 70, 3, 5
    X, Y = synth.make_data(n, d, seed=3)
    X, Y = np.array(X), np.array(Y)
    res = {}
    for method in synth.METHODS:
        th = synth.make_theta0(X, Y, method, m, het=True, seed=9)
        if method in ("GL", "GD", "GC"):      # globals use mean(gamma): make every mode share ONE gamma
            pass
        res[method] = (th, O.Model(d=d, k=1, m=m, method=method, heteroscedastic=True))
    # force a single common gamma so that all six parameterisations describe the same model
    g0 = 0.37
    out = {}
    for method, (th, model) in res.items():
        md = m * d
        G = {"GL": np.array([g0]), "VL": np.full(m, g0), "GD": np.full(d, g0), "VD": np.full(m * d, g0),
             "GC": (np.eye(d) * g0).reshape(-1, order="F"),
             "VC": np.repeat((np.eye(d) * g0)[:, :, None], m, axis=2).reshape(-1, order="F")}[method]
        th = th.copy()
        th[md:md + model.g_dim] = G
        th[md + model.g_dim + m + 1: md + model.g_dim + 2 * m + 1] = 0.05 * np.cos(np.arange(m))   # v != 0
        out[method] = (O.GPz(th, model, X, Y), model)
    f0 = out["VD"][0].nlogML
    for method in synth.METHODS:
        assert abs(out[method][0].nlogML - f0) <= 1e-13 * abs(f0), method
    md = m * d
    gVD = out["VD"][0].grad[md:md + m * d].reshape((m, d), order="F")
    gVC = out["VC"][0].grad[md:md + d * d * m].reshape((d, d, m), order="F")
    assert np.allclose(out["GL"][0].grad[md], gVD.sum(), rtol=1e-12, atol=1e-15)
    assert np.allclose(out["VL"][0].grad[md:md + m], gVD.sum(axis=1), rtol=1e-12, atol=1e-15)
    assert np.allclose(out["GD"][0].grad[md:md + d], gVD.sum(axis=0), rtol=1e-12, atol=1e-15)
    assert np.allclose(np.stack([np.diag(gVC[:, :, j]) for j in range(m)]), gVD, rtol=1e-12, atol=1e-15)
    gGC = out["GC"][0].grad[md:md + d * d].reshape((d, d), order="F")
    assert np.allclose(gGC, gVC.sum(axis=2), rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("method", synth.METHODS)
def test_T4_zero_psi_is_no_psi(method):
    model, theta, X, Y, _, omega = problem(method, True, False, False)
    n, d = X.shape
    Z = np.zeros((d, d, n)) if method[1] == "C" else np.zeros((n, d))
    a = O.GPz(theta, model, X, Y, None, omega)
    b = O.GPz(theta, model, X, Y, Z, omega)
    assert abs(a.nlogML - b.nlogML) <= 1e-13 * abs(a.nlogML)
    assert np.allclose(a.grad, b.grad, rtol=1e-10, atol=1e-13)


def _fitted_model(method, n=50, d=3, m=5):
    model, theta, X, Y, _, _ = problem(method, True, False, False, n=n, d=d, m=m)
    r = O.GPz(theta, model, X, Y, fit_only=True)
    model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(1)
    o2 = m * d + model.g_dim + m + 1
    model.best = dict(theta=theta, w=r.w, iSigma_w=r.iSigma_w, P=theta[:m * d].reshape((m, d), order="F"),
                      v=theta[o2:o2 + m].reshape(m, 1))
    return model, X, Y


@pytest.mark.parametrize("method", synth.METHODS)
def test_T5_noisy_predict_tends_to_full(method):
    model, X, _ = _fitted_model(method)
    n, d = X.shape
    Xt = X[:7]
    eps = 1e-12
    Psi = np.repeat((np.eye(d) * eps)[:, :, None], 7, axis=2) if method[1] == "C" else np.full((7, d), eps)
    mu0, s0, nu0, b0, g0, _ = O.predict(Xt, model, Psi=None)
    mu1, s1, nu1, b1, g1, _ = O.predict(Xt, model, Psi=Psi)
    assert np.allclose(mu0, mu1, rtol=1e-9, atol=1e-11)
    assert np.allclose(nu0, nu1, rtol=1e-8, atol=1e-10)
    assert np.allclose(b0, b1, rtol=1e-8, atol=1e-10)
    assert np.all(np.abs(g1) <= 1e-8)


def test_T6_fit_path_definitions():
    model, theta, X, Y, _, omega = problem("VD", True, False, False, n=90, m=7)
    r = O.GPz(theta, model, X, Y, None, omega, fit_only=True)
    PHI, _, lnB, _ = O.getPHI(X, None, theta, model, None)
    W = (np.exp(-lnB) * omega)[:, 0]
    m, d = model.m, model.d
    alpha = np.exp(theta[m * d + model.g_dim: m * d + model.g_dim + m])
    S = PHI.T @ (PHI * W[:, None]) + np.diag(alpha)
    assert np.allclose(S @ r.w[:, 0], PHI.T @ (W * Y[:, 0]), rtol=1e-9)
    assert np.allclose(r.iSigma_w[:, :, 0] @ S, np.eye(m), atol=1e-8)
    assert r.nlogML.shape == (1, 1) and r.grad == 0.0


def test_inv_logdet_and_dxy_and_fixpsi():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((9, 9))
    S = A @ A.T + np.eye(9)
    Xi, ld = O.inv_logdet(S)
    assert np.allclose(Xi @ S, np.eye(9), atol=1e-10)
    assert abs(ld - np.linalg.slogdet(S)[1]) < 1e-10
    # rank-deficient: pseudo-inverse, logdet over retained singular values only
    B = A[:, :4] @ A[:, :4].T
    Xi, ld = O.inv_logdet(B)
    assert np.allclose(Xi, np.linalg.pinv(B), atol=1e-8)
    X, P = rng.standard_normal((20, 3)), rng.standard_normal((6, 3))
    D = O.Dxy(X, P)
    assert np.allclose(D, ((X[:, None, :] - P[None]) ** 2).sum(-1), atol=1e-12)
    sd = np.array([2.0, 0.5, 1.5])
    psi_col = rng.random((20, 1))
    assert np.allclose(O.fixPsi(psi_col, 20, sd, "VD"), psi_col / sd[None] ** 2)
    C = O.fixPsi(rng.random((20, 3)), 20, sd, "VC")
    assert C.shape == (3, 3, 20) and np.allclose(C[0, 1, :], 0)
    cube = rng.random((3, 3, 20))
    assert np.allclose(O.fixPsi(cube, 20, sd, "GC")[:, :, 4], cube[:, :, 4] / np.outer(sd, sd))
    assert np.allclose(O.fixPsi(cube, 20, sd, "GD")[4], np.diag(cube[:, :, 4]) / sd ** 2)


def test_validation_stats_and_nan_groups():
    model, theta, X, Y, Psi, omega = problem("VD", True, True, True, n=80)
    tr = np.arange(80) % 4 != 0
    r = O.GPz(theta, model, X, Y, Psi, omega, training=tr, validation=~tr)
    for key in ("trainRMSE", "trainLL", "validRMSE", "validLL"):
        assert np.isfinite(r.stats[key])
    g = O.nan_groups(np.isnan(X))
    assert sum(int(x.sum()) for x in g) == 80 and len(g) >= 2


@pytest.mark.parametrize("with_psi", [False, True])
def test_T7_missing_input_prediction_cov_equals_diag_for_diagonal_gamma(with_psi):
    """predictMissing / predictNoisyMissing of the covariance family (predictCov.m:134-336) and of the diagonal family
    (predictDiag.m:127-295) are two independent restatements; with diagonal Gamma_j (and diagonal Psi) they describe the same
    model, so every output must agree.  Also: Psi -> 0 reduces predictNoisyMissing to predictMissing."""
    rng = np.random.default_rng(3)
    n, d, m, k = 6, 3, 4, 2
    X = rng.standard_normal((n, d))
    X[:, 1] = np.nan                                         # one group: dim 1 missing
    P = rng.standard_normal((m, d))
    gd = np.abs(rng.standard_normal((m, d))) + 0.5
    G = np.zeros((d, d, m))
    for j in range(m):
        G[:, :, j] = np.diag(gd[j])
    w, v, b = rng.standard_normal((m, k)), 0.3 * rng.standard_normal((m, k)), rng.standard_normal(k)
    iS = np.zeros((m, m, k))
    for o in range(k):
        A = rng.standard_normal((m, m))
        iS[:, :, o] = A @ A.T
    pri = np.abs(rng.standard_normal(m)) + 0.1
    pri /= pri.sum()
    Psi_d = Psi_c = None
    if with_psi:
        Psi_d = np.abs(rng.standard_normal((n, d))) * 0.2
        Psi_c = np.zeros((d, d, n))
        for t in range(n):
            Psi_c[:, :, t] = np.diag(Psi_d[t])
    rc = O._predictMissingCov(X, Psi_c, G, w, v, b, P, iS, pri)
    rd = O._predictMissingDiag(X, Psi_d, gd, w, v, b, P, iS, pri)
    for a, bb in zip(rc, rd):
        assert np.max(np.abs(a - bb)) <= 1e-12 * max(1.0, np.max(np.abs(bb)))
    if not with_psi:
        tiny = np.zeros((d, d, n))
        for t in range(n):
            tiny[:, :, t] = 1e-13 * np.eye(d)
        r0 = O._predictMissingCov(X, tiny, G, w, v, b, P, iS, pri)
        for a, bb in zip(r0, rc):
            assert np.max(np.abs(a - bb)) <= 1e-9 * max(1.0, np.max(np.abs(bb)))
