"""The oracle's predictNoisy for the covariance modes (oracle/gpz_oracle.py::_predictNoisyCov, following
GPz/predictCov.m:70-133) against an independent 40-digit evaluation of the same sums with mpmath, on a row of the
reference's own data file (tests/golden/sdss_cfg2.npz) at config 2's m = 100, d = 5.

Why: nu = sum_ij 2 Z_ij iSigma_w(i,j) cancels by ~1e4 here, so this is the quantity where an fp64 implementation shows
its conditioning.  The oracle must be accurate to ~eps x that amplification for the CUDA comparison in
tests/test_baseline_shapes.py to mean anything (it is: 1e-12 relative; the CUDA path read the other triangle of the
not-exactly-symmetric iSigma_w until this was checked, and differed at 2e-8)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from oracle import gpz_oracle as O  # noqa: E402

mp = pytest.importorskip("mpmath")


def test_predict_noisy_cov_against_40_digit_arithmetic():
    import make_sdss_fixture as F
    z = np.load(os.path.join(HERE, "golden", "sdss_cfg2.npz"))
    tr, va = z["training"], z["validation"]
    Xz, Yc, PsiC, theta = F.prepare(z["rows"], tr)
    m, d = 100, 5
    model = O.Model(d=d, k=1, m=m, method="VC", heteroscedastic=True)
    model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(1)
    o2 = m * d + model.g_dim + m + 1
    P = theta[:m * d].reshape((m, d), order="F")
    model.best = dict(theta=theta, w=z["w"], iSigma_w=z["iSigma_w"], P=P, v=theta[o2:o2 + m].reshape(m, 1))
    pick = np.flatnonzero(va)[:1]
    Xt = np.ascontiguousarray(Xz[pick])
    Pt = np.ascontiguousarray(PsiC[:, :, pick])
    mu, sigma, nu, be, ga, PHI = O.predict(Xt, model, Psi=Pt)

    mp.mp.dps = 40
    M = lambda a: mp.matrix(np.asarray(a, dtype=np.float64).tolist())            # noqa: E731
    Gamma = O.unpack_gamma(theta, model)
    iSw, w = z["iSigma_w"][:, :, 0], z["w"][:, 0]
    iS = [M(Gamma[:, :, i]).T * M(Gamma[:, :, i]) for i in range(m)]                # predictCov.m:84-87
    S = [mp.inverse(a) for a in iS]
    lnz = [-mp.log(mp.det(a)) / 2 for a in iS]
    Pm = [M(P[i:i + 1, :]) for i in range(m)]
    x, psi = M(Xt[0:1, :]), M(Pt[:, :, 0])
    nu_x, ga_x, mu_x = mp.mpf(0), mp.mpf(0), mp.mpf(0)
    for i in range(m):
        for j in range(i + 1):
            iC = iS[i] + iS[j]                                                     # :99-101
            C = mp.inverse(iC)
            c = (Pm[i] * iS[i] + Pm[j] * iS[j]) * C
            Dl = Pm[i] - Pm[j]
            Sij = S[i] + S[j]
            lnZ = lnz[i] + lnz[j] - (Dl * mp.inverse(Sij) * Dl.T)[0] / 2 - mp.log(mp.det(Sij)) / 2   # :105
            Dt = x - c
            CpP = psi + C
            lnN = -(Dt * mp.inverse(CpP) * Dt.T)[0] / 2 - mp.log(mp.det(CpP)) / 2               # :111
            Z = mp.exp(lnZ + lnN)
            f = 2 if j < i else 1                                                  # :115-125
            nu_x += f * Z * mp.mpf(float(iSw[i, j]))
            ga_x += f * Z * mp.mpf(float(w[i])) * mp.mpf(float(w[j]))
    # E[phi_i] (getPHI.m:80-88 with Psi) for mu = PHI w and gamma = sum - mu^2
    for i in range(m):
        Dt = x - Pm[i]
        SpP = psi + S[i]
        lnphi = -(Dt * mp.inverse(SpP) * Dt.T)[0] / 2 + mp.log(mp.det(S[i])) / 2 - mp.log(mp.det(SpP)) / 2
        mu_x += mp.exp(lnphi) * mp.mpf(float(w[i]))
    ga_x -= mu_x ** 2
    rel = lambda a, b: abs(float((mp.mpf(float(a)) - b) / b))                      # noqa: E731
    amp = float(sum(abs(float(iSw[i, j])) for i in range(m) for j in range(i + 1)))   # crude bound on sum |terms| / Z
    assert rel(mu[0, 0], mu_x) <= 1e-12
    assert rel(nu[0, 0], nu_x) <= 1e-10, (float(nu_x), nu[0, 0], amp)
    assert abs(float(mp.mpf(float(ga[0, 0])) - ga_x)) <= 1e-12 * max(1.0, float(mu_x) ** 2)


def _mp_nlogml(theta, meth, X, Y, omega, m, d, Psi=None):
    """GPz.m:20-110,233 restated in 40-digit arithmetic for k = 1, heteroscedastic, no Psi, no NaN; theta holds mp numbers.
    Only the OBJECTIVE: the gradient below comes from differencing it, which makes the check independent of the oracle's
    (and the CUDA path's) hand-derived gradient code."""
    n = X.shape[0]
    md = m * d
    P = [[theta[a * m + j] for a in range(d)] for j in range(m)]                       # P = reshape(theta(1:m*d), m, d)
    cov = meth[1] == "C"
    if meth == "VC":                                                                   # getPHI.m:26-40 / init.m:65-86
        g_dim = d * d * m
        Gam = [[[theta[md + b + a * d + j * d * d] for a in range(d)] for b in range(d)] for j in range(m)]   # Gamma(b, a, j)
    elif meth == "GC":
        g_dim = d * d
        Gam = [[[theta[md + b + a * d] for a in range(d)] for b in range(d)] for j in range(m)]
    elif meth == "VD":                                                                 # Gamma = reshape(., m, d)
        g_dim = md
        Gam = [[theta[md + j + a * m] for a in range(d)] for j in range(m)]
    elif meth == "GD":
        g_dim = d
        Gam = [[theta[md + a] for a in range(d)] for j in range(m)]
    elif meth == "VL":
        g_dim = m
        Gam = [[theta[md + j] for a in range(d)] for j in range(m)]
    else:                                                                              # GL
        g_dim = 1
        Gam = [[theta[md] for a in range(d)] for j in range(m)]
    o = md + g_dim
    lnA = theta[o:o + m]
    b = theta[o + m]
    v = theta[o + m + 1:o + m + 1 + m]
    lnT = theta[o + m + 1 + m:o + m + 1 + 2 * m]
    PHI = [[None] * m for _ in range(n)]
    Sig = None
    if cov:                                                                            # Sigma_j = (Gamma_j' Gamma_j)^-1, getPHI.m:73
        Sig = []
        for j in range(m):
            G = mp.matrix(Gam[j])
            Sig.append(mp.inverse(G.T * G))
    for i in range(n):
        obs = [a for a in range(d) if not np.isnan(X[i, a])]                           # getPHI.m:43-54: rows grouped by NaN pattern
        nmiss = d - len(obs)
        for j in range(m):
            dl = [mp.mpf(float(X[i, a])) - P[j][a] for a in obs]
            if cov:                                                                    # getPHI.m:73-88 on the observed block
                Soo = mp.matrix([[Sig[j][a, c] for c in obs] for a in obs])
                dv = mp.matrix([dl])
                if Psi is None:
                    e = -(dv * mp.inverse(Soo) * dv.T)[0] / 2
                else:
                    Sm = Soo + mp.matrix([[mp.mpf(float(Psi[a, c, i])) for c in obs] for a in obs])
                    e = -(dv * mp.inverse(Sm) * dv.T)[0] / 2 + mp.log(mp.det(Soo)) / 2 - mp.log(mp.det(Sm)) / 2
            else:                                                                      # getPHI.m:93-105, Sigma_ja = Gamma_ja^-2
                e = mp.mpf(0)
                for q, a in enumerate(obs):
                    if Psi is None:
                        e -= (dl[q] * Gam[j][a]) ** 2 / 2
                    else:
                        sg = 1 / Gam[j][a] ** 2
                        ps = mp.mpf(float(Psi[i, a]))
                        e += -dl[q] ** 2 / (ps + sg) / 2 - mp.log(1 + ps / sg) / 2
            PHI[i][j] = mp.exp(e - mp.mpf(nmiss) * mp.log(2) / 2)
    lnBi = [b + sum(PHI[i][j] * v[j] for j in range(m)) for i in range(n)]             # getPHI.m:116-125
    beta = [mp.exp(-t) for t in lnBi]
    wb = [beta[i] * mp.mpf(float(omega[i])) for i in range(n)]
    alpha = [mp.exp(t) for t in lnA]
    S = mp.matrix(m, m)
    r = mp.matrix(m, 1)
    for a in range(m):
        for c in range(m):
            S[a, c] = sum(wb[i] * PHI[i][a] * PHI[i][c] for i in range(n)) + (alpha[a] if a == c else 0)   # GPz.m:65
        r[a] = sum(wb[i] * PHI[i][a] * mp.mpf(float(Y[i])) for i in range(n))
    w = mp.lu_solve(S, r)                                                              # :70
    logdet = mp.log(mp.det(S))
    delta = [sum(PHI[i][j] * w[j] for j in range(m)) - mp.mpf(float(Y[i])) for i in range(n)]
    tau = [mp.exp(t) for t in lnT]
    nl = -sum(wb[i] * delta[i] ** 2 for i in range(n)) / 2 - sum(alpha[j] * w[j] ** 2 for j in range(m)) / 2 \
        + sum(lnA) / 2 - logdet / 2 - sum(lnBi[i] * mp.mpf(float(omega[i])) for i in range(n)) / 2            # :80-82
    nl += -sum(v[j] ** 2 * tau[j] for j in range(m)) / 2 + sum(lnT) / 2 - mp.mpf(m) * mp.log(2 * mp.pi) / 2   # :103
    nl -= mp.log(2 * mp.pi) * sum(mp.mpf(float(t)) for t in omega) / 2                                        # :110
    return -nl / n                                                                                            # :233


@pytest.mark.parametrize("meth,psi,nan", [("VC", False, False), ("VD", False, False), ("VC", True, False), ("VD", True, False),
                                          ("VC", False, True), ("VD", False, True), ("GC", True, False), ("GC", False, True),
                                          ("GL", True, False), ("VL", False, True), ("GD", True, False), ("VC", True, True), ("VD", True, True),
                                          ("GC", True, True)])
def test_oracle_objective_and_gradient_against_40_digit_differences(meth, psi, nan):
    """The oracle's nlogML and its analytic gradient (GPz.m:113-234 restated) against the 40-digit objective and its central
    differences (h = 1e-15: truncation 1e-30).  fp64 finite differences (the T1 identity tests) cannot see below ~1e-6."""
    from gpz_b200 import synth
    n, d, m = 36, 2, 5
    X, Y = synth.make_data(n, d, seed=11)
    X, Y = np.array(X), np.asarray(Y)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, meth, m, het=True, seed=12), 0.2, 13)
    omega = 0.5 + np.random.default_rng(14).random((n, 1))
    model = O.Model(d=d, k=1, m=m, method=meth, heteroscedastic=True)
    tr = np.ones(n, dtype=bool)
    if nan:                                                                            # two missing-input patterns + complete rows
        X[3:12, 0] = np.nan
        X[20:27, 1] = np.nan
    Psi = synth.make_psi(n, d, meth, seed=15) if psi else None       # VD: n x d variances; VC: d x d x n covariances (fixPsi.m form)
    ref = O.GPz(theta, model, X, Y, Psi, omega, tr, None)
    mp.mp.dps = 40
    th = [mp.mpf(float(t)) for t in theta]
    f0 = _mp_nlogml(th, meth, X, Y[:, 0], omega[:, 0], m, d, Psi)
    assert abs(float((mp.mpf(float(ref.nlogML)) - f0) / f0)) <= 1e-13
    h = mp.mpf(10) ** -15
    g = np.zeros(len(theta))
    for q in range(len(theta)):
        tp, tm = list(th), list(th)
        tp[q] += h
        tm[q] -= h
        g[q] = float((_mp_nlogml(tp, meth, X, Y[:, 0], omega[:, 0], m, d, Psi) - _mp_nlogml(tm, meth, X, Y[:, 0], omega[:, 0], m, d, Psi)) / (2 * h))
    err = np.max(np.abs(ref.grad - g)) / np.max(np.abs(g))
    assert err <= 1e-11, err


def test_train_and_validation_statistics_against_40_digit_arithmetic():
    """GPz.m:236-259: trainRMSE / trainLL on the training rows and validRMSE / validLL from a second getPHI on the validation
    rows (weights omega, heteroscedastic beta), the four numbers callBack.m reads -- oracle against 40 digits (VD, no Psi)."""
    from gpz_b200 import synth
    n, d, m = 48, 2, 5
    X, Y = synth.make_data(n, d, seed=61)
    X, Y = np.asarray(X), np.asarray(Y)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, "VD", m, het=True, seed=62), 0.2, 63)
    omega = 0.5 + np.random.default_rng(64).random((n, 1))
    tr = np.arange(n) % 3 != 0
    va = ~tr
    model = O.Model(d=d, k=1, m=m, method="VD", heteroscedastic=True)
    ref = O.GPz(theta, model, X, Y, None, omega, tr, va)
    mp.mp.dps = 40
    f = lambda t: mp.mpf(float(t))                                                    # noqa: E731
    md = m * d
    o_ = md + md
    G = [[f(theta[md + j + a * m]) for a in range(d)] for j in range(m)]
    Pm = [[f(theta[a * m + j]) for a in range(d)] for j in range(m)]
    lnA = [f(t) for t in theta[o_:o_ + m]]
    b = f(theta[o_ + m])
    v = [f(t) for t in theta[o_ + m + 1:o_ + 2 * m + 1]]

    def phi_row(i):
        return [mp.exp(-sum(((f(X[i, a]) - Pm[j][a]) * G[j][a]) ** 2 for a in range(d)) / 2) for j in range(m)]

    rows_t = [i for i in range(n) if tr[i]]
    PHI = {i: phi_row(i) for i in range(n)}
    beta = {i: mp.exp(-(b + sum(PHI[i][j] * v[j] for j in range(m)))) for i in range(n)}
    S = mp.matrix(m, m)
    r = mp.matrix(m, 1)
    for a in range(m):
        for c in range(m):
            S[a, c] = sum(beta[i] * f(omega[i, 0]) * PHI[i][a] * PHI[i][c] for i in rows_t) + (mp.exp(lnA[a]) if a == c else 0)
        r[a] = sum(beta[i] * f(omega[i, 0]) * PHI[i][a] * f(Y[i, 0]) for i in rows_t)
    w = mp.lu_solve(S, r)

    def stats(rows):
        nr = len(rows)
        delta = {i: sum(PHI[i][j] * w[j] for j in range(m)) - f(Y[i, 0]) for i in rows}
        rmse = mp.sqrt(sum(f(omega[i, 0]) * delta[i] ** 2 for i in rows) / nr)
        ll = sum(f(omega[i, 0]) * (-beta[i] * delta[i] ** 2 / 2 + mp.log(beta[i]) / 2) for i in rows) / nr - mp.log(2 * mp.pi) / 2
        return rmse, ll

    tr_rmse, tr_ll = stats(rows_t)
    va_rmse, va_ll = stats([i for i in range(n) if va[i]])
    rel = lambda a, x: abs(float((f(a) - x) / x))                                     # noqa: E731
    assert rel(ref.stats["trainRMSE"], tr_rmse) <= 1e-12 and rel(ref.stats["trainLL"], tr_ll) <= 1e-12
    assert rel(ref.stats["validRMSE"], va_rmse) <= 1e-12 and rel(ref.stats["validLL"], va_ll) <= 1e-12


def test_predict_full_and_noisy_diag_against_40_digit_arithmetic():
    """predictFull (predictDiag.m:58-74) and predictNoisy for the diagonal modes (predictDiag.m:75-125) of the oracle against
    40-digit evaluations of the same expressions: VD, m = 8, d = 3, 4 rows, w / iSigma_w from the oracle's fit."""
    from gpz_b200 import synth
    n, d, m = 60, 3, 8
    X, Y = synth.make_data(n, d, seed=21)
    X, Y = np.asarray(X), np.asarray(Y)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, "VD", m, het=True, seed=22), 0.2, 23)
    model = O.Model(d=d, k=1, m=m, method="VD", heteroscedastic=True)
    tr = np.ones(n, dtype=bool)
    fit = O.GPz(theta, model, X, Y, None, None, tr, None, fit_only=True)
    model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(1)
    md = m * d
    o = md + model.g_dim
    P = theta[:md].reshape((m, d), order="F")
    vv = theta[o + m + 1:o + 2 * m + 1]
    bb = theta[o + m]
    model.best = dict(theta=theta, w=fit.w, iSigma_w=fit.iSigma_w, P=P, v=vv.reshape(m, 1))
    Xt = X[:4]
    Psi = synth.make_psi(4, d, "VD", seed=24)
    mp.mp.dps = 40
    f = lambda t: mp.mpf(float(t))                                                    # noqa: E731
    G = [[f(theta[md + j + a * m]) for a in range(d)] for j in range(m)]
    Pm = [[f(P[j, a]) for a in range(d)] for j in range(m)]
    w = [f(t) for t in fit.w[:, 0]]
    v = [f(t) for t in vv]
    iSw = fit.iSigma_w[:, :, 0]
    rel = lambda a, b: abs(float((f(a) - b) / b))                                     # noqa: E731

    # ---- predictFull: mu = PHI w, nu = phi' iSigma_w phi, beta_i = exp(b + PHI v)
    mu, sigma, nu, be, ga, PHI = O.predict(Xt, model)
    for t in range(4):
        phi = [mp.exp(-sum(((f(Xt[t, a]) - Pm[j][a]) * G[j][a]) ** 2 for a in range(d)) / 2) for j in range(m)]
        mu_x = sum(phi[j] * w[j] for j in range(m))
        nu_x = sum(phi[i] * f(iSw[i, j]) * phi[j] for i in range(m) for j in range(m))
        be_x = mp.exp(f(bb) + sum(phi[j] * v[j] for j in range(m)))
        assert rel(mu[t, 0], mu_x) <= 1e-12 and rel(be[t, 0], be_x) <= 1e-12 and rel(nu[t, 0], nu_x) <= 1e-9
        assert ga[t, 0] == 0.0 and rel(sigma[t, 0], nu_x + be_x) <= 1e-11

    # ---- predictNoisy (diag)
    mu, sigma, nu, be, ga, PHI = O.predict(Xt, model, Psi=Psi)
    iS = [[G[j][a] ** 2 for a in range(d)] for j in range(m)]
    S = [[1 / iS[j][a] for a in range(d)] for j in range(m)]
    lnz = [-sum(mp.log(iS[j][a]) for a in range(d)) / 2 for j in range(m)]
    for t in range(4):
        x = [f(Xt[t, a]) for a in range(d)]
        ps = [f(Psi[t, a]) for a in range(d)]
        phi = [mp.exp(sum(-(x[a] - Pm[j][a]) ** 2 / (ps[a] + S[j][a]) / 2 - mp.log(1 + ps[a] / S[j][a]) / 2 for a in range(d)))
               for j in range(m)]                                                     # getPHI.m:102-105
        mu_x = sum(phi[j] * w[j] for j in range(m))
        ElnS = f(bb) + sum(phi[j] * v[j] for j in range(m))
        ga_x = nu_x = V_x = mp.mpf(0)
        for i in range(m):
            for j in range(i + 1):
                iC = [iS[i][a] + iS[j][a] for a in range(d)]
                C = [1 / c for c in iC]
                cc = [(Pm[i][a] * iS[i][a] + Pm[j][a] * iS[j][a]) / iC[a] for a in range(d)]
                lnZ = lnz[i] + lnz[j] - sum((Pm[i][a] - Pm[j][a]) ** 2 / (S[i][a] + S[j][a]) for a in range(d)) / 2 \
                    - sum(mp.log(S[i][a] + S[j][a]) for a in range(d)) / 2
                lnN = -sum((x[a] - cc[a]) ** 2 / (ps[a] + C[a]) for a in range(d)) / 2 - sum(mp.log(ps[a] + C[a]) for a in range(d)) / 2
                Z = mp.exp(lnZ + lnN)
                fac = 2 if j < i else 1
                ga_x += fac * Z * w[i] * w[j]
                V_x += fac * Z * v[i] * v[j]
                nu_x += fac * Z * f(iSw[i, j])
        V_x -= (ElnS - f(bb)) ** 2
        ga_x -= mu_x ** 2
        be_x = mp.exp(ElnS) * (1 + V_x / 2)
        assert rel(mu[t, 0], mu_x) <= 1e-12 and rel(be[t, 0], be_x) <= 1e-11
        assert rel(nu[t, 0], nu_x) <= 1e-9 and abs(float(f(ga[t, 0]) - ga_x)) <= 1e-12 * max(1.0, float(mu_x) ** 2)


@pytest.mark.parametrize("psi", [False, True])
def test_predict_missing_diag_against_40_digit_arithmetic(psi):
    """predictMissing (predictDiag.m:127-212) and predictNoisyMissing (predictDiag.m:213-295: the input-noise variances are
    added to the variances of the OBSERVED dims, :230-232 and :262-264) for the diagonal modes of the oracle against a 40-digit
    restatement: VD, m = 6, d = 3, three rows sharing one missing dimension, priors from the oracle's getPrior."""
    from gpz_b200 import synth
    n, d, m = 50, 3, 6
    X, Y = synth.make_data(n, d, seed=31)
    X, Y = np.array(X), np.asarray(Y)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, "VD", m, het=True, seed=32), 0.2, 33)
    model = O.Model(d=d, k=1, m=m, method="VD", heteroscedastic=True)
    tr = np.ones(n, dtype=bool)
    fit = O.GPz(theta, model, X, Y, None, None, tr, None, fit_only=True)
    pri = O.getPrior(X, None, theta, model, tr)
    model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(1)
    md = m * d
    o_ = md + model.g_dim
    P = theta[:md].reshape((m, d), order="F")
    vv = theta[o_ + m + 1:o_ + 2 * m + 1]
    bb = theta[o_ + m]
    model.best = dict(theta=theta, w=fit.w, iSigma_w=fit.iSigma_w, P=P, v=vv.reshape(m, 1), priors=pri)
    Xt = X[:3].copy()
    Xt[:, 1] = np.nan
    Psi = synth.make_psi(3, d, "VD", seed=34) if psi else None
    mu, sigma, nu, be, ga, PHI = O.predict(Xt, model, Psi=Psi)

    mp.mp.dps = 40
    f = lambda t: mp.mpf(float(t))                                                    # noqa: E731
    ob, un = [0, 2], [1]
    G = [[f(theta[md + j + a * m]) for a in range(d)] for j in range(m)]
    Pm = [[f(P[j, a]) for a in range(d)] for j in range(m)]
    w = [f(t) for t in fit.w[:, 0]]
    v = [f(t) for t in vv]
    pr = [f(t) for t in np.asarray(pri).reshape(-1)]
    iSw = fit.iSigma_w[:, :, 0]
    iS = [[G[j][a] ** 2 for a in range(d)] for j in range(m)]
    S = [[1 / iS[j][a] for a in range(d)] for j in range(m)]
    lnz = [-sum(mp.log(iS[j][a]) for a in range(d)) / 2 for j in range(m)]
    Nij = [[mp.exp(-sum((Pm[i][a] - Pm[j][a]) ** 2 / (S[i][a] + S[j][a]) for a in un) / 2
                   - sum(mp.log(S[i][a] + S[j][a]) for a in un) / 2) for j in range(m)] for i in range(m)]     # :161
    rel = lambda a, b: abs(float((f(a) - b) / b))                                     # noqa: E731
    for t in range(3):
        x = {a: f(Xt[t, a]) for a in ob}
        ps = {a: (f(Psi[t, a]) if psi else mp.mpf(0)) for a in ob}
        No = [mp.exp(-sum((x[a] - Pm[i][a]) ** 2 / (S[i][a] + ps[a]) for a in ob) / 2
                     - sum(mp.log(S[i][a] + ps[a]) for a in ob) / 2) for i in range(m)]
        Ex = [No[i] * pr[i] for i in range(m)]                                        # :145-151
        Pio = [e / sum(Ex) for e in Ex]                                               # :153-155
        phi = [mp.exp(lnz[i]) * No[i] * sum(Pio[j] * Nij[i][j] for j in range(m)) for i in range(m)]          # :163-164
        mu_x = sum(phi[i] * w[i] for i in range(m))
        ElnS = sum(phi[i] * v[i] for i in range(m))
        ga_x = nu_x = V_x = mp.mpf(0)
        for i in range(m):
            for j in range(i + 1):
                C = [1 / (iS[i][a] + iS[j][a]) for a in range(d)]
                cc = [(Pm[i][a] * iS[i][a] + Pm[j][a] * iS[j][a]) * C[a] for a in range(d)]
                No_p = mp.exp(-sum((x[a] - cc[a]) ** 2 / (C[a] + ps[a]) for a in ob) / 2
                              - sum(mp.log(C[a] + ps[a]) for a in ob) / 2)                                               # :180 / :263
                Nu = [mp.exp(-sum((Pm[l][a] - cc[a]) ** 2 / (S[l][a] + C[a]) for a in un) / 2
                             - sum(mp.log(S[l][a] + C[a]) for a in un) / 2) for l in range(m)]                          # :184
                Ec = sum(No_p * Nu[l] * Pio[l] for l in range(m))                                                        # :186-187
                Z = mp.exp(lnz[i] + lnz[j] - sum((Pm[i][a] - Pm[j][a]) ** 2 / (S[i][a] + S[j][a]) for a in range(d)) / 2
                           - sum(mp.log(S[i][a] + S[j][a]) for a in range(d)) / 2) * Ec                                  # :190
                fac = 2 if j < i else 1
                ga_x += fac * Z * w[i] * w[j]
                V_x += fac * Z * v[i] * v[j]
                nu_x += fac * Z * f(iSw[i, j])
        V_x -= ElnS ** 2                                                              # :204
        be_x = mp.exp(ElnS + f(bb)) * (1 + V_x / 2)                                   # :206-208
        ga_x -= mu_x ** 2
        for i in range(m):
            assert rel(PHI[t, i], phi[i]) <= 1e-12
        assert rel(mu[t, 0], mu_x) <= 1e-12 and rel(be[t, 0], be_x) <= 1e-11
        assert rel(nu[t, 0], nu_x) <= 1e-9 and abs(float(f(ga[t, 0]) - ga_x)) <= 1e-12 * max(1.0, float(mu_x) ** 2)


def test_get_prior_against_40_digit_arithmetic():
    """getPrior.m:7-20 (EM over the mixture weights of the normalised basis densities N of getPHI.m:98,114) in 40-digit
    arithmetic against the oracle: VD, no Psi, one missing-input pattern among the rows."""
    from gpz_b200 import synth
    n, d, m = 40, 2, 5
    X, Y = synth.make_data(n, d, seed=41)
    X, Y = np.array(X), np.asarray(Y)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, "VD", m, het=True, seed=42), 0.2, 43)
    X[5:11, 1] = np.nan
    model = O.Model(d=d, k=1, m=m, method="VD", heteroscedastic=True)
    tr = np.ones(n, dtype=bool)
    pri = np.asarray(O.getPrior(X, None, theta, model, tr)).reshape(-1)
    mp.mp.dps = 40
    f = lambda t: mp.mpf(float(t))                                                    # noqa: E731
    md = m * d
    G = [[f(theta[md + j + a * m]) for a in range(d)] for j in range(m)]
    Pm = [[f(theta[a * m + j]) for a in range(d)] for j in range(m)]
    N = []
    for i in range(n):
        ob = [a for a in range(d) if not np.isnan(X[i, a])]
        row = []
        for j in range(m):
            sg = {a: 1 / G[j][a] ** 2 for a in ob}
            lnphi = -sum((f(X[i, a]) - Pm[j][a]) ** 2 / sg[a] for a in ob) / 2 - mp.mpf(d - len(ob)) * mp.log(2) / 2
            lnN = lnphi - sum(mp.log(sg[a]) for a in ob) / 2 - mp.mpf(len(ob)) * mp.log(2 * mp.pi) / 2 \
                + mp.mpf(d - len(ob)) * mp.log(2) / 2                                 # getPHI.m:98
            row.append(mp.exp(lnN))
        N.append(row)
    prior = [mp.mpf(1) / m] * m
    for _ in range(100):
        old = list(prior)
        W = [[N[i][j] * prior[j] for j in range(m)] for i in range(n)]
        W = [[t / sum(r) for t in r] for r in W]
        prior = [sum(W[i][j] for i in range(n)) / n for j in range(m)]
        num = mp.sqrt(sum((a - b) ** 2 for a, b in zip(old, prior)))
        den = mp.sqrt(sum((a + b) ** 2 for a, b in zip(old, prior)))
        if num / den < mp.mpf(10) ** -10:
            break
    err = max(abs(float(f(pri[j]) - prior[j])) for j in range(m))
    assert err <= 1e-9, err
    assert abs(float(sum(prior)) - 1.0) <= 1e-30 and abs(pri.sum() - 1.0) <= 1e-12


@pytest.mark.parametrize("psi", [False, True])
def test_predict_missing_cov_against_40_digit_arithmetic(psi):
    """predictMissing (predictCov.m:134-232) and predictNoisyMissing (predictCov.m:233-336: Psi(o,o) joins the observed block,
    Psi_hat = T Psi_oo T' + Schur with T = [I; R'], :262-270) for the covariance modes of the oracle against a 40-digit restatement: VC, m = 4,
    d = 3, two rows with the middle dimension missing.  Per basis l the missing block is completed by regression on the
    observed one (R_l, X_hat_l, Psi_hat_l, :165-172); PHI_ti = e^{lnz_i} sum_j N(X_hat_tj; p_i, Sigma_i + Psi_hat_j) Pio_tj and the
    pair sums use N(X_hat_tl; c_ij, C_ij + Psi_hat_l) (:186-206)."""
    from gpz_b200 import synth
    n, d, m = 50, 3, 4
    X, Y = synth.make_data(n, d, seed=51)
    X, Y = np.array(X), np.asarray(Y)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, "VC", m, het=True, seed=52), 0.2, 53)
    model = O.Model(d=d, k=1, m=m, method="VC", heteroscedastic=True)
    tr = np.ones(n, dtype=bool)
    fit = O.GPz(theta, model, X, Y, None, None, tr, None, fit_only=True)
    pri = O.getPrior(X, None, theta, model, tr)
    model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(1)
    md = m * d
    o_ = md + model.g_dim
    P = theta[:md].reshape((m, d), order="F")
    vv = theta[o_ + m + 1:o_ + 2 * m + 1]
    bb = theta[o_ + m]
    model.best = dict(theta=theta, w=fit.w, iSigma_w=fit.iSigma_w, P=P, v=vv.reshape(m, 1), priors=pri)
    Xt = X[:2].copy()
    Xt[:, 1] = np.nan
    Psi = synth.make_psi(2, d, "VC", seed=54) if psi else None
    mu, sigma, nu, be, ga, PHI = O.predict(Xt, model, Psi=Psi)

    mp.mp.dps = 40
    f = lambda t: mp.mpf(float(t))                                                    # noqa: E731
    M = lambda a: mp.matrix(np.asarray(a, dtype=np.float64).tolist())                 # noqa: E731
    ob, un = [0, 2], [1]
    sub = lambda A, r, c: mp.matrix([[A[i, j] for j in c] for i in r])               # noqa: E731
    Gam = O.unpack_gamma(theta, model)
    iS = [M(Gam[:, :, i]).T * M(Gam[:, :, i]) for i in range(m)]
    S = [mp.inverse(a) for a in iS]
    lnz = [-mp.log(mp.det(a)) / 2 for a in iS]
    Pm = [M(P[i:i + 1, :]) for i in range(m)]
    w = [f(t) for t in fit.w[:, 0]]
    v = [f(t) for t in vv]
    pr = [f(t) for t in np.asarray(pri).reshape(-1)]
    iSw = fit.iSigma_w[:, :, 0]

    def lnN(delta, C):
        return -(delta * mp.inverse(C) * delta.T)[0] / 2 - mp.log(mp.det(C)) / 2

    rel = lambda a, b: abs(float((f(a) - b) / b))                                     # noqa: E731
    for t in range(2):
        xo = mp.matrix([[f(Xt[t, a]) for a in ob]])
        Ex, Xh, Ph = [], [], []
        for i in range(m):
            Soo = sub(S[i], ob, ob)
            dl = xo - sub(Pm[i], [0], ob)
            Poo = sub(M(Psi[:, :, t]), ob, ob) if psi else mp.zeros(len(ob), len(ob))
            Ex.append(mp.exp(lnN(dl, Soo + Poo)) * pr[i])                             # :163-164 / :260
            R = mp.inverse(Soo) * sub(S[i], ob, un)                                   # :166
            ph = mp.zeros(d, d)
            blk = sub(S[i], un, un) - sub(S[i], un, ob) * R                           # :168
            if psi:                                                                   # T Psi_oo T', T = [I; R']  (:264-268)
                order = ob + un
                Tm = mp.zeros(d, len(ob))
                for a_ in range(len(ob)):
                    Tm[a_, a_] = 1
                Rt = R.T
                for a_ in range(len(un)):
                    for b_ in range(len(ob)):
                        Tm[len(ob) + a_, b_] = Rt[a_, b_]
                TP = Tm * Poo * Tm.T
                for a_, ia in enumerate(order):
                    for b_, ib in enumerate(order):
                        ph[ia, ib] = TP[a_, b_]
            for a_, ia in enumerate(un):
                for b_, ib in enumerate(un):
                    ph[ia, ib] += blk[a_, b_]
            xh = mp.zeros(1, d)
            xu = dl * R + sub(Pm[i], [0], un)                                         # :170
            for a_, ia in enumerate(un):
                xh[0, ia] = xu[0, a_]
            for a_, ia in enumerate(ob):
                xh[0, ia] = xo[0, a_]
            Ph.append(ph)
            Xh.append(xh)
        Pio = [e / sum(Ex) for e in Ex]
        phi = [mp.exp(lnz[i]) * sum(mp.exp(lnN(Xh[j] - Pm[i], S[i] + Ph[j])) * Pio[j] for j in range(m)) for i in range(m)]
        mu_x = sum(phi[i] * w[i] for i in range(m))
        ElnS = sum(phi[i] * v[i] for i in range(m))
        ga_x = nu_x = V_x = mp.mpf(0)
        for i in range(m):
            for j in range(i + 1):
                iC = iS[i] + iS[j]
                C = mp.inverse(iC)
                c = (Pm[i] * iS[i] + Pm[j] * iS[j]) * C                               # :180-182
                Ec = sum(mp.exp(lnN(Xh[l] - c, C + Ph[l])) * Pio[l] for l in range(m))                       # :196-201
                Z = mp.exp(lnz[i] + lnz[j] + lnN(Pm[i] - Pm[j], S[i] + S[j])) * Ec                           # :203-204
                fac = 2 if j < i else 1
                ga_x += fac * Z * w[i] * w[j]
                V_x += fac * Z * v[i] * v[j]
                nu_x += fac * Z * f(iSw[i, j])
        V_x -= ElnS ** 2
        be_x = mp.exp(ElnS + f(bb)) * (1 + V_x / 2)
        ga_x -= mu_x ** 2
        for i in range(m):
            assert rel(PHI[t, i], phi[i]) <= 1e-11
        assert rel(mu[t, 0], mu_x) <= 1e-11 and rel(be[t, 0], be_x) <= 1e-10
        assert rel(nu[t, 0], nu_x) <= 1e-8 and abs(float(f(ga[t, 0]) - ga_x)) <= 1e-11 * max(1.0, float(mu_x) ** 2)


def _mp_nlogml_multi(theta, X, Y, omega, m, d, k):
    """GPz.m:20-110,233 for k outputs, VD, heteroscedastic, no Psi / NaN, in 40-digit arithmetic.  Keeps the reference's treatment of
    the constant: -1/2 m k ln(2 pi) is added to EACH of the k per-output terms (GPz.m:103) before they are summed (:110)."""
    n = X.shape[0]
    md = m * d
    P = [[theta[a * m + j] for a in range(d)] for j in range(m)]
    Gam = [[theta[md + j + a * m] for a in range(d)] for j in range(m)]
    o = 2 * md
    lnA = [[theta[o + j + c * m] for j in range(m)] for c in range(k)]
    b = [theta[o + m * k + c] for c in range(k)]
    v = [[theta[o + m * k + k + j + c * m] for j in range(m)] for c in range(k)]
    lnT = [[theta[o + m * k + k + m * k + j + c * m] for j in range(m)] for c in range(k)]
    PHI = [[mp.exp(-sum(((mp.mpf(float(X[i, a])) - P[j][a]) * Gam[j][a]) ** 2 for a in range(d)) / 2) for j in range(m)] for i in range(n)]
    total = mp.mpf(0)
    for c in range(k):
        lnBi = [b[c] + sum(PHI[i][j] * v[c][j] for j in range(m)) for i in range(n)]
        wb = [mp.exp(-lnBi[i]) * mp.mpf(float(omega[i])) for i in range(n)]
        alpha = [mp.exp(t) for t in lnA[c]]
        S = mp.matrix(m, m)
        r = mp.matrix(m, 1)
        for a in range(m):
            for e in range(m):
                S[a, e] = sum(wb[i] * PHI[i][a] * PHI[i][e] for i in range(n)) + (alpha[a] if a == e else 0)
            r[a] = sum(wb[i] * PHI[i][a] * mp.mpf(float(Y[i, c])) for i in range(n))
        w = mp.lu_solve(S, r)
        delta = [sum(PHI[i][j] * w[j] for j in range(m)) - mp.mpf(float(Y[i, c])) for i in range(n)]
        nl = -sum(wb[i] * delta[i] ** 2 for i in range(n)) / 2 - sum(alpha[j] * w[j] ** 2 for j in range(m)) / 2 \
            + sum(lnA[c]) / 2 - mp.log(mp.det(S)) / 2 - sum(lnBi[i] * mp.mpf(float(omega[i])) for i in range(n)) / 2
        nl += -sum(v[c][j] ** 2 * mp.exp(lnT[c][j]) for j in range(m)) / 2 + sum(lnT[c]) / 2 - mp.mpf(m * k) * mp.log(2 * mp.pi) / 2
        total += nl
    total -= mp.log(2 * mp.pi) * sum(mp.mpf(float(t)) for t in omega) / 2
    return -total / (n * k)


def test_two_outputs_objective_and_gradient_against_40_digit_differences():
    """k = 2: per-output Gram / solve / noise head, shared PHI and basis parameters, and the reference's constant
    (GPz.m:103: counted k times) -- the oracle's nlogML and gradient against the 40-digit objective and its differences."""
    from gpz_b200 import synth
    n, d, m, k = 30, 2, 4, 2
    X, Y = synth.make_data(n, d, seed=71, k=k)
    X, Y = np.asarray(X), np.asarray(Y)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, "VD", m, het=True, seed=72), 0.2, 73)
    omega = 0.5 + np.random.default_rng(74).random((n, 1))
    model = O.Model(d=d, k=k, m=m, method="VD", heteroscedastic=True)
    tr = np.ones(n, dtype=bool)
    ref = O.GPz(theta, model, X, Y, None, omega, tr, None)
    mp.mp.dps = 40
    th = [mp.mpf(float(t)) for t in theta]
    f0 = _mp_nlogml_multi(th, X, Y, omega[:, 0], m, d, k)
    assert abs(float((mp.mpf(float(ref.nlogML)) - f0) / f0)) <= 1e-13
    h = mp.mpf(10) ** -15
    g = np.zeros(len(theta))
    for q in range(len(theta)):
        tp, tm = list(th), list(th)
        tp[q] += h
        tm[q] -= h
        g[q] = float((_mp_nlogml_multi(tp, X, Y, omega[:, 0], m, d, k) - _mp_nlogml_multi(tm, X, Y, omega[:, 0], m, d, k)) / (2 * h))
    err = np.max(np.abs(ref.grad - g)) / np.max(np.abs(g))
    assert err <= 1e-11, err


def test_inv_logdet_against_40_digit_svd():
    """inv_logdet.m:3-15: SVD pseudo-inverse with tol = max(size) * eps(largest singular value) and log det over the retained
    singular values -- the oracle against a 40-digit SVD, for a well-conditioned SPD matrix and a rank-3 one of size 6."""
    rng = np.random.default_rng(81)
    mp.mp.dps = 40
    for rank in (6, 3):
        B = rng.standard_normal((6, rank))
        A = B @ B.T + (np.eye(6) if rank == 6 else 0.0)
        Xi, ld = O.inv_logdet(A)
        U, s, V = mp.svd_r(mp.matrix(A.tolist()))
        smax = max(s[i] for i in range(6))
        tol = 6 * float(np.spacing(float(smax)))
        keep = [i for i in range(6) if s[i] > tol]
        assert len(keep) == rank
        Xi_x = mp.zeros(6, 6)
        for i in keep:
            Xi_x += (V[i, :].T * U[:, i].T) / s[i]                                    # V(:,i) U(:,i)' / s_i
        ld_x = sum(mp.log(s[i]) for i in keep)
        err = max(abs(float(mp.mpf(float(Xi[a, b])) - Xi_x[a, b])) for a in range(6) for b in range(6))
        scale = max(abs(float(Xi_x[a, b])) for a in range(6) for b in range(6))
        assert err <= 1e-12 * scale * (1 if rank == 6 else float(smax / min(s[i] for i in keep)))
        assert abs(ld - float(ld_x)) <= 1e-12 * max(1.0, abs(float(ld_x)))
