"""CPU oracle for the GPz NLML/gradient/predict hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy fp64 restatement of the reference's MATLAB algorithm.  It exists so
that the CUDA path can be checked against something that follows the reference operation
by operation.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import it; the product (``gpz_b200``) never does.

PARITY PIN STATUS: **parity unpinned by the reference** -- the reference (pure MATLAB, no
MATLAB/Octave here) ships no tests, no golden vectors and no recorded outputs for this path
(SURVEY.md section 4 / 8c).  The arithmetic lives in MathWorks MATLAB built-ins (mtimes, svd,
inv, mrdivide, exp; version unpinned, not under /root/reference).  What pins this restatement
instead are the mathematical identities in tests/test_oracle_pins.py (finite-difference exact
gradient, independent dense-GP evidence, six-mode equivalence, Psi=0 identity, predictNoisy ->
predictFull limit, fit-path definitions), self-derived golden vectors in tests/golden/, and
tests/test_oracle_highprec.py: the objective of GPz.m restated on its own in 40-digit arithmetic
(all six modes, with Psi and with missing inputs) against this file's nlogML (1e-13) and, by
central differences of that objective, against its analytic gradient (1e-11); the predict
branches (Full, Noisy, Missing, NoisyMissing; diagonal and covariance modes) and getPrior against
40-digit restatements.

Every function cites the reference file:line it follows (paths under /root/reference/GPz).
MATLAB semantics kept: column-major reshapes (order='F'), 1-based slices translated to
0-based, logical masks, NaN == missing, groups of rows by NaN pattern in order of first
appearance.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

LN2 = math.log(2.0)
LN2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# model "struct"  (init.m:16-20, 86)
# --------------------------------------------------------------------------------------
@dataclass
class Model:
    d: int
    k: int
    m: int
    method: str            # 'GL','VL','GD','VD','GC','VC'
    heteroscedastic: bool
    g_dim: int = 0
    muX: np.ndarray | None = None
    sdX: np.ndarray | None = None
    muY: np.ndarray | None = None
    last: dict = field(default_factory=dict)
    best: dict = field(default_factory=dict)

    def __post_init__(self):
        if self.g_dim == 0:
            self.g_dim = g_dim_of(self.method, self.m, self.d)


def g_dim_of(method: str, m: int, d: int) -> int:
    """length(Gamma(:)) for each covariance mode (init.m:65-86)."""
    return {"GL": 1, "VL": m, "GD": d, "VD": m * d, "GC": d * d, "VC": d * d * m}[method]


def theta_len(model: Model) -> int:
    """init.m:87,97: [P(:); Gamma(:); lnAlpha(:); b(:); v(:); lnTau(:)]."""
    m, d, k = model.m, model.d, model.k
    p = m * d + model.g_dim + m * k + k
    if model.heteroscedastic:
        p += 2 * m * k
    return p


def _F(a, shape):
    return np.reshape(a, shape, order="F")


def unpack_gamma(theta, model: Model):
    """Gamma expansion per mode (getPHI.m:26-40).  Diag modes -> m x d, C modes -> d x d x m."""
    m, d = model.m, model.d
    md = m * d
    meth = model.method
    if meth == "GL":
        return np.full((m, d), theta[md])
    if meth == "VL":
        return np.repeat(theta[md:md + m].reshape(m, 1), d, axis=1)
    if meth == "GD":
        return np.repeat(theta[md:md + d].reshape(1, d), m, axis=0)
    if meth == "VD":
        return _F(theta[md:md + md], (m, d))
    if meth == "GC":
        G = _F(theta[md:md + d * d], (d, d))
        return np.repeat(G[:, :, None], m, axis=2)
    if meth == "VC":
        return _F(theta[md:md + d * d * m], (d, d, m))
    raise ValueError(meth)


def nan_groups(missing: np.ndarray):
    """Rows grouped by identical NaN pattern, in order of first appearance
    (getPHI.m:43-54, GPz.m:118-129, predict.m:45-56).  Returns list of boolean masks."""
    n = missing.shape[0]
    todo = np.ones(n, dtype=bool)
    groups = []
    while todo.any():
        first = int(np.argmax(todo))
        grp = np.zeros(n, dtype=bool)
        same = (missing[todo] == missing[first]).all(axis=1)
        grp[np.flatnonzero(todo)[same]] = True
        groups.append(grp)
        todo[grp] = False
    return groups


# --------------------------------------------------------------------------------------
# Dxy.m:1-10
# --------------------------------------------------------------------------------------
def Dxy(X, Y):
    """Pairwise squared Euclidean distance through the quadratic expansion (Dxy.m:3-7)."""
    xx = np.sum(X ** 2, axis=1)[:, None]
    yy = np.sum(Y ** 2, axis=1)[None, :]
    yb = X @ Y.T
    return np.abs(np.abs(yy + (xx - 2.0 * yb)))


# --------------------------------------------------------------------------------------
# inv_logdet.m:1-15
# --------------------------------------------------------------------------------------
def inv_logdet(A):
    """SVD pseudo-inverse + log-determinant over retained singular values (inv_logdet.m:3-15)."""
    U, s, Vt = np.linalg.svd(A, full_matrices=False)
    tol = max(A.shape) * np.spacing(np.max(np.abs(s)))      # max(size(X))*eps(norm(s,inf))
    r = int(np.sum(s > tol))
    U, s, Vt = U[:, :r], s[:r], Vt[:r, :]
    Xi = (Vt.T / s[None, :]) @ U.T
    return Xi, float(np.sum(np.log(s)))


# --------------------------------------------------------------------------------------
# fixPsi.m:1-55
# --------------------------------------------------------------------------------------
def fixPsi(Psi, n, sdX, method):
    """Normalise the input-noise layout: -> n x d (L/D modes) or d x d x n (C modes), scaled by
    the z-scoring of X (fixPsi.m:4-54)."""
    if Psi is None or np.size(Psi) == 0:
        return None
    sdX = np.asarray(sdX, dtype=np.float64).reshape(-1)
    d = sdX.size
    Psi = np.asarray(Psi, dtype=np.float64)
    shp = Psi.shape + (1,) * (3 - Psi.ndim)
    cube = (shp[0] == d and shp[1] == d and shp[2] == n)
    outer = np.outer(sdX, sdX)
    if method[1] == "C":
        new = np.zeros((d, d, n))
        if not cube:
            if shp[1] == 1:
                P1 = Psi.reshape(-1)
                for i in range(n):
                    new[:, :, i] = (np.eye(d) * P1[i]) / outer
            else:
                for i in range(n):
                    new[:, :, i] = np.diag(Psi[i, :] / sdX ** 2)
        else:
            for i in range(n):
                new[:, :, i] = Psi[:, :, i] / outer
        return new
    if not cube:
        if shp[1] == 1:
            return np.repeat(Psi.reshape(n, 1), d, axis=1) / sdX[None, :] ** 2
        return Psi / sdX[None, :] ** 2
    new = np.zeros((n, d))
    for i in range(n):
        new[i, :] = np.diag(Psi[:, :, i] / outer)
    return new


# --------------------------------------------------------------------------------------
# getPHI.m:1-127
# --------------------------------------------------------------------------------------
def getPHI(X, Psi, theta, model: Model, selection=None, want_N=True):
    """Design matrix PHI (n x m), expanded Gamma, ln-noise lnBeta_i (n x k) and normalised
    densities N (getPHI.m:1-127).  Per-group / per-basis loops as in the reference."""
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    if selection is None or np.size(selection) == 0:
        selection = np.ones(X.shape[0], dtype=bool)
    selection = np.asarray(selection, dtype=bool).reshape(-1)
    n = int(selection.sum())
    d, m, k = model.d, model.m, model.k
    meth = model.method
    X = X[selection, :]                                                 # :14
    if Psi is not None:
        Psi = Psi[:, :, selection] if meth[1] == "C" else Psi[selection, :]   # :16-22
    P = _F(theta[:m * d], (m, d))                                       # :24
    Gamma = unpack_gamma(theta, model)                                  # :26-40
    missing = np.isnan(X)
    groups = nan_groups(missing)                                        # :43-54
    lnPHI = np.zeros((n, m))
    lnN = np.zeros((n, m))
    for grp in groups:                                                  # :60
        first = int(np.argmax(grp))
        u = np.isnan(X[first, :])
        o = ~u
        nu_, no_ = int(u.sum()), int(o.sum())
        idx = np.flatnonzero(grp)
        for j in range(m):                                              # :67
            Delta = X[:, o] - P[j, o][None, :]                          # :69
            if meth[1] == "C":
                Sigma = np.linalg.inv(Gamma[:, :, j].T @ Gamma[:, :, j])     # :73
                Soo = Sigma[np.ix_(o, o)]
                if Psi is None:
                    Dg = Delta[grp, :]
                    lnPHI[grp, j] = -0.5 * np.sum(np.linalg.solve(Soo.T, Dg.T).T * Dg, axis=1) \
                        - 0.5 * nu_ * LN2                               # :76  (Delta/Soo)
                    lnN[grp, j] = lnPHI[grp, j] - 0.5 * np.sum(np.log(np.linalg.svd(Soo, compute_uv=False))) \
                        - 0.5 * no_ * LN2PI + 0.5 * nu_ * LN2           # :77
                else:
                    lsS = np.sum(np.log(np.linalg.svd(Soo, compute_uv=False)))
                    for i in idx:                                       # :82
                        PpS = Psi[np.ix_(o, o, [i])][:, :, 0] + Soo     # :84
                        Di = Delta[i, :]
                        lnPHI[i, j] = -0.5 * np.sum(np.linalg.solve(PpS.T, Di) * Di) + 0.5 * lsS \
                            - 0.5 * np.sum(np.log(np.linalg.svd(PpS, compute_uv=False))) - 0.5 * nu_ * LN2   # :86
                        lnN[i, j] = lnPHI[i, j] - 0.5 * lsS - 0.5 * no_ * LN2PI + 0.5 * nu_ * LN2            # :87
            else:
                Sigma = Gamma[j, o] ** -2.0                             # :93
                if Psi is None:
                    lnPHI[grp, j] = -0.5 * np.sum(Delta[grp, :] ** 2 / Sigma[None, :], axis=1) - 0.5 * nu_ * LN2   # :97
                else:
                    Pg = Psi[np.ix_(grp, o)]
                    PpS = Pg + Sigma[None, :]                           # :102
                    lnPHI[grp, j] = -0.5 * np.sum(Delta[grp, :] ** 2 / PpS, axis=1) \
                        - 0.5 * np.sum(np.log(1.0 + Pg / Sigma[None, :]), axis=1) - 0.5 * nu_ * LN2          # :104
                lnN[grp, j] = lnPHI[grp, j] - 0.5 * np.sum(np.log(Sigma)) - 0.5 * no_ * LN2PI + 0.5 * nu_ * LN2  # :98,105
    PHI = np.exp(lnPHI)                                                 # :113
    N = np.exp(lnN) if want_N else None                                 # :114
    g_dim = model.g_dim
    off = m * d + g_dim + m * k
    b = theta[off:off + k].reshape(1, k)                                # :117
    lnBeta_i = np.repeat(b, n, axis=0)                                  # :119
    if model.heteroscedastic:
        v = _F(theta[off + k:off + k + m * k], (m, k))                  # :122
        lnBeta_i = lnBeta_i + PHI @ v                                   # :124
    return PHI, Gamma, lnBeta_i, N


# --------------------------------------------------------------------------------------
# getPrior.m:1-22
# --------------------------------------------------------------------------------------
def getPrior(X, Psi, theta, model: Model, selection=None):
    """EM for the mixture weights of the bases (getPrior.m:5-20); N does not change between iterations."""
    m = model.m
    prior = np.ones((1, m)) / m
    _, _, _, N = getPHI(X, Psi, theta, model, selection)
    for _ in range(100):
        old = prior
        w = N * prior
        w = w / np.sum(w, axis=1, keepdims=True)
        prior = np.mean(w, axis=0, keepdims=True)
        if np.linalg.norm(old - prior) / np.linalg.norm(old + prior) < 1e-10:
            break
    return prior


# --------------------------------------------------------------------------------------
# GPz.m:1-263
# --------------------------------------------------------------------------------------
@dataclass
class GPzResult:
    nlogML: object          # scalar (eval) or 1 x k array (fit exit, un-normalised: GPz.m:84-87)
    grad: object            # p-vector, or 0 on the fit exit
    w: np.ndarray
    iSigma_w: np.ndarray    # m x m x k
    PHI: np.ndarray
    stats: dict             # trainRMSE, trainLL, validRMSE, validLL  (the four globals, GPz.m:3-7)


def GPz(theta, model: Model, X, Y, Psi=None, omega=None, training=None, validation=None, fit_only=False):
    """NLML objective + gradient (GPz.m:1-263).  ``fit_only=True`` is the nargout>2 early exit
    (GPz.m:84-87).  The four global side-channel scalars are returned in ``.stats``."""
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    k, m = model.k, model.m
    meth = model.method
    n_all, d = X.shape
    if training is None or np.size(training) == 0:
        training = np.ones(n_all, dtype=bool)                           # :16-18
    training = np.asarray(training, dtype=bool).reshape(-1)
    if omega is None or np.size(omega) == 0:
        omega = np.ones((n_all, 1))                                     # :20-22
    omega = np.asarray(omega, dtype=np.float64).reshape(n_all, -1)
    n = int(training.sum())                                             # :24
    g_dim = model.g_dim
    P = _F(theta[:m * d], (m, d))                                       # :28
    PHI, Gamma, lnBeta_i, _ = getPHI(X, Psi, theta, model, training)    # :30
    o1 = m * d + g_dim
    lnAlpha = _F(theta[o1:o1 + m * k], (m, k))                          # :32
    Yt = Y[training, :]
    om = omega[training, :]
    beta = np.exp(-lnBeta_i)                                            # :43
    beta_i = 1.0 / beta                                                 # :44
    df = -beta                                                          # :45
    omega_x_beta = beta * om                                            # :48
    alpha = np.exp(lnAlpha)                                             # :50
    da = alpha
    nu = np.zeros((n, k))
    w = np.zeros((m, k))
    iSigma_w = np.zeros((m, m, k))
    logdet = np.zeros(k)
    dwda = np.zeros((m, k))
    dlnPHI = np.zeros((n, m))
    dlnAlpha = np.zeros((m, k))
    for i in range(k):                                                  # :61
        BxPHI = PHI * omega_x_beta[:, [i]]                              # :63
        SIGMA = BxPHI.T @ PHI + np.diag(alpha[:, i])                    # :65
        iS, logdet[i] = inv_logdet(SIGMA)                               # :67
        iSigma_w[:, :, i] = iS
        nu[:, i] = np.sum(PHI * (PHI @ iS), axis=1)                     # :69
        w[:, i] = (iS @ BxPHI.T) @ Yt[:, i]                             # :70 (left-to-right)
        dwda[:, i] = -iS @ (da[:, i] * w[:, i])                         # :71
        dlnPHI = dlnPHI - BxPHI @ iS                                    # :72
        dlnAlpha[:, i] = -0.5 * np.diag(iS) * da[:, i]                  # :73
    delta = PHI @ w - Yt                                                # :77
    omega_beta_x_delta = omega_x_beta * delta                           # :79
    nlogML = -0.5 * np.sum(omega_beta_x_delta * delta, axis=0) - 0.5 * np.sum(alpha * w ** 2, axis=0) \
        + 0.5 * np.sum(lnAlpha, axis=0) - 0.5 * logdet                  # :81
    nlogML = nlogML + 0.5 * np.sum(-lnBeta_i * om, axis=0)              # :82
    stats = {}
    if fit_only:                                                        # :84-87
        return GPzResult(nlogML.reshape(1, k), 0.0, w, iSigma_w, PHI, stats)
    dlnAlpha = dlnAlpha - (PHI.T @ omega_beta_x_delta) * dwda - alpha * w * dwda - 0.5 * da * w ** 2 + 0.5   # :89
    dlnPHI = dlnPHI - omega_beta_x_delta @ w.T                          # :90
    dbeta = 0.5 * df * (beta_i - (delta ** 2 + nu)) * om                # :93
    db = np.sum(dbeta, axis=0)                                          # :94
    if model.heteroscedastic:                                           # :96
        o2 = o1 + m * k + k
        v = _F(theta[o2:o2 + m * k], (m, k))                            # :98
        lnTau = _F(theta[o2 + m * k:o2 + 2 * m * k], (m, k))            # :100
        tau = np.exp(lnTau)
        nlogML = nlogML - 0.5 * np.sum(v ** 2 * tau, axis=0) + 0.5 * np.sum(lnTau, axis=0) \
            - 0.5 * m * k * LN2PI                                       # :103 (m*k on every column)
        dv = PHI.T @ dbeta - v * tau                                    # :104
        dlnTau = -0.5 * tau * v ** 2 + 0.5                              # :105
        dlnPHI = dlnPHI + dbeta @ v.T                                   # :106
    nlogML = np.sum(nlogML) - 0.5 * LN2PI * np.sum(om)                  # :110
    dPHI = dlnPHI * PHI                                                 # :113
    dP = np.zeros_like(P)
    dGamma = np.zeros_like(Gamma)
    Xt = X[training, :]
    missing = np.isnan(Xt)                                              # :118
    groups = nan_groups(missing)                                        # :120-129
    lst = np.flatnonzero(training)                                      # :131
    for grp in groups:                                                  # :133
        first = int(np.argmax(grp))
        u = missing[first, :]
        o = ~u
        gidx = np.flatnonzero(grp)
        for j in range(m):                                              # :135
            Delta = Xt[:, o] - P[j, o][None, :]                         # :142
            dPj = dPHI[grp, j]
            if meth[1] == "C":
                G = Gamma[:, :, j]
                iSigma = G.T @ G                                        # :146
                Sigma = np.linalg.inv(iSigma)                           # :147
                Soo = Sigma[np.ix_(o, o)]
                if u.any():
                    GuuGuo = np.linalg.solve(iSigma[np.ix_(u, u)], iSigma[np.ix_(u, o)])   # :156,178
                else:
                    GuuGuo = np.zeros((0, int(o.sum())))
                Gproj = G[:, o] - G[:, u] @ GuuGuo
                if Psi is None:
                    iSoo = np.linalg.inv(Soo)                           # :151
                    Dg = Delta[grp, :]
                    dP[j, o] += (dPj @ Dg) @ iSoo                       # :152
                    diSoo = -0.5 * (Dg * dPj[:, None]).T @ Dg           # :154
                    dGo = 2.0 * Gproj @ diSoo                           # :157
                    dGamma[:, o, j] += dGo                              # :158
                    dGamma[:, u, j] -= dGo @ GuuGuo.T                   # :159
                else:
                    iSoo = np.linalg.inv(Soo)
                    for t in gidx:                                      # :168
                        iPS = np.linalg.inv(Soo + Psi[np.ix_(o, o, [lst[t]])][:, :, 0])   # :170
                        Dt = Delta[t, :][None, :]
                        dP[j, o] += (dPHI[t, j] * Dt @ iPS).reshape(-1)                  # :172
                        dSoo = 0.5 * (iSoo - iPS + iPS @ (Dt.T @ Dt) @ iPS)              # :174
                        diSoo = -Soo @ dSoo @ Soo                                        # :176
                        dGo = 2.0 * Gproj @ diSoo                                        # :179
                        dGamma[:, o, j] += dPHI[t, j] * dGo                              # :180
                        dGamma[:, u, j] -= dPHI[t, j] * dGo @ GuuGuo.T                   # :181
            else:
                Sigma = Gamma[j, o] ** -2.0                             # :189
                if Psi is None:
                    Dg = Delta[grp, :]
                    dP[j, o] += (dPj @ Dg) / Sigma                      # :192
                    dGamma[j, o] -= Gamma[j, o] * np.sum(Dg ** 2 * dPj[:, None], axis=0)   # :194
                else:
                    Pg = Psi[np.ix_(lst[grp], o)]
                    PpS = Pg + Sigma[None, :]                           # :200
                    Dg = Delta[grp, :]
                    dP[j, o] += dPj @ (Dg / PpS)                        # :202
                    PxiS = 1.0 / (1.0 + Pg / Sigma[None, :])            # :204
                    dGamma[j, o] -= Gamma[j, o] * (dPj @ (Dg * PxiS) ** 2
                                                   - dPj @ (PxiS * Sigma[None, :] - Sigma[None, :]))   # :206
    if meth == "GL":                                                    # :215-225
        dGamma = np.array([dGamma.sum()])
    elif meth == "VL":
        dGamma = dGamma.sum(axis=1)
    elif meth == "GD":
        dGamma = dGamma.sum(axis=0)
    elif meth == "GC":
        dGamma = dGamma.sum(axis=2)
    grad = np.concatenate([dP.reshape(-1, order="F"), np.asarray(dGamma).reshape(-1, order="F"),
                           dlnAlpha.reshape(-1, order="F"), db.reshape(-1)])        # :227
    if model.heteroscedastic:
        grad = np.concatenate([grad, dv.reshape(-1, order="F"), dlnTau.reshape(-1, order="F")])   # :230
    nlogML = -nlogML / (n * k)                                          # :233
    grad = -grad / (n * k)                                              # :234
    stats["trainRMSE"] = math.sqrt(np.sum(delta ** 2 * om) / (n * k))   # :236
    stats["trainLL"] = float(np.sum((-0.5 * beta * delta ** 2 + 0.5 * np.log(beta)) * om) / (n * k) - 0.5 * LN2PI)  # :237
    stats["validRMSE"] = float("nan")
    stats["validLL"] = float("nan")
    if validation is not None and np.size(validation) > 0:              # :239
        validation = np.asarray(validation, dtype=bool).reshape(-1)
        nv = int(validation.sum())
        PHIv, _, lnBv, _ = getPHI(X, Psi, theta, model, validation)     # :243
        betav = np.exp(-lnBv)
        # GPz.m:250-252 also forms nu on the validation rows and never uses it; skipped here.
        deltav = PHIv @ w - Y[validation, :]                            # :254-255
        omv = omega[validation, :]
        stats["validRMSE"] = math.sqrt(np.sum(deltav ** 2 * omv) / (nv * k))        # :258
        stats["validLL"] = float(np.sum((-0.5 * betav * deltav ** 2 + 0.5 * np.log(betav)) * omv) / (nv * k)
                                 - 0.5 * LN2PI)                         # :259
    return GPzResult(float(nlogML), grad, w, iSigma_w, PHI, stats)


# --------------------------------------------------------------------------------------
# predict.m / predictDiag.m / predictCov.m   (Full and Noisy sub-paths)
# --------------------------------------------------------------------------------------
def _predictFull(X, theta, w, iSigma_w, model):
    """predictDiag.m:58-74 == predictCov.m:53-69."""
    n = X.shape[0]
    k = w.shape[1]
    PHI, _, ElnS, _ = getPHI(X, None, theta, model, None, want_N=False)
    mu = PHI @ w
    nu = np.zeros((n, k))
    for out in range(k):
        nu[:, out] = np.sum(PHI * (PHI @ iSigma_w[:, :, out]), axis=1)
    return mu, nu, np.exp(ElnS), np.zeros((n, k)), PHI


def _predictNoisyDiag(X, Psi, Gamma, w, v, b, P, iSigma_w, theta, model):
    """predictDiag.m:75-125 (basis-pair loop; the diagonal pair is subtracted once, :117-119)."""
    n, d = X.shape
    m, k = w.shape
    PHI, _, ElnS, _ = getPHI(X, Psi, theta, model, None, want_N=False)
    mu = PHI @ w
    nu = np.zeros((n, k))
    gamma = np.zeros((n, k))
    VlnS = np.zeros((n, k))
    iSigma = Gamma ** 2
    Sigma = Gamma ** -2.0
    lnz = -0.5 * np.sum(np.log(iSigma), axis=1)
    for i in range(m):
        Z = None
        for j in range(i + 1):
            iCij = iSigma[i, :] + iSigma[j, :]
            Cij = 1.0 / iCij
            cij = (P[i, :] * iSigma[i, :] + P[j, :] * iSigma[j, :]) / iCij
            lnZij = lnz[i] + lnz[j] - 0.5 * np.sum((P[i, :] - P[j, :]) ** 2 / (Sigma[i, :] + Sigma[j, :])) \
                - 0.5 * np.sum(np.log(Sigma[i, :] + Sigma[j, :]))
            Delta = X - cij[None, :]
            CpP = Psi + Cij[None, :]
            lnNxc = -0.5 * np.sum(Delta ** 2 / CpP, axis=1) - 0.5 * np.sum(np.log(CpP), axis=1)
            Z = np.exp(lnZij + lnNxc)[:, None]
            gamma += 2.0 * Z * (w[i, :] * w[j, :])[None, :]
            VlnS += 2.0 * Z * (v[i, :] * v[j, :])[None, :]
            nu += 2.0 * Z * iSigma_w[i, j, :][None, :]
        gamma -= Z * (w[i, :] * w[i, :])[None, :]            # j == i after the inner loop
        VlnS -= Z * (v[i, :] * v[i, :])[None, :]
        nu -= Z * iSigma_w[i, i, :][None, :]
    VlnS = VlnS - (ElnS - b.reshape(1, k)) ** 2
    gamma = gamma - mu ** 2
    beta_i = np.exp(ElnS) * (1.0 + 0.5 * VlnS)
    return mu, nu, beta_i, gamma, PHI


def _predictNoisyCov(X, Psi, Gamma, w, v, b, P, iSigma_w, theta, model):
    """predictCov.m:70-133 (per-sample, per-pair d x d inverses)."""
    n, d = X.shape
    m, k = w.shape
    PHI, _, ElnS, _ = getPHI(X, Psi, theta, model, None, want_N=False)
    mu = PHI @ w
    nu = np.zeros((n, k))
    gamma = np.zeros((n, k))
    VlnS = np.zeros((n, k))
    iSigma = np.zeros((d, d, m))
    Sigma = np.zeros((d, d, m))
    lnz = np.zeros(m)
    for i in range(m):
        iSigma[:, :, i] = Gamma[:, :, i].T @ Gamma[:, :, i]
        Sigma[:, :, i] = np.linalg.inv(iSigma[:, :, i])
        lnz[i] = -0.5 * np.sum(np.log(np.linalg.svd(iSigma[:, :, i], compute_uv=False)))
    for i in range(m):
        for j in range(i + 1):
            iCij = iSigma[:, :, i] + iSigma[:, :, j]
            Cij = np.linalg.inv(iCij)
            cij = np.linalg.solve(iCij.T, (P[i, :] @ iSigma[:, :, i] + P[j, :] @ iSigma[:, :, j]))
            Dl = P[i, :] - P[j, :]
            Sij = Sigma[:, :, i] + Sigma[:, :, j]
            lnZij = lnz[i] + lnz[j] - 0.5 * Dl @ np.linalg.solve(Sij.T, Dl) \
                - 0.5 * np.sum(np.log(np.linalg.svd(Sij, compute_uv=False)))
            for t in range(n):
                Dt = X[t, :] - cij
                CpP = Psi[:, :, t] + Cij
                lnNxc = -0.5 * Dt @ np.linalg.solve(CpP.T, Dt) - 0.5 * np.sum(np.log(np.linalg.svd(CpP, compute_uv=False)))
                Z = math.exp(lnZij + lnNxc)
                f = 2.0 if j < i else 1.0                  # 2*Z for every pair, minus Z once for j==i
                gamma[t, :] += f * Z * (w[i, :] * w[j, :])
                VlnS[t, :] += f * Z * (v[i, :] * v[j, :])
                nu[t, :] += f * Z * iSigma_w[i, j, :]
    VlnS = VlnS - (ElnS - b.reshape(1, k)) ** 2
    gamma = gamma - mu ** 2
    beta_i = np.exp(ElnS) * (1.0 + 0.5 * VlnS)
    return mu, nu, beta_i, gamma, PHI


def _predictMissingDiag(X, Psi, Gamma, w, v, b, P, iSigma_w, priors):
    """predictMissing (Psi is None, predictDiag.m:127-212) and predictNoisyMissing (predictDiag.m:213-295) for one
    group of rows sharing a NaN pattern.  The two reference functions differ only in Psi being added to the
    variances of the observed dims (:230-232, :262-264)."""
    o = ~np.isnan(X[0, :])
    u = ~o
    n = X.shape[0]
    m, k = w.shape
    iSigma = Gamma ** 2
    Sigma = Gamma ** -2.0
    lnz = -0.5 * np.sum(np.log(iSigma), axis=1)                         # :141 (== +0.5*sum(log(Sigma)), :224)
    Ps = np.zeros((n, int(o.sum()))) if Psi is None else Psi[:, o]
    No = np.zeros((n, m))
    Ex = np.zeros((n, m))
    for i in range(m):                                                  # :145-151 / :228-235
        Delta = X[:, o] - P[i, o][None, :]
        SpP = Ps + Sigma[i, o][None, :]
        No[:, i] = np.exp(-0.5 * np.sum(Delta ** 2 / SpP, axis=1) - 0.5 * np.sum(np.log(SpP), axis=1))
        Ex[:, i] = No[:, i] * priors.reshape(-1)[i]
    Pio = Ex / np.sum(Ex, axis=1, keepdims=True)                        # :153-155
    ii = np.arange(m * m) % m                                           # :157-158
    jj = np.arange(m * m) // m
    Nij = np.exp(-0.5 * np.sum((P[ii][:, u] - P[jj][:, u]) ** 2 / (Sigma[ii][:, u] + Sigma[jj][:, u]), axis=1)
                 - 0.5 * np.sum(np.log(Sigma[ii][:, u] + Sigma[jj][:, u]), axis=1))          # :161
    Nmat = Nij.reshape((m, m), order="F")                               # Nmat[i, j]
    PHI = No * (Pio @ Nmat.T)                                           # :163  sum_j No(:,i) Pio(:,j) Nij(i,j)
    PHI = PHI * np.exp(lnz)[None, :]                                    # :164
    mu = PHI @ w
    ElnS = PHI @ v
    gamma = np.zeros((n, k))
    nu = np.zeros((n, k))
    VlnS = np.zeros((n, k))
    for i in range(m):                                                  # :173
        Z = None
        for j in range(i + 1):
            Cij = 1.0 / (iSigma[i, :] + iSigma[j, :])
            cij = (P[i, :] * iSigma[i, :] + P[j, :] * iSigma[j, :]) * Cij
            Delta = X[:, o] - cij[o][None, :]
            CpP = Ps + Cij[o][None, :]
            No_p = np.exp(-0.5 * np.sum(Delta ** 2 / CpP, axis=1) - 0.5 * np.sum(np.log(CpP), axis=1))   # :180 / :263
            Dl = P[:, u] - cij[u][None, :]
            CpS = Sigma[:, u] + Cij[u][None, :]
            Nu = np.exp(-0.5 * np.sum(Dl ** 2 / CpS, axis=1) - 0.5 * np.sum(np.log(CpS), axis=1))          # :184
            EcCij = np.sum(np.outer(No_p, Nu) * Pio, axis=1)            # :186-187
            Dp = P[i, :] - P[j, :]
            Z = (np.exp(lnz[i] + lnz[j] - 0.5 * np.sum(Dp ** 2 / (Sigma[i, :] + Sigma[j, :]))
                        - 0.5 * np.sum(np.log(Sigma[i, :] + Sigma[j, :]))) * EcCij)[:, None]               # :190
            gamma += 2.0 * Z * (w[i, :] * w[j, :])[None, :]
            VlnS += 2.0 * Z * (v[i, :] * v[j, :])[None, :]
            nu += 2.0 * Z * iSigma_w[i, j, :][None, :]
        gamma -= Z * (w[i, :] * w[i, :])[None, :]                       # :197-199
        VlnS -= Z * (v[i, :] * v[i, :])[None, :]
        nu -= Z * iSigma_w[i, i, :][None, :]
    VlnS = VlnS - ElnS ** 2                                             # :204
    ElnS = ElnS + b.reshape(1, k)                                       # :206
    beta_i = np.exp(ElnS) * (1.0 + 0.5 * VlnS)                          # :208
    gamma = gamma - mu ** 2                                             # :210
    return mu, nu, beta_i, gamma, PHI


def _lnN(delta, S):
    """-1/2 delta S^-1 delta' - 1/2 ln det S (the reference writes ln det as sum(log(svd(S))))."""
    return -0.5 * delta @ np.linalg.solve(S.T, delta) - 0.5 * np.sum(np.log(np.linalg.svd(S, compute_uv=False)))


def _predictMissingCov(X, Psi, Gamma, w, v, b, P, iSigma_w, priors):
    """predictMissing (Psi is None, predictCov.m:134-232) and predictNoisyMissing (predictCov.m:233-336) for one group of rows
    sharing a NaN pattern.  The two reference functions differ in Psi(o,o) being added to Sigma(o,o) in the responsibilities
    (:267) and propagated into Psi_hat = T Psi_oo T' + Schur complement (:270-275)."""
    o = ~np.isnan(X[0, :])
    u = ~o
    io, iu = np.nonzero(o)[0], np.nonzero(u)[0]
    n, d = X.shape
    m, k = w.shape
    pri = np.asarray(priors, dtype=np.float64).reshape(-1)
    iSigma = np.zeros((m, d, d))
    Sigma = np.zeros((m, d, d))
    lnz = np.zeros(m)
    R = []
    Schur = np.zeros((m, d, d))
    for i in range(m):                                                  # :156-172 / :257-279
        iSigma[i] = Gamma[:, :, i].T @ Gamma[:, :, i]
        Sigma[i] = np.linalg.inv(iSigma[i])
        lnz[i] = -0.5 * np.sum(np.log(np.linalg.svd(iSigma[i], compute_uv=False)))
        Soo = Sigma[i][np.ix_(io, io)]
        R.append(np.linalg.solve(Soo, Sigma[i][np.ix_(io, iu)]))
        Schur[i][np.ix_(iu, iu)] = Sigma[i][np.ix_(iu, iu)] - Sigma[i][np.ix_(iu, io)] @ R[i]
    Ex = np.zeros((n, m))
    X_hat = np.zeros((n, m, d))
    Psi_hat = np.zeros((n, m, d, d))
    for t in range(n):
        for i in range(m):
            Soo = Sigma[i][np.ix_(io, io)]
            Delta = X[t, io] - P[i, io]
            PS = Soo if Psi is None else Soo + Psi[np.ix_(io, io, [t])][:, :, 0]
            Ex[t, i] = math.exp(_lnN(Delta, PS)) * pri[i]               # :163 / :267
            X_hat[t, i, iu] = Delta @ R[i] + P[i, iu]                   # :169-170 / :277-278
            X_hat[t, i, io] = X[t, io]
            Psi_hat[t, i] = Schur[i]
            if Psi is not None:                                         # :270-275
                T = np.vstack([np.eye(len(io)), R[i].T])
                full = T @ Psi[np.ix_(io, io, [t])][:, :, 0] @ T.T
                idx = np.concatenate([io, iu])
                Psi_hat[t, i][np.ix_(idx, idx)] += full
    Pio = Ex / np.sum(Ex, axis=1, keepdims=True)                        # :175 / :282
    PHI = np.zeros((n, m))
    gamma = np.zeros((n, k))
    nu = np.zeros((n, k))
    VlnS = np.zeros((n, k))
    for i in range(m):                                                  # :177-222 / :284-326
        for j in range(i + 1):
            iCij = iSigma[i] + iSigma[j]
            Cij = np.linalg.inv(iCij)
            cij = np.linalg.solve(iCij.T, P[i, :] @ iSigma[i] + P[j, :] @ iSigma[j])
            Dp = P[i, :] - P[j, :]
            lnZij = lnz[i] + lnz[j] + _lnN(Dp, Sigma[i] + Sigma[j])
            for t in range(n):
                # each ordered pair once: the reference adds both orders for j < i and adds-then-subtracts the j == i term
                PHI[t, i] += math.exp(_lnN(X_hat[t, j] - P[i, :], Sigma[i] + Psi_hat[t, j])) * Pio[t, j]
                if j < i:
                    PHI[t, j] += math.exp(_lnN(X_hat[t, i] - P[j, :], Sigma[j] + Psi_hat[t, i])) * Pio[t, i]
                Ec = 0.0
                for l in range(m):
                    Ec += math.exp(_lnN(X_hat[t, l] - cij, Cij + Psi_hat[t, l])) * Pio[t, l]
                Z = math.exp(lnZij) * Ec
                f = 2.0 if j < i else 1.0
                gamma[t, :] += f * Z * (w[i, :] * w[j, :])
                VlnS[t, :] += f * Z * (v[i, :] * v[j, :])
                nu[t, :] += f * Z * iSigma_w[i, j, :]
    PHI = PHI * np.exp(lnz)[None, :]                                    # :224 / :328
    mu = PHI @ w
    ElnS = PHI @ v
    VlnS = VlnS - ElnS ** 2
    ElnS = ElnS + b.reshape(1, k)
    beta_i = np.exp(ElnS) * (1.0 + 0.5 * VlnS)
    gamma = gamma - mu ** 2
    return mu, nu, beta_i, gamma, PHI


def predict(X, model: Model, which="best", Psi=None, selection=None):
    """predict.m:1-75: rows grouped by NaN pattern; Full / Noisy for complete rows (both mode families),
    Missing / NoisyMissing for the diagonal modes (predictDiag.m:127-295) and the covariance modes
    (predictCov.m:134-336)."""
    st = model.best if which == "best" else model.last
    n_all = X.shape[0]
    if selection is None:
        selection = np.ones(n_all, dtype=bool)
    selection = np.asarray(selection, dtype=bool).reshape(-1)
    n = int(selection.sum())
    k, m, d = model.k, model.m, model.d
    meth = model.method
    X = X[selection, :]
    if Psi is not None:
        Psi = np.asarray(Psi, dtype=np.float64)
        if Psi.ndim == 3:
            Psi = Psi[:, :, selection]
        else:
            Psi = Psi.reshape(n_all, -1)[selection, :]
    Xz = (X - model.muX[None, :]) / model.sdX[None, :]                  # :35-36
    theta, w, iSigma_w, P = st["theta"], st["w"], st["iSigma_w"], st["P"]
    Psi = fixPsi(Psi, n, model.sdX, meth)                               # :43
    v = st["v"] if model.heteroscedastic else np.zeros((m, k))
    Gamma = unpack_gamma(theta, model)
    off = m * d + model.g_dim + m * k
    b = theta[off:off + k]
    mu, nu, beta_i, gamma = (np.zeros((n, k)) for _ in range(4))
    PHI = np.zeros((n, m))
    for grp in nan_groups(np.isnan(Xz)):                                # predict.m:45-69
        Xg = Xz[grp]
        full = not np.isnan(Xg[0]).any()
        if full and Psi is None:
            r = _predictFull(Xg, theta, w, iSigma_w, model)
        elif full and meth[1] == "C":
            r = _predictNoisyCov(Xg, Psi[:, :, grp], Gamma, w, v, b, P, iSigma_w, theta, model)
        elif full:
            r = _predictNoisyDiag(Xg, Psi[grp], Gamma, w, v, b, P, iSigma_w, theta, model)
        elif meth[1] == "C":
            r = _predictMissingCov(Xg, None if Psi is None else Psi[:, :, grp], Gamma, w, v, b, P, iSigma_w, st["priors"])
        else:
            r = _predictMissingDiag(Xg, None if Psi is None else Psi[grp], Gamma, w, v, b, P, iSigma_w, st["priors"])
        mu[grp], nu[grp], beta_i[grp], gamma[grp], PHI[grp] = r
    sigma = nu + beta_i + gamma                                         # :72
    mu = mu + model.muY.reshape(1, k)                                   # :73
    return mu, sigma, nu, beta_i, gamma, PHI


# --------------------------------------------------------------------------------------
# theta packing as init.m:54-101 produces it (given centres P and per-basis gamma)
# --------------------------------------------------------------------------------------
def pack_theta_init(P, gamma, Yc_var, method, heteroscedastic=True):
    """theta0 = [P(:); Gamma(:); lnAlpha(:); b(:); v(:)=0; lnTau(:)=0]  (init.m:54-55, 65-101)."""
    m, d = P.shape
    Yc_var = np.asarray(Yc_var, dtype=np.float64).reshape(-1)
    k = Yc_var.size
    b = np.log(Yc_var)
    lnAlpha = np.repeat(-np.log(Yc_var).reshape(1, k), m, axis=0)
    gamma = np.asarray(gamma, dtype=np.float64).reshape(-1)
    if method == "GL":
        G = np.array([gamma.mean()])
    elif method == "VL":
        G = gamma.copy()
    elif method == "GD":
        G = np.full(d, gamma.mean())
    elif method == "VD":
        G = np.repeat(gamma.reshape(m, 1), d, axis=1)
    elif method == "GC":
        G = np.eye(d) * gamma.mean()
    elif method == "VC":
        G = np.zeros((d, d, m))
        for j in range(m):
            G[:, :, j] = np.eye(d) * gamma[j]
    else:
        raise ValueError(method)
    parts = [P.reshape(-1, order="F"), np.asarray(G).reshape(-1, order="F"), lnAlpha.reshape(-1, order="F"), b]
    if heteroscedastic:
        parts += [np.zeros(m * k), np.zeros(m * k)]
    return np.concatenate(parts)


def init_gamma(Xz, P, m):
    """gamma_j = sqrt(0.5 * m^(1/d) / mean_i Dxy(X,P)_ij)  (init.m:62)."""
    d = Xz.shape[1]
    return np.sqrt(0.5 * m ** (1.0 / d) / np.mean(Dxy(Xz, P), axis=0))
