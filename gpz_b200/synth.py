"""Synthetic (n, d, m) problems for tests and bench (SURVEY.md 8d "Synthetic inputs").

Host-side NumPy only.  theta0 follows the reference's own initialisation (init.m:54-101):
centres P ~ U(-sqrt3, sqrt3) (init.m:58-59 with an identity PCA rotation, valid for
already z-scored X), gamma_j = sqrt(0.5 * m^(1/d) / mean_i ||x_i - p_j||^2) (init.m:62),
Gamma isotropic per mode (init.m:65-84), lnAlpha = -ln var(y), b = ln var(y), v = 0, lnTau = 0.
"""
from __future__ import annotations

import numpy as np

METHODS = ("GL", "VL", "GD", "VD", "GC", "VC")


def g_dim_of(method: str, m: int, d: int) -> int:
    return {"GL": 1, "VL": m, "GD": d, "VD": m * d, "GC": d * d, "VC": d * d * m}[method]


def theta_len(method: str, m: int, d: int, k: int, het: bool) -> int:
    return m * d + g_dim_of(method, m, d) + m * k + k + (2 * m * k if het else 0)


def make_data(n: int, d: int, seed: int = 0, k: int = 1, dtype=np.float64):
    """X ~ N(0,1)^{n x d}; y = sin(x1) + 0.5 x2 x3 + heteroscedastic noise, centred."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    x1 = X[:, 0]
    x2 = X[:, 1 % d]
    x3 = X[:, 2 % d]
    f = np.sin(x1) + 0.5 * x2 * x3
    s = 0.05 + 0.2 / (1.0 + np.exp(-x1))
    Y = np.empty((n, k))
    for c in range(k):
        Y[:, c] = (1.0 + 0.5 * c) * f + s * rng.standard_normal(n)
    Y -= Y.mean(axis=0, keepdims=True)
    return np.asfortranarray(X.astype(dtype)), np.asfortranarray(Y.astype(dtype))


def make_psi(n: int, d: int, method: str, seed: int = 1):
    """Input-noise covariances in the layout fixPsi produces: n x d (L/D) or d x d x n (C)."""
    rng = np.random.default_rng(seed)
    if method[1] == "C":
        A = rng.standard_normal((n, d, d))
        Psi = np.einsum("nij,nkj->nik", A, A) * (0.1 / d) + 0.01 * np.eye(d)[None]
        return np.asfortranarray(np.transpose(Psi, (1, 2, 0)))
    return np.asfortranarray(rng.gamma(1.0, 0.25, size=(n, d)) * 0.2)


def mean_sqdist(X: np.ndarray, P: np.ndarray) -> np.ndarray:
    """mean_i ||x_i - p_j||^2 without the n x m matrix (same expansion Dxy.m:3-7 uses)."""
    mx = X.mean(axis=0)
    mxx = float(np.mean(np.sum(X * X, axis=1)))
    return mxx - 2.0 * (P @ mx) + np.sum(P * P, axis=1)


def make_theta0(X, Y, method: str, m: int, het: bool = True, seed: int = 2):
    n, d = X.shape
    k = Y.shape[1]
    rng = np.random.default_rng(seed)
    P = (rng.random((m, d)) - 0.5) * np.sqrt(12.0)
    gamma = np.sqrt(0.5 * m ** (1.0 / d) / mean_sqdist(X, P))
    var = Y.var(axis=0, ddof=1)
    b = np.log(var)
    lnAlpha = np.repeat(-np.log(var).reshape(1, k), m, axis=0)
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
        G[np.arange(d), np.arange(d), :] = gamma[None, :]
    else:
        raise ValueError(method)
    parts = [P.reshape(-1, order="F"), G.reshape(-1, order="F"), lnAlpha.reshape(-1, order="F"), b]
    if het:
        parts += [np.zeros(m * k), np.zeros(m * k)]
    return np.concatenate(parts)


def perturb_theta(theta: np.ndarray, scale: float = 0.05, seed: int = 3) -> np.ndarray:
    """A generic (non-isotropic) point near theta0, as a line search would visit."""
    rng = np.random.default_rng(seed)
    return theta + scale * rng.standard_normal(theta.shape)
