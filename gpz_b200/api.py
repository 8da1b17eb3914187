"""Host-side mirror of the reference's MATLAB API for the hot path: init / train / predict / GPz /
getPHI / inv_logdet / Dxy with the reference's names, argument meaning and error behaviour
(GPz/init.m:1, GPz/train.m:1, GPz/predict.m:1, GPz/GPz.m:1, GPz/getPHI.m:1).

In production the host stays MATLAB and calls the same C ABI through matlab/gpz_b200_mex.cpp
(INTEGRATION.md); MATLAB is not available in this environment, so this module is the executable
stand-in.  Everything numerical on the hot path runs in libgpz_b200.so on the GPU -- this file only
normalises inputs (z-scoring, fixPsi), packs theta and drives a host L-BFGS, exactly the parts the
reference keeps on the host.  There is no CPU fallback for the objective.
"""
from __future__ import annotations

import time

import numpy as np

from . import _lib as L

__all__ = ["init", "train", "predict", "GPz", "getPHI", "inv_logdet", "Dxy", "fixPsi", "Objective"]


# ------------------------------------------------------------------------------------------------
# host preprocessing the reference also does on the host
# ------------------------------------------------------------------------------------------------
def fixPsi(Psi, n, sdX, method):
    """Input-noise layout normalisation (GPz/fixPsi.m:1-55): -> n x d for ?L/?D, d x d x n for ?C."""
    if Psi is None or np.size(Psi) == 0:
        return None
    sdX = np.asarray(sdX, dtype=np.float64).reshape(-1)
    d = sdX.size
    Psi = np.asarray(Psi, dtype=np.float64)
    cube = Psi.ndim == 3 and Psi.shape == (d, d, n)
    if Psi.ndim == 1:
        Psi = Psi.reshape(n, 1)
    if method[1] == "C":
        if cube:
            return np.asfortranarray(Psi / np.outer(sdX, sdX)[:, :, None])
        out = np.zeros((d, d, n), order="F")
        diag = np.repeat(Psi, d, axis=1) if Psi.shape[1] == 1 else Psi
        out[np.arange(d), np.arange(d), :] = (diag / sdX[None, :] ** 2).T
        return out
    if cube:
        return np.asfortranarray(np.stack([Psi[a, a, :] for a in range(d)], axis=1) / sdX[None, :] ** 2)
    if Psi.shape[1] == 1:
        Psi = np.repeat(Psi, d, axis=1)
    return np.asfortranarray(Psi / sdX[None, :] ** 2)


def _pca(X):
    """NaN-aware PCA used to rotate the random centres (GPz/pca.m:1-48 with th=1)."""
    n, d = X.shape
    miss = np.isnan(X)
    X0 = np.where(miss, 0.0, X)
    counts = n - miss.sum(axis=0)
    mu = X0.sum(axis=0) / counts
    Xc = np.where(miss, 0.0, X0 - mu[None, :])
    M = miss.astype(np.float64)
    sig = n * (Xc.T @ Xc) / (n - M.T @ M)
    S, U = np.linalg.eigh(sig)
    S = np.abs(S)
    order = np.argsort(-S)
    U, S = U[:, order], S[order]
    Ssd = np.sqrt(S / (n - 1))
    Ti = np.diag(Ssd) @ U.T
    return mu, sig / n, Ti


def _fill_linear(X, mu, Sigma):
    """Conditional-mean fill of missing entries (GPz/fillLinear.m:1-30)."""
    X = X.copy()
    miss = np.isnan(X)
    pats = np.unique(miss, axis=0)
    for u in pats:
        if not u.any():
            continue
        rows = (miss == u[None, :]).all(axis=1)
        o = ~u
        Delta = X[np.ix_(rows, o)] - mu[o][None, :]
        X[np.ix_(rows, u)] = Delta @ np.linalg.solve(Sigma[np.ix_(o, o)], Sigma[np.ix_(o, u)]) + mu[u][None, :]
    return X


def _theta_offsets(model):
    m, d, k = model["m"], model["d"], model["k"]
    oG = m * d
    oA = oG + model["g_dim"]
    oB = oA + m * k
    oV = oB + k
    return oG, oA, oB, oV, oV + m * k


def _c_model(model):
    return L.make_model(model["d"], model["k"], model["m"], model["method"], model["heteroscedastic"])


class Objective:
    """f = @(params) GPz(params,model,X,Y,Psi,omega,training,validation)  (GPz/train.m:40) with the
    normalised data resident on the GPU.  Calling it returns (nlogML, grad); the four statistics the
    reference passes through globals (GPz.m:3-7) are left in ``.stats``."""

    def __init__(self, model, Xz, Yc, Psi=None, omega=None, training=None, validation=None, device=0):
        self.ctx = L.Context(_c_model(model), Xz, Yc, Psi, omega, training, validation, device=device)
        self.stats = {}
        self.evals = 0

    def __call__(self, theta):
        f, g, self.stats = self.ctx.eval(theta)
        self.evals += 1
        return f, g

    def fit(self, theta):
        _, w, iS = self.ctx.fit(theta, want_nlogML=False)
        return w, iS

    def close(self):
        self.ctx.close()


# ------------------------------------------------------------------------------------------------
# reference-signature functions
# ------------------------------------------------------------------------------------------------
def GPz(theta, model, X, Y, Psi=None, omega=None, training=None, validation=None, nargout=2, device=0):
    """[nlogML,grad,w,iSigma_w] = GPz(theta,model,X,Y,Psi,omega,training,validation) (GPz/GPz.m:1).
    nargout<=2 -> (nlogML, grad, stats); nargout>2 -> the fit exit (nlogML 1 x k un-normalised, 0, w, iSigma_w)
    of GPz.m:84-87.  One-shot convenience: uploads the data each call; use Objective for loops."""
    ctx = L.Context(_c_model(model), X, Y, Psi, omega, training, validation, device=device)
    try:
        if nargout > 2:
            nl, w, iS = ctx.fit(theta)
            return nl, 0.0, w, iS
        return ctx.eval(theta)
    finally:
        ctx.close()


def getPHI(X, Psi, theta, model, selection=None, device=0):
    """[PHI,~,lnBeta_i] = getPHI(X,Psi,theta,model,selection) (GPz/getPHI.m:1)."""
    ctx = L.Context(_c_model(model), X, np.zeros((X.shape[0], model["k"])), Psi, None, selection, None, device=device)
    try:
        return ctx.phi(theta, 0)
    finally:
        ctx.close()


def getPrior(X, Psi, theta, model, selection=None, device=0):
    """prior = getPrior(X,Sx,theta,model,set) (GPz/getPrior.m:1): EM for the mixture weights of the bases, on the device."""
    ctx = L.Context(_c_model(model), X, np.zeros((X.shape[0], model["k"])), Psi, None, selection, None, device=device)
    try:
        return ctx.get_prior(theta)
    finally:
        ctx.close()


def inv_logdet(X, device=0):
    """[Xi,logdet] = inv_logdet(X) (GPz/inv_logdet.m:1) for symmetric positive definite X."""
    return L.inv_logdet(X, device)


def Dxy(X, Y, device=0):
    """D = Dxy(X,Y) (GPz/Dxy.m:1)."""
    return L.dxy(X, Y, device)


def init(X, Y, method, m, heteroscedastic=True, normalize=True, omega=None, training=None, Psi=None, seed=None, device=0):
    """model = init(X,Y,method,m,...) (GPz/init.m:1-124): z-scoring statistics, PCA-rotated random
    centres, length-scale heuristic through Dxy, theta packing and the first fit of w / iSigma_w."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64).reshape(X.shape[0], -1)
    n, d = X.shape
    k = Y.shape[1]
    if d == 1:
        method = method[0] + "L"                                   # init.m:12-14
    training = np.ones(n, dtype=bool) if training is None else np.asarray(training, dtype=bool).reshape(-1)
    omega = np.ones(n) if omega is None else np.asarray(omega, dtype=np.float64).reshape(-1)
    model = dict(d=d, k=k, m=int(m), method=method, heteroscedastic=bool(heteroscedastic))
    if normalize:                                                  # init.m:22-36
        miss = np.isnan(X)
        X0 = np.where(miss, 0.0, X)
        cnt = (~miss).sum(axis=0)
        muX = X0.sum(axis=0) / cnt
        sdX = np.sqrt((X0 ** 2).sum(axis=0) / cnt - muX ** 2)
    else:
        muX, sdX = np.zeros(d), np.ones(d)
    muY = Y[training].mean(axis=0)
    model.update(muX=muX, sdX=sdX, muY=muY)
    Yc = Y - muY[None, :]
    Xz = (X - muX[None, :]) / sdX[None, :]
    Psi = fixPsi(Psi, n, sdX, method)
    var = Yc[training].var(axis=0, ddof=1)
    b = np.log(var)                                                # init.m:54
    lnAlpha = np.repeat(-np.log(var)[None, :], m, axis=0)          # init.m:55
    mu, sigmas, Vi = _pca(Xz[training])
    rng = np.random.default_rng(seed)
    P = (rng.random((m, d)) - 0.5) * np.sqrt(12.0)                 # init.m:58
    P = P @ Vi + mu[None, :]                                       # init.m:59
    Xl = _fill_linear(Xz[training], mu, sigmas)                    # init.m:61
    meanD = L.dxy_colmean(Xl, P, device)                           # init.m:62: mean(Dxy(Xl,P)), reduced on the GPU
    gamma = np.sqrt(0.5 * m ** (1.0 / d) / meanD)
    if method == "GL":
        G = np.array([gamma.mean()])
    elif method == "VL":
        G = gamma.copy()
    elif method == "GD":
        G = np.full(d, gamma.mean())
    elif method == "VD":
        G = np.repeat(gamma[:, None], d, axis=1)
    elif method == "GC":
        G = np.eye(d) * gamma.mean()
    elif method == "VC":
        G = np.zeros((d, d, m))
        G[np.arange(d), np.arange(d), :] = gamma[None, :]
    else:
        raise ValueError(f"unknown method {method!r}")
    model["g_dim"] = int(G.size)                                   # init.m:86
    parts = [P.reshape(-1, order="F"), G.reshape(-1, order="F"), lnAlpha.reshape(-1, order="F"), b]
    last = {}
    if heteroscedastic:                                            # init.m:92-101
        parts += [np.zeros(m * k), np.zeros(m * k)]
        last["v"] = np.zeros((m, k))
    theta = np.concatenate(parts)
    obj = Objective(model, Xz, Yc, Psi, omega, training, None, device=device)
    try:
        w, iS = obj.fit(theta)                                     # init.m:104
    finally:
        obj.close()
    last.update(theta=theta, w=w, iSigma_w=iS, priors=np.ones(m) / m, P=P)
    best = dict(last)
    best["LL"] = -np.inf
    model["last"], model["best"] = last, best
    return model


def train(model, X, Y, maxIter=200, maxAttempts=np.inf, omega=None, training=None, validation=None, Psi=None,
          display=True, device=0, priors_wanted=True):
    """model = train(model,X,Y,...) (GPz/train.m:1-81).  The optimiser is minFunc's L-BFGS as train.m:42-48 configures
    it, run device-resident by gpz_train (gpz_b200/csrc/train.cu) together with the best-theta / early-stopping rule of
    GPz/callBack.m; this function only prints callBack.m's table and re-fits w / iSigma_w / priors for the last and
    the best theta (train.m:53-79)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64).reshape(X.shape[0], -1)
    n, d = X.shape
    m, k = model["m"], model["k"]
    Yc = Y - model["muY"][None, :]
    Xz = (X - model["muX"][None, :]) / model["sdX"][None, :]
    Psi = fixPsi(Psi, n, model["sdX"], model["method"])
    obj = Objective(model, Xz, Yc, Psi, omega, training, validation, device=device)
    training_only = validation is None
    clock = dict(t=time.time())

    def callback(it):                                               # the table of callBack.m:14-34
        if display:
            if it["iter"] == 1:
                print("\tIter\tlogML/n\t\tTrain RMSE\tTrain MLL\t" + ("" if training_only else "Valid RMSE\tValid MLL\t") + "Time")
            row = f"\t{it['iter']}\t{-it['f']:1.5e}\t{it['trainRMSE']:1.5e}\t{it['trainLL']:1.5e}"
            if not training_only:
                mark = ("[", "]") if it["improved"] else (" ", "")
                row += f"\t{it['validRMSE']:1.5e}\t{mark[0]}{it['validLL']:1.5e}{mark[1]}"
            print(row + f"\t{time.time() - clock['t']:f}")
        clock["t"] = time.time()
        return False

    try:
        theta, best_theta, best_valid, info = obj.ctx.train(
            model["last"]["theta"], model["best"]["theta"], model["best"]["LL"], callback=callback, max_iter=int(maxIter),
            max_attempts=float(maxAttempts), training_only=1 if training_only else 0)
        if display:
            print(info["message"] if info["reason"] != 7 else "No improvment after maximum number of attempts")
        oG, oA, oB, oV, oT = _theta_offsets(model)
        for name, th in (("last", theta), ("best", best_theta)):
            w, iS = obj.fit(th)                                     # train.m:53,69
            # the optimised parameters are stored BEFORE the priors are computed: a failure of the EM below must not cost
            # the training result
            model[name].update(theta=th.copy(), w=w, iSigma_w=iS, P=th[:m * d].reshape((m, d), order="F"))
            if model["heteroscedastic"]:
                model[name]["v"] = th[oV:oV + m * k].reshape((m, k), order="F")
            model[name]["priors"] = np.ones(m) / m
            if priors_wanted:
                try:
                    model[name]["priors"] = obj.ctx.get_prior(th)   # train.m:59,74 (getPrior.m)
                except L.GpzError as e:
                    import warnings
                    warnings.warn(f"getPrior failed ({e}); model.{name}.priors left uniform")
        info["best_valid"] = best_valid                            # train.m never writes best.LL back (it stays init's -inf)
        model["evals"] = obj.evals + info["fun_evals"]
        model["train_info"] = info
    finally:
        obj.close()
    return model


def predict(X, model, whichSet="best", Psi=None, selection=None, device=0):
    """[mu,sigma,nu,beta_i,gamma,PHI,w,iSigma_w] = predict(X,model,...) (GPz/predict.m:1-75).  Rows with missing
    values go through predictMissing / predictNoisyMissing (both mode families; they need model.best['priors'])."""
    st = model["best"] if whichSet == "best" else model["last"]
    X = np.asarray(X, dtype=np.float64)
    n_all = X.shape[0]
    sel = np.ones(n_all, dtype=bool) if selection is None else np.asarray(selection, dtype=bool).reshape(-1)
    X = X[sel]
    n = X.shape[0]
    if Psi is not None:
        Psi = np.asarray(Psi, dtype=np.float64)
        Psi = Psi[:, :, sel] if Psi.ndim == 3 else Psi.reshape(n_all, -1)[sel]
    Xz = (X - model["muX"][None, :]) / model["sdX"][None, :]
    Psi = fixPsi(Psi, n, model["sdX"], model["method"])
    mu, nu, beta_i, gamma, PHI = L.predict_core(_c_model(model), st["theta"], st["w"], st["iSigma_w"], Xz, Psi,
                                                want_phi=True, device=device, priors=st.get("priors"))
    sigma = nu + beta_i + gamma                                     # predict.m:72
    mu = mu + model["muY"][None, :]                                 # predict.m:73
    return mu, sigma, nu, beta_i, gamma, PHI, st["w"], st["iSigma_w"]
