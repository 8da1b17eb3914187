"""Two-sweep, row-sharded form of the NLML+gradient evaluation in NumPy (test infrastructure).

It restates the algorithm the CUDA path uses (partial sums over a row shard -> allreduce #1 -> m x m
solve on every rank -> partial sums -> allreduce #2 -> assembly) on top of the oracle's getPHI, for the
diagonal covariance modes without Psi / NaN.  tests/test_multi.py runs it over torch.distributed (gloo,
world_size 2) and checks it against the unsharded oracle GPz()."""
import numpy as np

from oracle import gpz_oracle as O


def sweep1(theta, model, X, Y, omega):
    PHI, Gamma, lnB, _ = O.getPHI(X, None, theta, model, None, want_N=False)
    beta = np.exp(-lnB)
    ob = beta * omega
    k, m = model.k, model.m
    S = np.stack([PHI.T @ (PHI * ob[:, [o]]) for o in range(k)])            # k x m x m
    r = np.stack([PHI.T @ (ob[:, o] * Y[:, o]) for o in range(k)])          # k x m
    sc = np.concatenate([np.sum(omega * lnB, axis=0), [omega.sum(), X.shape[0]]])
    return dict(PHI=PHI, Gamma=Gamma, lnB=lnB, beta=beta, ob=ob), np.concatenate([S.ravel(), r.ravel(), sc])


def solve(theta, model, red1):
    k, m, d = model.k, model.m, model.d
    S = red1[:k * m * m].reshape(k, m, m)
    r = red1[k * m * m:k * m * m + k * m].reshape(k, m)
    sc = red1[k * m * m + k * m:]
    oA = m * d + model.g_dim
    alpha = np.exp(theta[oA:oA + m * k].reshape((m, k), order="F"))
    iS, logdet, w, dwda = [], [], [], []
    for o in range(k):
        L = np.linalg.cholesky(S[o] + np.diag(alpha[:, o]))
        Li = np.linalg.inv(L)
        inv = Li.T @ Li
        iS.append(inv)
        logdet.append(2.0 * np.sum(np.log(np.diag(L))))
        w.append(inv @ r[o])
        dwda.append(-inv @ (alpha[:, o] * w[-1]))
    return dict(iS=iS, logdet=np.array(logdet), w=np.array(w).T, dwda=np.array(dwda).T, alpha=alpha, sc=sc)


def sweep2(theta, model, X, Y, omega, loc, sol):
    k, m, d = model.k, model.m, model.d
    PHI, Gamma, beta, ob, lnB = loc["PHI"], loc["Gamma"], loc["beta"], loc["ob"], loc["lnB"]
    oV = m * d + model.g_dim + m * k + k
    v = theta[oV:oV + m * k].reshape((m, k), order="F") if model.heteroscedastic else np.zeros((m, k))
    P = theta[:m * d].reshape((m, d), order="F")
    delta = PHI @ sol["w"] - Y
    dlnPHI = np.zeros_like(PHI)
    nu = np.zeros_like(delta)
    for o in range(k):
        T = PHI @ sol["iS"][o]
        nu[:, o] = np.sum(PHI * T, axis=1)
        dlnPHI -= ob[:, [o]] * T
    obd = ob * delta
    dbeta = -0.5 * omega * (1.0 - beta * (delta ** 2 + nu))
    dlnPHI += -obd @ sol["w"].T + dbeta @ v.T
    dPHI = dlnPHI * PHI
    dP = np.zeros((m, d))
    dG = np.zeros((m, d))
    for j in range(m):
        Delta = X - P[j][None, :]
        dP[j] = (dPHI[:, j] @ Delta) * Gamma[j] ** 2
        dG[j] = -Gamma[j] * (dPHI[:, j] @ Delta ** 2)
    dGm = {"GL": np.array([dG.sum()]), "VL": dG.sum(axis=1), "GD": dG.sum(axis=0), "VD": dG.reshape(-1, order="F")}[model.method]
    q = PHI.T @ obd
    dvraw = PHI.T @ dbeta
    sc = np.concatenate([np.sum(obd * delta, axis=0), dbeta.sum(axis=0),
                         [np.sum(omega * delta ** 2), np.sum(omega * (-0.5 * beta * delta ** 2 - 0.5 * lnB))]])
    return np.concatenate([dP.reshape(-1, order="F"), dGm, q.reshape(-1, order="F"), dvraw.reshape(-1, order="F"), sc])


def assemble(theta, model, sol, red2):
    k, m, d = model.k, model.m, model.d
    md, g = m * d, model.g_dim
    dP, dG = red2[:md], red2[md:md + g]
    q = red2[md + g:md + g + m * k].reshape((m, k), order="F")
    dvraw = red2[md + g + m * k:md + g + 2 * m * k].reshape((m, k), order="F")
    sc = red2[md + g + 2 * m * k:]
    sc1 = sol["sc"]
    n = sc1[k + 1]
    oA = md + g
    lnAlpha = theta[oA:oA + m * k].reshape((m, k), order="F")
    alpha, w, dwda = sol["alpha"], sol["w"], sol["dwda"]
    nl = -0.5 * sc[:k] - 0.5 * np.sum(alpha * w ** 2, axis=0) + 0.5 * lnAlpha.sum(axis=0) - 0.5 * sol["logdet"] - 0.5 * sc1[:k]
    dlnAlpha = np.stack([-0.5 * np.diag(sol["iS"][o]) * alpha[:, o] for o in range(k)], axis=1) \
        - q * dwda - alpha * w * dwda - 0.5 * alpha * w ** 2 + 0.5
    parts = [dP, dG, dlnAlpha.reshape(-1, order="F"), sc[k:2 * k]]
    if model.heteroscedastic:
        oV = oA + m * k + k
        v = theta[oV:oV + m * k].reshape((m, k), order="F")
        lnTau = theta[oV + m * k:oV + 2 * m * k].reshape((m, k), order="F")
        tau = np.exp(lnTau)
        nl = nl - 0.5 * np.sum(v ** 2 * tau, axis=0) + 0.5 * lnTau.sum(axis=0) - 0.5 * m * k * O.LN2PI
        parts += [(dvraw - v * tau).reshape(-1, order="F"), (-0.5 * tau * v ** 2 + 0.5).reshape(-1, order="F")]
    tot = nl.sum() - 0.5 * O.LN2PI * sc1[k]
    f = -tot / (n * k)
    grad = -np.concatenate(parts) / (n * k)
    stats = dict(trainRMSE=np.sqrt(sc[2 * k] / (n * k)), trainLL=sc[2 * k + 1] / (n * k) - 0.5 * O.LN2PI)
    return f, grad, stats


def shard_bounds(n, rank, world):
    """Contiguous row shard of rank `rank` (the partition bench.py and the multi-GPU tests use)."""
    return (n * rank) // world, (n * (rank + 1)) // world
