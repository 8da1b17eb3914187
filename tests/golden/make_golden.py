"""Generates tests/golden/*.npz: seeded inputs and the ORACLE's outputs (self-derived golden vectors --
the reference itself is MATLAB and cannot run here).  Run from the repo root:
    python tests/golden/make_golden.py
The cases mirror BASELINE.json's configs at sizes the oracle finishes in seconds."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gpz_b200 import synth  # noqa: E402
from oracle import gpz_oracle as O  # noqa: E402

CASES = {
    # name: (n, d, m, method, het, psi, nan, k)
    "cfg1_sinc_VL": (200, 1, 25, "VL", True, True, False, 1),      # demo_sinc.m: d=1 forces ?L, heteroscedastic, Psi n x 1
    "cfg2_photoz_VC_psi": (240, 5, 40, "VC", True, True, False, 1),  # demo_photoz.m shape (d=5, VC, Psi -> 5x5xn), reduced n,m
    "cfg3_VD": (600, 10, 48, "VD", True, False, False, 1),
    "cfg4_VC": (600, 10, 40, "VC", True, False, False, 1),
    "cfg5_GC_psi": (150, 6, 20, "GC", True, True, False, 1),
    "GL_nan": (300, 4, 16, "GL", True, False, True, 1),
    "GD_psi_nan_k2": (300, 3, 12, "GD", False, True, True, 2),
    "VC_nan": (400, 4, 24, "VC", True, False, True, 1),
    "VC_psi_nan": (300, 3, 12, "VC", True, True, True, 1),        # covariance mode + missing inputs + input noise
    "VC_predict_missing": (240, 3, 8, "VC", True, True, False, 1),  # + predictMissing / predictNoisyMissing (predictCov.m:134-336)
    "VD_predict_missing": (240, 3, 8, "VD", True, True, False, 1),  # + the diagonal family (predictDiag.m:127-295)
}


def build(name):
    n, d, m, method, het, psi, nan, k = CASES[name]
    seed = abs(hash(name)) % 1000 if False else sum(map(ord, name))
    X, Y = synth.make_data(n, d, seed=seed, k=k)
    X, Y = np.array(X), np.array(Y)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, method, m, het=het, seed=seed + 1), 0.1, seed + 2)
    Psi = np.array(synth.make_psi(n, d, method, seed=seed + 3)) if psi else None
    rng = np.random.default_rng(seed + 4)
    if nan:
        X[rng.random(n) < 0.2, 0] = np.nan
        X[rng.random(n) < 0.1, d - 1] = np.nan
    omega = 0.5 + rng.random((n, 1))
    tr = np.arange(n) % 4 != 0
    va = ~tr
    model = O.Model(d=d, k=k, m=m, method=method, heteroscedastic=het)
    r = O.GPz(theta, model, X, Y, Psi, omega, tr, va)
    fit = O.GPz(theta, model, X, Y, Psi, omega, tr, None, fit_only=True)
    out = dict(X=X, Y=Y, omega=omega, training=tr, validation=va, theta=theta, nlogML=r.nlogML, grad=r.grad,
               stats=np.array([r.stats[s] for s in ("trainRMSE", "trainLL", "validRMSE", "validLL")]),
               w=fit.w, iSigma_w=fit.iSigma_w, fit_nlogML=fit.nlogML,
               meta=np.array([n, d, m, k, int(het)]), method=np.array(method))
    if Psi is not None:
        out["Psi"] = Psi
    if not nan:
        model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(k)
        o2 = m * d + model.g_dim + m * k + k
        model.best = dict(theta=theta, w=fit.w, iSigma_w=fit.iSigma_w, P=theta[:m * d].reshape((m, d), order="F"),
                          v=(theta[o2:o2 + m * k].reshape((m, k), order="F") if het else np.zeros((m, k))))
        Xt = X[va][:40]
        mu, sigma, nu, be, ga, _ = O.predict(Xt, model, Psi=None)
        out.update(pred_X=Xt, pred_mu=mu, pred_nu=nu, pred_beta_i=be, pred_gamma=ga)
        if Psi is not None:
            Pt = Psi[:, :, va][:, :, :40] if method[1] == "C" else Psi[va][:40]
            mu, sigma, nu, be, ga, _ = O.predict(Xt, model, Psi=Pt)
            out.update(predn_Psi=Pt, predn_mu=mu, predn_nu=nu, predn_beta_i=be, predn_gamma=ga)
    if name.endswith("predict_missing"):
        model.best["priors"] = O.getPrior(X, None, theta, model, tr)
        Xm = X[va][:24].copy()
        Xm[4:14, 1] = np.nan
        Xm[10:20, 2] = np.nan
        Pm = Psi[:, :, va][:, :, :24] if method[1] == "C" else Psi[va][:24]
        mu, sigma, nu, be, ga, ph = O.predict(Xm, model, Psi=None)
        out.update(pm_X=Xm, pm_priors=model.best["priors"], pm_mu=mu, pm_nu=nu, pm_beta_i=be, pm_gamma=ga, pm_PHI=ph)
        mu, sigma, nu, be, ga, ph = O.predict(Xm, model, Psi=Pm)
        out.update(pmn_Psi=Pm, pmn_mu=mu, pmn_nu=nu, pmn_beta_i=be, pmn_gamma=ga, pmn_PHI=ph)
    return out


if __name__ == "__main__":
    only = sys.argv[1:]
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(here, name + ".npz"), **build(name))
        print("wrote", name)
