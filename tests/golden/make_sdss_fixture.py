"""Generates tests/golden/sdss_cfg2.npz: BASELINE.json config 2 on the reference's OWN data file.

Input  : /root/reference/data/sdss_sample.csv (300 000 x 11: 5 magnitudes, 5 magnitude errors, z_spec;
         md5 9d5bc62281303c83319fcd8ad9783f8a, checked below).  It is the one reference-held artefact on this path.
Recipe : demo_photoz.m:42-62 -- Y = last column, X = the magnitudes, Psi = (magnitude errors).^2 (inputNoise = true),
         then init.m:22-52 -- z-score X with the column statistics, centre Y on the training rows, fixPsi -> 5 x 5 x n --
         and init.m:54-101 for theta0 (PCA-rotated uniform centres, gamma through Dxy), perturbed like a line-search point.
         d = 5, m = 100, VC, heteroscedastic, as BASELINE.json states config 2.
Rows   : the first NROWS rows of the file, even rows training / odd rows validation (the demo draws its split with
         MATLAB's rng; any fixed split exercises the same path).
Output : the raw rows (so the test rebuilds X, Psi, Y itself), the masks, theta, and the ORACLE's nlogML / gradient /
         statistics / fit (w, iSigma_w) on them.  The oracle needs ~10 min for these 2 x 8 000 rows (scalar loop over
         n x m of getPHI.m:80-88 / GPz.m:166-184), which is why its outputs are stored instead of recomputed on the GPU box.
Run from the repo root:  python tests/golden/make_sdss_fixture.py
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gpz_oracle as O  # noqa: E402

SRC = "/root/reference/data/sdss_sample.csv"
MD5 = "9d5bc62281303c83319fcd8ad9783f8a"
NROWS = 16000
M = 100
METHOD = "VC"


def prepare(rows, training, m=M, method=METHOD, seed=7):
    """demo_photoz.m:42-62 + init.m:22-101 on the host (shared with tests/test_baseline_shapes.py)."""
    Y = rows[:, -1:].copy()
    X = rows[:, :5].copy()
    Psi = rows[:, 5:10] ** 2
    n, d = X.shape
    muX, sdX = X.mean(axis=0), X.std(axis=0)                   # init.m:26-31 (population std)
    Xz = (X - muX) / sdX
    muY = Y[training].mean(axis=0)
    Yc = Y - muY
    PsiC = O.fixPsi(Psi, n, sdX, method)                        # init.m:52
    Xt = Xz[training]
    mu = Xt.mean(axis=0)                                        # pca.m with no missing values
    C = np.cov(Xt.T, bias=True)
    S, U = np.linalg.eigh(C)
    order = np.argsort(-S)
    S, U = S[order], U[:, order]
    Vi = np.diag(np.sqrt(S * Xt.shape[0] / (Xt.shape[0] - 1))) @ U.T
    rng = np.random.default_rng(seed)
    P = (rng.random((m, d)) - 0.5) * np.sqrt(12.0)              # init.m:58
    P = P @ Vi + mu[None, :]                                    # init.m:59
    gamma = O.init_gamma(Xt, P, m)                              # init.m:62
    theta0 = O.pack_theta_init(P, gamma, Yc[training].var(axis=0, ddof=1), method, True)
    theta = theta0 + 0.05 * np.random.default_rng(seed + 1).standard_normal(theta0.shape)
    return Xz, Yc, PsiC, theta


def main():
    with open(SRC, "rb") as fh:
        raw = fh.read()
    assert hashlib.md5(raw).hexdigest() == MD5, "sdss_sample.csv is not the file SURVEY.md recorded"
    rows = np.loadtxt(SRC, delimiter=",", max_rows=NROWS)
    assert rows.shape == (NROWS, 11)
    training = np.arange(NROWS) % 2 == 0
    validation = ~training
    Xz, Yc, PsiC, theta = prepare(rows, training)
    model = O.Model(d=5, k=1, m=M, method=METHOD, heteroscedastic=True)
    t0 = time.time()
    r = O.GPz(theta, model, Xz, Yc, PsiC, None, training, validation)
    fit = O.GPz(theta, model, Xz, Yc, PsiC, None, training, None, fit_only=True)
    print(f"oracle: {time.time() - t0:.0f} s, nlogML {r.nlogML:.15g}")
    out = os.path.join(ROOT, "tests", "golden", "sdss_cfg2.npz")
    np.savez_compressed(out, rows=rows, training=training, validation=validation, theta=theta, nlogML=r.nlogML, grad=r.grad,
                        stats=np.array([r.stats[s] for s in ("trainRMSE", "trainLL", "validRMSE", "validLL")]),
                        w=fit.w, iSigma_w=fit.iSigma_w, fit_nlogML=fit.nlogML, source_md5=np.array(MD5),
                        cond=np.linalg.cond(fit.iSigma_w[:, :, 0]))
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
