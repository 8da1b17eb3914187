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
