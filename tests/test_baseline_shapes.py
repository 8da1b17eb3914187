"""GPU parity at the TRUE BASELINE.json shapes (d, m, covariance mode), against the CPU oracle.

SURVEY.md 8d "Parity run": configs 3-5 at a sub-sampled n (<= 2e4 rows) with the real d and m, config 2 on rows of the
reference's own data file (data/sdss_sample.csv, see tests/golden/make_sdss_fixture.py), each on both GEMM engines
(int8 tcgen05 digit GEMMs / fp64 DMMA).  The benchmarked configuration is therefore checked against the oracle, not
only against the library's other engine.

Tolerance.  Stated (BASELINE.json): 1e-5 relative.  Asserted here: err <= C_COND * cond(SIGMA) * eps, with cond(SIGMA)
MEASURED in the test from the oracle's own iSigma_w -- every quantity downstream of the m x m inverse (w, nu, the
gradient) carries cond * eps from the inversion alone, in the oracle (SVD pseudo-inverse) as much as here (Cholesky).
The measured error / (cond * eps) ratio is printed so the margin is visible in the log.
"""
import functools
import os
import time

import numpy as np
import pytest

from gpz_b200 import _lib as L
from gpz_b200 import synth
from oracle import gpz_oracle as O

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps
C_COND = 64.0          # err <= C_COND * cond * eps  (and never above the stated 1e-5)
STATED = 1e-5
HERE = os.path.dirname(os.path.abspath(__file__))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def grad_blocks(m, d, k, g_dim, het, g):
    o = [0, m * d, m * d + g_dim, m * d + g_dim + m * k, m * d + g_dim + m * k + k]
    names = ["dP", "dGamma", "dlnAlpha", "db"]
    if het:
        o += [o[-1] + m * k, o[-1] + 2 * m * k]
        names += ["dv", "dlnTau"]
    return {nm: g[o[i]:o[i + 1]] for i, nm in enumerate(names)}


def check(tag, model, cond, f, g, st, rf, rg, rstats):
    """err <= C_COND * cond * eps (floored at 1e-12: plain summation-order noise), and within the stated 1e-5."""
    bound = min(STATED, max(1e-12, C_COND * cond * EPS))
    worst = abs(f - rf) / abs(rf)
    assert worst <= bound, (tag, "nlogML", worst, bound)
    gb = grad_blocks(model.m, model.d, model.k, model.g_dim, model.heteroscedastic, g)
    rb = grad_blocks(model.m, model.d, model.k, model.g_dim, model.heteroscedastic, rg)
    errs = {nm: rel(gb[nm], rb[nm]) for nm in gb}
    for nm, e in errs.items():
        assert e <= bound, (tag, nm, e, bound, cond)
    for i, key in enumerate(("trainRMSE", "trainLL", "validRMSE", "validLL")):
        if not np.isnan(rstats[i]):
            assert abs(st[key] - rstats[i]) <= bound * max(1.0, abs(rstats[i])), (tag, key)
    worst = max(worst, max(errs.values()))
    print(f"[{tag}] cond(SIGMA) = {cond:.3e}; worst relative error {worst:.2e} = {worst / (cond * EPS):.3f} x cond*eps "
          f"(bound {bound:.2e}); per block " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))


# ------------------------------------------------------------------------------------------------ configs 3, 4 / headline
SHAPES = {
    # name: (d, m, method, n_train + n_valid rows)
    "cfg3_d10_m500_VD": (10, 500, "VD", 20000),
    "headline_cfg4_d10_m1000_VC": (10, 1000, "VC", 20000),
}


@functools.lru_cache(maxsize=None)
def shape_problem(name):
    d, m, method, n = SHAPES[name]
    X, Y = synth.make_data(n, d, seed=0)                       # the bench's own generator and seed (bench.py)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, method, m, het=True, seed=1), 0.01, 5)
    tr = np.arange(n) % 5 != 0
    model = O.Model(d=d, k=1, m=m, method=method, heteroscedastic=True)
    t0 = time.time()
    ref = O.GPz(theta, model, np.array(X), np.array(Y), None, None, tr, ~tr)
    fit = O.GPz(theta, model, np.array(X), np.array(Y), None, None, tr, None, fit_only=True)
    cond = float(np.linalg.cond(fit.iSigma_w[:, :, 0]))
    print(f"[{name}] oracle {time.time() - t0:.1f} s")
    return model, theta, X, Y, tr, ref, fit, cond


@pytest.mark.parametrize("engine", ["int8", "fp64"])
@pytest.mark.parametrize("name", list(SHAPES))
def test_true_shape_eval_fit_predict_against_oracle(name, engine):
    """NLML + gradient + statistics, the fit exit (GPz.m:84-87) and predictFull at the real (d, m, mode) of BASELINE
    configs 3 and 4 (= the benchmarked headline), 16 000 training + 4 000 validation rows."""
    model, theta, X, Y, tr, ref, fit, cond = shape_problem(name)
    gm = L.make_model(model.d, 1, model.m, model.method, True)
    ctx = L.Context(gm, X, Y, None, None, tr, ~tr)
    ctx.set_option("ozaki_slices", 7 if engine == "int8" else 0)
    f, g, st = ctx.eval(theta)
    f2, g2, _ = ctx.eval(theta)
    assert f == f2 and np.array_equal(g, g2)
    assert ctx.last_timing()["int8_slices"] == (7 if engine == "int8" else 0)
    check(f"{name}/{engine}", model, cond, f, g, st, ref.nlogML, ref.grad, [ref.stats[s] for s in ("trainRMSE", "trainLL", "validRMSE", "validLL")])
    bound = min(STATED, C_COND * cond * EPS)
    nl, w, iS = ctx.fit(theta)
    assert rel(nl, fit.nlogML) <= bound and rel(w, fit.w) <= bound and rel(iS, fit.iSigma_w) <= bound, \
        (rel(nl, fit.nlogML), rel(w, fit.w), rel(iS, fit.iSigma_w), bound)
    ctx.close()
    # predict (predict.m:25-73 -> predictFull) on 3 000 held-out rows with the ORACLE's w / iSigma_w
    m, d = model.m, model.d
    model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(1)
    o2 = m * d + model.g_dim + m + 1
    model.best = dict(theta=theta, w=fit.w, iSigma_w=fit.iSigma_w, P=theta[:m * d].reshape((m, d), order="F"),
                      v=theta[o2:o2 + m].reshape(m, 1))
    Xt = np.array(X)[~tr][:3000]
    mu, sigma, nu, be, ga, _ = O.predict(Xt, model)
    mu2, nu2, be2, ga2, _ = L.predict_core(gm, theta, fit.w, fit.iSigma_w, Xt, None)
    assert rel(mu2, mu) <= bound and rel(be2, be) <= bound and rel(nu2 + be2 + ga2, sigma) <= bound, \
        (rel(mu2, mu), rel(be2, be), rel(nu2 + be2 + ga2, sigma), bound)
    print(f"[{name}/{engine}] predict: mu {rel(mu2, mu):.1e}, sigma {rel(nu2 + be2 + ga2, sigma):.1e}, nu {rel(nu2, nu):.1e}")


# ------------------------------------------------------------------------------------------------ config 5: d = 32, GC + Psi
@functools.lru_cache(maxsize=None)
def cfg5_problem():
    n, d, m = 500, 32, 256
    X, Y = synth.make_data(n, d, seed=0)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, "GC", m, het=True, seed=1), 0.01, 5)
    Psi = synth.make_psi(n, d, "GC", seed=3)
    tr = np.arange(n) % 5 != 0
    model = O.Model(d=d, k=1, m=m, method="GC", heteroscedastic=True)
    t0 = time.time()
    ref = O.GPz(theta, model, np.array(X), np.array(Y), np.array(Psi), None, tr, ~tr)
    fit = O.GPz(theta, model, np.array(X), np.array(Y), np.array(Psi), None, tr, None, fit_only=True)
    cond = float(np.linalg.cond(fit.iSigma_w[:, :, 0]))
    print(f"[cfg5] oracle {time.time() - t0:.1f} s")
    return model, theta, X, Y, Psi, tr, ref, cond


@pytest.mark.parametrize("path", ["fast_int8", "fast_fp64", "generic"])
def test_cfg5_true_dimension_gc_psi_against_oracle(path):
    """BASELINE config 5 at its real input dimension d = 32 (GC + per-sample 32 x 32 input-noise covariances), m = 256
    (two basis tiles, so the int8 engine is the default), 400 + 100 rows: the per-row-factorisation fast path
    (gcpsi.cu) on both engines and the literal per-(sample, basis) kernels (getPHI.m:80-88, GPz.m:166-184)."""
    model, theta, X, Y, Psi, tr, ref, cond = cfg5_problem()
    gm = L.make_model(model.d, 1, model.m, "GC", True)
    ctx = L.Context(gm, X, Y, Psi, None, tr, ~tr)
    ctx.set_option("gc_fast", 0 if path == "generic" else 1)
    ctx.set_option("ozaki_slices", 0 if path == "fast_fp64" else 7)
    f, g, st = ctx.eval(theta)
    f2, g2, _ = ctx.eval(theta)
    assert f == f2 and np.array_equal(g, g2)
    check(f"cfg5_d32_GC_psi/{path}", model, cond, f, g, st, ref.nlogML, ref.grad,
          [ref.stats[s] for s in ("trainRMSE", "trainLL", "validRMSE", "validLL")])
    ctx.close()


# ------------------------------------------------------------------------------------------------ config 2: the reference's data file
def sdss_inputs():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_sdss_fixture as F
    z = np.load(os.path.join(HERE, "golden", "sdss_cfg2.npz"))
    assert str(z["source_md5"]) == F.MD5
    tr, va = z["training"], z["validation"]
    Xz, Yc, PsiC, theta = F.prepare(z["rows"], tr)              # demo_photoz.m:42-62 + init.m:22-101, on the stored rows
    assert np.array_equal(theta, z["theta"])
    return z, Xz, Yc, PsiC, theta, tr, va


@pytest.mark.parametrize("engine", ["auto", "int8"])
def test_cfg2_photoz_on_reference_data_rows(engine):
    """BASELINE config 2 (demo_photoz.m: d = 5, m = 100, VC, magnitude errors as 5 x 5 x n input noise) on 8 000 training
    + 8 000 validation rows of the reference's data/sdss_sample.csv against the oracle's stored outputs."""
    z, Xz, Yc, PsiC, theta, tr, va = sdss_inputs()
    model = O.Model(d=5, k=1, m=100, method="VC", heteroscedastic=True)
    gm = L.make_model(5, 1, 100, "VC", True)
    ctx = L.Context(gm, Xz, Yc, PsiC, None, tr, va)
    if engine == "int8":
        ctx.set_option("ozaki_slices", 7)
    f, g, st = ctx.eval(theta)
    f2, g2, _ = ctx.eval(theta)
    assert f == f2 and np.array_equal(g, g2)
    cond = float(z["cond"])
    check(f"cfg2_sdss/{engine}", model, cond, f, g, st, float(z["nlogML"]), z["grad"], z["stats"])
    bound = min(STATED, C_COND * cond * EPS)
    nl, w, iS = ctx.fit(theta)
    assert rel(nl, z["fit_nlogML"]) <= bound and rel(w, z["w"]) <= bound and rel(iS, z["iSigma_w"]) <= bound
    ctx.close()


def test_cfg2_photoz_small_live_oracle():
    """The same recipe on the first 1 200 stored rows with the oracle run live (guards the stored outputs against drift
    of either side)."""
    z, *_ = sdss_inputs()
    import make_sdss_fixture as F
    rows = z["rows"][:1200]
    tr = np.arange(1200) % 2 == 0
    Xz, Yc, PsiC, theta = F.prepare(rows, tr)
    model = O.Model(d=5, k=1, m=100, method="VC", heteroscedastic=True)
    ref = O.GPz(theta, model, Xz, Yc, PsiC, None, tr, ~tr)
    cond = float(np.linalg.cond(ref.iSigma_w[:, :, 0]))
    ctx = L.Context(L.make_model(5, 1, 100, "VC", True), Xz, Yc, PsiC, None, tr, ~tr)
    f, g, st = ctx.eval(theta)
    check("cfg2_sdss_1200", model, cond, f, g, st, ref.nlogML, ref.grad, [ref.stats[s] for s in ("trainRMSE", "trainLL", "validRMSE", "validLL")])
    ctx.close()


def test_cfg2_photoz_predict_noisy_on_reference_rows():
    """demo_photoz.m:88 `predict(X,model,'Psi',Psi,'selection',testing)`: predictNoisy of the covariance modes
    (predictCov.m:62-132) at config 2's m = 100, d = 5 on 24 stored rows of the reference's data file with their magnitude
    errors, using the ORACLE's stored fit (w, iSigma_w).  24 rows: the oracle's pair loop is O(n m^2)."""
    z, Xz, Yc, PsiC, theta, tr, va = sdss_inputs()
    m, d = 100, 5
    model = O.Model(d=d, k=1, m=m, method="VC", heteroscedastic=True)
    model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(1)
    o2 = m * d + model.g_dim + m + 1
    model.best = dict(theta=theta, w=z["w"], iSigma_w=z["iSigma_w"], P=theta[:m * d].reshape((m, d), order="F"),
                      v=theta[o2:o2 + m].reshape(m, 1))
    pick = np.flatnonzero(va)[:24]
    Xt, Pt = np.ascontiguousarray(Xz[pick]), np.ascontiguousarray(PsiC[:, :, pick])
    t0 = time.time()
    mu, sigma, nu, be, ga, PHI = O.predict(Xt, model, Psi=Pt)
    gm = L.make_model(d, 1, m, "VC", True)
    mu2, nu2, be2, ga2, PHI2 = L.predict_core(gm, theta, z["w"], z["iSigma_w"], Xt, Pt, want_phi=True)
    bound = min(STATED, C_COND * float(z["cond"]) * EPS)
    errs = dict(mu=rel(mu2, mu), nu=rel(nu2, nu), beta_i=rel(be2, be), gamma=rel(ga2, ga), sigma=rel(nu2 + be2 + ga2, sigma),
                PHI=rel(PHI2, PHI))
    print(f"[cfg2_sdss/predictNoisy] oracle {time.time() - t0:.0f} s; " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) <= bound, (errs, bound)
    assert (sigma > 0).all() and (ga >= 0).all()
