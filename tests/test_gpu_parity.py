"""GPU parity: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same
seeded inputs.  Stated tolerance (BASELINE.json north_star): 1e-5 relative in fp64 for nlogML, the
gradient and predict means/variances; the bar used here is 1e-9 (block-wise, max-norm relative),
so the stated tolerance has four digits of margin."""
import itertools

import numpy as np
import pytest

from gpz_b200 import _lib as L
from gpz_b200 import synth
from oracle import gpz_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def grad_blocks(model, g):
    m, d, k = model.m, model.d, model.k
    o = [0, m * d, m * d + model.g_dim, m * d + model.g_dim + m * k, m * d + model.g_dim + m * k + k]
    names = ["dP", "dGamma", "dlnAlpha", "db"]
    if model.heteroscedastic:
        o += [o[-1] + m * k, o[-1] + 2 * m * k]
        names += ["dv", "dlnTau"]
    return {nm: g[o[i]:o[i + 1]] for i, nm in enumerate(names)}


def problem(method, het, psi, nan, n=300, d=3, m=20, k=1, seed=0, valid=True):
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
    tr = np.arange(n) % 5 != 0
    va = ~tr if valid else None
    return model, theta, X, np.array(Y), Psi, omega, tr, va


def run_both(model, theta, X, Y, Psi, omega, tr, va, chunk_rows=None, digits=None):
    """digits: None = the library's automatic choice of GEMM engine (fp64 DMMA kernels for one 128-wide basis tile on few
    rows, int8 digit GEMMs otherwise); 7 = force the int8 tcgen05 path; 0 = force the fp64 DMMA path."""
    ref = O.GPz(theta, model, X, Y, Psi, omega, tr, va)
    gm = L.make_model(model.d, model.k, model.m, model.method, model.heteroscedastic)
    ctx = L.Context(gm, X, Y, Psi, omega, tr, va)
    if chunk_rows:
        ctx.set_option("chunk_rows", chunk_rows)
    if digits is not None:
        ctx.set_option("ozaki_slices", digits)
    f, g, st = ctx.eval(theta)
    f2, g2, _ = ctx.eval(theta)
    assert f == f2 and np.array_equal(g, g2), "evaluation is not bit-reproducible"
    return ref, f, g, st, ctx


def assert_eval_matches(model, ref, f, g, st, tol=TOL):
    assert abs(f - ref.nlogML) <= tol * abs(ref.nlogML), (f, ref.nlogML)
    gb, rb = grad_blocks(model, g), grad_blocks(model, ref.grad)
    for nm in gb:
        assert rel(gb[nm], rb[nm]) <= tol, (nm, rel(gb[nm], rb[nm]))
    for key in ("trainRMSE", "trainLL", "validRMSE", "validLL"):
        if np.isnan(ref.stats[key]):
            assert np.isnan(st[key])
        else:
            assert abs(st[key] - ref.stats[key]) <= tol * max(1.0, abs(ref.stats[key])), key


# all 48 combinations, including covariance modes with missing inputs AND input noise (getPHI.m:80-88 on the observed dims,
# GPz.m:166-184 with the GuuGuo correction)
COMBOS = list(itertools.product(synth.METHODS, (True, False), (False, True), (False, True)))


@pytest.mark.parametrize("engine", ["int8", "auto"])
@pytest.mark.parametrize("method,het,psi,nan", COMBOS)
def test_eval_matches_oracle(method, het, psi, nan, engine):
    """engine int8: the tcgen05 digit GEMMs forced on (what large problems use); auto: at this size the fp64 DMMA kernels."""
    model, theta, X, Y, Psi, omega, tr, va = problem(method, het, psi, nan)
    ref, f, g, st, ctx = run_both(model, theta, X, Y, Psi, omega, tr, va, digits=7 if engine == "int8" else None)
    assert ctx.last_timing()["int8_slices"] == (7 if engine == "int8" else 0)
    assert_eval_matches(model, ref, f, g, st)
    ctx.close()


@pytest.mark.parametrize("method", ["VD", "VC", "GL"])
def test_eval_multi_tile_m(method):
    """m > 128: several Gram tiles, several Cholesky panels with a ragged last one."""
    model, theta, X, Y, Psi, omega, tr, va = problem(method, True, False, False, n=2500, d=4, m=150, seed=7)
    ref, f, g, st, ctx = run_both(model, theta, X, Y, Psi, omega, tr, va)
    # m=150 bases on 2000 rows: cond(SIGMA) ~ 1e7, so the oracle (SVD pseudo-inverse), the fp64 DMMA Gram and the int8-slice
    # Gram legitimately differ at the cond*eps ~ 1e-9 level; 5e-9 here, 1e-9 on the well-conditioned cases
    assert_eval_matches(model, ref, f, g, st, tol=5e-9)
    ctx.close()


def test_eval_m_not_a_multiple_of_the_gram_tile():
    """m = 300 -> MP = 384: the second 256-row tile of the digit GEMMs hangs over the end of the operand (TMA zero fill),
    3 column tiles, ragged Cholesky panels."""
    model, theta, X, Y, Psi, omega, tr, va = problem("VC", True, False, False, n=6000, d=3, m=300, seed=9)
    ref, f, g, st, ctx = run_both(model, theta, X, Y, Psi, omega, tr, va)
    assert ctx.last_timing()["int8_slices"] == 7              # more than one basis tile: the digit GEMMs are the default
    assert_eval_matches(model, ref, f, g, st, tol=2e-8)       # m=300 bases in 3-d: cond(SIGMA) is large
    ctx.close()


@pytest.mark.parametrize("digits", [7, None])
@pytest.mark.parametrize("method,psi", [("VD", False), ("VC", False), ("VD", True)])
def test_eval_row_chunked_equals_resident(method, psi, digits):
    model, theta, X, Y, Psi, omega, tr, va = problem(method, True, psi, False, n=5000, d=3, m=20, seed=3)
    ref, f, g, st, ctx = run_both(model, theta, X, Y, Psi, omega, tr, va, chunk_rows=1024, digits=digits)
    assert_eval_matches(model, ref, f, g, st)
    ctx.close()


@pytest.mark.parametrize("method,psi", [("VC", False), ("VC", True), ("GC", False)])
def test_eval_row_chunked_with_missing_input_pattern_groups(method, psi):
    """Covariance modes with NaN inputs: rows are grouped by missing pattern; with row chunks (PHI not resident) every
    (chunk, pattern group) piece gets its own moment GEMM (round 1 refused this combination)."""
    model, theta, X, Y, Psi, omega, tr, va = problem(method, True, psi, True, n=5000, d=3, m=20, seed=4)
    ref, f, g, st, ctx = run_both(model, theta, X, Y, Psi, omega, tr, va, chunk_rows=1024)
    assert_eval_matches(model, ref, f, g, st)
    pr = ctx.get_prior(theta)                                   # getPrior.m with PHI rebuilt per row chunk
    ctx.close()
    gm = L.make_model(model.d, model.k, model.m, model.method, model.heteroscedastic)
    ctx = L.Context(gm, X, Y, Psi, omega, tr, va)
    assert rel(pr, ctx.get_prior(theta)) <= 1e-12
    ctx.close()


@pytest.mark.parametrize("digits", [7, None])
def test_eval_two_outputs_row_chunked(digits):
    """k = 2 outputs with row chunks: one Gram accumulation per output across the chunks (round 1 refused this combination)."""
    model, theta, X, Y, Psi, omega, tr, va = problem("VD", True, False, False, n=5000, d=3, m=20, k=2, seed=12)
    ref, f, g, st, ctx = run_both(model, theta, X, Y, Psi, omega, tr, va, chunk_rows=1024, digits=digits)
    assert_eval_matches(model, ref, f, g, st)
    ctx.close()


def test_eval_two_outputs():
    model, theta, X, Y, Psi, omega, tr, va = problem("VD", True, False, False, k=2, seed=11)
    ref, f, g, st, ctx = run_both(model, theta, X, Y, Psi, omega, tr, va)
    assert_eval_matches(model, ref, f, g, st)
    ctx.close()


@pytest.mark.parametrize("method", synth.METHODS)
def test_fit_and_phi(method):
    model, theta, X, Y, Psi, omega, tr, va = problem(method, True, False, False, seed=5)
    ref = O.GPz(theta, model, X, Y, Psi, omega, tr, None, fit_only=True)
    gm = L.make_model(model.d, model.k, model.m, model.method, True)
    ctx = L.Context(gm, X, Y, Psi, omega, tr, va)
    nl, w, iS = ctx.fit(theta)
    assert rel(nl, ref.nlogML) <= TOL
    assert rel(w, ref.w) <= 1e-8
    assert rel(iS, ref.iSigma_w) <= 1e-8
    PHI, lnb = ctx.phi(theta, 0)
    PHIr, _, lnbr, _ = O.getPHI(X, Psi, theta, model, tr)
    assert rel(PHI, PHIr) <= 1e-12 and rel(lnb, lnbr) <= 1e-12
    PHIv, lnbv = ctx.phi(theta, 1)
    PHIvr, _, lnbvr, _ = O.getPHI(X, Psi, theta, model, va)
    assert rel(PHIv, PHIvr) <= 1e-12 and rel(lnbv, lnbvr) <= 1e-12
    ctx.close()


def test_inv_logdet_and_dxy():
    rng = np.random.default_rng(0)
    for m in (5, 64, 100, 200):
        A = rng.standard_normal((m, 2 * m))
        S = A @ A.T + np.eye(m)
        Xi, ld = L.inv_logdet(S)
        Xr, ldr = O.inv_logdet(S)
        assert rel(Xi, Xr) <= 1e-9 and abs(ld - ldr) <= 1e-10 * abs(ldr)
    Xi, ld = L.inv_logdet(-np.eye(4))           # not SPD: the truncating SVD of inv_logdet.m:3-15 (round 1 returned NaN here)
    assert rel(Xi, -np.eye(4)) <= 1e-14 and abs(ld) <= 1e-14
    X, P = rng.standard_normal((300, 4)), rng.standard_normal((17, 4))
    assert rel(L.dxy(X, P), O.Dxy(X, P)) <= 1e-13
    # init.m:62 takes the column means; gpz_dxy_colmean reduces them on the device (ragged chunk: 9001 rows)
    X2 = rng.standard_normal((9001, 4))
    assert rel(L.dxy_colmean(X2, P), O.Dxy(X2, P).mean(axis=0)) <= 1e-13
    assert rel(L.dxy_colmean(X[:1], P), O.Dxy(X[:1], P)[0]) <= 1e-13


@pytest.mark.parametrize("method,psi", [(m, p) for m in synth.METHODS for p in (False, True)])
def test_predict_matches_oracle(method, psi):
    n, d, m = 200, 3, 12
    model, theta, X, Y, _, omega, tr, _ = problem(method, True, False, False, n=n, d=d, m=m, seed=9)
    r = O.GPz(theta, model, X, Y, None, omega, tr, None, fit_only=True)
    model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(1)
    o2 = m * d + model.g_dim + m + 1
    model.best = dict(theta=theta, w=r.w, iSigma_w=r.iSigma_w, P=theta[:m * d].reshape((m, d), order="F"),
                      v=theta[o2:o2 + m].reshape(m, 1))
    Xt = X[:57]
    Psi = synth.make_psi(57, d, method, seed=4) if psi else None
    mu, sigma, nu, be, ga, PHI = O.predict(Xt, model, Psi=Psi)
    gm = L.make_model(d, 1, m, method, True)
    mu2, nu2, be2, ga2, PHI2 = L.predict_core(gm, theta, r.w, r.iSigma_w, Xt, Psi, want_phi=True)
    assert rel(mu2, mu) <= TOL and rel(nu2, nu) <= 1e-8 and rel(be2, be) <= TOL
    assert np.max(np.abs(ga2 - ga)) <= 1e-9 * max(1.0, np.max(np.abs(mu)) ** 2)
    assert rel(PHI2, PHI) <= 1e-12


@pytest.mark.parametrize("method", ["VD", "VC", "GL"])
def test_moment_gemm_warp_variants_match_the_oracle(method):
    """The back-projection moment GEMM dPHI'F (gemm.cu: atb_dphi_kernel) has an 8-warp / 16-row-stage and a 16-warp / 32-row-stage
    form, chosen by shape; both are forced here (`moment_warps`) on a problem with a ragged last stage and compared with the oracle."""
    n, d, m = 3001, 3, 150
    model, theta, X, Y, _, omega, tr, va = problem(method, True, False, False, n=n, d=d, m=m, seed=41)
    ref = O.GPz(theta, model, X, Y, None, omega, tr, va)
    gm = L.make_model(d, 1, m, method, True)
    ctx = L.Context(gm, X, Y, None, omega, tr, va)
    try:
        for warps in (8, 16):
            ctx.set_option("moment_warps", warps)
            f, g, st = ctx.eval(theta)
            f2, g2, _ = ctx.eval(theta)
            assert f == f2 and np.array_equal(g, g2)
            assert rel(f, ref.nlogML) <= 1e-10 and rel(g, ref.grad) <= 2e-9, (warps, rel(f, ref.nlogML), rel(g, ref.grad))
    finally:
        ctx.set_option("moment_warps", 0)
        ctx.close()


@pytest.mark.parametrize("m", [200, 300, 470])
def test_solve_schedules_agree_with_the_oracle(m):
    """csrc/solve.cu has three schedules of the same blocked Cholesky + inverse (`solve_lookahead` 0: in-stream, 1: look-ahead
    factorisation on two streams, 2: + W = L^-1 and Sinv = W'W built block row by block row on two more streams).  The fit exit
    (GPz.m:84-87: w, iSigma_w from inv_logdet.m) must match the oracle under each, at block counts 4, 5 and 8 with a ragged
    last block, and the process-wide option is restored."""
    n, d = 4000, 3
    model, theta, X, Y, _, omega, tr, _ = problem("VD", True, False, False, n=n, d=d, m=m, seed=31)
    ref = O.GPz(theta, model, X, Y, None, omega, tr, None, fit_only=True)
    cond = np.linalg.cond(ref.iSigma_w[:, :, 0])
    bound = min(1e-5, 64 * cond * np.finfo(float).eps)
    gm = L.make_model(d, 1, m, "VD", True)
    ctx = L.Context(gm, X, Y, None, omega, tr, None)
    try:
        out = {}
        for mode in (0, 1, 2):
            ctx.set_option("solve_lookahead", mode)
            nl, w, iS = ctx.fit(theta)
            nl2, w2, iS2 = ctx.fit(theta)
            assert np.array_equal(iS, iS2) and np.array_equal(w, w2)            # each schedule is deterministic
            assert rel(iS, ref.iSigma_w) <= bound and rel(w, ref.w) <= bound and rel(nl, ref.nlogML) <= bound, \
                (mode, rel(iS, ref.iSigma_w), rel(w, ref.w), bound)
            assert rel(iS, iS.transpose(1, 0, 2)) <= bound
            out[mode] = iS
        assert rel(out[2], out[0]) <= bound and rel(out[1], out[0]) <= bound
    finally:
        ctx.set_option("solve_lookahead", 2)
        ctx.close()


@pytest.mark.parametrize("method", ["VD", "VC"])
def test_pairwise_predict_reads_the_lower_triangle_of_iSigma_w(method):
    """predictDiag.m:113 / :193 / :279 and predictCov.m:119 / :209 / :313 read iSigma_w(i,j,:) with j <= i only.  The stored
    inverse is symmetric up to rounding (cond * eps), and nu = sum 2 Z_ij iSigma_w(i,j) cancels by ~cond, so reading the
    other triangle shows up at 1e-8 (it did, on config 2's rows).  Here the upper triangle is replaced outright."""
    n, d, m = 200, 3, 12
    model, theta, X, Y, _, omega, tr, _ = problem(method, True, False, False, n=n, d=d, m=m, seed=9)
    r = O.GPz(theta, model, X, Y, None, omega, tr, None, fit_only=True)
    pri = O.getPrior(X, None, theta, model, tr)
    model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(1)
    o2 = m * d + model.g_dim + m + 1
    bad = r.iSigma_w.copy()
    iu = np.triu_indices(m, 1)
    bad[iu[0], iu[1], :] *= -3.0
    model.best = dict(theta=theta, w=r.w, iSigma_w=bad, P=theta[:m * d].reshape((m, d), order="F"),
                      v=theta[o2:o2 + m].reshape(m, 1), priors=pri)
    Xt = X[:40].copy()
    Xt[10:20, 1] = np.nan                                   # two groups: complete rows (predictNoisy) and one NaN pattern
    Psi = synth.make_psi(40, d, method, seed=4)
    mu, sigma, nu, be, ga, PHI = O.predict(Xt, model, Psi=Psi)
    gm = L.make_model(d, 1, m, method, True)
    _, nu_bad, _, _, _ = L.predict_core(gm, theta, r.w, bad, Xt, Psi, priors=pri)
    _, nu_ok, _, _, _ = L.predict_core(gm, theta, r.w, np.tril(bad.transpose(2, 0, 1)).transpose(1, 2, 0) +
                                       np.tril(bad.transpose(2, 0, 1), -1).transpose(2, 1, 0), Xt, Psi, priors=pri)
    assert np.array_equal(nu_bad, nu_ok)                     # the upper triangle is never read
    assert rel(nu_bad, nu) <= 64 * np.linalg.cond(r.iSigma_w[:, :, 0]) * np.finfo(float).eps, rel(nu_bad, nu)


def test_cov_mode_with_missing_inputs_and_psi_ignores_psi_of_missing_dims():
    """Only Psi(o,o) enters (getPHI.m:84): garbage in the rows / columns of the missing dims must not matter."""
    model, theta, X, Y, Psi, omega, tr, va = problem("VC", True, True, True, n=400, d=3, m=12, seed=4)
    gm = L.make_model(model.d, 1, model.m, "VC", True)
    ctx = L.Context(gm, X, Y, Psi, omega, tr, va)
    f, g, st = ctx.eval(theta)
    ctx.close()
    Psi2 = np.array(Psi, copy=True)
    for i in range(X.shape[0]):
        for a in np.nonzero(np.isnan(X[i]))[0]:
            Psi2[a, :, i] = 1e300
            Psi2[:, a, i] = np.nan
    ctx = L.Context(gm, X, Y, Psi2, omega, tr, va)
    f2, g2, st2 = ctx.eval(theta)
    ctx.close()
    assert f2 == f and np.array_equal(g2, g)


@pytest.mark.parametrize("method", ["VC", "GC"])
def test_cov_modes_with_missing_inputs_phi_and_multi_tile(method):
    """getPHI.m:76 / GPz.m:151-159: rows grouped by NaN pattern, marginal precision per pattern and basis."""
    model, theta, X, Y, Psi, omega, tr, va = problem(method, True, False, True, n=2200, d=4, m=150, seed=13)
    ref, f, g, st, ctx = run_both(model, theta, X, Y, None, omega, tr, va)
    assert_eval_matches(model, ref, f, g, st, tol=1e-8)
    for which, sel in ((0, tr), (1, va)):
        PHI, lnb = ctx.phi(theta, which)
        PHIr, _, lnbr, _ = O.getPHI(X, None, theta, model, sel)
        assert rel(PHI, PHIr) <= 1e-11 and rel(lnb, lnbr) <= 1e-11      # rows come back in selection order
    ctx.close()


def test_large_n_properties():
    """Full-size-style checks that need no oracle: shard-sum consistency (two half contexts vs one)
    is covered in test_gpu_multi.py; here: bit-reproducibility and finite outputs at n=2e5."""
    n, d, m = 200_000, 10, 256
    X, Y = synth.make_data(n, d, seed=1)
    theta = synth.make_theta0(X, Y, "VC", m, het=True, seed=2)
    gm = L.make_model(d, 1, m, "VC", True)
    ctx = L.Context(gm, X, Y)
    f1, g1, st = ctx.eval(theta)
    f2, g2, _ = ctx.eval(theta)
    assert np.isfinite(f1) and np.isfinite(g1).all() and f1 == f2 and np.array_equal(g1, g2)
    # directional finite difference of the objective agrees with the analytic gradient
    u = np.random.default_rng(0).standard_normal(theta.size)
    u /= np.linalg.norm(u)
    h = 1e-5
    fp, _, _ = ctx.eval(theta + h * u)
    fm, _, _ = ctx.eval(theta - h * u)
    assert abs((fp - fm) / (2 * h) - g1 @ u) <= 1e-5 * max(1.0, np.linalg.norm(g1))
    ctx.close()


def test_headline_workload_properties_at_full_size():
    """BASELINE.json's headline size (n=1e6, d=10, m=1000, VC), where the oracle cannot run: size-independent properties.
    (1) bit-reproducibility; (2) the analytic gradient against a central difference of the objective along a random
    direction; (3) the int8-digit tensor-core path against the pure fp64 DMMA path within the stated 1e-5 (cond(SIGMA) ~ 1e8-1e9
    here, so two fp64-accurate evaluations differ at the 1e-7 level in the gradient; the objective itself to 1e-12)."""
    n, d, m = 1_000_000, 10, 1000
    X, Y = synth.make_data(n, d, seed=0)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, "VC", m, het=True, seed=1), 0.01, 5)
    gm = L.make_model(d, 1, m, "VC", True)
    ctx = L.Context(gm, X, Y)
    f1, g1, st = ctx.eval(theta)
    f2, g2, _ = ctx.eval(theta)
    assert np.isfinite(f1) and np.isfinite(g1).all() and f1 == f2 and np.array_equal(g1, g2)
    u = np.random.default_rng(0).standard_normal(theta.size)
    u /= np.linalg.norm(u)
    h = 1e-5
    fp, _, _ = ctx.eval(theta + h * u)
    fm, _, _ = ctx.eval(theta - h * u)
    assert abs((fp - fm) / (2 * h) - g1 @ u) <= 1e-5 * max(1.0, np.linalg.norm(g1))
    ctx.close()
    ctx = L.Context(gm, X, Y)
    ctx.set_option("ozaki_slices", 0)
    f0, g0, _ = ctx.eval(theta)
    ctx.close()
    assert abs(f1 - f0) <= 1e-11 * abs(f0)
    md = m * d
    for sl in (slice(0, md), slice(md, md + d * d * m), slice(md + d * d * m, None)):
        assert np.max(np.abs(g1[sl] - g0[sl])) <= 1e-5 * np.max(np.abs(g0[sl]))


@pytest.mark.parametrize("method", ["VD", "VC", "GL", "GC"])
def test_tensor_core_phi_and_fused_backproj_agree_with_direct_kernels(method):
    """The fast paths (PHI = exp(F W) with F W on the int8 tensor cores [2] or on the DMMA pipe [1], fused dPHI
    back-projection GEMM) against the direct-difference kernels + materialised dPHI, and all against the oracle."""
    model, theta, X, Y, Psi, omega, tr, va = problem(method, True, False, False, n=3000, d=5, m=140, seed=21)
    ref = O.GPz(theta, model, X, Y, None, omega, tr, va)
    ref_fit = O.GPz(theta, model, X, Y, None, omega, tr, None, fit_only=True)
    gm = L.make_model(model.d, 1, model.m, method, True)
    out = {}
    for tp, fb, sc in ((1, 1, 1), (2, 1, 1), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 0), (2, 0, 0)):
        ctx = L.Context(gm, X, Y, None, omega, tr, va)
        ctx.set_option("tensor_phi", tp)
        ctx.set_option("fused_backproj", fb)
        ctx.set_option("spare_column", sc)
        out[(tp, fb, sc)] = ctx.eval(theta)
        # m=140 bases on 2400 rows: SIGMA is ill-conditioned here, so the oracle's SVD pseudo-inverse and the
        # Cholesky inverse differ at the 1e-9 level in dlnAlpha; the stated tolerance is 1e-5
        assert_eval_matches(model, ref, *out[(tp, fb, sc)], tol=1e-7)
        nl, w, iS = ctx.fit(theta)
        assert rel(w, ref_fit.w) <= 1e-6 and rel(nl, ref_fit.nlogML) <= 1e-9
        ctx.close()
    f0, g0, _ = out[(1, 1, 1)]
    for key, (f, g, _) in out.items():
        assert abs(f - f0) <= 1e-11 * abs(f0) and rel(g, g0) <= 1e-8, key


def test_unnormalised_inputs_with_large_offset():
    """normalize=false style data (x ~ 50 +- 1): only x - p enters, the library stores X relative to its
    column means so the monomial expansion does not cancel."""
    model, theta, X, Y, Psi, omega, tr, va = problem("VC", True, False, False, n=1500, d=4, m=30, seed=31)
    X = X + 50.0
    theta = theta.copy()
    theta[:model.m * model.d] += 50.0
    ref, f, g, st, ctx = run_both(model, theta, X, Y, None, omega, tr, va)
    assert_eval_matches(model, ref, f, g, st, tol=1e-8)
    ctx.close()


@pytest.mark.parametrize("method,psi,nan", [("VD", False, False), ("VD", True, True), ("GL", False, True), ("VC", False, False),
                                            ("VC", False, True), ("GC", True, False)])
def test_densities_and_get_prior(method, psi, nan):
    """getPHI's 4th output N (getPHI.m:114) and getPrior.m's EM for the basis priors."""
    model, theta, X, Y, Psi, omega, tr, va = problem(method, True, psi, nan, n=400, d=3, m=10, seed=17)
    gm = L.make_model(model.d, 1, model.m, method, True)
    ctx = L.Context(gm, X, Y, Psi, omega, tr, va)
    PHI, lnb, N = ctx.phi(theta, 0, want_N=True)
    _, _, _, Nr = O.getPHI(X, Psi, theta, model, tr)
    assert rel(N, Nr) <= 1e-11
    prior = ctx.get_prior(theta)
    pr = O.getPrior(X, Psi, theta, model, tr)
    assert abs(prior.sum() - 1.0) <= 1e-12 and rel(prior, pr.reshape(-1)) <= 1e-8
    ctx.close()


@pytest.mark.parametrize("method,psi", [("VD", False), ("VD", True), ("GL", False), ("VL", True), ("GD", False),
                                        ("VC", False), ("VC", True), ("GC", False), ("GC", True)])
def test_predict_with_missing_inputs(method, psi):
    """predict.m:45-69 grouping; predictMissing / predictNoisyMissing for the diagonal (predictDiag.m:127-295) and the
    covariance modes (predictCov.m:134-336) with priors from getPrior."""
    n, d, m = 300, 3, 9
    model, theta, X, Y, _, omega, tr, _ = problem(method, True, False, False, n=n, d=d, m=m, seed=23)
    r = O.GPz(theta, model, X, Y, None, omega, tr, None, fit_only=True)
    pri = O.getPrior(X, None, theta, model, tr)
    model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(1)
    o2 = m * d + model.g_dim + m + 1
    model.best = dict(theta=theta, w=r.w, iSigma_w=r.iSigma_w, P=theta[:m * d].reshape((m, d), order="F"),
                      v=theta[o2:o2 + m].reshape(m, 1), priors=pri)
    Xt = X[:60].copy()
    Xt[5:25, 1] = np.nan
    Xt[15:40, 2] = np.nan
    Xt[50:55, 0] = np.nan
    Psi = synth.make_psi(60, d, method, seed=4) if psi else None
    mu, sigma, nu, be, ga, PHI = O.predict(Xt, model, Psi=Psi)
    gm = L.make_model(d, 1, m, method, True)
    mu2, nu2, be2, ga2, PHI2 = L.predict_core(gm, theta, r.w, r.iSigma_w, Xt, Psi, want_phi=True, priors=pri)
    assert rel(mu2, mu) <= TOL and rel(PHI2, PHI) <= 1e-10
    assert rel(nu2, nu) <= 1e-8 and rel(be2, be) <= 1e-8
    assert np.max(np.abs(ga2 - ga)) <= 1e-9 * max(1.0, np.max(np.abs(mu)) ** 2)
    with pytest.raises(L.GpzError):
        L.predict_core(gm, theta, r.w, r.iSigma_w, Xt, Psi)          # missing rows without priors: loud failure


@pytest.mark.parametrize("method,slices,tol", [("VC", 7, 1e-9), ("VD", 7, 1e-9), ("GL", 7, 1e-7), ("VC", 6, 1e-7)])
def test_int8_tensor_core_tgemm_matches_fp64(method, slices, tol):
    """T = PHI*iSigma and the Gram as error-free base-256 digit GEMMs on tcgen05 (ozaki.cu, ozmma.cu) against the fp64
    DMMA path and the oracle."""
    model, theta, X, Y, Psi, omega, tr, va = problem(method, True, False, False, n=3000, d=5, m=140, seed=21)
    ref = O.GPz(theta, model, X, Y, None, omega, tr, va)
    gm = L.make_model(model.d, 1, model.m, method, True)
    res = {}
    for oz in (0, slices, -slices):
        ctx = L.Context(gm, X, Y, None, omega, tr, va)
        ctx.set_option("ozaki_slices", abs(oz))
        ctx.set_option("ozaki_gram", 1 if oz > 0 else 0)      # -slices: int8 T-GEMM with the fp64 DMMA Gram
        res[oz] = ctx.eval(theta)
        f2, g2, _ = ctx.eval(theta)
        assert f2 == res[oz][0] and np.array_equal(g2, res[oz][1])
        ctx.close()
    assert_eval_matches(model, ref, *res[slices], tol=max(tol, 1e-7))
    f0, g0, _ = res[0]
    f2, g2, _ = res[-slices]
    assert abs(f2 - f0) <= 1e-12 * abs(f0)          # the objective value does not depend on T
    for key in (slices, -slices):
        f1, g1, _ = res[key]
        assert abs(f1 - f0) <= 1e-10 * abs(f0)
        gb0, gb1 = grad_blocks(model, g0), grad_blocks(model, g1)
        for nm in gb0:
            assert rel(gb1[nm], gb0[nm]) <= tol, (key, nm, rel(gb1[nm], gb0[nm]))


def _exact_matmul_nt(A, B):
    """A @ B.T in exact rational arithmetic (every double is an integer times a power of two)."""
    from fractions import Fraction
    fa = [[Fraction(float(v)) for v in row] for row in A]
    fb = [[Fraction(float(v)) for v in row] for row in B]
    return [[sum(x * y for x, y in zip(ra, rb)) for rb in fb] for ra in fa]


@pytest.mark.parametrize("M,N,K,digits", [(37, 29, 200, 7), (300, 130, 128, 7), (64, 48, 1000, 6), (20, 260, 333, 5)])
def test_digit_gemm_against_exact_arithmetic(M, N, K, digits):
    """gpz_dgemm_nt (digit extraction + the hand-written tcgen05 kernel of ozmma.cu) against exact rational arithmetic:
    ragged sizes (TMA zero fill), rows of very different magnitude (per-row scales), sign changes."""
    rng = np.random.default_rng(100 + digits)
    A = rng.standard_normal((M, K)) * np.exp(rng.uniform(-20, 20, (M, 1)))
    B = rng.standard_normal((N, K)) * np.exp(rng.uniform(-20, 20, (N, 1)))
    A[:, ::7] *= 1e-6                      # wide dynamic range inside the rows as well
    C = L.dgemm_nt(A, B, digits=digits)
    exact = _exact_matmul_nt(A[:12], B[:10])                     # a corner is enough for the exact check (pure Python)
    amax, bmax = np.max(np.abs(A), axis=1), np.max(np.abs(B), axis=1)
    bound = 2.0 ** (-8 * digits + 4) * K                        # dropped digit pairs + digit rounding, relative to the scales
    for i in range(12):
        for j in range(10):
            err = abs(float(exact[i][j] - __import__("fractions").Fraction(float(C[i, j]))))
            assert err <= bound * amax[i] * bmax[j] + 2.0 ** -52 * abs(float(exact[i][j])), (i, j, err)
    ref = A @ B.T                                               # and the whole matrix against the fp64 BLAS result
    tol = max(2.0 ** (-8 * digits + 4), 1e-15) * K
    assert np.all(np.abs(C - ref) <= tol * np.outer(amax, bmax) + 1e-13 * np.abs(ref))
    C2 = L.dgemm_nt(A, B, digits=digits)
    assert np.array_equal(C, C2)                                # static work partition: bit-reproducible


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (257, 129, 130), (3, 1000, 16384), (1025, 7, 127), (2, 2, 4097)])
def test_digit_gemm_extreme_shapes(M, N, K):
    """Single rows / columns, one-past-a-tile sizes, the largest K one int32 level accumulator may take (16384)."""
    rng = np.random.default_rng(M * 7919 + N * 31 + K)
    A = rng.standard_normal((M, K))
    B = rng.standard_normal((N, K))
    C = L.dgemm_nt(A, B)
    ref = A @ B.T
    amax, bmax = np.max(np.abs(A), axis=1), np.max(np.abs(B), axis=1)
    assert np.all(np.abs(C - ref) <= 1e-15 * K * np.outer(amax, bmax) + 1e-13 * np.abs(ref))


def test_digit_gemm_propagates_non_finite_inputs():
    A = np.ones((8, 128)); B = np.ones((8, 128))
    A[3, 5] = np.nan
    assert np.all(np.isnan(L.dgemm_nt(A, B)))
    with pytest.raises(L.GpzError):
        L.dgemm_nt(np.ones((4, 20000)), np.ones((4, 20000)))    # K beyond the exact-accumulation bound


@pytest.mark.parametrize("case", ["rank_deficient", "gap_to_1e-18", "indefinite", "odd_m_rank_deficient"])
def test_inv_logdet_reference_svd_semantics(case):
    """GPz/inv_logdet.m:3-15 pseudo-inverts by SVD and DROPS singular values <= m eps(max s); logdet sums the kept ones.
    gpz_inv_logdet takes the Cholesky route while that cannot differ and runs the truncating SVD itself (one-sided Jacobi on
    the device) when the factorisation fails or the factor shows cond(X) near 1 / (m eps): rank-deficient, numerically
    singular and indefinite inputs then follow the reference instead of returning NaN (round 1)."""
    rng = np.random.default_rng(7)
    m = 300 if case != "odd_m_rank_deficient" else 77
    Q, _ = np.linalg.qr(rng.standard_normal((m, m)))
    if case in ("rank_deficient", "odd_m_rank_deficient"):
        B = rng.standard_normal((m, m // 3))
        X = B @ B.T
    elif case == "gap_to_1e-18":
        lam = np.concatenate([np.logspace(0, -4, m // 2), 1e-18 * (1.0 + rng.random(m - m // 2))])
        X = (Q * lam) @ Q.T
    else:
        lam = np.concatenate([np.logspace(0, -3, m // 2), -np.logspace(0, -3, m - m // 2)])
        X = (Q * lam) @ Q.T
    X = 0.5 * (X + X.T)
    Xi0, ld0 = O.inv_logdet(X)
    Xi, ld = L.inv_logdet(X)
    assert np.all(np.isfinite(Xi)) and np.isfinite(ld)
    s = np.linalg.svd(X, compute_uv=False)
    kept = s[s > m * np.spacing(s.max())]
    tol = 1e-9 * kept.max() / kept.min()          # pinv accuracy ~ cond of the KEPT part x eps, with margin
    assert rel(Xi, Xi0) <= max(tol, 1e-10), (case, rel(Xi, Xi0), tol)
    assert abs(ld - ld0) <= 1e-9 * max(1.0, abs(ld0)), (case, ld, ld0)


@pytest.mark.parametrize("slices", [0, 7])
def test_non_finite_parameters_return_nan(slices):
    """Error convention: numerical trouble returns NaN with status 0, minFunc's isLegal handles it
    (WolfeLineSearch.m:53) -- also on the int8 digit path, where NaN cannot be expressed in digits."""
    model, theta, X, Y, Psi, omega, tr, va = problem("VC", True, False, False, n=2000, d=4, m=60, seed=5)
    gm = L.make_model(model.d, 1, model.m, "VC", True)
    ctx = L.Context(gm, X, Y, None, omega, tr, va)
    ctx.set_option("ozaki_slices", slices)
    f, g, st = ctx.eval(theta)
    assert np.isfinite(f) and np.all(np.isfinite(g))
    bad = theta.copy()
    bad[3] = np.nan
    f, g, st = ctx.eval(bad)
    assert np.isnan(f) and np.all(np.isnan(g))
    f2, g2, _ = ctx.eval(theta)                                  # and the context recovers
    assert np.isfinite(f2) and np.all(np.isfinite(g2))
    ctx.close()


@pytest.mark.parametrize("d,chunk", [(3, None), (6, 1024), (12, None)])
def test_gc_psi_fast_path_matches_generic_kernels_and_oracle(d, chunk):
    """GC + Psi (BASELINE config 5): one d x d factorisation per row + GEMMs (gcpsi.cu) against the per-(i,j) kernels that
    follow getPHI.m:80-88 / GPz.m:166-184 literally, and against the oracle; also row-chunked (features rebuilt per chunk)."""
    n, m = 2600, 40
    model, theta, X, Y, Psi, omega, tr, va = problem("GC", True, True, False, n=n, d=d, m=m, seed=31)
    gm = L.make_model(d, 1, m, "GC", True)
    out = {}
    for fast in (1, 0):
        ctx = L.Context(gm, X, Y, Psi, omega, tr, va)
        ctx.set_option("gc_fast", fast)
        if chunk:
            ctx.set_option("chunk_rows", chunk)
        out[fast] = ctx.eval(theta)
        f2, g2, _ = ctx.eval(theta)
        assert f2 == out[fast][0] and np.array_equal(g2, out[fast][1])
        ctx.close()
    (f1, g1, s1), (f0, g0, s0) = out[1], out[0]
    assert abs(f1 - f0) <= 1e-11 * abs(f0)
    gb1, gb0 = grad_blocks(model, g1), grad_blocks(model, g0)
    for nm in gb0:
        assert rel(gb1[nm], gb0[nm]) <= 1e-8, (nm, rel(gb1[nm], gb0[nm]))
    for key in s0:
        assert abs(s1[key] - s0[key]) <= 1e-10 * max(1.0, abs(s0[key]))
    if d <= 6:
        ref = O.GPz(theta, model, X, Y, Psi, omega, tr, va)
        assert_eval_matches(model, ref, f1, g1, s1, tol=1e-8)


def test_two_devices_in_one_process():
    """Opt-in shared-memory attributes and the resident-cluster count are per device: a second context on another GPU of the
    same process must give the same bits."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    model, theta, X, Y, Psi, omega, tr, va = problem("VC", True, False, False, n=3000, d=4, m=150, seed=17)
    gm = L.make_model(model.d, 1, model.m, "VC", True)
    out = []
    for dev in (0, 1):
        ctx = L.Context(gm, X, Y, None, omega, tr, va, device=dev)
        out.append(ctx.eval(theta))
        ctx.close()
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("ngpus", [1, 2])
def test_single_process_multi_gpu_context(ngpus):
    """gpz_create_multi (one caller thread -- the MATLAB interpreter holding the closure of train.m:40 -- drives
    N GPUs): rows split inside the library, one worker thread and one NCCL rank per device.  Against the plain one-GPU
    context: eval, the fit exit, getPrior and a short device-resident training run."""
    import torch
    if torch.cuda.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    model, theta, X, Y, Psi, omega, tr, va = problem("VC", True, False, False, n=6000, d=4, m=150, seed=19)
    gm = L.make_model(model.d, 1, model.m, "VC", True)
    one = L.Context(gm, X, Y, None, omega, tr, va)
    f0, g0, st0 = one.eval(theta)
    nl0, w0, iS0 = one.fit(theta)
    pr0 = one.get_prior(theta)
    th0, bt0, bv0, info0 = one.train(theta, theta, float("nan"), max_iter=4, training_only=0)
    one.close()
    mc = L.MultiContext(gm, X, Y, None, omega, tr, va, ngpus=ngpus)
    f, g, st = mc.eval(theta)
    f2, g2, _ = mc.eval(theta)
    assert f == f2 and np.array_equal(g, g2)
    tol = 0.0 if ngpus == 1 else 1e-9          # one device: the same kernels on the same rows -> the same bits
    assert abs(f - f0) <= tol * abs(f0) and rel(g, g0) <= max(tol, 0.0)
    for key in st0:
        assert abs(st[key] - st0[key]) <= max(tol, 0.0) * max(1.0, abs(st0[key])) + (0.0 if ngpus == 1 else 1e-12)
    nl, w, iS = mc.fit(theta)
    assert rel(nl, nl0) <= max(tol, 0.0) and rel(w, w0) <= max(10 * tol, 0.0) and rel(iS, iS0) <= max(10 * tol, 0.0)
    assert rel(mc.get_prior(theta), pr0) <= max(tol, 0.0)
    th, bt, bv, info = mc.train(theta, theta, float("nan"), max_iter=4, training_only=0)
    assert info["iterations"] == info0["iterations"] and info["fun_evals"] == info0["fun_evals"]
    assert rel(th, th0) <= max(1e3 * tol, 0.0) and abs(bv - bv0) <= max(1e3 * tol, 0.0) * abs(bv0)
    mc.close()


@pytest.mark.parametrize("method,psi,nan", [("VD", False, False), ("VC", True, True), ("GL", False, True)])
def test_small_problem_graph_replay_is_bit_identical(method, psi, nan):
    """Launch-bound sizes replay the whole evaluation as one CUDA graph (api.cu eval_device): same bits as plain launches,
    for a sequence of different thetas, including a non-finite one in the middle."""
    model, theta, X, Y, Psi, omega, tr, va = problem(method, True, psi, nan, n=400, d=3, m=20, seed=21)
    gm = L.make_model(model.d, model.k, model.m, model.method, model.heteroscedastic)
    thetas = [synth.perturb_theta(theta, 0.05, s) for s in range(6)]
    thetas[3] = thetas[3].copy()
    thetas[3][0] = np.nan
    outs = {}
    for graph in (0, 1, -1):
        ctx = L.Context(gm, X, Y, Psi, omega, tr, va)
        ctx.set_option("graph", graph)
        outs[graph] = [ctx.eval(th) for th in thetas]
        replays = ctx.graph_replays()
        assert replays == (0 if graph == 0 else len(thetas) - 1), (graph, replays)
        assert ctx.last_timing()["total"] > 0
        ctx.close()
    for a, b, c in zip(outs[0], outs[1], outs[-1]):
        assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1], equal_nan=True)
        assert np.array_equal(a[0], c[0], equal_nan=True) and np.array_equal(a[1], c[1], equal_nan=True)
        assert all(np.array_equal(a[2][k], b[2][k], equal_nan=True) for k in a[2])
    ref = O.GPz(thetas[5], model, X, Y, Psi, omega, tr, va)
    assert abs(outs[1][5][0] - ref.nlogML) <= TOL * abs(ref.nlogML)


@pytest.mark.parametrize("n,d,m", [(2, 1, 1), (3, 2, 1), (7, 2, 3), (33, 1, 5), (40, 2, 129), (129, 3, 128), (1025, 2, 7)])
@pytest.mark.parametrize("method", ["VL", "VD", "GC", "VC"])
def test_eval_edge_shapes(n, d, m, method):
    """Tiny and ragged extents: fewer rows than a warp / a tile / bases (m > n: SIGMA carried by the prior), one basis, one
    input dimension, m on both sides of the 128-wide tile, n one past the 1024-row digit chunk; no validation set."""
    rng = np.random.default_rng(100 + n + m)
    X = rng.standard_normal((n, d))
    Y = (np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1)))
    Y = Y - Y.mean()
    theta = synth.perturb_theta(synth.make_theta0(X, Y + 1e-3 * np.arange(n)[:, None], method, m, het=True, seed=1), 0.05, 2)
    model = O.Model(d=d, k=1, m=m, method=method, heteroscedastic=True)
    for digits in ((None, 7) if m > 1 else (None,)):
        ref, f, g, st, ctx = run_both(model, theta, X, Y, None, None, None, None, digits=digits)
        # m >= n: SIGMA = alpha + a rank-n Gram, conditioning set by the prior; the SVD pseudo-inverse of the oracle and the
        # Cholesky inverse differ by cond * eps there
        assert_eval_matches(model, ref, f, g, st, tol=1e-9 if m < n else 1e-7)
        ctx.close()


@pytest.mark.parametrize("method", ["VD", "VC"])
def test_predict_full_through_the_int8_engine(method):
    """m > 128: nu = rowsum((PHI iSigma) .* PHI) of predictFull runs through the tcgen05 digit GEMM, as T does in the
    objective (ragged: 700 rows, 150 of 256 padded bases); a non-finite theta gives NaN, not garbage digits."""
    n, d, m = 1500, 3, 150
    model, theta, X, Y, _, omega, tr, _ = problem(method, True, False, False, n=n, d=d, m=m, seed=13)
    r = O.GPz(theta, model, X, Y, None, omega, tr, None, fit_only=True)
    model.muX, model.sdX, model.muY = np.zeros(d), np.ones(d), np.zeros(1)
    o2 = m * d + model.g_dim + m + 1
    model.best = dict(theta=theta, w=r.w, iSigma_w=r.iSigma_w, P=theta[:m * d].reshape((m, d), order="F"),
                      v=theta[o2:o2 + m].reshape(m, 1))
    Xt = X[:700]
    mu, sigma, nu, be, ga, PHI = O.predict(Xt, model)
    gm = L.make_model(d, 1, m, method, True)
    mu2, nu2, be2, ga2, PHI2 = L.predict_core(gm, theta, r.w, r.iSigma_w, Xt, None, want_phi=True)
    # m = 150 bases in 3-d: iSigma_w has entries of mixed sign up to ~1e7 while nu ~ 1e-2: nu is a cancelling sum and the
    # fp64 oracle itself only carries it to cond * eps
    assert rel(mu2, mu) <= 1e-8 and rel(nu2, nu) <= 1e-6 and rel(be2, be) <= TOL and rel(PHI2, PHI) <= 1e-12
    bad = theta.copy()
    bad[0] = np.nan
    _, nu3, _, _, _ = L.predict_core(gm, bad, r.w, r.iSigma_w, Xt, None)
    assert np.isnan(nu3).all()


def test_demo_sinc_script_size_against_oracle():
    """Config 1' (SURVEY 8): demo_sinc.m as the script really runs it -- ~5 250 training rows of 7 500, d = 1 (so VL),
    m = 100, heteroscedastic, input noise Psi n x 1 -- at FULL size against the oracle."""
    rng = np.random.default_rng(11)
    n, m = 7500, 100
    X = rng.uniform(-10, 10, (n, 1))
    Y = (np.sinc(X[:, 0] / np.pi) + (0.05 + 0.2 * (1 + np.sin(X[:, 0] / 3)) / 2) * rng.standard_normal(n)).reshape(n, 1)
    Xz, Yc = (X - X.mean()) / X.std(), Y - Y.mean()
    Psi = rng.gamma(1.0, 0.25, (n, 1)) ** 2 / X.var()
    tr = rng.random(n) < 0.7
    va = ~tr & (rng.random(n) < 0.5)
    theta = synth.perturb_theta(synth.make_theta0(Xz, Yc, "VL", m, het=True, seed=3), 0.05, 4)
    model = O.Model(d=1, k=1, m=m, method="VL", heteroscedastic=True)
    ref, f, g, st, ctx = run_both(model, theta, Xz, Yc, Psi, None, tr, va)
    assert_eval_matches(model, ref, f, g, st, tol=1e-8)       # 100 bases on a line: cond(SIGMA) ~ 1e9
    ctx.close()


def test_demo_photoz_full_size_properties():
    """Config 2 at FULL size (60 000 training + 60 000 validation rows, d = 5, m = 100, VC, Psi 5 x 5 x n), where the oracle needs
    a quarter of an hour: (1) bit-reproducibility; (2) analytic gradient against a central difference of the objective;
    (3) with Psi -> 0 the per-(sample, basis) Cholesky kernels of getPHI.m:80-88 / GPz.m:166-184 must agree with the
    tensor-core no-Psi path (identity T4 of the oracle pins, here between two unrelated kernel families)."""
    n, d, m = 120_000, 5, 100
    X, Y = synth.make_data(n, d, seed=0)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, "VC", m, het=True, seed=1), 0.05, 2)
    Psi = synth.make_psi(n, d, "VC", seed=3)
    tr = np.arange(n) % 2 == 0
    gm = L.make_model(d, 1, m, "VC", True)
    ctx = L.Context(gm, X, Y, Psi, None, tr, ~tr)
    f1, g1, st1 = ctx.eval(theta)
    f2, g2, _ = ctx.eval(theta)
    assert np.isfinite(f1) and np.isfinite(g1).all() and f1 == f2 and np.array_equal(g1, g2)
    assert np.isfinite(st1["validLL"]) and st1["validRMSE"] > 0
    u = np.random.default_rng(0).standard_normal(theta.size)
    u /= np.linalg.norm(u)
    h = 1e-5
    fp, _, _ = ctx.eval(theta + h * u)
    fm, _, _ = ctx.eval(theta - h * u)
    assert abs((fp - fm) / (2 * h) - g1 @ u) <= 1e-6 * max(1.0, np.linalg.norm(g1))
    ctx.close()
    ctx = L.Context(gm, X, Y, Psi * 1e-30, None, tr, ~tr)
    fa, ga, sta = ctx.eval(theta)
    ctx.close()
    ctx = L.Context(gm, X, Y, None, None, tr, ~tr)
    fb, gb, stb = ctx.eval(theta)
    ctx.close()
    assert abs(fa - fb) <= 1e-10 * abs(fb) and rel(ga, gb) <= 1e-8
    assert abs(sta["validLL"] - stb["validLL"]) <= 1e-10 * abs(stb["validLL"])
