"""matlab/gpz_b200_mex.cpp EXECUTED: the gateway is linked against a minimal libmx mock (tests/mex_stub/mex_mock.cpp,
tests/mex_harness.py) and driven the way matlab/GPz.m / train.m / predict.m drive it.  The MEX convention it follows is the
reference's own (minFunc_2012/minFunc/mex/lbfgsProdC.c:7-44: plain mexFunction, mxGetPr in, mxCreateDoubleMatrix out,
an error raised on misuse).  CPU tests: the harness loads and usage errors are raised before anything touches a device.
GPU tests: every command against the ctypes path, bit for bit, and the size checks that protect MATLAB's heap."""
import os

import numpy as np
import pytest

import mex_harness as H
from gpz_b200 import _lib as L
from gpz_b200 import synth

MODEL = dict(d=3, k=1, m=20, method="VC", heteroscedastic=1.0)


def problem(n=600, seed=5):
    d, m = MODEL["d"], MODEL["m"]
    X, Y = synth.make_data(n, d, seed=seed)
    theta = synth.perturb_theta(synth.make_theta0(X, Y, "VC", m, het=True, seed=seed + 1), 0.05, seed + 2)
    omega = 0.5 + np.random.default_rng(seed + 3).random(n)
    tr = np.arange(n) % 4 != 0
    return np.array(X), np.array(Y), theta, omega, tr, ~tr


def test_harness_builds_and_usage_errors_need_no_device():
    assert os.path.exists(H.PATH), "run `make` (or __graft_entry__.build())"
    with pytest.raises(H.MexError, match="unknown command"):
        H.call("nope")
    with pytest.raises(H.MexError, match="invalid context handle"):
        H.call("eval", 12345.0, np.zeros(5))
    with pytest.raises(H.MexError, match="X must be n x model.d"):
        H.call("create", MODEL, np.zeros((5, 4)), np.zeros((5, 1)))
    with pytest.raises(H.MexError, match="Y must be n x model.k"):
        H.call("create", MODEL, np.zeros((5, 3)), np.zeros((4, 1)))
    with pytest.raises(H.MexError, match="model.m is missing"):
        H.call("create", dict(d=3, k=1, method="VC", heteroscedastic=1.0), np.zeros((5, 3)), np.zeros((5, 1)))
    with pytest.raises(H.MexError, match="theta has 3 elements"):
        H.call("predict", MODEL, np.zeros(3), np.zeros(20), np.zeros((20, 20)), np.zeros((4, 3)), None, None, nlhs=5)
    with pytest.raises(H.MexError, match="usage"):
        H.call("eval")
    H.call("destroy", 999.0)                     # unknown handles are ignored, as a double destroy would be


@pytest.mark.gpu
def test_gateway_commands_match_the_ctypes_path_bit_for_bit():
    X, Y, theta, omega, tr, va = problem()
    gm = L.make_model(MODEL["d"], 1, MODEL["m"], "VC", True)
    ctx = L.Context(gm, X, Y, None, omega, tr, va)
    h = H.call("create", MODEL, X, Y, None, omega, tr, va)
    # eval: [f, g, stats]
    f, g, st = H.call("eval", h, theta, nlhs=3)
    f0, g0, st0 = ctx.eval(theta)
    assert f.item() == f0 and np.array_equal(g.reshape(-1), g0)
    assert np.array_equal(st.reshape(-1), [st0[k] for k in ("trainRMSE", "trainLL", "validRMSE", "validLL")])
    # fit exit (GPz.m:84-87): [nl, w, iSigma_w], sized from the handle's model
    nl, w, iS = H.call("fit", h, theta, MODEL, nlhs=3)
    nl0, w0, iS0 = ctx.fit(theta)
    assert np.array_equal(nl.reshape(-1), np.asarray(nl0).reshape(-1)) and np.array_equal(w, w0) and np.array_equal(iS.reshape(iS0.shape), iS0)
    # phi with and without the optional arguments
    PHI, lnb, N = H.call("phi", h, theta, 0.0, MODEL, nlhs=3)
    PHI0, lnb0, N0 = ctx.phi(theta, 0, want_N=True)
    assert np.array_equal(PHI, PHI0) and np.array_equal(lnb, lnb0) and np.array_equal(N, N0)
    PHIv = H.call("phi", h, theta, 1.0)
    assert PHIv.shape == (int(va.sum()), MODEL["m"])
    assert np.array_equal(H.call("phi", h, theta), PHI0)                 # 'which' and 'model' are optional
    pr = H.call("get_prior", h, theta)
    assert np.array_equal(pr.reshape(-1), ctx.get_prior(theta))
    # train: the device-resident minFunc loop; inputs are never written
    th_in = theta.copy()
    th, best, bv, info = H.call("train", h, th_in, th_in, None, 6.0, np.inf, 0.0, 1.0, nlhs=4)
    assert np.array_equal(th_in, theta)
    th0, best0, bv0, info0 = ctx.train(theta, theta, float("nan"), max_iter=6, max_attempts=float("inf"), training_only=0)
    assert np.array_equal(th.reshape(-1), th0) and np.array_equal(best.reshape(-1), best0) and bv.item() == bv0
    assert int(info[0, 0]) == info0["iterations"] and int(info[0, 1]) == info0["fun_evals"]
    table = H.printed()
    assert "Iter" in table and "Valid MLL" in table and table.count("\n") >= info0["iterations"]
    # predict (stateless)
    Xt = X[va][:50]
    mu, nu, be, ga, ph = H.call("predict", MODEL, best.reshape(-1), w, iS, Xt, None, pr, nlhs=5)
    mu0, nu0, be0, ga0, ph0 = L.predict_core(gm, best0, w0, iS0, Xt, None, want_phi=True, priors=pr.reshape(-1))
    assert np.array_equal(mu, mu0) and np.array_equal(nu, nu0) and np.array_equal(be, be0) and np.array_equal(ph, ph0)
    # init-side helpers
    S = np.cov(np.random.default_rng(0).standard_normal((40, 200)))
    Xi, ld = H.call("inv_logdet", S, nlhs=2)
    Xi0, ld0 = L.inv_logdet(S)
    assert np.array_equal(Xi, Xi0) and ld.item() == ld0
    P = np.random.default_rng(1).standard_normal((7, 3))
    assert np.array_equal(H.call("dxy", X[:30], P), L.dxy(X[:30], P))
    assert np.array_equal(H.call("dxy_colmean", X, P).reshape(-1), L.dxy_colmean(X, P))
    H.call("destroy", h)
    with pytest.raises(H.MexError, match="invalid context handle"):
        H.call("eval", h, theta)
    ctx.close()


@pytest.mark.gpu
def test_gateway_rejects_mismatched_sizes_instead_of_corrupting_memory():
    """a stale handle with another model, or a theta of the wrong length, must raise -- the library reads and
    writes sizes derived from the context's own model."""
    X, Y, theta, omega, tr, va = problem(n=300)
    h = H.call("create", MODEL, X, Y, None, omega, tr, va)
    other = dict(MODEL, m=24)
    for cmd, args in (("eval", (h, theta[:-1])), ("fit", (h, np.concatenate([theta, [0.0]]))), ("phi", (h, theta[:5])),
                      ("get_prior", (h, theta[:7])), ("train", (h, theta[:-2], theta[:-2], None, 3.0, np.inf, 0.0))):
        with pytest.raises(H.MexError, match="theta has"):
            H.call(cmd, *args)
    for cmd, args in (("fit", (h, theta, other)), ("phi", (h, theta, 0.0, other)), ("get_prior", (h, theta, other))):
        with pytest.raises(H.MexError, match="model argument differs"):
            H.call(cmd, *args)
    with pytest.raises(H.MexError, match="which must be"):
        H.call("phi", h, theta, 2.0)
    f = H.call("eval", h, theta)                                          # the context is still usable
    assert np.isfinite(f.item())
    H.call("destroy", h)


@pytest.mark.gpu
def test_gateway_multi_gpu_handle():
    """create(..., ngpus): one caller thread, rows split over the devices inside the library (gpz_create_multi); eval / fit /
    get_prior / train go through the same commands, phi is refused on such a handle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    X, Y, theta, omega, tr, va = problem(n=2000)
    h1 = H.call("create", MODEL, X, Y, None, omega, tr, va, 1.0)
    h2 = H.call("create", MODEL, X, Y, None, omega, tr, va, 2.0)
    f1, g1, s1 = H.call("eval", h1, theta, nlhs=3)
    f2, g2, s2 = H.call("eval", h2, theta, nlhs=3)
    assert abs(f1.item() - f2.item()) <= 1e-10 * abs(f1.item()) and np.max(np.abs(g1 - g2)) <= 1e-9 * np.max(np.abs(g1))
    _, w1, _ = H.call("fit", h1, theta, nlhs=3)
    _, w2, _ = H.call("fit", h2, theta, nlhs=3)
    assert np.max(np.abs(w1 - w2)) <= 1e-8 * np.max(np.abs(w1))
    with pytest.raises(H.MexError, match="multi-GPU handle"):
        H.call("phi", h2, theta)
    H.call("destroy", h1)
    H.call("destroy", h2)
