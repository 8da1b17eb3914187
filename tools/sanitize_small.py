"""Small end-to-end pass for compute-sanitizer: create, eval (plain + graph replay), train, fit, get_prior, predict."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpz_b200 import _lib as L, synth

for meth, n, d, m in (("VD", 700, 3, 20), ("VC", 1300, 2, 150)):
    X, Y = synth.make_data(n, d, seed=0)
    th = synth.make_theta0(X, Y, meth, m, het=True, seed=1)
    va = np.arange(n) % 5 == 4
    gm = L.make_model(d, 1, m, meth, True)
    ctx = L.Context(gm, X, Y, training=~va, validation=va)
    for _ in range(3):
        f, g, st = ctx.eval(th)
    x, best, bv, info = ctx.train(th, th, -np.inf, max_iter=6, training_only=0)
    nl, w, iS = ctx.fit(best)
    pr = ctx.get_prior(best)
    ctx.close()
    mu, nu, be, ga, _ = L.predict_core(gm, best, w, iS, X[:300])
    print(meth, f, info["f"], info["fun_evals"], float(nu.mean()), float(pr.sum()), flush=True)
print(L.dxy_colmean(np.random.default_rng(0).standard_normal((5000, 3)), np.zeros((4, 3))))
