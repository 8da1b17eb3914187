"""How many base-256 digits do the int8 GEMMs need?  Gradient / NLML of one evaluation with different digit counts against the
7-digit result, next to the difference of the plain fp64 DMMA path (whose own rounding is the yardstick)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpz_b200 import _lib as L, synth

n, d, m, meth = int(sys.argv[1]) if len(sys.argv) > 1 else 200000, 10, 1000, "VC"
X, Y = synth.make_data(n, d, seed=0)
th = synth.perturb_theta(synth.make_theta0(X, Y, meth, m, het=True, seed=1), 0.01, 3)
res = {}
for name, opts in [("fp64", {"ozaki_slices": 0}), ("s7g7", {"ozaki_slices": 7}), ("s7g6", {"ozaki_slices": 7, "ozaki_gram_slices": 6}),
                   ("s7g5", {"ozaki_slices": 7, "ozaki_gram_slices": 5}), ("s7g4", {"ozaki_slices": 7, "ozaki_gram_slices": 4}),
                   ("s6g6", {"ozaki_slices": 6}), ("s6g5", {"ozaki_slices": 6, "ozaki_gram_slices": 5}), ("s5g5", {"ozaki_slices": 5})]:
    ctx = L.Context(L.make_model(d, 1, m, meth, True), X, Y)
    for k, v in opts.items():
        ctx.set_option(k, float(v))
    f, g, st = ctx.eval(th)
    res[name] = (f, g.copy())
    ctx.close()
f0, g0 = res["s7g7"]
md = m * d
blocks = {"dP": slice(0, md), "dGamma": slice(md, md + d * d * m), "rest": slice(md + d * d * m, None)}
for name, (f, g) in res.items():
    line = [f"{name}: f rel diff {abs(f - f0) / abs(f0):.2e}"]
    for bn, sl in blocks.items():
        line.append(f"{bn} {np.max(np.abs(g[sl] - g0[sl])) / np.max(np.abs(g0[sl])):.2e}")
    print("  ".join(line), flush=True)
