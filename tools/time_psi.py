"""Timing of the input-noise (Psi) paths: photoz-like VC+Psi and GC+Psi at d=32 (scratch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpz_b200 import _lib as L, synth
for meth, n, d, m in (("VC", 60000, 5, 100), ("VD", 60000, 5, 100), ("GC", 4000, 32, 2000), ("GC", 20000, 10, 500)):
    X, Y = synth.make_data(n, d, seed=0)
    th = synth.make_theta0(X, Y, meth, m, het=True, seed=1)
    Psi = synth.make_psi(n, d, meth, seed=2)
    ctx = L.Context(L.make_model(d, 1, m, meth, True), X, Y, Psi)
    f, g, st = ctx.eval(th)
    ts = []
    for _ in range(3):
        t = time.time(); f, g, st = ctx.eval(th); ts.append(time.time() - t)
    print(meth, "+Psi n=%d d=%d m=%d: eval %.1f ms" % (n, d, m, 1e3 * min(ts)), {k: round(float(v), 2) for k, v in ctx.last_timing().items() if k in ("phi", "gram", "solve", "tgemm", "backproj")}, flush=True)
    ctx.close()
