"""getPrior (train.m:59,74) and fit at the headline shape."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpz_b200 import _lib as L, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
d, m = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (10, 1000)
meth = "VC"
X, Y = synth.make_data(n, d, seed=0)
th = synth.make_theta0(X, Y, meth, m, het=True, seed=1)
ctx = L.Context(L.make_model(d, 1, m, meth, True), X, Y)
ctx.eval(th)
for rep in range(2):
    t = time.perf_counter(); pr = ctx.get_prior(th); t1 = time.perf_counter() - t
    t = time.perf_counter(); ctx.fit(th); t2 = time.perf_counter() - t
    print(f"get_prior {1e3 * t1:.1f} ms (sum {pr.sum():.6f}, max {pr.max():.3e}), fit {1e3 * t2:.1f} ms", flush=True)
ctx.close()
