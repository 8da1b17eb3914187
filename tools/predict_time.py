"""predict() throughput at the headline shape: n test rows, d=10, m=1000, VC (host buffers in and out)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpz_b200 import _lib as L, synth

n, d, m, meth = (int(sys.argv[1]) if len(sys.argv) > 1 else 1000000), 10, 1000, "VC"
X, Y = synth.make_data(n, d, seed=0)
th = synth.make_theta0(X, Y, meth, m, het=True, seed=1)
gm = L.make_model(d, 1, m, meth, True)
ctx = L.Context(gm, X[:200000], Y[:200000])
_, w, iS = ctx.fit(th)
ctx.close()
L.predict_core(gm, th, w, iS, X[:1000])
for rep in range(2):
    t = time.perf_counter()
    mu, nu, be, ga, _ = L.predict_core(gm, th, w, iS, X)
    dt = time.perf_counter() - t
    print(f"predict {n} rows: {1e3 * dt:.1f} ms wall ({n / dt / 1e6:.2f} M rows/s), nu mean {nu.mean():.3e}", flush=True)
