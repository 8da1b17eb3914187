"""BASELINE config 2 shape (demo_photoz: n=60000 train + 60000 valid, d=5, m=100, VC + per-sample input noise Psi)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpz_b200 import _lib as L, synth

n, d, m = 120000, 5, 100
X, Y = synth.make_data(n, d, seed=0)
th = synth.make_theta0(X, Y, "VC", m, het=True, seed=1)
Psi = synth.make_psi(n, d, "VC", seed=3)
tr = np.arange(n) % 2 == 0
ctx = L.Context(L.make_model(d, 1, m, "VC", True), X, Y, Psi, None, tr, ~tr)
f, g, st = ctx.eval(th)
ts = []
for _ in range(5):
    t = time.time(); f, g, st = ctx.eval(th); ts.append(time.time() - t)
print("photoz-shape VC+Psi n_train=60000 n_valid=60000 d=5 m=100: eval %.2f ms f=%.6f" % (1e3 * min(ts), f),
      {k: round(float(v), 2) for k, v in ctx.last_timing().items() if k != "i8_gemms_ops"}, st, flush=True)
ctx.close()
