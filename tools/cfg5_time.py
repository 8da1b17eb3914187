"""BASELINE config 5 (per-GPU share): n=2.5e5 (of 1e6 over 4 GPUs), d=32, m=2000, GC + per-sample input noise Psi."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpz_b200 import _lib as L, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
d, m = 32, 2000
X, Y = synth.make_data(n, d, seed=0)
th = synth.make_theta0(X, Y, "GC", m, het=True, seed=1)
rng = np.random.default_rng(3)
Psi = np.empty((d, d, n))
for s0 in range(0, n, 20000):
    s1 = min(n, s0 + 20000)
    A = rng.standard_normal((s1 - s0, d, d))
    P_ = np.einsum("iab,icb->iac", A, A) * (0.1 / d) + 0.01 * np.eye(d)[None]
    Psi[:, :, s0:s1] = np.transpose(P_, (1, 2, 0))
t0 = time.time()
ctx = L.Context(L.make_model(d, 1, m, "GC", True), X, Y, Psi)
print("create %.1f s" % (time.time() - t0), flush=True)
f, g, st = ctx.eval(th)
ts = []
for _ in range(3):
    t = time.time(); f, g, st = ctx.eval(th); ts.append(time.time() - t)
print("cfg5 per-GPU n=%d d=%d m=%d GC+Psi: eval %.1f ms  f=%.6f finite_grad=%s" % (n, d, m, 1e3 * min(ts), f, bool(np.isfinite(g).all())),
      {k: round(float(v), 2) for k, v in ctx.last_timing().items() if k not in ("i8_gemms_ops",)}, flush=True)
ctx.close()
