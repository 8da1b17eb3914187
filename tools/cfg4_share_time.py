import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from gpz_b200 import _lib as L, synth
n, d, m = 1250000, 10, 1000
X, Y = synth.make_data(n, d, seed=0)
th = synth.make_theta0(X, Y, "VC", m, het=True, seed=1)
ctx = L.Context(L.make_model(d, 1, m, "VC", True), X, Y)
f, g, st = ctx.eval(th)
ts = []
for _ in range(5):
    t = time.time(); f, g, st = ctx.eval(th); ts.append(time.time() - t)
print("cfg4 per-GPU share n=1.25e6 d=10 m=1000 VC: eval %.1f ms" % (1e3 * min(ts)), {k: round(float(v), 1) for k, v in ctx.last_timing().items() if k != "i8_gemms_ops"})
ctx.close()
