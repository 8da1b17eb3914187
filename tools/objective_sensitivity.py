import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_train import _gpz_case
for method in ("GL", "VD", "VC"):
    ctx, th = _gpz_case(method)
    f0, g0, _ = ctx.eval(th)
    rng = np.random.default_rng(0)
    dfs = []
    for r in range(5):
        th2 = np.where(rng.random(th.size) < 0.5, np.nextafter(th, np.inf), np.nextafter(th, -np.inf))
        f1, g1, _ = ctx.eval(th2)
        dfs.append((abs(f1 - f0), np.max(np.abs(g1 - g0)) / np.max(np.abs(g0))))
    print(method, "f0", f0, "|df| for 1-ulp theta changes:", ["%.1e / g %.1e" % d for d in dfs])
    ctx.close()
