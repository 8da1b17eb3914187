"""Latency of one evaluation at small sizes (launch-bound regime): wall ms, device phases, launches; with
OZ=0/7 both GEMM paths (fp64 DMMA vs int8 digit GEMMs) to place the automatic switch."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpz_b200 import _lib as L, synth

cases = [("VL", 3000, 1, 100), ("VC", 60000, 5, 100), ("VD", 3000, 2, 25), ("VD", 20000, 3, 100), ("VD", 200000, 3, 100),
         ("VD", 10000, 3, 500), ("VD", 40000, 3, 500), ("VD", 3000, 3, 1000), ("VD", 20000, 3, 1000)]
for meth, n, d, m in cases:
    X, Y = synth.make_data(n, d, seed=0)
    th = synth.make_theta0(X, Y, meth, m, het=True, seed=1)
    va = np.arange(n) % 5 == 4
    out = []
    for oz in (None, 0, 7):
        ctx = L.Context(L.make_model(d, 1, m, meth, True), X, Y, training=~va, validation=va)
        if oz is not None:
            ctx.set_option("ozaki_slices", oz)
        ctx.eval(th)
        l0 = ctx.launch_count()
        ts = []
        for _ in range(10):
            t = time.perf_counter(); ctx.eval(th); ts.append(time.perf_counter() - t)
        nl = (ctx.launch_count() - l0) / 10
        tm = ctx.last_timing()
        out.append("%s: %.3f ms (%d launches; gram %.3f solve %.3f tgemm %.3f bp %.3f)" % (
            {None: "auto", 0: "dmma", 7: "int8"}[oz], 1e3 * min(ts), nl, tm["gram"], tm["solve"], tm["tgemm"], tm["backproj"]))
        ctx.close()
    MP = (m + 127) // 128 * 128
    print(meth, n, d, m, "n_tr*MP^2=%.1e" % (0.8 * n * MP * MP), " | ".join(out), flush=True)
