"""ncu launch-list driver: a few evaluations at m = 1000, d = 10, VC on few rows with prep_block = 128 then 32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpz_b200 import _lib as L, synth
n, d, m = 20000, 10, 1000
X, Y = synth.make_data(n, d, seed=0)
th = synth.make_theta0(X, Y, "VC", m, het=True, seed=1)
ctx = L.Context(L.make_model(d, 1, m, "VC", True), X, Y)
for blk in (128, 32, 64):
    ctx.set_option("prep_block", blk)
    for _ in range(3):
        f, g, st = ctx.eval(th)
    print(blk, f, ctx.last_timing())
ctx.close()
