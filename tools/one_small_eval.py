import sys, os
sys.path.insert(0, ".")
import numpy as np
from gpz_b200 import _lib as L, synth
meth, n, d, m = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
X, Y = synth.make_data(n, d, seed=0)
th = synth.make_theta0(X, Y, meth, m, het=True, seed=1)
va = np.arange(n) % 5 == 4
ctx = L.Context(L.make_model(d, 1, m, meth, True), X, Y, training=~va, validation=va)
for _ in range(3): ctx.eval(th)
ctx.close()
