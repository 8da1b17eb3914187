import sys
sys.path.insert(0, ".")
import numpy as np
from gpz_b200 import _lib as L, synth
X, Y = synth.make_data(400, 3, seed=0)
th = synth.make_theta0(X, Y, "VD", 20, het=True, seed=1)
ctx = L.Context(L.make_model(3, 1, 20, "VD", True), X, Y)
ctx.set_option("graph", 1)
for i in range(4):
    print(ctx.eval(th)[0], ctx.graph_replays())
