"""One warm + N evaluations of a bench workload -- the command ncu wraps (profiles/README.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gpz_b200 import _lib as L

name = sys.argv[1] if len(sys.argv) > 1 else "target"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n, d, m, method, X, Y, theta0 = bench.make_problem(name)
ctx = L.Context(L.make_model(d, 1, m, method, True), X, Y)
for th in bench.thetas_for(theta0, 1 + reps):
    f, g, st = ctx.eval(th)
print(name, "nlogML", f, ctx.last_timing(), "launches/eval", ctx.launch_count() // (1 + reps))
ctx.close()
