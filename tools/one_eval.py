"""One warm + N evaluations of a bench workload (this rank's share for the weak-scaling ones) -- the command ncu wraps
(profiles/README.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gpz_b200 import _lib as L

name = sys.argv[1] if len(sys.argv) > 1 else "target"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
_, d, m, method, _ = bench.WORKLOADS[name]
sh = bench.make_shard(name, 0, 1)
ctx = L.Context(L.make_model(d, 1, m, method, True), sh["X"], sh["Y"], sh["Psi"])
for th in bench.thetas_for(sh["theta0"], 1 + reps):
    f, g, st = ctx.eval(th)
print(name, "nlogML", f, {k: round(float(v), 3) for k, v in ctx.last_timing().items()}, {k: round(float(v), 3) for k, v in ctx.kernel_timing().items()},
      "launches/eval", ctx.launch_count() // (1 + reps))
ctx.close()
