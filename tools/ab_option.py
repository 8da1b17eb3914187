"""A/B of a process-wide or per-context option on one box, interleaved so that clock drift cancels:
    python tools/ab_option.py <workload> <option> <value A> <value B> [reps]
Prints the mean phase times (gpz_last_timing) of each setting.  Only options that may change after the first evaluation."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from gpz_b200 import _lib as L  # noqa: E402

name, opt, va, vb = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 6
_, d, m, method, _ = bench.WORKLOADS[name]
sh = bench.make_shard(name, 0, 1)
X, Y, Psi, theta0 = sh["X"], sh["Y"], sh["Psi"], sh["theta0"]
ctx = L.Context(L.make_model(d, 1, m, method, True), X, Y, Psi)
ths = bench.thetas_for(theta0, 2 * reps + 2)
ctx.eval(ths[0])
ctx.eval(ths[1])
acc = {va: [], vb: []}
res = {}
for r in range(reps):
    for v in (va, vb):
        ctx.set_option(opt, v)
        f, g, _ = ctx.eval(ths[2 + r])
        acc[v].append(dict(ctx.last_timing(), **{"k_" + k: x for k, x in ctx.kernel_timing().items()}))
        res.setdefault(r, []).append((f, g))
for v in (va, vb):
    keys = [k for k in acc[v][0] if k not in ("i8_gemms_ops", "int8_slices", "int8_gram")]
    print(opt, v, {k: round(float(np.mean([t[k] for t in acc[v]])), 3) for k in keys})
print("bit-identical f and gradient between the two settings:", all(a[0] == b[0] and np.array_equal(a[1], b[1]) for a, b in res.values()))
ctx.close()
