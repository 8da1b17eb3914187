"""A/B of an option that must be set before the first evaluation (one context per setting, interleaved repetitions):
    python tools/ab_create_option.py <workload> <option> <value A> <value B> [reps]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from gpz_b200 import _lib as L  # noqa: E402

name, opt, va, vb = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
_, d, m, method, _ = bench.WORKLOADS[name]
sh = bench.make_shard(name, 0, 1)
X, Y, Psi, theta0 = sh["X"], sh["Y"], sh["Psi"], sh["theta0"]
ths = bench.thetas_for(theta0, reps + 2)
ctxs = {}
for v in (va, vb):
    c = L.Context(L.make_model(d, 1, m, method, True), X, Y, Psi)
    c.set_option(opt, v)
    c.eval(ths[0])
    ctxs[v] = c
acc = {va: [], vb: []}
out = {}
for r in range(reps):
    for v in (va, vb):
        out[v] = ctxs[v].eval(ths[1 + r])
        t = ctxs[v].last_timing()
        t.update(ctxs[v].kernel_timing())
        acc[v].append(t)
for v in (va, vb):
    keys = [k for k in acc[v][0] if k not in ("i8_gemms_ops", "int8_slices", "int8_gram")]
    print(opt, v, {k: round(float(np.mean([t[k] for t in acc[v]])), 3) for k in keys})
fa, ga, _ = out[va]
fb, gb, _ = out[vb]
print("agreement: f", abs(fa - fb) / abs(fb), "g", float(np.max(np.abs(ga - gb)) / np.max(np.abs(gb))))
