"""CPU arm at the row counts BASELINE.md section 4 asks for: the NumPy restatement of the reference (oracle/gpz_oracle.py, all
host threads) timed at n in {2e4, 5e4, 1e5} rows of the headline shape (d=10, m=1000, VC, heteroscedastic), the linear model
t(n) = a n + c fitted through the two largest, and its residual at the smallest.  One evaluation each after a warm-up at 5 000
rows.  Run on the GPU box's host cores:  python tools/cpu_linearity.py gpurun_out/cpu_linearity.json [workload]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gpz_b200 import synth  # noqa: E402
from oracle import gpz_oracle as O  # noqa: E402


def one(name, ns):
    n, d, m, method = bench.WORKLOADS[name][:4]
    X, Y = synth.make_data(ns, d, seed=0)
    theta = bench.thetas_for(synth.make_theta0(X, Y, method, m, het=True, seed=1), 1)[0]
    model = O.Model(d=d, k=1, m=m, method=method, heteroscedastic=True)
    X, Y = np.array(X), np.array(Y)
    t0 = time.perf_counter()
    O.GPz(theta, model, X, Y)
    return time.perf_counter() - t0


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/cpu_linearity.json"
    name = sys.argv[2] if len(sys.argv) > 2 else "target"
    sizes = [int(s) for s in sys.argv[3].split(",")] if len(sys.argv) > 3 else [20000, 50000, 100000]
    bench.use_all_host_threads()
    one(name, 5000)
    ts = {ns: one(name, ns) for ns in sizes}
    (n1, t1), (n2, t2) = [(ns, ts[ns]) for ns in sizes[-2:]]
    a = (t2 - t1) / (n2 - n1)
    c = t2 - a * n2
    n_full = bench.WORKLOADS[name][0]
    pred0 = a * sizes[0] + c
    out = {"workload": name, "cores": os.cpu_count(), "seconds": {str(k): v for k, v in ts.items()}, "a_s_per_row": a, "c_s": c,
           "residual_at_smallest": (ts[sizes[0]] - pred0) / ts[sizes[0]],
           "extrapolated_s_per_eval_at_full_n": a * n_full + c, "extrapolated_evals_per_s": 1.0 / (a * n_full + c),
           "how": "oracle/gpz_oracle.py GPz (reference operation sequence), OpenBLAS on all host cores, one evaluation per size after a warm-up"}
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
