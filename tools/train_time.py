"""Where the time of a training run goes: device-resident optimiser (gpz_train) against a host-driven loop
(oracle minFunc restatement calling gpz_eval per evaluation).  usage: train_time.py [n m d method iters]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gpz_b200 import _lib  # noqa: E402
from oracle import gpz_oracle as O  # noqa: E402
from oracle import minfunc_oracle as MO  # noqa: E402

n, m, d = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (60000, 100, 5)
method = sys.argv[4] if len(sys.argv) > 4 else "VC"
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 60
host_too = (sys.argv[6] != "0") if len(sys.argv) > 6 else True
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d))
Y = (np.sin(X[:, 0]) + 0.2 * X[:, -1] ** 2 + 0.1 * rng.standard_normal(n)).reshape(n, 1)
Y -= Y.mean()
P = X[rng.choice(n, m, replace=False)] + 0.05 * rng.standard_normal((m, d))
theta0 = O.pack_theta_init(P, O.init_gamma(X[:5000], P, m), float(np.var(Y)), method, True)
va = np.arange(n) % 5 == 4
ctx = _lib.Context(_lib.make_model(d, 1, m, method, True), X, Y, training=~va, validation=va)
ctx.eval(theta0)
t0 = time.time()
xd, best, bv, info = ctx.train(theta0, theta0, -np.inf, max_iter=iters, training_only=0)
t_dev = time.time() - t0
print(f"p={theta0.size} device: {info['iterations']} it, {info['fun_evals']} evals, total {info['ms_total']:.1f} ms, in evals "
      f"{info['ms_eval']:.1f} ms, optimiser+sync overhead {(info['ms_total'] - info['ms_eval']) / info['iterations']:.3f} ms/it, "
      f"f={info['f']:.6f} ({info['message']}) wall {t_dev:.2f}s")
if host_too:
    def fs(th):
        f, g, st = ctx.eval(th)
        return f, g, (st["trainRMSE"], st["trainLL"], st["validRMSE"], st["validLL"])
    t0 = time.time()
    xo, bo, bvo, flag, io = MO.train_loop(fs, theta0, theta0, -np.inf, max_iter=iters, training_only=False)
    t_host = time.time() - t0
    print(f"host loop (numpy two-loop + gpz_eval): {io['iterations']} it, {io['funcCount']} evals, {1e3 * t_host:.1f} ms; "
          f"|theta_dev - theta_host| = {np.max(np.abs(xd - xo)):.2e}, bestLL {bv:.6f} vs {bvo:.6f}")
ctx.close()
