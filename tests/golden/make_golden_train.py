"""Golden training trajectory (tests/golden/train_VD.npz): the oracle's minFunc restatement + callBack.m rule driving the
oracle's GPz objective (self-derived, like every golden vector here -- no MATLAB run exists to compare with).  Guards both
restatements against drift and gives the CUDA optimiser a fixed target.   python tests/golden/make_golden_train.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from gpz_b200 import synth  # noqa: E402
from oracle import gpz_oracle as O  # noqa: E402
from oracle import minfunc_oracle as MO  # noqa: E402


def build(max_iter=12, max_attempts=4):
    n, d, m, method = 500, 2, 10, "VD"
    X, Y = synth.make_data(n, d, seed=77)
    X, Y = np.array(X), np.array(Y)
    theta0 = synth.make_theta0(X, Y, method, m, het=True, seed=78)
    tr = np.arange(n) % 3 != 0
    va = ~tr
    model = O.Model(d=d, k=1, m=m, method=method, heteroscedastic=True)

    def fun_stats(th):
        r = O.GPz(th, model, X, Y, None, None, tr, va)
        return r.nlogML, r.grad, tuple(r.stats[s] for s in ("trainRMSE", "trainLL", "validRMSE", "validLL"))

    log = []
    x, best, bv, flag, info = MO.train_loop(fun_stats, theta0, theta0, -np.inf, max_iter=max_iter, max_attempts=max_attempts,
                                            training_only=False, log=log)
    return dict(X=X, Y=Y, training=tr, validation=va, theta0=theta0, theta_last=x, theta_best=best, best_valid=bv, exitflag=flag,
                iterations=info["iterations"], fun_evals=info["funcCount"], f=np.array([e["f"] for e in log]),
                t=np.array([e["t"] for e in log]), validLL=np.array([e["stats"][3] for e in log]),
                improved=np.array([e["improved"] for e in log]), evals=np.array([e["fun_evals"] for e in log]),
                meta=np.array([n, d, m, 1, 1, max_iter, max_attempts]), method=np.array(method))


if __name__ == "__main__":
    out = build()
    np.savez_compressed(os.path.join(HERE, "train_VD.npz"), **out)
    print("train_VD:", out["iterations"], "iterations,", out["fun_evals"], "evaluations, f:", out["f"][:3], "...", out["f"][-1])
