"""init -> train -> predict through gpz_b200.api at a mid-size problem (the flow of demo_photoz.m without input noise)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpz_b200 import api, synth

n, d, m = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (300000, 5, 500)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 30
X, Y = synth.make_data(n, d, seed=0)
X, Y = np.array(X) * 2.0 + 1.0, np.array(Y) + 3.0
tr = np.arange(n) % 5 < 2
va = np.arange(n) % 5 == 2
te = np.arange(n) % 5 > 2
t = time.perf_counter(); model = api.init(X, Y, "VC", m, training=tr, seed=1); t_init = time.perf_counter() - t
t = time.perf_counter(); model = api.train(model, X, Y, maxIter=iters, training=tr, validation=va, display=False); t_train = time.perf_counter() - t
t = time.perf_counter(); mu, sigma, *_ = api.predict(X, model, selection=te); t_pred = time.perf_counter() - t
info = model["train_info"]
rmse = float(np.sqrt(np.mean((mu[:, 0] - Y[te, 0]) ** 2)))
print(f"n={n} d={d} m={m}: init {t_init:.2f} s, train {t_train:.2f} s ({info['iterations']} it, {info['fun_evals']} evals, optimiser call "
      f"{info['ms_total']:.0f} ms of which evals {info['ms_eval']:.0f} ms; rest = 2 x (fit + getPrior)), predict {int(te.sum())} rows {t_pred:.2f} s; "
      f"test RMSE {rmse:.4f}, mean sigma {float(np.mean(sigma)):.4f}, best validLL {info['best_valid']:.4f}")
