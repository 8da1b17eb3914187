"""How fast two minFunc runs separate on the GL objective: device optimiser vs oracle optimiser vs the oracle
optimiser started 1 ulp away (all three on the same GPU objective)."""
import sys

import numpy as np

sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_train import _gpz_case  # noqa: E402
from oracle import minfunc_oracle as MO  # noqa: E402

method = sys.argv[1] if len(sys.argv) > 1 else "GL"
ctx, th0 = _gpz_case(method)


def fs(th):
    f, g, st = ctx.eval(th)
    return f, g, (st["trainRMSE"], st["trainLL"], st["validRMSE"], st["validLL"])


logs = []
for start in (th0, np.nextafter(th0, np.inf)):
    log = []
    MO.train_loop(fs, start, start, -np.inf, max_iter=14, training_only=False, log=log)
    logs.append(log)
its = []
ctx.train(th0, th0, -np.inf, callback=lambda it: its.append(it) and False, max_iter=14, training_only=0)
print("it  evals(o,o',dev)  t_oracle        |f_o - f_o'|   |f_o - f_dev|   |t_o - t_o'|  |t_o - t_dev|")
for a, b, c in zip(logs[0], logs[1], its):
    print(f"{a['i']:2d}  {a['fun_evals']:3d} {b['fun_evals']:3d} {c['fun_evals']:3d}   {a['t']:.9f}   {abs(a['f'] - b['f']):.2e}      {abs(a['f'] - c['f']):.2e}"
          f"      {abs(a['t'] - b['t']):.2e}     {abs(a['t'] - c['t']):.2e}")
ctx.close()
