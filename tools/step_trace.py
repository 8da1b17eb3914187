"""Per-evaluation device time over a long back-to-back run at the headline size (power-cap dynamics)."""
import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpz_b200 import _lib as L, synth

n, d, m = 1000000, 10, 1000
X, Y = synth.make_data(n, d, seed=0)
th = synth.make_theta0(X, Y, "VC", m, het=True, seed=1)
ctx = L.Context(L.make_model(d, 1, m, "VC", True), X, Y)
dev = torch.device("cuda", 0)
d_th = torch.from_numpy(th).to(dev)
d_out = torch.empty(th.size + 5, dtype=torch.float64, device=dev)
st = torch.cuda.ExternalStream(ctx.stream(), device=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 80
ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
time.sleep(float(sys.argv[2]) if len(sys.argv) > 2 else 0.0)
ev[0].record(st)
for i in range(N):
    ctx.eval_dev(d_th.data_ptr(), d_out.data_ptr())
    ev[i + 1].record(st)
ctx.sync()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(N)]
print("per-eval ms:", " ".join(f"{v:.1f}" for v in ms))
print("mean first 10 %.2f, 10-20 %.2f, 20-40 %.2f, last 20 %.2f" % (np.mean(ms[:10]), np.mean(ms[10:20]), np.mean(ms[20:40]), np.mean(ms[-20:])))
q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active", "--format=csv,noheader"], capture_output=True, text=True)
print(q.stdout.strip())
ctx.close()
