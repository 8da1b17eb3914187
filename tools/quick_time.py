"""Scratch timing script (not the bench): eval time and phase split at a few sizes + fp64 DGEMM peak."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpz_b200 import _lib as L, synth

def dgemm_peak(N=8192, reps=5):
    a = torch.randn(N, N, dtype=torch.float64, device="cuda"); b = torch.randn(N, N, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * N**3 / best / 1e9

if not os.environ.get("NO_PEAK"): print("cuBLAS DGEMM 8192^3 TF/s:", dgemm_peak())
cfgs = [("VD", 100000, 10, 500), ("VC", 100000, 10, 1000), ("VC", 1000000, 10, 1000), ("VD", 1000000, 10, 500)]
if len(sys.argv) > 1: cfgs = [cfgs[int(i)] for i in sys.argv[1].split(",")]
for meth, n, d, m in cfgs:
    X, Y = synth.make_data(n, d, seed=0)
    th = synth.make_theta0(X, Y, meth, m, het=True, seed=1)
    t0 = time.time(); ctx = L.Context(L.make_model(d, 1, m, meth, True), X, Y); t1 = time.time()
    if os.environ.get("GEMM_WARPS"): ctx.set_option("gemm_warps", float(os.environ["GEMM_WARPS"]))
    for opt in ("spare_column", "tensor_phi", "fused_backproj", "ozaki_slices", "ozaki_gram", "ozaki_gram_slices"):
        if os.environ.get(opt.upper()): ctx.set_option(opt, float(os.environ[opt.upper()]))
    f, g, st = ctx.eval(th)
    ts = []
    for _ in range(6):
        t = time.time(); f, g, st = ctx.eval(th); ts.append(time.time() - t)
    print(meth, n, d, m, "create %.2fs" % (t1 - t0), "eval %.1f ms" % (1e3 * min(ts)), "f=%.6f" % f, {k: round(v, 3) for k, v in ctx.last_timing().items()},
          "GEMM TF/s (4nm^2): %.1f" % (4 * n * m * m / min(ts) / 1e12), flush=True)
    ctx.close()
