"""Round-2 kernels under compute-sanitizer: the N = 256 dual-level digit GEMM (m = 300 -> three column tiles), the persistent
TMA-fed PHI kernel, the blocked look-ahead solve, the row-chunked path with NaN-pattern groups, GC + Psi through the digit
GEMMs (K > 128), predictNoisy (cov), the SVD branch of gpz_inv_logdet, and gpz_kernel_timing.  Small sizes: the sanitizer
slows the kernels 10-50 x."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gpz_b200 import _lib as L, synth

# VC m = 300, int8 engine, resident and row-chunked, with a NaN pattern group
n, d, m = 5000, 3, 300
X, Y = synth.make_data(n, d, seed=0)
th = synth.perturb_theta(synth.make_theta0(X, Y, "VC", m, het=True, seed=1), 0.05, 3)
va = np.arange(n) % 5 == 4
gm = L.make_model(d, 1, m, "VC", True)
Xn = X.copy()
Xn[100:400, 1] = np.nan
for chunk in (0, 2048):
    for data in (X, Xn):
        ctx = L.Context(gm, data, Y, training=~va, validation=va)
        ctx.set_option("ozaki_slices", 7)
        if chunk:
            ctx.set_option("chunk_rows", chunk)
        f, g, st = ctx.eval(th)
        f2, g2, _ = ctx.eval(th)
        assert f == f2 and np.array_equal(g, g2)
        nl, w, iS = ctx.fit(th)
        kt = ctx.kernel_timing()
        ctx.close()
        print("VC m=300 chunk", chunk, "nan", data is Xn, f, float(np.abs(g).max()), kt, flush=True)
mu, nu, be, ga, _ = L.predict_core(gm, th, w, iS, X[:200])
Psi = synth.make_psi(40, d, "VC", seed=4)
mu2, nu2, be2, ga2, _ = L.predict_core(gm, th, w, iS, X[:40], Psi)
print("predict", float(nu.mean()), float(nu2.mean()), float(ga2.mean()), flush=True)

# GC + Psi, d = 12 (K = 91 / 13 d^2 > 128 -> the digit-GEMM route of gcpsi.cu)
n, d, m = 600, 12, 40
X, Y = synth.make_data(n, d, seed=2)
th = synth.perturb_theta(synth.make_theta0(X, Y, "GC", m, het=True, seed=1), 0.05, 3)
Psi = synth.make_psi(n, d, "GC", seed=5)
for opt in (1, 0):
    ctx = L.Context(L.make_model(d, 1, m, "GC", True), X, Y, Psi, None, np.ones(n, bool), None)
    ctx.set_option("gc_int8", opt)
    f, g, st = ctx.eval(th)
    print("GC+Psi d=12 gc_int8", opt, f, float(np.abs(g).max()), flush=True)
    ctx.close()

# inv_logdet: Cholesky branch and the SVD (rank-deficient) branch
rng = np.random.default_rng(0)
A = rng.standard_normal((70, 70))
S = A @ A.T + 70 * np.eye(70)
Xi, ld = L.inv_logdet(S)
B = rng.standard_normal((70, 30))
Xi2, ld2 = L.inv_logdet(B @ B.T)
print("inv_logdet", float(np.abs(Xi @ S - np.eye(70)).max()), ld, ld2, flush=True)
