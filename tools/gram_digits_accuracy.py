"""Error of a Gram-like product S = (w*PHI)' PHI (K = n = 16384 rows, real PHI values) for 4..7 base-256 digits, next to the
error of the fp64 BLAS product; reference in 80-bit long double."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpz_b200 import _lib as L, synth
from oracle import gpz_oracle as O

n, d, m = 16384, 10, 128
X, Y = synth.make_data(n, d, seed=0)
th = synth.make_theta0(X, Y, "VC", m, het=True, seed=1)
model = O.Model(d=d, k=1, m=m, method="VC", heteroscedastic=True)
PHI = O.getPHI(np.array(X), None, th, model, None, want_N=False)[0]
w = np.exp(np.random.default_rng(0).standard_normal(n))
A = np.ascontiguousarray((PHI * w[:, None]).T)       # m x n
B = np.ascontiguousarray(PHI.T)
ref = (A.astype(np.longdouble) @ B.astype(np.longdouble).T)
scale = np.max(np.abs(ref))
def err(C): return float(np.max(np.abs(C.astype(np.longdouble) - ref)) / scale)
print("fp64 BLAS          max err / max|S| = %.2e" % err(A @ B.T))
for dg in (4, 5, 6, 7):
    print("digits = %d         max err / max|S| = %.2e" % (dg, err(L.dgemm_nt(A, B, digits=dg))))
