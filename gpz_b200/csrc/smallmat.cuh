// Tiny dense d x d helpers (d <= 32) used per basis / per (sample,basis) pair in the covariance
// modes.  Matrices are addressed through an accessor so the same code runs on thread-local arrays
// and on [entry][MP]-strided global arrays (coalesced across threads that walk bases j).
#pragma once
#include "common.cuh"

namespace gpz {

struct LocalMat {            // row-major d x d in a thread-local array
    double* p;
    int d;
    __device__ __forceinline__ double& operator()(int a, int b) const { return p[a * d + b]; }
};
struct StridedMat {          // entry (a,b) at p[(a*d+b)*s]
    double* p;
    int64_t s;
    int d;
    __device__ __forceinline__ double& operator()(int a, int b) const { return p[(static_cast<int64_t>(a) * d + b) * s]; }
};

// in-place Cholesky of the lower triangle; returns false on a non-positive pivot.
// *half_logdet = sum log L_aa  (= 0.5 ln det A)
template <class M>
__device__ inline bool chol_lower(M A, int d, double* half_logdet) {
    double hl = 0.0;
    for (int c = 0; c < d; ++c) {
        double s = A(c, c);
        for (int q = 0; q < c; ++q) s -= A(c, q) * A(c, q);
        if (!(s > 0.0)) return false;
        const double l = sqrt(s);
        hl += log(l);
        A(c, c) = l;
        const double il = 1.0 / l;
        for (int r = c + 1; r < d; ++r) {
            double v = A(r, c);
            for (int q = 0; q < c; ++q) v -= A(r, q) * A(c, q);
            A(r, c) = v * il;
        }
    }
    *half_logdet = hl;
    return true;
}

// in-place inverse of a lower-triangular matrix
template <class M>
__device__ inline void tri_inv_lower(M L, int d) {
    for (int c = 0; c < d; ++c) {
        L(c, c) = 1.0 / L(c, c);
        for (int r = c + 1; r < d; ++r) {
            double s = 0.0;
            for (int q = c; q < r; ++q) s += L(r, q) * L(q, c);
            L(r, c) = -s / L(r, r);      // L(r,r) not yet inverted (r > c)
        }
    }
}
// note: tri_inv_lower walks columns left to right; L(r,r) for r>c is still the original pivot.

// W lower triangular -> full symmetric W' W, in place
template <class M>
__device__ inline void ltl_full(M W, int d) {
    for (int b = 0; b < d; ++b) {
        for (int a = 0; a <= b; ++a) {
            double s = 0.0;
            for (int c = b; c < d; ++c) s += W(c, a) * W(c, b);
            if (a == b) W(a, a) = s; else W(a, b) = s;
        }
    }
    for (int b = 0; b < d; ++b)
        for (int a = 0; a < b; ++a) W(b, a) = W(a, b);
}

// SPD inverse in place (full symmetric result); returns false if not positive definite
template <class M>
__device__ inline bool spd_inv(M A, int d, double* half_logdet) {
    if (!chol_lower(A, d, half_logdet)) return false;
    tri_inv_lower(A, d);
    ltl_full(A, d);
    return true;
}

}  // namespace gpz
