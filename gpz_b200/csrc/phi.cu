// Design-matrix kernels: PHI = exp(lnPHI) for the six covariance modes (GPz/getPHI.m:60-113), the
// fused ln-noise row-dots (getPHI.m:116-125), theta unpacking (getPHI.m:24-40) and Dxy (Dxy.m:3-7).
//
// Layout: X is [d][n] (MATLAB column-major n x d, coalesced over rows), PHI is row-major [n][MP]
// (basis index fastest, zero in the padded columns j >= m), per-basis parameters are [..][MP].
// A warp owns RT rows and sweeps the bases 64 at a time (2 per lane), so PHI is written with
// 16-byte stores, 512 contiguous bytes per warp and row, and the row-dots PHI_i.v / PHI_i.w
// are finished inside the warp with shuffles.
#include "internal.cuh"
#include "smallmat.cuh"

namespace gpz {

constexpr double kLn2 = 0.69314718055994530942;
int g_prep_block = 32;          // threads per CTA of the per-basis theta -> parameter kernels ("prep_block" option: 32, 64 or 128)

// ------------------------------------------------------------------------------------------------
// theta -> per-basis arrays
// ------------------------------------------------------------------------------------------------
// Both kernels below: blockDim = (bases, PREP_TY).  threadIdx.x walks the bases (all arrays are [..][MP], so accesses stay coalesced),
// threadIdx.y splits the ROWS of the per-basis d x d work; what needs a whole matrix (Cholesky, the constant term) is done by the
// y = 0 thread after a block barrier.  One thread per basis did O(d^3) dependent work: 55 + 85 us at m = 1000, d = 10
// (profiles/r02z_small_launches.csv), the largest n-independent part of the 8-GPU step.  Every output element is still produced by
// the same expression in the same order.
constexpr int PREP_TY = 8;

__global__ void prep_kernel(const double* __restrict__ th, Params P, int need_sigma) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int ty = threadIdx.y, TY = blockDim.y;
    const int d = P.d, m = P.m, MP = P.MP, k = P.k, dp = P.dp;
    const bool valid = j < MP;
    const bool live = j < m;
    const double* thG = th + P.oG;
    // centres are stored relative to the dataset shift (X was shifted by the same constant at upload)
    // the centre is also kept locally (covariance modes: d <= 32) or recomputed (diagonal modes: any d) instead of read back
    double pl[32];
    if (valid) {
        for (int a = 0; a < d; ++a) {
            const double pa = live ? th[a * m + j] - P.xshift[a] : 0.0;
            if (a < 32) pl[a] = pa;
            if (ty == 0) P.Pt[a * MP + j] = pa;
        }
    }
    if (valid && !mode_is_cov(P.mode)) {
        for (int a = ty; a < d; a += TY) {
            double gv = 0.0;
            if (live) {
                switch (P.mode) {
                    case GL: gv = thG[0]; break;
                    case VL: gv = thG[j]; break;
                    case GD: gv = thG[a]; break;
                    default: gv = thG[a * m + j]; break;     // VD: Gamma(j,a), column-major m x d
                }
            }
            P.Gt[a * MP + j] = gv;
            const double pa = live ? th[a * m + j] - P.xshift[a] : 0.0;
            P.Ct[a * MP + j] = live ? gv * pa : 0.0;
            if (P.Wc != nullptr) {          // lnPHI = sum_a [ -1/2 g^2 x^2 + g^2 p x - 1/2 g^2 p^2 ]
                P.Wc[static_cast<int64_t>(1 + a) * MP + j] = live ? gv * gv * pa : 0.0;
                P.Wc[static_cast<int64_t>(1 + d + a) * MP + j] = live ? -0.5 * gv * gv : 0.0;
            }
        }
        if (P.Wc != nullptr) {
            if (ty == 0) {
                double c0 = 0.0;
                for (int a = 0; a < d; ++a) {
                    double gv = 0.0;
                    if (live) {
                        switch (P.mode) {
                            case GL: gv = thG[0]; break;
                            case VL: gv = thG[j]; break;
                            case GD: gv = thG[a]; break;
                            default: gv = thG[a * m + j]; break;
                        }
                    }
                    const double gp = live ? gv * (th[a * m + j] - P.xshift[a]) : 0.0;      // = P.Ct[a * MP + j], recomputed instead of read back
                    c0 = fma(gp, gp, c0);
                }
                P.Wc[j] = live ? -0.5 * c0 : 0.0;
            }
            for (int r = 1 + 2 * d + ty; r < P.KQ; r += TY) P.Wc[static_cast<int64_t>(r) * MP + j] = 0.0;
        }
    } else if (valid) {
        const double* G = (P.mode == GC) ? thG : thG + static_cast<int64_t>(d) * d * j;   // Gamma_j(b,a) = G[b + a*d]
        for (int b = ty; b < d; b += TY) {
            double c = 0.0;
            for (int a = 0; a < dp; ++a) {
                const double gv = (live && a < d) ? G[b + a * d] : 0.0;
                P.Gam[(static_cast<int64_t>(b) * dp + a) * MP + j] = gv;
                if (live && a < d) c += gv * pl[a];
            }
            P.Ct[b * MP + j] = c;
        }
        for (int a = ty; a < d; a += TY)
            for (int b = 0; b < d; ++b) {
                double s = 0.0;
                if (live)
                    for (int c = 0; c < d; ++c) s += G[c + a * d] * G[c + b * d];
                P.Aj[(static_cast<int64_t>(a) * d + b) * MP + j] = live ? s : (a == b ? 1.0 : 0.0);
            }
    }
    if (mode_is_cov(P.mode) && need_sigma) {
        __syncthreads();                                  // the rows of Aj written by the other y threads of this block
        if (valid && ty == 0) {
            StridedMat S{P.Sj + j, MP, d};
            StridedMat A{P.Aj + j, MP, d};
            for (int a = 0; a < d; ++a)
                for (int b = 0; b < d; ++b) S(a, b) = A(a, b);
            double hl = 0.0;
            const bool ok = spd_inv(S, d, &hl);
            P.lndS[j] = ok ? -2.0 * hl : nan("");
        }
    }
    if (!valid || ty != 0) return;
    for (int o = 0; o < k; ++o) {
        P.alpha[o * MP + j] = live ? exp(th[P.oA + static_cast<int64_t>(o) * m + j]) : 1.0;
        if (P.het) {
            P.v[o * MP + j] = live ? th[P.oV + static_cast<int64_t>(o) * m + j] : 0.0;
            P.tau[o * MP + j] = live ? exp(th[P.oT + static_cast<int64_t>(o) * m + j]) : 0.0;
        } else {
            P.v[o * MP + j] = 0.0;
            P.tau[o * MP + j] = 0.0;
        }
    }
    if (j < k) P.bk[j] = th[P.oB + j];
}

// cov modes, per missing-input pattern g (observed set o, missing u) and basis j:
//   M  = (Sigma_j(o,o))^-1 = A_oo - A_ou A_uu^-1 A_uo      (A = Gamma_j' Gamma_j, getPHI.m:73-76)
//   G  = A_uu^-1 A_uo                                       (GPz.m:156)
//   W  = coefficients of  -1/2 (x-p)_o' M (x-p)_o - 1/2 |u| ln 2  in the monomials [1, x_a, x_a x_b (a<=b)]
template <int DMAX>
__global__ void __launch_bounds__(1024)
prep_patterns_kernel(Params P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int ty = threadIdx.y, TY = blockDim.y;
    const int g = blockIdx.y;
    const int d = P.d, m = P.m, MP = P.MP;
    const bool valid = j < MP, live = j < m;
    const unsigned char* ob = P.obs + g * d;
    double* Mg = P.Mg + static_cast<int64_t>(g) * d * d * MP;
    double* Gg = P.Gg + static_cast<int64_t>(g) * d * d * MP;
    double* Wg = P.Wc != nullptr ? P.Wc + static_cast<int64_t>(g) * P.KQ * MP : nullptr;
    if (valid && !live) {
        if (ty == 0) P.lndM[static_cast<int64_t>(g) * MP + j] = 0.0;
        for (int e = ty; e < d * d; e += TY) Mg[static_cast<int64_t>(e) * MP + j] = Gg[static_cast<int64_t>(e) * MP + j] = 0.0;
        if (Wg != nullptr)
            for (int r = ty; r < P.KQ; r += TY) Wg[static_cast<int64_t>(r) * MP + j] = 0.0;
    }
    int ui[DMAX];
    int nu = 0;
    for (int a = 0; a < d; ++a)
        if (!ob[a]) ui[nu++] = a;
    double U[DMAX * DMAX];            // A_uu, then its inverse (nu x nu); every y thread forms its own copy (nu is small)
    LocalMat Um{U, nu};
    bool ok = true;
    if (live) {
        for (int r = 0; r < nu; ++r)
            for (int c = 0; c < nu; ++c) Um(r, c) = P.Aj[(static_cast<int64_t>(ui[r]) * d + ui[c]) * MP + j];
        double hl = 0.0;
        if (nu > 0) ok = spd_inv(Um, nu, &hl);
        // G(e,b) for e in u, b in o; zero elsewhere: rows a of G split over y
        for (int a = ty; a < d; a += TY)
            for (int b = 0; b < d; ++b) Gg[(static_cast<int64_t>(a) * d + b) * MP + j] = 0.0;
    }
    __syncthreads();                                       // (zero fill before the rows that other y threads overwrite)
    if (live) {
        for (int r = ty; r < nu; r += TY)
            for (int b = 0; b < d; ++b) {
                if (!ob[b]) continue;
                double s = 0.0;
                for (int c = 0; c < nu; ++c) s += Um(r, c) * P.Aj[(static_cast<int64_t>(ui[c]) * d + b) * MP + j];
                Gg[(static_cast<int64_t>(ui[r]) * d + b) * MP + j] = ok ? s : nan("");
            }
    }
    __syncthreads();                                       // G complete
    // M(a,b) = A(a,b) - sum_{e in u} A(a,e) G(e,b)   for a,b in o: rows a split over y
    if (live) {
        for (int a = ty; a < d; a += TY)
            for (int b = 0; b < d; ++b) {
                double v = 0.0;
                if (ob[a] && ob[b]) {
                    v = P.Aj[(static_cast<int64_t>(a) * d + b) * MP + j];
                    for (int r = 0; r < nu; ++r)
                        v -= P.Aj[(static_cast<int64_t>(a) * d + ui[r]) * MP + j] * Gg[(static_cast<int64_t>(ui[r]) * d + b) * MP + j];
                }
                Mg[(static_cast<int64_t>(a) * d + b) * MP + j] = v;
            }
    }
    __syncthreads();                                       // M complete
    if (!live) return;
    if (ty == 0) {   // ln det M over the observed block (for the normalised densities N)
        int oi[DMAX];
        int no = 0;
        for (int a = 0; a < d; ++a)
            if (ob[a]) oi[no++] = a;
        LocalMat Om{U, no};
        for (int r = 0; r < no; ++r)
            for (int c = 0; c <= r; ++c) Om(r, c) = Mg[(static_cast<int64_t>(oi[r]) * d + oi[c]) * MP + j];
        double hm = 0.0;
        const bool okm = no == 0 || chol_lower(Om, no, &hm);
        P.lndM[static_cast<int64_t>(g) * MP + j] = okm ? 2.0 * hm : nan("");
    }
    if (Wg == nullptr) return;
    double pl[DMAX];
    for (int a = 0; a < d; ++a) pl[a] = P.Pt[a * MP + j];
    // rows a of the coefficient table split over y; the y = 1 thread (y = 0 is busy with the Cholesky) also forms the constant
    for (int a = ty; a < d; a += TY) {
        double bv = 0.0;
        for (int b = 0; b < d; ++b) bv += Mg[(static_cast<int64_t>(a) * d + b) * MP + j] * pl[b];
        Wg[static_cast<int64_t>(1 + a) * MP + j] = bv;
        int idx = 1 + d + a * d - a * (a - 1) / 2;                                   // first entry of row a in the packed a <= b list
        for (int b = a; b < d; ++b, ++idx) {
            const double av = Mg[(static_cast<int64_t>(a) * d + b) * MP + j];
            Wg[static_cast<int64_t>(idx) * MP + j] = (a == b) ? -0.5 * av : -av;
        }
    }
    const int tc = TY > 1 ? 1 : 0;
    if (ty == tc) {
        double c0 = 0.0;
        for (int a = 0; a < d; ++a) {
            double bv = 0.0;
            for (int b = 0; b < d; ++b) bv += Mg[(static_cast<int64_t>(a) * d + b) * MP + j] * pl[b];
            c0 += bv * pl[a];
        }
        Wg[j] = -0.5 * c0 - 0.5 * nu * kLn2;
    }
    for (int r = 1 + d + d * (d + 1) / 2 + ty; r < P.KQ; r += TY) Wg[static_cast<int64_t>(r) * MP + j] = 0.0;
}

int prep_params(const double* d_theta, const Params& P, int need_sigma, cudaStream_t st, int64_t* launches) {
    // 32 bases x PREP_TY row threads per CTA: the per-basis work is latency-bound, so few bases per CTA spread it over many SMs
    const int bt = g_prep_block;
    const dim3 blk(static_cast<unsigned>(bt), PREP_TY);
    prep_kernel<<<static_cast<unsigned>(ceil_div(P.MP, bt)), blk, 0, st>>>(d_theta, P, need_sigma);
    GPZ_KERNEL_CHECK();
    ++*launches;
    if (mode_is_cov(P.mode) && P.Mg != nullptr) {
        dim3 grid(static_cast<unsigned>(ceil_div(P.MP, bt)), static_cast<unsigned>(P.npat));
        if (P.d <= 8) prep_patterns_kernel<8><<<grid, blk, 0, st>>>(P);
        else if (P.d <= 16) prep_patterns_kernel<16><<<grid, blk, 0, st>>>(P);
        else prep_patterns_kernel<32><<<grid, blk, 0, st>>>(P);
        GPZ_KERNEL_CHECK();
        ++*launches;
    }
    return GPZ_OK;
}

// ------------------------------------------------------------------------------------------------
// PHI build
// ------------------------------------------------------------------------------------------------
constexpr int PHI_RT = 8;            // rows per warp
constexpr int PHI_WARPS = 8;
constexpr int PHI_ROWS = PHI_RT * PHI_WARPS;   // rows per CTA

// KIND 0: diag modes, no Psi, no NaN   lnPHI = -1/2 sum_a (g x - g p)^2            getPHI.m:93-98
// KIND 1: diag modes, generic (NaN aware, optional Psi)                             getPHI.m:93-105
// KIND 2: cov modes, no Psi, no NaN    lnPHI = -1/2 |Gamma_j x - Gamma_j p|^2       getPHI.m:73-77
template <int KIND, bool PSI>
__global__ void __launch_bounds__(PHI_WARPS * 32)
phi_kernel(Params P, const double* __restrict__ X, const double* __restrict__ Psi, int64_t n, int64_t r0, int64_t r1,
           double* __restrict__ Phi, DotSpec dots) {
    extern __shared__ __align__(16) double sm[];
    const int d = P.d, dp = P.dp, MP = P.MP, m = P.m;
    double* xs = sm;                          // [PHI_ROWS][dp]   (row-major so x pairs are 16-byte loads)
    double* ps = xs + PHI_ROWS * dp;          // [PHI_ROWS][dp]   Psi rows (KIND 1 with PSI)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t rb = r0 + static_cast<int64_t>(blockIdx.x) * PHI_ROWS;

    for (int e = tid; e < PHI_ROWS * dp; e += blockDim.x) {
        const int a = e / PHI_ROWS, r = e % PHI_ROWS;       // consecutive threads -> consecutive rows (coalesced)
        const int64_t gi = rb + r;
        double xv = 0.0, pv = 0.0;
        if (gi < r1 && a < d) {
            xv = X[a * n + gi];
            if (PSI) pv = Psi[a * n + gi];
        }
        xs[r * dp + a] = xv;
        if (PSI) ps[r * dp + a] = pv;
    }
    __syncthreads();

    const int rw = warp * PHI_RT;
    double dot[PHI_RT][2];
#pragma unroll
    for (int r = 0; r < PHI_RT; ++r) dot[r][0] = dot[r][1] = 0.0;

    for (int jc = 0; jc < MP; jc += 64) {
        const int j = jc + 2 * lane;
        double acc[PHI_RT][2];
#pragma unroll
        for (int r = 0; r < PHI_RT; ++r) acc[r][0] = acc[r][1] = 0.0;

        if (KIND == 0) {
            for (int a = 0; a < d; ++a) {
                const double2 gv = __ldg(reinterpret_cast<const double2*>(P.Gt + a * MP + j));
                const double2 cv = __ldg(reinterpret_cast<const double2*>(P.Ct + a * MP + j));
#pragma unroll
                for (int r = 0; r < PHI_RT; ++r) {
                    const double x = xs[(rw + r) * dp + a];
                    const double y0 = fma(gv.x, x, -cv.x), y1 = fma(gv.y, x, -cv.y);
                    acc[r][0] = fma(y0, y0, acc[r][0]);
                    acc[r][1] = fma(y1, y1, acc[r][1]);
                }
            }
        } else if (KIND == 1) {
            double lg[PHI_RT][2], pr[PHI_RT][2];
#pragma unroll
            for (int r = 0; r < PHI_RT; ++r) { lg[r][0] = lg[r][1] = 0.0; pr[r][0] = pr[r][1] = 1.0; }
            for (int a = 0; a < d; ++a) {
                const double2 gv = __ldg(reinterpret_cast<const double2*>(P.Gt + a * MP + j));
                const double2 pv = __ldg(reinterpret_cast<const double2*>(P.Pt + a * MP + j));
                const double g0 = gv.x * gv.x, g1 = gv.y * gv.y;
#pragma unroll
                for (int r = 0; r < PHI_RT; ++r) {
                    const double x = xs[(rw + r) * dp + a];
                    if (x != x) {                 // missing dim: no distance term, -1/2 ln2 (getPHI.m:97)
                        lg[r][0] += kLn2;
                        lg[r][1] += kLn2;
                        continue;
                    }
                    const double d0 = x - pv.x, d1 = x - pv.y;
                    if (PSI) {
                        const double psi = ps[(rw + r) * dp + a];
                        const double s0 = fma(psi, g0, 1.0), s1 = fma(psi, g1, 1.0);   // 1 + Psi/Sigma
                        acc[r][0] += d0 * d0 * g0 / s0;
                        acc[r][1] += d1 * d1 * g1 / s1;
                        pr[r][0] *= s0;
                        pr[r][1] *= s1;
                    } else {
                        acc[r][0] = fma(d0 * d0, g0, acc[r][0]);
                        acc[r][1] = fma(d1 * d1, g1, acc[r][1]);
                    }
                }
                if (PSI && ((a & 7) == 7 || a == d - 1)) {       // fold the product before it can overflow
#pragma unroll
                    for (int r = 0; r < PHI_RT; ++r) {
                        lg[r][0] += log(pr[r][0]);
                        lg[r][1] += log(pr[r][1]);
                        pr[r][0] = pr[r][1] = 1.0;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < PHI_RT; ++r) { acc[r][0] += lg[r][0]; acc[r][1] += lg[r][1]; }
        } else {
            for (int b = 0; b < d; ++b) {
                const double2 cv = __ldg(reinterpret_cast<const double2*>(P.Ct + b * MP + j));
                double s[PHI_RT][2];
#pragma unroll
                for (int r = 0; r < PHI_RT; ++r) { s[r][0] = -cv.x; s[r][1] = -cv.y; }
                for (int a = 0; a < dp; a += 2) {
                    const double2 ga = __ldg(reinterpret_cast<const double2*>(P.Gam + (static_cast<int64_t>(b) * dp + a) * MP + j));
                    const double2 gb = __ldg(reinterpret_cast<const double2*>(P.Gam + (static_cast<int64_t>(b) * dp + a + 1) * MP + j));
#pragma unroll
                    for (int r = 0; r < PHI_RT; ++r) {
                        const double2 xx = *reinterpret_cast<const double2*>(xs + (rw + r) * dp + a);
                        s[r][0] = fma(ga.x, xx.x, s[r][0]);
                        s[r][1] = fma(ga.y, xx.x, s[r][1]);
                        s[r][0] = fma(gb.x, xx.y, s[r][0]);
                        s[r][1] = fma(gb.y, xx.y, s[r][1]);
                    }
                }
#pragma unroll
                for (int r = 0; r < PHI_RT; ++r) {
                    acc[r][0] = fma(s[r][0], s[r][0], acc[r][0]);
                    acc[r][1] = fma(s[r][1], s[r][1], acc[r][1]);
                }
            }
        }

        double2 v0 = make_double2(0.0, 0.0), v1 = make_double2(0.0, 0.0);
        if (dots.n > 0) v0 = __ldg(reinterpret_cast<const double2*>(dots.vec[0] + j));
        if (dots.n > 1) v1 = __ldg(reinterpret_cast<const double2*>(dots.vec[1] + j));
#pragma unroll
        for (int r = 0; r < PHI_RT; ++r) {
            const int64_t gi = rb + rw + r;
            const double p0 = (j < m) ? exp(-0.5 * acc[r][0]) : 0.0;
            const double p1 = (j + 1 < m) ? exp(-0.5 * acc[r][1]) : 0.0;
            if (Phi != nullptr && gi < r1)
                *reinterpret_cast<double2*>(Phi + (gi - r0) * MP + j) = make_double2(p0, p1);
            dot[r][0] = fma(p0, v0.x, fma(p1, v0.y, dot[r][0]));
            dot[r][1] = fma(p0, v1.x, fma(p1, v1.y, dot[r][1]));
        }
    }
    if (dots.n > 0) {
#pragma unroll
        for (int r = 0; r < PHI_RT; ++r) {
            const double s0 = warp_sum(dot[r][0]);
            const double s1 = warp_sum(dot[r][1]);
            const int64_t gi = rb + rw + r;
            if (lane == 0 && gi < r1) {
                dots.out[0][gi] = s0;
                if (dots.n > 1) dots.out[1][gi] = s1;
            }
        }
    }
}

// cov modes + Psi: one thread per (sample, basis) pair                                getPHI.m:80-88
//   lnPHI = -1/2 D (Psi_i+Sigma_j)^{-1} D' + 1/2 ln|Sigma_j| - 1/2 ln|Psi_i+Sigma_j|
template <int DMAX>
__global__ void __launch_bounds__(128)
phi_cov_psi_kernel(Params P, const double* __restrict__ X, const double* __restrict__ Psi, int64_t n, int64_t r0,
                   int64_t r1, int pat, double* __restrict__ Phi) {
    const int d = P.d, MP = P.MP, m = P.m;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gi = r0 + blockIdx.y;
    if (j >= MP || gi >= r1) return;
    double val = 0.0;
    if (j < m) {
        // rows with missing inputs (getPHI.m:80-88 on the observed dims o): the (o,o) blocks are embedded in d x d with the
        // identity on the missing dims, so the factorisation below needs no gather and ln|S| = ln|S(o,o)|
        const unsigned char* ob = P.obs + pat * d;
        double S[DMAX * DMAX];
        double z[DMAX];
        LocalMat Sm{S, d};
        const double* psi = Psi + gi * d * d;
        int nu = 0;
        for (int a = 0; a < d; ++a) {
            for (int b = 0; b <= a; ++b)
                Sm(a, b) = (ob[a] && ob[b]) ? psi[a + b * d] + P.Sj[(static_cast<int64_t>(a) * d + b) * MP + j] : (a == b ? 1.0 : 0.0);
            z[a] = ob[a] ? X[a * n + gi] - P.Pt[a * MP + j] : 0.0;
            nu += !ob[a];
        }
        double hl = 0.0;
        if (chol_lower(Sm, d, &hl)) {
            double q = 0.0;
            for (int a = 0; a < d; ++a) {          // forward substitution L y = Delta
                double s = z[a];
                for (int b = 0; b < a; ++b) s -= Sm(a, b) * z[b];
                s /= Sm(a, a);
                z[a] = s;
                q += s * s;
            }
            // + 1/2 ln|Sigma_j(o,o)| = - 1/2 ln det of the marginal precision (prep_patterns_kernel)
            val = exp(-0.5 * q - 0.5 * P.lndM[static_cast<int64_t>(pat) * MP + j] - hl - 0.5 * nu * kLn2);
        } else {
            val = nan("");
        }
    }
    if (Phi != nullptr) Phi[(gi - r0) * MP + j] = val;
}

template <int KIND, bool PSI>
static int launch_phi(const Params& P, const RowData& R, int64_t r0, int64_t r1, double* Phi, const DotSpec& dots,
                      cudaStream_t st) {
    const size_t smem = sizeof(double) * PHI_ROWS * P.dp * 2;
    if (smem > 48 * 1024) {
        static PerDeviceOnce once;
        if (once.need()) {
            GPZ_CUDA(cudaFuncSetAttribute(phi_kernel<KIND, PSI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        }
    }
    const int64_t nblk = ceil_div(r1 - r0, PHI_ROWS);
    phi_kernel<KIND, PSI><<<static_cast<unsigned>(nblk), PHI_WARPS * 32, smem, st>>>(P, R.X, R.Psi, R.n, r0, r1, Phi, dots);
    GPZ_KERNEL_CHECK();
    return GPZ_OK;
}

__global__ void sum_parts_kernel(const double* __restrict__ part, int ntn, int64_t stride, int64_t count, double* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (int t = 0; t < ntn; ++t) s += part[static_cast<int64_t>(t) * stride + i];
    out[i] = s;
}

int phi_build(const Params& P, const RowData& R, int64_t r0, int64_t r1, double* Phi, const DotSpec& dots,
              double* dot_scratch, cudaStream_t st, int64_t* launches) {
    if (r1 <= r0) return GPZ_OK;
    int rc = GPZ_OK;
    const bool psi = R.Psi != nullptr;
    if (R.F != nullptr && P.Wc != nullptr && !psi && !R.has_nan) {
        // tensor-core path: PHI = exp(F W), row-dot partials per 128-column tile, then an ordered sum
        const int ntn = P.MP / TILE;
        const size_t ng = R.g_pat.empty() ? 1 : R.g_pat.size();
        for (size_t g = 0; g < ng; ++g) {        // one launch per missing-input pattern group (one group without NaN)
            int64_t s0 = r0, s1 = r1;
            int pat = 0;
            if (!R.g_pat.empty()) {
                s0 = R.g_r0[g] > r0 ? R.g_r0[g] : r0;
                s1 = R.g_r1[g] < r1 ? R.g_r1[g] : r1;
                pat = R.g_pat[g];
            }
            if (s1 <= s0) continue;
            const int64_t rows = s1 - s0;
            double* p0 = dot_scratch;
            double* p1 = dot_scratch + static_cast<int64_t>(ntn) * rows;
            if (R.FD8 != nullptr && P.KQ <= 128)       // quadratic forms on the int8 tensor cores (ozmma.cu), exp in its epilogue
                rc = ozaki_phi(R.FD8 + s0 * R.phi_digits * 128, R.eaF + s0, P.Wc + static_cast<int64_t>(pat) * P.KQ * P.MP, P.KQ, P.MP, P.m,
                               R.phi_digits, rows, R.WD8, R.ebW, Phi != nullptr ? Phi + (s0 - r0) * P.MP : nullptr, dots.n, dots.vec[0],
                               dots.vec[1], p0, p1, rows, R.ycol != nullptr ? R.ycol + s0 : nullptr, R.flag, st, launches);
            else
                rc = phi_gemm(R.F + s0 * P.QP, P.QP, P.KQ, P.q, P.Wc + static_cast<int64_t>(pat) * P.KQ * P.MP, P.MP, P.m, rows,
                              Phi != nullptr ? Phi + (s0 - r0) * P.MP : nullptr, dots.n, dots.vec[0], dots.vec[1], p0, p1, rows,
                              R.ycol != nullptr ? R.ycol + s0 : nullptr, st, launches);
            if (rc) return rc;
            for (int q = 0; q < dots.n; ++q) {
                sum_parts_kernel<<<static_cast<unsigned>(ceil_div(rows, 256)), 256, 0, st>>>(q == 0 ? p0 : p1, ntn, rows, rows, dots.out[q] + s0);
                GPZ_KERNEL_CHECK();
                ++*launches;
            }
        }
        return GPZ_OK;
    }
    if (!mode_is_cov(P.mode)) {
        if (!psi && !R.has_nan) rc = launch_phi<0, false>(P, R, r0, r1, Phi, dots, st);
        else if (!psi) rc = launch_phi<1, false>(P, R, r0, r1, Phi, dots, st);
        else rc = launch_phi<1, true>(P, R, r0, r1, Phi, dots, st);
        ++*launches;
        return rc;
    }
    const bool patterned = R.has_nan || R.g_pat.size() > 1 || (R.g_pat.size() == 1 && R.g_pat[0] != 0);
    if (patterned && !psi) {
        set_error("covariance modes with missing inputs need the tensor-core PHI path (option tensor_phi=1) when there is no Psi");
        return GPZ_ERR_USAGE;
    }
    if (!psi) {
        rc = launch_phi<2, false>(P, R, r0, r1, Phi, dots, st);
        ++*launches;
        return rc;
    }
    if (Phi == nullptr) {
        set_error("phi_build: cov+Psi needs a PHI buffer");
        return GPZ_ERR_USAGE;
    }
    const int64_t rows = r1 - r0;
    if (P.mode == GC && !patterned && R.gcF != nullptr) {
        // one covariance for all bases: Psi_i + Sigma depends on the row only -> per-row features, then PHI = exp(F W) (gcpsi.cu)
        if ((rc = gc_features(P, R, r0, r1, st, launches))) return rc;
        const int KQ = gc_feature_width(P.d), ntn = P.MP / TILE;
        double* p0 = dot_scratch;
        double* p1 = dot_scratch + static_cast<int64_t>(ntn) * rows;
        const int kvalid = 1 + P.d + P.d * (P.d + 1) / 2;
        if (R.gc_digits > 0) {
            // K = 561 at d = 32: the product F W as an error-free digit GEMM on the int8 tensor cores (ozmma.cu, stored directly),
            // exp and the row dots in one memory-bound pass afterwards (an exp epilogue starves under the tensor pipe)
            const int K128 = static_cast<int>(round_up(KQ, 128));
            if ((rc = ozaki_row_digits(R.gcF, KQ, kvalid, K128, rows, R.gc_digits, R.gcA8, R.gcEa, R.flag, st, launches))) return rc;
            if ((rc = ozaki_transpose_digits(R.gcW, P.MP, kvalid, P.m, K128, P.MP, R.gc_digits, R.gcWD8, R.gcEbW, R.flag, st, launches))) return rc;
            if ((rc = ozmma_gemm_rows(R.gcA8, R.gcEa, rows, R.gcWD8, R.gcEbW, P.m, K128, R.gc_digits, Phi, P.MP, st, launches))) return rc;
            DotSpec ds = dots;
            for (int q = 0; q < ds.n; ++q) ds.out[q] = dots.out[q] + r0;
            return exp_rows_inplace(Phi, P.MP, P.m, P.MP, rows, ds, st, launches);
        }
        if ((rc = phi_gemm(R.gcF, KQ, KQ, kvalid, R.gcW, P.MP, P.m, rows, Phi, dots.n, dots.vec[0], dots.vec[1], p0, p1, rows, nullptr, st, launches)))
            return rc;
        for (int q = 0; q < dots.n; ++q) {
            sum_parts_kernel<<<static_cast<unsigned>(ceil_div(rows, 256)), 256, 0, st>>>(q == 0 ? p0 : p1, ntn, rows, rows, dots.out[q] + r0);
            GPZ_KERNEL_CHECK();
            ++*launches;
        }
        return GPZ_OK;
    }
    const size_t ng = R.g_pat.empty() ? 1 : R.g_pat.size();
    for (size_t g = 0; g < ng; ++g) {             // one pass per missing-input pattern group (one group without NaN)
        int64_t s0 = r0, s1 = r1;
        int pat = 0;
        if (!R.g_pat.empty()) {
            s0 = R.g_r0[g] > r0 ? R.g_r0[g] : r0;
            s1 = R.g_r1[g] < r1 ? R.g_r1[g] : r1;
            pat = R.g_pat[g];
        }
        for (int64_t c0 = s0; c0 < s1; c0 += 65535) {     // gridDim.y limit
            const int64_t c1 = (c0 + 65535 < s1) ? c0 + 65535 : s1;
            dim3 grid(static_cast<unsigned>(ceil_div(P.MP, 128)), static_cast<unsigned>(c1 - c0));
            double* out = Phi + (c0 - r0) * P.MP;
            if (P.d <= 8) phi_cov_psi_kernel<8><<<grid, 128, 0, st>>>(P, R.X, R.Psi, R.n, c0, c1, pat, out);
            else if (P.d <= 16) phi_cov_psi_kernel<16><<<grid, 128, 0, st>>>(P, R.X, R.Psi, R.n, c0, c1, pat, out);
            else phi_cov_psi_kernel<32><<<grid, 128, 0, st>>>(P, R.X, R.Psi, R.n, c0, c1, pat, out);
            GPZ_KERNEL_CHECK();
            ++*launches;
        }
    }
    if (dots.n > 0) return rowdot(Phi, P.MP, P.m, rows, DotSpec{dots.n, {dots.vec[0], dots.vec[1]},
                                  {dots.out[0] + r0, dots.n > 1 ? dots.out[1] + r0 : nullptr}}, st, launches);
    return GPZ_OK;
}

// ------------------------------------------------------------------------------------------------
// N_ij = PHI_ij * exp(c_ij),  c_ij = -1/2 ln|Sigma_j(o,o)| - 1/2 |o| ln 2pi + 1/2 |u| ln 2   (o/u = observed/missing dims of row i)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
density_kernel(Params P, const double* __restrict__ X, int64_t n, int64_t r0, int64_t r1, int pat,
               const double* __restrict__ Phi, double* __restrict__ N) {
    const int j = blockIdx.x * 128 + threadIdx.x;
    const int d = P.d, MP = P.MP;
    const int64_t rb = r0 + static_cast<int64_t>(blockIdx.y) * 64;
    const bool cov = mode_is_cov(P.mode);
    const double hl2pi = 0.91893853320467274178;    // 1/2 ln 2pi
    double cj = 0.0;
    int no_pat = d;
    if (cov) {
        no_pat = 0;
        for (int a = 0; a < d; ++a) no_pat += P.obs[pat * d + a];
        cj = 0.5 * P.lndM[static_cast<int64_t>(pat) * MP + j] - hl2pi * no_pat + 0.5 * kLn2 * (d - no_pat);
    }
    for (int r = 0; r < 64; ++r) {
        const int64_t i = rb + r;
        if (i >= r1) break;
        double c = cj;
        if (!cov) {
            int nu = 0;
            for (int a = 0; a < d; ++a) {
                const double x = X[a * n + i];
                if (x != x) { ++nu; continue; }
                c += log(fabs(P.Gt[a * MP + j])) - hl2pi;
            }
            c += 0.5 * kLn2 * nu;
        }
        const int64_t off = (i - r0) * MP + j;
        N[off] = (j < P.m) ? Phi[off] * exp(c) : 0.0;
    }
}

int phi_to_density(const Params& P, const RowData& R, int64_t r0, int64_t r1, const double* Phi, double* N, cudaStream_t st,
                   int64_t* launches) {
    const size_t ng = R.g_pat.empty() ? 1 : R.g_pat.size();
    for (size_t g = 0; g < ng; ++g) {
        int64_t s0 = r0, s1 = r1;
        int pat = 0;
        if (!R.g_pat.empty()) {
            s0 = R.g_r0[g] > r0 ? R.g_r0[g] : r0;
            s1 = R.g_r1[g] < r1 ? R.g_r1[g] : r1;
            pat = R.g_pat[g];
        }
        if (s1 <= s0) continue;
        dim3 grid(static_cast<unsigned>(P.MP / 128), static_cast<unsigned>(ceil_div(s1 - s0, 64)));
        density_kernel<<<grid, 128, 0, st>>>(P, R.X, R.n, s0, s1, pat, Phi + (s0 - r0) * P.MP, N + (s0 - r0) * P.MP);
        GPZ_KERNEL_CHECK();
        ++*launches;
    }
    return GPZ_OK;
}

// ------------------------------------------------------------------------------------------------
// row dots on a stored PHI: out_q[i] = sum_j PHI[i][j] vec_q[j]   (warp per row)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rowdot_kernel(const double* __restrict__ Phi, int64_t ld, int m, int64_t n, DotSpec dots) {
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    double s0 = 0.0, s1 = 0.0;
    const double* row = Phi + i * ld;
    for (int j = 2 * lane; j < m; j += 64) {
        const double2 p = *reinterpret_cast<const double2*>(row + j);      // padded columns are zero
        const double2 a = __ldg(reinterpret_cast<const double2*>(dots.vec[0] + j));
        s0 = fma(p.x, a.x, fma(p.y, a.y, s0));
        if (dots.n > 1) {
            const double2 b = __ldg(reinterpret_cast<const double2*>(dots.vec[1] + j));
            s1 = fma(p.x, b.x, fma(p.y, b.y, s1));
        }
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0) {
        dots.out[0][i] = s0;
        if (dots.n > 1) dots.out[1][i] = s1;
    }
}

int rowdot(const double* Phi, int64_t ld, int m, int64_t n, const DotSpec& dots, cudaStream_t st, int64_t* launches) {
    if (n <= 0 || dots.n <= 0) return GPZ_OK;
    rowdot_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, st>>>(Phi, ld, m, n, dots);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

// ------------------------------------------------------------------------------------------------
// Dxy (Dxy.m:3-7): D = | xx + yy - 2 X Y' |, column-major n x m output
// ------------------------------------------------------------------------------------------------
__global__ void dxy_kernel(const double* __restrict__ X, int64_t n, const double* __restrict__ Y, int m, int d,
                           double* __restrict__ D) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= n) return;
    double xx = 0.0, yy = 0.0, xy = 0.0;
    for (int a = 0; a < d; ++a) {
        const double x = X[a * n + i], y = Y[a * m + j];
        xx = fma(x, x, xx);
        yy = fma(y, y, yy);
        xy = fma(x, y, xy);
    }
    D[static_cast<int64_t>(j) * n + i] = fabs(yy + (xx - 2.0 * xy));
}

int dxy_device(const double* X, int64_t n, const double* Y, int m, int d, double* D, cudaStream_t st) {
    if (n <= 0 || m <= 0) return GPZ_OK;
    dim3 grid(static_cast<unsigned>(ceil_div(n, 256)), static_cast<unsigned>(m));
    dxy_kernel<<<grid, 256, 0, st>>>(X, n, Y, m, d, D);
    GPZ_KERNEL_CHECK();
    return GPZ_OK;
}

// mean(Dxy(X,Y)) (init.m:62) without materialising the n x m matrix: block (c, j) sums |xx + yy - 2xy| over a chunk of
// rows in a fixed order, the finish kernel adds the chunks in order and divides by n.
constexpr int DXM_ROWS = 4096;
__global__ void __launch_bounds__(256) dxy_colsum_kernel(const double* __restrict__ X, int64_t n, const double* __restrict__ Y, int m,
                                                         int d, double* __restrict__ part) {
    __shared__ double sh[8];
    const int j = blockIdx.y;
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * DXM_ROWS, i1 = min(n, i0 + DXM_ROWS);
    double acc = 0.0;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
        double xx = 0.0, yy = 0.0, xy = 0.0;
        for (int a = 0; a < d; ++a) {
            const double x = X[a * n + i], y = Y[a * m + j];
            xx = fma(x, x, xx);
            yy = fma(y, y, yy);
            xy = fma(x, y, xy);
        }
        acc += fabs(yy + (xx - 2.0 * xy));
    }
    const double r = block_sum<256>(acc, sh);
    if (threadIdx.x == 0) part[static_cast<int64_t>(j) * gridDim.x + blockIdx.x] = r;
}
__global__ void dxy_colmean_finish_kernel(const double* __restrict__ part, int m, int nchunk, int64_t n, double* __restrict__ mean) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double s = 0.0;
    for (int c = 0; c < nchunk; ++c) s += part[static_cast<int64_t>(j) * nchunk + c];
    mean[j] = s / static_cast<double>(n);
}
int dxy_colmean_device(const double* X, int64_t n, const double* Y, int m, int d, double* part, double* mean, cudaStream_t st) {
    const int nchunk = static_cast<int>(ceil_div(n, DXM_ROWS));
    dxy_colsum_kernel<<<dim3(nchunk, m), 256, 0, st>>>(X, n, Y, m, d, part);
    GPZ_KERNEL_CHECK();
    dxy_colmean_finish_kernel<<<static_cast<int>(ceil_div(m, 128)), 128, 0, st>>>(part, m, nchunk, n, mean);
    GPZ_KERNEL_CHECK();
    return GPZ_OK;
}
int64_t dxy_colmean_chunks(int64_t n) { return ceil_div(n, DXM_ROWS); }

// row-major [n][ld] -> column-major n x m (MATLAB) through a 32x32 smem tile
__global__ void transpose_kernel(const double* __restrict__ src, int64_t ld, int64_t n, int m, double* __restrict__ dst) {
    __shared__ double tile[32][33];
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * 32;
    const int j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int64_t i = i0 + r;
        const int j = j0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < n && j < m) ? src[i * ld + j] : 0.0;
    }
    __syncthreads();
    for (int c = threadIdx.y; c < 32; c += 8) {
        const int j = j0 + c;
        const int64_t i = i0 + threadIdx.x;
        if (i < n && j < m) dst[static_cast<int64_t>(j) * n + i] = tile[threadIdx.x][c];
    }
}

int transpose_out(const double* src, int64_t ld, int64_t n, int m, double* dst, cudaStream_t st) {
    if (n <= 0 || m <= 0) return GPZ_OK;
    dim3 grid(static_cast<unsigned>(ceil_div(n, 32)), static_cast<unsigned>(ceil_div(m, 32)));
    transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(src, ld, n, m, dst);
    GPZ_KERNEL_CHECK();
    return GPZ_OK;
}

}  // namespace gpz
