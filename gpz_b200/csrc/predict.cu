// predictNoisy for the diagonal modes (GPz/predictDiag.m:75-125): input-noise aware predictive
// variance terms through the m(m+1)/2 basis-pair sum
//   Z_ij(x) = exp(lnZ_ij) * N(x; c_ij, C_ij + Psi),   gamma += f Z w_i w_j,  VlnS += f Z v_i v_j,
//   nu += f Z iSigma_w(i,j)          (f = 2 off the diagonal, 1 on it: predictDiag.m:113-119)
// A pair table (C_ij, c_ij, lnZ_ij and the three weights) is built once per call; the row kernel
// keeps x_i, Psi_i in registers/local memory and streams the table through L1 (every thread of a CTA
// reads the same pair, so the loads are broadcasts).
#include <vector>

#include "internal.cuh"

namespace gpz {

struct PairTab {
    int64_t npairs;
    int d, k;
    double* C;      // [d][npairs]
    double* c;      // [d][npairs]
    double* lnZ;    // [npairs]   (includes -1/2 sum ln(C) so that the row kernel needs one log)
    double* ww;     // [k][npairs]
    double* vv;     // [k][npairs]
    double* ss;     // [k][npairs]
};

__global__ void __launch_bounds__(128)
pair_table_kernel(Params P, const double* __restrict__ w, const double* __restrict__ Sinv, PairTab T) {
    const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= T.npairs) return;
    // q -> (i >= j)
    int64_t i = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(q) + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= q) ++i;
    while (i * (i + 1) / 2 > q) --i;
    const int64_t j = q - i * (i + 1) / 2;
    const int d = P.d, MP = P.MP;
    double lz = 0.0;
    for (int a = 0; a < d; ++a) {
        const double gi = P.Gt[a * MP + i], gj = P.Gt[a * MP + j];
        const double isi = gi * gi, isj = gj * gj;            // iSigma
        const double si = 1.0 / isi, sj = 1.0 / isj;          // Sigma
        const double pi_ = P.Pt[a * MP + i], pj = P.Pt[a * MP + j];
        const double C = 1.0 / (isi + isj);
        T.C[a * T.npairs + q] = C;
        T.c[a * T.npairs + q] = (pi_ * isi + pj * isj) * C;
        const double dl = pi_ - pj;
        lz += -0.5 * log(isi) - 0.5 * log(isj) - 0.5 * dl * dl / (si + sj) - 0.5 * log(si + sj);
    }
    T.lnZ[q] = lz;
    const double f = (i == j) ? 1.0 : 2.0;
    for (int o = 0; o < P.k; ++o) {
        T.ww[o * T.npairs + q] = f * w[o * MP + i] * w[o * MP + j];
        T.vv[o * T.npairs + q] = f * P.v[o * MP + i] * P.v[o * MP + j];
        T.ss[o * T.npairs + q] = f * Sinv[(static_cast<int64_t>(o) * MP + j) * MP + i];   // iSigma_w(i,j), i >= j, as the reference reads it
    }
}

template <int DMAX, int KMAX>
__global__ void __launch_bounds__(128)
predict_noisy_kernel(Params P, const double* __restrict__ X, const double* __restrict__ Psi, int64_t n, PairTab T,
                     const double* __restrict__ ElnS, const double* __restrict__ mu, double* __restrict__ nu,
                     double* __restrict__ beta_i, double* __restrict__ gamma) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int d = P.d, k = P.k;
    double x[DMAX], ps[DMAX];
    const bool live = i < n;
    for (int a = 0; a < d; ++a) {
        x[a] = live ? X[a * n + i] : 0.0;
        ps[a] = live ? Psi[a * n + i] : 1.0;
    }
    double g[KMAX], vl[KMAX], nv[KMAX];
#pragma unroll
    for (int o = 0; o < KMAX; ++o) g[o] = vl[o] = nv[o] = 0.0;
    for (int64_t q = 0; q < T.npairs; ++q) {
        double quad = 0.0, prod = 1.0, lsum = 0.0;
        for (int a = 0; a < d; ++a) {
            const double cp = __ldg(T.C + a * T.npairs + q) + ps[a];
            const double dl = x[a] - __ldg(T.c + a * T.npairs + q);
            quad += dl * dl / cp;
            prod *= cp;
            if ((a & 7) == 7) { lsum += log(prod); prod = 1.0; }
        }
        lsum += log(prod);
        const double Z = exp(__ldg(T.lnZ + q) - 0.5 * quad - 0.5 * lsum);
#pragma unroll
        for (int o = 0; o < KMAX; ++o) {
            if (o < k) {
                g[o] = fma(Z, __ldg(T.ww + o * T.npairs + q), g[o]);
                vl[o] = fma(Z, __ldg(T.vv + o * T.npairs + q), vl[o]);
                nv[o] = fma(Z, __ldg(T.ss + o * T.npairs + q), nv[o]);
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int o = 0; o < KMAX; ++o) {
        if (o < k) {
            const double e = ElnS[o * n + i];
            const double m_ = mu[o * n + i];
            const double dv = e - P.bk[o];
            const double V = vl[o] - dv * dv;                    // predictDiag.m:122
            gamma[o * n + i] = g[o] - m_ * m_;                   // :123
            beta_i[o * n + i] = exp(e) * (1.0 + 0.5 * V);        // :124
            nu[o * n + i] = nv[o];
        }
    }
}

int predict_noisy_diag(const Params& P, const RowData& R, const double* w, const double* Sinv, const double* ElnS,
                       const double* mu, double* nu, double* beta_i, double* gamma, cudaStream_t st, int64_t* launches) {
    if (P.k > 4) {
        set_error("predictNoisy: k > 4 outputs not supported");
        return GPZ_ERR_USAGE;
    }
    if (P.d > 32) {
        set_error("predictNoisy: d > 32 not supported");
        return GPZ_ERR_USAGE;
    }
    PairTab T;
    T.npairs = static_cast<int64_t>(P.m) * (P.m + 1) / 2;
    T.d = P.d;
    T.k = P.k;
    double* buf = nullptr;
    const int64_t per = 2LL * P.d + 1 + 3LL * P.k;
    GPZ_CUDA(cudaMalloc(&buf, sizeof(double) * per * T.npairs));
    T.C = buf;
    T.c = T.C + static_cast<int64_t>(P.d) * T.npairs;
    T.lnZ = T.c + static_cast<int64_t>(P.d) * T.npairs;
    T.ww = T.lnZ + T.npairs;
    T.vv = T.ww + static_cast<int64_t>(P.k) * T.npairs;
    T.ss = T.vv + static_cast<int64_t>(P.k) * T.npairs;
    pair_table_kernel<<<static_cast<unsigned>(ceil_div(T.npairs, 128)), 128, 0, st>>>(P, w, Sinv, T);
    ++*launches;
    const unsigned nb = static_cast<unsigned>(ceil_div(R.n, 128));
    if (P.d <= 8) predict_noisy_kernel<8, 4><<<nb, 128, 0, st>>>(P, R.X, R.Psi, R.n, T, ElnS, mu, nu, beta_i, gamma);
    else if (P.d <= 16) predict_noisy_kernel<16, 4><<<nb, 128, 0, st>>>(P, R.X, R.Psi, R.n, T, ElnS, mu, nu, beta_i, gamma);
    else predict_noisy_kernel<32, 4><<<nb, 128, 0, st>>>(P, R.X, R.Psi, R.n, T, ElnS, mu, nu, beta_i, gamma);
    ++*launches;
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(buf);
    if (e != cudaSuccess) {
        set_error("predict_noisy_diag: %s", cudaGetErrorString(e));
        return GPZ_ERR_CUDA;
    }
    return GPZ_OK;
}

}  // namespace gpz

// ================================================================================================
// predictMissing / predictNoisyMissing for the diagonal modes (GPz/predictDiag.m:127-295): rows of ONE
// missing-input pattern (observed set o, missing u).  Expected basis under the mixture prior over bases,
// then the basis-pair sum with the missing dims integrated out analytically.
//   No_il   = N(x_o; p_l(o), Sigma_l(o) [+Psi_i])              Pio = No .* prior / rowsum      (:145-155)
//   Nmat_lj = N(p_l(u); p_j(u), Sigma_l(u)+Sigma_j(u))         PHI = No .* (Pio Nmat') e^{lnz} (:161-164)
//   pair (a>=b): C, c (:177-178), NU_l = N(p_l(u); c(u), Sigma_l(u)+C(u)) (:183-184), Q = Pio NU (GEMM),
//   Z = e^{lnZ_ab} N(x_o; c(o), C(o)[+Psi_i]) Q_i,ab           gamma, VlnS, nu += f Z {w w, v v, iSigma}  (:186-199)
// ================================================================================================
namespace gpz {

__global__ void __launch_bounds__(128)
pm_no_kernel(Params P, const double* __restrict__ X, const double* __restrict__ Psi, int64_t n, const unsigned char* __restrict__ ob,
             const double* __restrict__ prior, double* __restrict__ No, double* __restrict__ Pio) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = P.d, m = P.m, MP = P.MP;
    double ey = 0.0;
    for (int l = 0; l < MP; ++l) {
        double val = 0.0;
        if (l < m) {
            double q = 0.0, lp = 0.0;
            for (int a = 0; a < d; ++a) {
                if (!ob[a]) continue;
                const double g = P.Gt[a * MP + l];
                const double s = 1.0 / (g * g) + (Psi ? Psi[a * n + i] : 0.0);
                const double dl = X[a * n + i] - P.Pt[a * MP + l];
                q += dl * dl / s;
                lp += log(s);
            }
            val = exp(-0.5 * q - 0.5 * lp);
        }
        No[i * MP + l] = val;
        const double ex = val * (l < m ? prior[l] : 0.0);
        Pio[i * MP + l] = ex;
        ey += ex;
    }
    for (int l = 0; l < MP; ++l) Pio[i * MP + l] /= ey;
}

__global__ void __launch_bounds__(128)
pm_nmat_kernel(Params P, const unsigned char* __restrict__ ob, double* __restrict__ Nmat) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    const int d = P.d, m = P.m, MP = P.MP;
    if (l >= MP) return;
    double val = 0.0;
    if (l < m && j < m) {
        double q = 0.0, lp = 0.0;
        for (int a = 0; a < d; ++a) {
            if (ob[a]) continue;
            const double gl = P.Gt[a * MP + l], gj = P.Gt[a * MP + j];
            const double s = 1.0 / (gl * gl) + 1.0 / (gj * gj);
            const double dl = P.Pt[a * MP + l] - P.Pt[a * MP + j];
            q += dl * dl / s;
            lp += log(s);
        }
        val = exp(-0.5 * q - 0.5 * lp);
    }
    Nmat[static_cast<int64_t>(j) * MP + l] = val;
}

// PHI = No .* T .* e^{lnz}  (T = Pio Nmat')
__global__ void pm_phi_kernel(Params P, int64_t n, const double* __restrict__ No, const double* __restrict__ T, double* __restrict__ Phi) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int MP = P.MP;
    if (e >= n * MP) return;
    const int l = static_cast<int>(e % MP);
    double v = 0.0;
    if (l < P.m) {
        double lnz = 0.0;
        for (int a = 0; a < P.d; ++a) lnz -= log(fabs(P.Gt[a * MP + l]));
        v = No[e] * T[e] * exp(lnz);
    }
    Phi[e] = v;
}

struct MPairTab {
    int64_t npairs;
    double* C;      // [d][npairs]
    double* c;      // [d][npairs]
    double* cst;    // [npairs]
    double* ww;     // [k][npairs]
    double* vv;
    double* ss;
    double* NU;     // [MP][npairs]  (row l, column pair)
};

__global__ void __launch_bounds__(128)
pm_pairs_kernel(Params P, const unsigned char* __restrict__ ob, const double* __restrict__ w, const double* __restrict__ Sinv,
                MPairTab T) {
    const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= T.npairs) return;
    int64_t i = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(q) + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= q) ++i;
    while (i * (i + 1) / 2 > q) --i;
    const int64_t j = q - i * (i + 1) / 2;
    const int d = P.d, m = P.m, MP = P.MP;
    double lz = 0.0;
    for (int a = 0; a < d; ++a) {
        const double gi = P.Gt[a * MP + i], gj = P.Gt[a * MP + j];
        const double isi = gi * gi, isj = gj * gj, si = 1.0 / isi, sj = 1.0 / isj;
        const double C = 1.0 / (isi + isj);
        T.C[a * T.npairs + q] = C;
        T.c[a * T.npairs + q] = (P.Pt[a * MP + i] * isi + P.Pt[a * MP + j] * isj) * C;
        const double dl = P.Pt[a * MP + i] - P.Pt[a * MP + j];
        lz += -0.5 * log(isi) - 0.5 * log(isj) - 0.5 * dl * dl / (si + sj) - 0.5 * log(si + sj);
    }
    T.cst[q] = lz;
    const double f = (i == j) ? 1.0 : 2.0;
    for (int o = 0; o < P.k; ++o) {
        T.ww[o * T.npairs + q] = f * w[o * MP + i] * w[o * MP + j];
        T.vv[o * T.npairs + q] = f * P.v[o * MP + i] * P.v[o * MP + j];
        T.ss[o * T.npairs + q] = f * Sinv[(static_cast<int64_t>(o) * MP + j) * MP + i];   // iSigma_w(i,j), i >= j, as the reference reads it
    }
    for (int l = 0; l < MP; ++l) {
        double val = 0.0;
        if (l < m) {
            double qd = 0.0, lp = 0.0;
            for (int a = 0; a < d; ++a) {
                if (ob[a]) continue;
                const double g = P.Gt[a * MP + l];
                const double s = 1.0 / (g * g) + T.C[a * T.npairs + q];
                const double dl = P.Pt[a * MP + l] - T.c[a * T.npairs + q];
                qd += dl * dl / s;
                lp += log(s);
            }
            val = exp(-0.5 * qd - 0.5 * lp);
        }
        T.NU[static_cast<int64_t>(l) * T.npairs + q] = val;
    }
}

template <int KMAX>
__global__ void __launch_bounds__(128)
pm_rows_kernel(Params P, const double* __restrict__ X, const double* __restrict__ Psi, int64_t n, int64_t r0, int64_t r1,
               const unsigned char* __restrict__ ob, MPairTab T, const double* __restrict__ Q /*[rows][npairs]*/,
               const double* __restrict__ elns0 /*[k][n] = PHI v*/, const double* __restrict__ mu, double* __restrict__ nu,
               double* __restrict__ beta_i, double* __restrict__ gamma) {
    const int64_t i = r0 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= r1) return;
    const int d = P.d, k = P.k;
    double g[KMAX], vl[KMAX], nv[KMAX];
#pragma unroll
    for (int o = 0; o < KMAX; ++o) g[o] = vl[o] = nv[o] = 0.0;
    const double* Qi = Q + (i - r0) * T.npairs;
    for (int64_t q = 0; q < T.npairs; ++q) {
        double quad = 0.0, lp = 0.0;
        for (int a = 0; a < d; ++a) {
            if (!ob[a]) continue;
            const double cp = __ldg(T.C + a * T.npairs + q) + (Psi ? Psi[a * n + i] : 0.0);
            const double dl = X[a * n + i] - __ldg(T.c + a * T.npairs + q);
            quad += dl * dl / cp;
            lp += log(cp);
        }
        const double Z = exp(__ldg(T.cst + q) - 0.5 * quad - 0.5 * lp) * Qi[q];
#pragma unroll
        for (int o = 0; o < KMAX; ++o) {
            if (o < k) {
                g[o] = fma(Z, __ldg(T.ww + o * T.npairs + q), g[o]);
                vl[o] = fma(Z, __ldg(T.vv + o * T.npairs + q), vl[o]);
                nv[o] = fma(Z, __ldg(T.ss + o * T.npairs + q), nv[o]);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < KMAX; ++o) {
        if (o < k) {
            const double e0 = elns0[o * n + i];
            const double m_ = mu[o * n + i];
            const double V = vl[o] - e0 * e0;                              // predictDiag.m:204
            gamma[o * n + i] = g[o] - m_ * m_;                             // :210
            beta_i[o * n + i] = exp(e0 + P.bk[o]) * (1.0 + 0.5 * V);       // :206-208
            nu[o * n + i] = nv[o];
        }
    }
}

// X, Psi: [d][n] device (shifted like P.Pt); ob: [d] device; prior: [MP] device; outputs [k][n] device, Phi [n][MP] device
int predict_missing_diag(const Params& P, const double* X, const double* Psi, int64_t n, const unsigned char* ob,
                         const double* prior, const double* w, const double* Sinv, double* mu, double* nu, double* beta_i,
                         double* gamma, double* Phi, cudaStream_t st, int64_t* launches) {
    if (P.k > 4) {
        set_error("predictMissing: k > 4 outputs not supported");
        return GPZ_ERR_USAGE;
    }
    const int64_t MP = P.MP;
    const int64_t npairs = static_cast<int64_t>(P.m) * (P.m + 1) / 2;
    int64_t chunk = static_cast<int64_t>(1.0e9 / (8.0 * static_cast<double>(npairs)));
    if (chunk < 1) chunk = 1;
    if (chunk > n) chunk = n;
    double *No = nullptr, *Pio = nullptr, *Nmat = nullptr, *Tm = nullptr, *tab = nullptr, *Q = nullptr, *elns0 = nullptr;
    std::vector<double*> bufs;
    auto A = [&](double** p, int64_t cnt) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), sizeof(double) * static_cast<size_t>(cnt > 0 ? cnt : 1));
        if (e != cudaSuccess) {
            set_error("predict_missing_diag: cudaMalloc: %s", cudaGetErrorString(e));
            return static_cast<int>(GPZ_ERR_CUDA);
        }
        bufs.push_back(*p);
        return static_cast<int>(GPZ_OK);
    };
    auto done = [&](int rc) {
        cudaStreamSynchronize(st);
        for (double* b : bufs) cudaFree(b);
        return rc;
    };
    int rc;
    const int64_t per = 2LL * P.d + 1 + 3LL * P.k + MP;
    if ((rc = A(&No, n * MP)) || (rc = A(&Pio, n * MP)) || (rc = A(&Nmat, MP * MP)) || (rc = A(&Tm, n * MP)) ||
        (rc = A(&tab, per * npairs)) || (rc = A(&Q, chunk * npairs)) || (rc = A(&elns0, P.k * n)))
        return done(rc);
    MPairTab T;
    T.npairs = npairs;
    T.C = tab;
    T.c = T.C + static_cast<int64_t>(P.d) * npairs;
    T.cst = T.c + static_cast<int64_t>(P.d) * npairs;
    T.ww = T.cst + npairs;
    T.vv = T.ww + static_cast<int64_t>(P.k) * npairs;
    T.ss = T.vv + static_cast<int64_t>(P.k) * npairs;
    T.NU = T.ss + static_cast<int64_t>(P.k) * npairs;
    pm_no_kernel<<<static_cast<unsigned>(ceil_div(n, 128)), 128, 0, st>>>(P, X, Psi, n, ob, prior, No, Pio);
    pm_nmat_kernel<<<dim3(static_cast<unsigned>(MP / 128), static_cast<unsigned>(MP)), 128, 0, st>>>(P, ob, Nmat);
    *launches += 2;
    // T = Pio * Nmat'   (B(k=j, col=l) = Nmat[l][j]; Nmat is stored [j][l] so B(k,col) = Nmat_store[k*MP + col] transposed twice = symmetric)
    if ((rc = sgemm(static_cast<int>(n), P.m, P.m, 1.0, Pio, MP, 1, Nmat, MP, 1, 0.0, Tm, MP, 0, st, launches))) return done(rc);
    pm_phi_kernel<<<static_cast<unsigned>(ceil_div(n * MP, 256)), 256, 0, st>>>(P, n, No, Tm, Phi);
    ++*launches;
    for (int o = 0; o < P.k; ++o)
        if ((rc = rowdot(Phi, MP, P.m, n, DotSpec{2, {w + o * MP, P.v + o * MP}, {mu + o * n, elns0 + o * n}}, st, launches))) return done(rc);
    pm_pairs_kernel<<<static_cast<unsigned>(ceil_div(npairs, 128)), 128, 0, st>>>(P, ob, w, Sinv, T);
    ++*launches;
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t r1 = (r0 + chunk < n) ? r0 + chunk : n;
        // Q(rows x npairs) = Pio[rows] (rows x m) * NU (m x npairs)
        if ((rc = sgemm(static_cast<int>(r1 - r0), static_cast<int>(npairs), P.m, 1.0, Pio + r0 * MP, MP, 1, T.NU, npairs, 1, 0.0, Q,
                        npairs, 0, st, launches))) return done(rc);
        pm_rows_kernel<4><<<static_cast<unsigned>(ceil_div(r1 - r0, 128)), 128, 0, st>>>(P, X, Psi, n, r0, r1, ob, T, Q, elns0, mu, nu,
                                                                                        beta_i, gamma);
        ++*launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("predict_missing_diag: %s", cudaGetErrorString(e));
        return done(GPZ_ERR_CUDA);
    }
    return done(GPZ_OK);
}

}  // namespace gpz

// ================================================================================================
// predictNoisy for the covariance modes (GPz/predictCov.m:70-133): per basis pair (a >= b)
//   iC = iSigma_a + iSigma_b, C = iC^-1, c = (p_a iSigma_a + p_b iSigma_b) C,
//   lnZ = lnz_a + lnz_b - 1/2 dp (Sigma_a+Sigma_b)^-1 dp' - 1/2 ln|Sigma_a+Sigma_b|            (:101-107)
// and per (sample, pair) a d x d Cholesky of C + Psi_i for N(x_i; c, C + Psi_i)                   (:109-113)
// ================================================================================================
#include "smallmat.cuh"

namespace gpz {

struct CPairTab {
    int64_t npairs;
    double* C;      // [d*d][npairs]
    double* c;      // [d][npairs]
    double* lnZ;    // [npairs]
    double* ww;     // [k][npairs]
    double* vv;
    double* ss;
};

template <int DMAX>
__global__ void __launch_bounds__(64)
cpair_table_kernel(Params P, const double* __restrict__ w, const double* __restrict__ Sinv, CPairTab T) {
    const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= T.npairs) return;
    int64_t i = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(q) + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= q) ++i;
    while (i * (i + 1) / 2 > q) --i;
    const int64_t j = q - i * (i + 1) / 2;
    const int d = P.d, MP = P.MP;
    double M[DMAX * DMAX], dp[DMAX], t[DMAX];
    LocalMat Mm{M, d};
    // C = (iSigma_i + iSigma_j)^-1
    for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b)
            Mm(a, b) = P.Aj[(static_cast<int64_t>(a) * d + b) * MP + i] + P.Aj[(static_cast<int64_t>(a) * d + b) * MP + j];
    double hl = 0.0;
    bool ok = spd_inv(Mm, d, &hl);
    for (int a = 0; a < d; ++a) {
        double s = 0.0;                       // t = p_i iSigma_i + p_j iSigma_j   (row vector)
        for (int b = 0; b < d; ++b)
            s += P.Pt[b * MP + i] * P.Aj[(static_cast<int64_t>(b) * d + a) * MP + i] + P.Pt[b * MP + j] * P.Aj[(static_cast<int64_t>(b) * d + a) * MP + j];
        t[a] = s;
    }
    for (int a = 0; a < d; ++a) {
        double s = 0.0;
        for (int b = 0; b < d; ++b) {
            s += t[b] * Mm(b, a);
            T.C[(static_cast<int64_t>(a) * d + b) * T.npairs + q] = Mm(a, b);
        }
        T.c[a * T.npairs + q] = s;
    }
    // lnZ: Cholesky of Sigma_i + Sigma_j
    for (int a = 0; a < d; ++a) {
        for (int b = 0; b <= a; ++b)
            Mm(a, b) = P.Sj[(static_cast<int64_t>(a) * d + b) * MP + i] + P.Sj[(static_cast<int64_t>(a) * d + b) * MP + j];
        dp[a] = P.Pt[a * MP + i] - P.Pt[a * MP + j];
    }
    double hs = 0.0, quad = 0.0;
    ok = chol_lower(Mm, d, &hs) && ok;
    for (int a = 0; a < d; ++a) {
        double s = dp[a];
        for (int b = 0; b < a; ++b) s -= Mm(a, b) * dp[b];
        s /= Mm(a, a);
        dp[a] = s;
        quad += s * s;
    }
    T.lnZ[q] = ok ? 0.5 * P.lndS[i] + 0.5 * P.lndS[j] - 0.5 * quad - hs : nan("");
    const double f = (i == j) ? 1.0 : 2.0;
    for (int o = 0; o < P.k; ++o) {
        T.ww[o * T.npairs + q] = f * w[o * MP + i] * w[o * MP + j];
        T.vv[o * T.npairs + q] = f * P.v[o * MP + i] * P.v[o * MP + j];
        T.ss[o * T.npairs + q] = f * Sinv[(static_cast<int64_t>(o) * MP + j) * MP + i];   // iSigma_w(i,j), i >= j, as the reference reads it
    }
}

template <int DMAX, int KMAX>
__global__ void __launch_bounds__(64)
predict_noisy_cov_kernel(Params P, const double* __restrict__ X, const double* __restrict__ Psi, int64_t n, CPairTab T,
                         const double* __restrict__ ElnS, const double* __restrict__ mu, double* __restrict__ nu,
                         double* __restrict__ beta_i, double* __restrict__ gamma) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = P.d, k = P.k;
    double x[DMAX], S[DMAX * DMAX], z[DMAX];
    LocalMat Sm{S, d};
    for (int a = 0; a < d; ++a) x[a] = X[a * n + i];
    const double* psi = Psi + i * d * d;
    double g[KMAX], vl[KMAX], nv[KMAX];
#pragma unroll
    for (int o = 0; o < KMAX; ++o) g[o] = vl[o] = nv[o] = 0.0;
    for (int64_t q = 0; q < T.npairs; ++q) {
        for (int a = 0; a < d; ++a) {
            for (int b = 0; b <= a; ++b) Sm(a, b) = psi[a + b * d] + __ldg(T.C + (static_cast<int64_t>(a) * d + b) * T.npairs + q);
            z[a] = x[a] - __ldg(T.c + a * T.npairs + q);
        }
        double hl = 0.0, quad = 0.0;
        if (!chol_lower(Sm, d, &hl)) {
            g[0] = nan("");
            continue;
        }
        for (int a = 0; a < d; ++a) {
            double s = z[a];
            for (int b = 0; b < a; ++b) s -= Sm(a, b) * z[b];
            s /= Sm(a, a);
            z[a] = s;
            quad += s * s;
        }
        const double Z = exp(__ldg(T.lnZ + q) - 0.5 * quad - hl);
#pragma unroll
        for (int o = 0; o < KMAX; ++o) {
            if (o < k) {
                g[o] = fma(Z, __ldg(T.ww + o * T.npairs + q), g[o]);
                vl[o] = fma(Z, __ldg(T.vv + o * T.npairs + q), vl[o]);
                nv[o] = fma(Z, __ldg(T.ss + o * T.npairs + q), nv[o]);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < KMAX; ++o) {
        if (o < k) {
            const double e = ElnS[o * n + i];
            const double m_ = mu[o * n + i];
            const double dv = e - P.bk[o];
            const double V = vl[o] - dv * dv;                    // predictCov.m:131
            gamma[o * n + i] = g[o] - m_ * m_;                   // :132
            beta_i[o * n + i] = exp(e) * (1.0 + 0.5 * V);        // :133
            nu[o * n + i] = nv[o];
        }
    }
}

// needs P.Sj / P.lndS (prep_params with need_sigma = 1); X shifted like P.Pt; Psi [n][d*d]
int predict_noisy_cov(const Params& P, const RowData& R, const double* w, const double* Sinv, const double* ElnS,
                      const double* mu, double* nu, double* beta_i, double* gamma, cudaStream_t st, int64_t* launches) {
    if (P.k > 4 || P.d > 32) {
        set_error("predictNoisy (cov): k <= 4 and d <= 32 supported");
        return GPZ_ERR_USAGE;
    }
    CPairTab T;
    T.npairs = static_cast<int64_t>(P.m) * (P.m + 1) / 2;
    double* buf = nullptr;
    const int64_t per = static_cast<int64_t>(P.d) * P.d + P.d + 1 + 3LL * P.k;
    GPZ_CUDA(cudaMalloc(&buf, sizeof(double) * per * T.npairs));
    T.C = buf;
    T.c = T.C + static_cast<int64_t>(P.d) * P.d * T.npairs;
    T.lnZ = T.c + static_cast<int64_t>(P.d) * T.npairs;
    T.ww = T.lnZ + T.npairs;
    T.vv = T.ww + static_cast<int64_t>(P.k) * T.npairs;
    T.ss = T.vv + static_cast<int64_t>(P.k) * T.npairs;
    const unsigned nbp = static_cast<unsigned>(ceil_div(T.npairs, 64)), nbr = static_cast<unsigned>(ceil_div(R.n, 64));
    if (P.d <= 8) {
        cpair_table_kernel<8><<<nbp, 64, 0, st>>>(P, w, Sinv, T);
        predict_noisy_cov_kernel<8, 4><<<nbr, 64, 0, st>>>(P, R.X, R.Psi, R.n, T, ElnS, mu, nu, beta_i, gamma);
    } else if (P.d <= 16) {
        cpair_table_kernel<16><<<nbp, 64, 0, st>>>(P, w, Sinv, T);
        predict_noisy_cov_kernel<16, 4><<<nbr, 64, 0, st>>>(P, R.X, R.Psi, R.n, T, ElnS, mu, nu, beta_i, gamma);
    } else {
        cpair_table_kernel<32><<<nbp, 64, 0, st>>>(P, w, Sinv, T);
        predict_noisy_cov_kernel<32, 4><<<nbr, 64, 0, st>>>(P, R.X, R.Psi, R.n, T, ElnS, mu, nu, beta_i, gamma);
    }
    *launches += 2;
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(buf);
    if (e != cudaSuccess) {
        set_error("predict_noisy_cov: %s", cudaGetErrorString(e));
        return GPZ_ERR_CUDA;
    }
    return GPZ_OK;
}

// ================================================================================================
// predictMissing / predictNoisyMissing for the covariance modes (GPz/predictCov.m:134-336), one group of rows sharing a
// missing-input pattern (o observed, u missing).  Per basis l: R_l = Sigma_l(o,o)^-1 Sigma_l(o,u), the Schur complement
// Sigma_l(u,u) - Sigma_l(u,o) R_l; per (sample t, basis l): responsibility Pio, completed input X_hat, its covariance
// Psi_hat = T Psi_oo T' + Schur (T = [I; R']); then PHI_ti = e^{lnz_i} sum_j N(X_hat_tj; p_i, Sigma_i + Psi_hat_tj) Pio_tj and,
// per basis pair, Z = e^{lnZ_ij} sum_l N(X_hat_tl; c_ij, C_ij + Psi_hat_tl) Pio_tl.  All (o,o)/(u,u) blocks are embedded
// in d x d arrays.  Cost O(n m^3 d^3): correct, sized for the reference's use (demo-scale m), not optimised.
// ================================================================================================
template <int DMAX>
__global__ void __launch_bounds__(64)
pmc_basis_kernel(Params P, const unsigned char* __restrict__ ob, double* __restrict__ Rt, double* __restrict__ Sch) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = P.d, MP = P.MP;
    if (l >= P.m) return;
    double S[DMAX * DMAX];
    LocalMat Sm{S, d};
    for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b)
            Sm(a, b) = (ob[a] && ob[b]) ? P.Sj[(static_cast<int64_t>(a) * d + b) * MP + l] : (a == b ? 1.0 : 0.0);
    double hl;
    const bool ok = spd_inv(Sm, d, &hl);
    for (int a = 0; a < d; ++a)
        for (int e = 0; e < d; ++e) {
            double r = 0.0;
            if (ob[a] && !ob[e])
                for (int b = 0; b < d; ++b)
                    if (ob[b]) r += Sm(a, b) * P.Sj[(static_cast<int64_t>(b) * d + e) * MP + l];
            Rt[(static_cast<int64_t>(a) * d + e) * MP + l] = ok ? r : nan("");
        }
    for (int e = 0; e < d; ++e)
        for (int f = 0; f < d; ++f) {
            double v = 0.0;
            if (!ob[e] && !ob[f]) {
                v = P.Sj[(static_cast<int64_t>(e) * d + f) * MP + l];
                for (int a = 0; a < d; ++a)
                    if (ob[a]) v -= P.Sj[(static_cast<int64_t>(e) * d + a) * MP + l] * Rt[(static_cast<int64_t>(a) * d + f) * MP + l];
            }
            Sch[(static_cast<int64_t>(e) * d + f) * MP + l] = v;
        }
}

// Ex [rows][MP] (un-normalised responsibilities), XH [rows][MP][d], PH [rows][MP][d*d] (only with Psi)
template <int DMAX>
__global__ void __launch_bounds__(128)
pmc_rows_kernel(Params P, const double* __restrict__ X, const double* __restrict__ Psi, int64_t n, int64_t r0, int64_t r1,
                const unsigned char* __restrict__ ob, const double* __restrict__ prior, const double* __restrict__ Rt,
                const double* __restrict__ Sch, double* __restrict__ Ex, double* __restrict__ XH, double* __restrict__ PH) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t t = r0 + blockIdx.y;
    const int d = P.d, MP = P.MP;
    if (l >= MP || t >= r1) return;
    const int64_t tl = (t - r0) * MP + l;
    if (l >= P.m) {
        Ex[tl] = 0.0;
        return;
    }
    double S[DMAX * DMAX], dl[DMAX], z[DMAX];
    LocalMat Sm{S, d};
    const double* psi = Psi != nullptr ? Psi + t * d * d : nullptr;
    for (int a = 0; a < d; ++a) {
        for (int b = 0; b <= a; ++b)
            Sm(a, b) = (ob[a] && ob[b]) ? P.Sj[(static_cast<int64_t>(a) * d + b) * MP + l] + (psi != nullptr ? psi[a + b * d] : 0.0)
                                        : (a == b ? 1.0 : 0.0);
        dl[a] = ob[a] ? X[a * n + t] - P.Pt[a * MP + l] : 0.0;
        z[a] = dl[a];
    }
    double hl = 0.0, q = 0.0;
    const bool ok = chol_lower(Sm, d, &hl);
    for (int a = 0; a < d; ++a) {
        double s = z[a];
        for (int b = 0; b < a; ++b) s -= Sm(a, b) * z[b];
        s /= Sm(a, a);
        z[a] = s;
        q += s * s;
    }
    Ex[tl] = ok ? exp(-0.5 * q - hl) * prior[l] : nan("");                         // predictCov.m:163 / :267
    double* xh = XH + tl * d;
    for (int a = 0; a < d; ++a) {
        double v;
        if (ob[a]) v = X[a * n + t];
        else {
            v = P.Pt[a * MP + l];                                                   // :169-170 / :277-278
            for (int b = 0; b < d; ++b)
                if (ob[b]) v += dl[b] * Rt[(static_cast<int64_t>(b) * d + a) * MP + l];
        }
        xh[a] = v;
    }
    if (PH == nullptr) return;
    // Psi_hat = [Psi_oo, Psi_oo R; R' Psi_oo, R' Psi_oo R + Schur]                 :270-275
    double* ph = PH + tl * d * d;
    for (int a = 0; a < d; ++a)            // Q(a,e) = sum_{b in o} Psi(a,b) R(b,e), a in o, e in u  (kept in S)
        for (int e = 0; e < d; ++e) {
            double s = 0.0;
            if (ob[a] && !ob[e])
                for (int b = 0; b < d; ++b)
                    if (ob[b]) s += psi[a + b * d] * Rt[(static_cast<int64_t>(b) * d + e) * MP + l];
            Sm(a, e) = s;
        }
    for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) {
            double v;
            if (ob[a] && ob[b]) v = psi[a + b * d];
            else if (ob[a]) v = Sm(a, b);
            else if (ob[b]) v = Sm(b, a);
            else {
                v = Sch[(static_cast<int64_t>(a) * d + b) * MP + l];
                for (int c = 0; c < d; ++c)
                    if (ob[c]) v += Rt[(static_cast<int64_t>(c) * d + a) * MP + l] * Sm(c, b);
            }
            ph[a * d + b] = v;
        }
}

__global__ void pmc_norm_kernel(double* __restrict__ Ex, int64_t rows, int m, int MP) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= rows) return;
    double s = 0.0;
    for (int l = 0; l < m; ++l) s += Ex[t * MP + l];
    for (int l = 0; l < m; ++l) Ex[t * MP + l] /= s;                                // :175 / :282
}

// ln N(delta; 0, S) with S(a,b) = A(a,b) + B(a,b) given through accessors; thread-local Cholesky
template <int DMAX, class FA, class FB, class FD>
__device__ __forceinline__ double pmc_lnN(int d, FA A, FB B, FD delta, bool* ok) {
    double S[DMAX * DMAX], z[DMAX];
    LocalMat Sm{S, d};
    for (int a = 0; a < d; ++a) {
        for (int b = 0; b <= a; ++b) Sm(a, b) = A(a, b) + B(a, b);
        z[a] = delta(a);
    }
    double hl = 0.0, q = 0.0;
    if (!chol_lower(Sm, d, &hl)) {
        *ok = false;
        return 0.0;
    }
    for (int a = 0; a < d; ++a) {
        double s = z[a];
        for (int b = 0; b < a; ++b) s -= Sm(a, b) * z[b];
        s /= Sm(a, a);
        q += s * s;
        z[a] = s;
    }
    return -0.5 * q - hl;
}

// PHI_ti = exp(lnz_i) sum_j N(X_hat_tj - p_i; Sigma_i + Psi_hat_tj) Pio_tj                    :186-199,224 / :293-305,328
template <int DMAX>
__global__ void __launch_bounds__(128)
pmc_phi_kernel(Params P, int64_t rows, const double* __restrict__ Pio, const double* __restrict__ XH, const double* __restrict__ PH,
               const double* __restrict__ Sch, double* __restrict__ Phi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t t = blockIdx.y;
    const int d = P.d, MP = P.MP, m = P.m;
    if (i >= MP || t >= rows) return;
    double acc = 0.0;
    bool ok = true;
    if (i < m) {
        for (int j = 0; j < m; ++j) {
            const int64_t tj = t * MP + j;
            const double* xh = XH + tj * d;
            const double* ph = PH != nullptr ? PH + tj * d * d : nullptr;
            const double ln = pmc_lnN<DMAX>(
                d, [&](int a, int b) { return P.Sj[(static_cast<int64_t>(a) * d + b) * MP + i]; },
                [&](int a, int b) { return ph != nullptr ? ph[a * d + b] : Sch[(static_cast<int64_t>(a) * d + b) * MP + j]; },
                [&](int a) { return xh[a] - P.Pt[a * MP + i]; }, &ok);
            acc += exp(ln) * Pio[tj];
        }
        acc *= exp(0.5 * P.lndS[i]);
    }
    Phi[t * MP + i] = ok ? acc : nan("");
}

// one block per sample: threads walk the basis pairs, fixed-order block reduction                 :179-222 / :286-326
template <int DMAX, int KMAX>
__global__ void __launch_bounds__(128)
pmc_pairs_kernel(Params P, int64_t n, int64_t r0, int64_t rows, CPairTab T, const double* __restrict__ Pio, const double* __restrict__ XH,
                 const double* __restrict__ PH, const double* __restrict__ Sch, const double* __restrict__ elns0,
                 const double* __restrict__ mu, double* __restrict__ nu, double* __restrict__ beta_i, double* __restrict__ gamma) {
    __shared__ double red[8];
    const int64_t t = blockIdx.x;
    if (t >= rows) return;
    const int d = P.d, MP = P.MP, m = P.m, k = P.k;
    double g[KMAX], vl[KMAX], nv[KMAX];
#pragma unroll
    for (int o = 0; o < KMAX; ++o) g[o] = vl[o] = nv[o] = 0.0;
    bool ok = true;
    for (int64_t q = threadIdx.x; q < T.npairs; q += 128) {
        double Ec = 0.0;
        for (int l = 0; l < m; ++l) {
            const int64_t tl = t * MP + l;
            const double* xh = XH + tl * d;
            const double* ph = PH != nullptr ? PH + tl * d * d : nullptr;
            const double ln = pmc_lnN<DMAX>(
                d, [&](int a, int b) { return T.C[(static_cast<int64_t>(a) * d + b) * T.npairs + q]; },
                [&](int a, int b) { return ph != nullptr ? ph[a * d + b] : Sch[(static_cast<int64_t>(a) * d + b) * MP + l]; },
                [&](int a) { return xh[a] - T.c[a * T.npairs + q]; }, &ok);
            Ec += exp(ln) * Pio[tl];
        }
        const double Z = exp(T.lnZ[q]) * Ec;
#pragma unroll
        for (int o = 0; o < KMAX; ++o)
            if (o < k) {
                g[o] = fma(Z, T.ww[o * T.npairs + q], g[o]);
                vl[o] = fma(Z, T.vv[o * T.npairs + q], vl[o]);
                nv[o] = fma(Z, T.ss[o * T.npairs + q], nv[o]);
            }
    }
    const int64_t gt = r0 + t;
#pragma unroll
    for (int o = 0; o < KMAX; ++o) {
        if (o >= k) break;
        const double G = block_sum<128>(ok ? g[o] : nan(""), red);
        const double V = block_sum<128>(vl[o], red);
        const double N = block_sum<128>(nv[o], red);
        if (threadIdx.x == 0) {
            const double e0 = elns0[o * n + gt], m_ = mu[o * n + gt];
            gamma[o * n + gt] = G - m_ * m_;
            beta_i[o * n + gt] = exp(e0 + P.bk[o]) * (1.0 + 0.5 * (V - e0 * e0));
            nu[o * n + gt] = N;
        }
    }
}

template <int DMAX>
static void pmc_launch(const Params& P, const double* X, const double* Psi, int64_t n, int64_t r0, int64_t r1, const unsigned char* ob,
                       const double* prior, const double* Rt, const double* Sch, double* Ex, double* XH, double* PH, double* Phi,
                       const CPairTab& T, const double* elns0, const double* mu, double* nu, double* beta_i, double* gamma, int stage,
                       cudaStream_t st) {
    const int64_t rows = r1 - r0;
    const unsigned gm = static_cast<unsigned>(ceil_div(P.MP, 128));
    if (stage == 0) {              // per-(sample, basis) tables of this chunk
        pmc_rows_kernel<DMAX><<<dim3(gm, static_cast<unsigned>(rows)), 128, 0, st>>>(P, X, Psi, n, r0, r1, ob, prior, Rt, Sch, Ex, XH, PH);
        pmc_norm_kernel<<<static_cast<unsigned>(ceil_div(rows, 128)), 128, 0, st>>>(Ex, rows, P.m, P.MP);
    } else if (stage == 1) {
        pmc_phi_kernel<DMAX><<<dim3(gm, static_cast<unsigned>(rows)), 128, 0, st>>>(P, rows, Ex, XH, PH, Sch, Phi + r0 * P.MP);
    } else {
        pmc_pairs_kernel<DMAX, 4><<<static_cast<unsigned>(rows), 128, 0, st>>>(P, n, r0, rows, T, Ex, XH, PH, Sch, elns0, mu, nu, beta_i, gamma);
    }
}

// needs P.Sj / P.lndS (prep_params with need_sigma = 1); X [d][n] NaN-free (missing dims zero-filled); Psi [n][d*d] or null;
// ob: device [d] observed mask; Phi [n][MP] out
int predict_missing_cov(const Params& P, const double* X, const double* Psi, int64_t n, const unsigned char* ob, const double* prior,
                        const double* w, const double* Sinv, double* mu, double* nu, double* beta_i, double* gamma, double* Phi,
                        cudaStream_t st, int64_t* launches) {
    if (P.k > 4 || P.d > 32) {
        set_error("predictMissing (cov): k <= 4 and d <= 32 supported");
        return GPZ_ERR_USAGE;
    }
    const int64_t MP = P.MP, dd = static_cast<int64_t>(P.d) * P.d;
    std::vector<void*> bufs;
    auto A = [&](double** p, int64_t cnt) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), sizeof(double) * static_cast<size_t>(cnt > 0 ? cnt : 1));
        if (e != cudaSuccess) {
            set_error("predict_missing_cov: cudaMalloc: %s", cudaGetErrorString(e));
            return static_cast<int>(GPZ_ERR_CUDA);
        }
        bufs.push_back(*p);
        return static_cast<int>(GPZ_OK);
    };
    auto done = [&](int rc) {
        cudaStreamSynchronize(st);
        for (void* b : bufs) cudaFree(b);
        return rc;
    };
    // rows per pass: keep the per-(sample, basis) tables below ~1 GB; gridDim.y limit
    int64_t chunk = static_cast<int64_t>(1.0e9 / (8.0 * static_cast<double>(MP) * static_cast<double>(dd + P.d + 1)));
    if (chunk < 1) chunk = 1;
    if (chunk > 65535) chunk = 65535;
    if (chunk > n) chunk = n;
    CPairTab T;
    T.npairs = static_cast<int64_t>(P.m) * (P.m + 1) / 2;
    const int64_t per = dd + P.d + 1 + 3LL * P.k;
    double *tab = nullptr, *Rt = nullptr, *Sch = nullptr, *Ex = nullptr, *XH = nullptr, *PH = nullptr, *elns0 = nullptr;
    int rc;
    if ((rc = A(&tab, per * T.npairs)) || (rc = A(&Rt, dd * MP)) || (rc = A(&Sch, dd * MP)) || (rc = A(&Ex, chunk * MP)) ||
        (rc = A(&XH, chunk * MP * P.d)) || (rc = A(&elns0, P.k * n)))
        return done(rc);
    if (Psi != nullptr && (rc = A(&PH, chunk * MP * dd))) return done(rc);
    T.C = tab;
    T.c = T.C + dd * T.npairs;
    T.lnZ = T.c + static_cast<int64_t>(P.d) * T.npairs;
    T.ww = T.lnZ + T.npairs;
    T.vv = T.ww + static_cast<int64_t>(P.k) * T.npairs;
    T.ss = T.vv + static_cast<int64_t>(P.k) * T.npairs;
    const unsigned nbp = static_cast<unsigned>(ceil_div(T.npairs, 64)), nbb = static_cast<unsigned>(ceil_div(P.m, 64));
    if (P.d <= 8) {
        cpair_table_kernel<8><<<nbp, 64, 0, st>>>(P, w, Sinv, T);
        pmc_basis_kernel<8><<<nbb, 64, 0, st>>>(P, ob, Rt, Sch);
    } else if (P.d <= 16) {
        cpair_table_kernel<16><<<nbp, 64, 0, st>>>(P, w, Sinv, T);
        pmc_basis_kernel<16><<<nbb, 64, 0, st>>>(P, ob, Rt, Sch);
    } else {
        cpair_table_kernel<32><<<nbp, 64, 0, st>>>(P, w, Sinv, T);
        pmc_basis_kernel<32><<<nbb, 64, 0, st>>>(P, ob, Rt, Sch);
    }
    *launches += 2;
    auto run = [&](int stage, int64_t r0, int64_t r1) {
        if (P.d <= 8) pmc_launch<8>(P, X, Psi, n, r0, r1, ob, prior, Rt, Sch, Ex, XH, PH, Phi, T, elns0, mu, nu, beta_i, gamma, stage, st);
        else if (P.d <= 16) pmc_launch<16>(P, X, Psi, n, r0, r1, ob, prior, Rt, Sch, Ex, XH, PH, Phi, T, elns0, mu, nu, beta_i, gamma, stage, st);
        else pmc_launch<32>(P, X, Psi, n, r0, r1, ob, prior, Rt, Sch, Ex, XH, PH, Phi, T, elns0, mu, nu, beta_i, gamma, stage, st);
        *launches += stage == 0 ? 2 : 1;
    };
    const bool one_pass = chunk >= n;
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {          // PHI for all rows
        const int64_t r1 = (r0 + chunk < n) ? r0 + chunk : n;
        run(0, r0, r1);
        run(1, r0, r1);
    }
    for (int o = 0; o < P.k; ++o)                        // mu = PHI w, E ln S = PHI v: the pair sums need them
        if ((rc = rowdot(Phi, MP, P.m, n, DotSpec{2, {w + o * MP, P.v + o * MP}, {mu + o * n, elns0 + o * n}}, st, launches))) return done(rc);
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t r1 = (r0 + chunk < n) ? r0 + chunk : n;
        if (!one_pass) run(0, r0, r1);                   // the tables of this chunk were overwritten: rebuild (cheap next to the pair sums)
        run(2, r0, r1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("predict_missing_cov: %s", cudaGetErrorString(e));
        return done(GPZ_ERR_CUDA);
    }
    return done(GPZ_OK);
}

}  // namespace gpz
