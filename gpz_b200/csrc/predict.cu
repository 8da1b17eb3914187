// predictNoisy for the diagonal modes (GPz/predictDiag.m:75-125): input-noise aware predictive
// variance terms through the m(m+1)/2 basis-pair sum
//   Z_ij(x) = exp(lnZ_ij) * N(x; c_ij, C_ij + Psi),   gamma += f Z w_i w_j,  VlnS += f Z v_i v_j,
//   nu += f Z iSigma_w(i,j)          (f = 2 off the diagonal, 1 on it: predictDiag.m:113-119)
// A pair table (C_ij, c_ij, lnZ_ij and the three weights) is built once per call; the row kernel
// keeps x_i, Psi_i in registers/local memory and streams the table through L1 (every thread of a CTA
// reads the same pair, so the loads are broadcasts).
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
        T.ss[o * T.npairs + q] = f * Sinv[(static_cast<int64_t>(o) * MP + i) * MP + j];
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
