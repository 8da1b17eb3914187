// Back-projection of dPHI = dlnPHI .* PHI onto the basis centres P and the length-scale / precision
// parameters Gamma (GPz/GPz.m:113-225).
//
//   no Psi, no NaN (all six modes): the per-basis sums  sum_i dPHI_ij {1, x_i, x_i x_i'}  are ONE
//     tensor-core GEMM  R = dPHI' F  against a row-feature matrix F (gemm.cu:atb_general); the
//     centred moments  sum_i dPHI_ij Delta_i , sum_i dPHI_ij Delta_i' Delta_i  follow by expansion
//     (GPz.m:152-154 cov modes, GPz.m:192-194 diag modes).
//   diag modes with Psi and/or NaN: element-wise sweep (GPz.m:200-206), thread per basis.
//   cov modes with Psi: per (sample,basis) d x d inverse (GPz.m:166-184), thread per basis.
// Every sum over rows is taken in a fixed order (row slabs -> ordered final sum).
#include "internal.cuh"
#include "smallmat.cuh"

namespace gpz {

int feature_count(const Params& P) {
    return mode_is_cov(P.mode) ? 1 + P.d + P.d * (P.d + 1) / 2 : 1 + 2 * P.d;
}

// F[i][c], row-major [rows][QP]:  diag: [1, x_a, x_a^2]   cov: [1, x_a, x_a x_b (a<=b)]
__global__ void __launch_bounds__(256)
features_kernel(const double* __restrict__ X, int64_t n, int64_t r0, int64_t r1, int d, int cov, int QP,
                double* __restrict__ F) {
    extern __shared__ double xs[];                 // [32][d]
    const int64_t rb = r0 + static_cast<int64_t>(blockIdx.x) * 32;
    for (int e = threadIdx.x; e < 32 * d; e += blockDim.x) {
        const int a = e / 32, r = e % 32;
        const int64_t gi = rb + r;
        xs[r * d + a] = (gi < r1) ? X[a * n + gi] : 0.0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * QP; e += blockDim.x) {
        const int r = e / QP, c = e % QP;
        const int64_t gi = rb + r;
        if (gi >= r1) continue;
        double v = 0.0;
        if (c == 0) v = 1.0;
        else if (c <= d) v = xs[r * d + c - 1];
        else if (!cov) {
            if (c <= 2 * d) { const double x = xs[r * d + c - 1 - d]; v = x * x; }
        } else {
            int q = c - 1 - d, a = 0;
            while (a < d && q >= d - a) { q -= d - a; ++a; }
            if (a < d) v = xs[r * d + a] * xs[r * d + a + q];
        }
        F[(gi - r0) * QP + c] = v;
    }
}

int build_features(const Params& P, const double* X, int64_t n, int64_t r0, int64_t r1, double* F, cudaStream_t st,
                   int64_t* launches) {
    if (r1 <= r0) return GPZ_OK;
    features_kernel<<<static_cast<unsigned>(ceil_div(r1 - r0, 32)), 256, sizeof(double) * 32 * P.d, st>>>(
        X, n, r0, r1, P.d, mode_is_cov(P.mode) ? 1 : 0, P.QP, F);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

// ------------------------------------------------------------------------------------------------
// reduce per-basis Gamma gradients by mode (GPz.m:215-225) -- fixed-order block sums
//   diag: full[a*m + j]      cov: full[b + a*d + d*d*j]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mode_reduce_kernel(Params P, const double* __restrict__ full, double* __restrict__ out) {
    __shared__ double sh[8];
    const int d = P.d, m = P.m;
    const int o = blockIdx.x;
    double s = 0.0;
    switch (P.mode) {
        case GL:
            for (int e = threadIdx.x; e < m * d; e += 256) s += full[e];
            break;
        case VL:
            for (int a = threadIdx.x; a < d; a += 256) s += full[a * m + o];
            break;
        case GD:
            for (int j = threadIdx.x; j < m; j += 256) s += full[o * m + j];
            break;
        case GC:
            for (int j = threadIdx.x; j < m; j += 256) s += full[o + static_cast<int64_t>(d) * d * j];
            break;
        default:
            break;
    }
    s = block_sum<256>(s, sh);
    if (threadIdx.x == 0) out[o] = s;
}

__global__ void copy_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

int mode_reduce(const Params& P, const double* full, double* dG, cudaStream_t st, int64_t* launches) {
    if (P.mode == VD || P.mode == VC) {
        copy_kernel<<<static_cast<unsigned>(ceil_div(P.g_dim, 256)), 256, 0, st>>>(full, dG, P.g_dim);
    } else {
        mode_reduce_kernel<<<static_cast<unsigned>(P.g_dim), 256, 0, st>>>(P, full, dG);
    }
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

// ------------------------------------------------------------------------------------------------
// moments R = dPHI' F  ->  dP, per-basis dGamma
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
moments_diag_kernel(Params P, const double* __restrict__ Rm, int QP, double* __restrict__ dP, double* __restrict__ full) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = P.d, m = P.m, MP = P.MP;
    if (j >= m) return;
    const double* R = Rm + static_cast<int64_t>(j) * QP;
    const double M0 = R[0];
    for (int a = 0; a < d; ++a) {
        const double p = P.Pt[a * MP + j], gv = P.Gt[a * MP + j];
        const double S1 = R[1 + a], S2 = R[1 + d + a];
        const double sP = S1 - p * M0;                       // sum dPHI * Delta
        const double sG = S2 - 2.0 * p * S1 + p * p * M0;    // sum dPHI * Delta^2
        dP[a * m + j] = gv * gv * sP;                        // GPz.m:192
        full[a * m + j] = -gv * sG;                          // GPz.m:194
    }
}

// cov modes, one missing-input pattern (observed set o, missing set u):                       GPz.m:146-159
//   iSoo = (Sigma_j(o,o))^-1 = Mg;  dP_j(o) += M1d(o) iSoo;  diSoo = -1/2 M2d(o,o);
//   GuuGuo = Gg;  dGo = 2 (Gamma(:,o) - Gamma(:,u) GuuGuo) diSoo;  dGamma(:,o) += dGo;  dGamma(:,u) -= dGo GuuGuo'
// thread = (basis j, row c of dGamma_j): blockIdx.y = c, consecutive threads = consecutive bases (coalesced parameter loads).
// (One thread per basis did all d rows in sequence: 1000 threads x d^3 dependent steps = 0.39 ms at m = 1000, d = 10 --
// a fixed cost that does not shrink with the number of GPUs.)
template <int DMAX>
__global__ void __launch_bounds__(128)
moments_cov_kernel(Params P, int pat, const double* __restrict__ Rm, int QP, double* __restrict__ dP,
                   double* __restrict__ full, int accumulate) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    const int d = P.d, m = P.m, MP = P.MP, dp = P.dp;
    if (j >= m) return;
    const unsigned char* ob = P.obs + pat * d;
    const double* Mg = P.Mg + static_cast<int64_t>(pat) * d * d * MP;
    const double* Gg = P.Gg + static_cast<int64_t>(pat) * d * d * MP;
    const double* R = Rm + static_cast<int64_t>(j) * QP;
    const double M0 = R[0];
    double m1[DMAX], pj[DMAX];
    for (int a = 0; a < d; ++a) {
        pj[a] = P.Pt[a * MP + j];
        m1[a] = ob[a] ? R[1 + a] - pj[a] * M0 : 0.0;
    }
    {                      // dP_j(a) for a = c
        const int a = c;
        double s = 0.0;
        if (ob[a])
            for (int b = 0; b < d; ++b)
                if (ob[b]) s += m1[b] * Mg[(static_cast<int64_t>(b) * d + a) * MP + j];
        const int64_t o = a * m + j;
        dP[o] = (accumulate ? dP[o] : 0.0) + s;
    }
    // Gproj(c,b) = Gamma(c,b) - sum_{e in u} Gamma(c,e) G(e,b)   (b in o)
    double dGo[DMAX];      // row c of dGo
    {
        for (int a = 0; a < d; ++a) dGo[a] = 0.0;
        for (int b = 0; b < d; ++b) {
            if (!ob[b]) continue;
            double gp = P.Gam[(static_cast<int64_t>(c) * dp + b) * MP + j];
            for (int e = 0; e < d; ++e)
                if (!ob[e]) gp -= P.Gam[(static_cast<int64_t>(c) * dp + e) * MP + j] * Gg[(static_cast<int64_t>(e) * d + b) * MP + j];
            const double pb = pj[b];
            for (int a = 0; a < d; ++a) {
                if (!ob[a]) continue;
                const int lo = b < a ? b : a, hi = b < a ? a : b;
                const int idx = 1 + d + lo * d - lo * (lo - 1) / 2 + (hi - lo);
                const double m2 = R[idx] - pb * R[1 + a] - R[1 + b] * pj[a] + pj[a] * pb * M0;
                dGo[a] -= gp * m2;                      // 2 * Gproj * (-1/2 M2d)
            }
        }
        for (int a = 0; a < d; ++a) {
            double val;
            if (ob[a]) {
                val = dGo[a];
            } else {                                    // dGamma(c, e in u) -= sum_{a in o} dGo(c,a) G(e,a)
                val = 0.0;
                for (int b = 0; b < d; ++b)
                    if (ob[b]) val -= dGo[b] * Gg[(static_cast<int64_t>(a) * d + b) * MP + j];
            }
            const int64_t o = c + a * d + static_cast<int64_t>(d) * d * j;
            full[o] = (accumulate ? full[o] : 0.0) + val;
        }
    }
}

int moments_to_grad(const Params& P, int pat, const double* Rm, int QP, double* dP, double* full, int accumulate,
                    cudaStream_t st, int64_t* launches) {
    const unsigned nb = static_cast<unsigned>(ceil_div(P.m, 128));
    if (!mode_is_cov(P.mode)) moments_diag_kernel<<<nb, 128, 0, st>>>(P, Rm, QP, dP, full);
    else if (P.d <= 8) moments_cov_kernel<8><<<dim3(nb, P.d), 128, 0, st>>>(P, pat, Rm, QP, dP, full, accumulate);
    else if (P.d <= 16) moments_cov_kernel<16><<<dim3(nb, P.d), 128, 0, st>>>(P, pat, Rm, QP, dP, full, accumulate);
    else moments_cov_kernel<32><<<dim3(nb, P.d), 128, 0, st>>>(P, pat, Rm, QP, dP, full, accumulate);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

// ------------------------------------------------------------------------------------------------
// diag modes, generic (Psi and/or NaN): partial[slab][2][d][MP]
//   accP_ja = sum_i dPHI_ij Delta g^2 / s          (s = 1 + Psi_ia g^2)        GPz.m:202
//   accG_ja = sum_i dPHI_ij [ (Delta/s)^2 + Psi_ia / s ]                        GPz.m:204-206
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
backproj_diag_kernel(Params P, const double* __restrict__ X, const double* __restrict__ Psi, int64_t n, int64_t r0,
                     int64_t r1, int64_t rows_per_slab, const double* __restrict__ dPhi, int64_t ld,
                     double* __restrict__ partial, int accumulate) {
    extern __shared__ double sm[];
    const int d = P.d, MP = P.MP;
    double* accP = sm;                    // [d][128]
    double* accG = accP + d * 128;        // [d][128]
    double* xs = accG + d * 128;          // [32][d]
    double* ps = xs + 32 * d;             // [32][d]
    const int tid = threadIdx.x;
    const int j = blockIdx.x * 128 + tid;
    const int64_t sb = r0 + static_cast<int64_t>(blockIdx.y) * rows_per_slab;
    int64_t se = sb + rows_per_slab;
    if (se > r1) se = r1;
    for (int a = 0; a < d; ++a) accP[a * 128 + tid] = accG[a * 128 + tid] = 0.0;
    for (int64_t rb = sb; rb < se; rb += 32) {
        __syncthreads();
        for (int e = tid; e < 32 * d; e += 128) {
            const int a = e / 32, r = e % 32;
            const int64_t gi = rb + r;
            xs[r * d + a] = (gi < se) ? X[a * n + gi] : 0.0;
            ps[r * d + a] = (gi < se && Psi != nullptr) ? Psi[a * n + gi] : 0.0;
        }
        __syncthreads();
        const int nr = (se - rb < 32) ? static_cast<int>(se - rb) : 32;
        for (int r = 0; r < nr; ++r) {
            const double dphi = dPhi[(rb + r - r0) * ld + j];
            for (int a = 0; a < d; ++a) {
                const double x = xs[r * d + a];
                if (x != x) continue;
                const double gv = P.Gt[a * MP + j];
                const double g2 = gv * gv;
                const double dl = x - P.Pt[a * MP + j];
                const double psi = ps[r * d + a];
                const double is = 1.0 / fma(psi, g2, 1.0);
                const double dr = dl * is;
                accP[a * 128 + tid] += dphi * dr * g2;
                accG[a * 128 + tid] += dphi * fma(dr, dr, psi * is);
            }
        }
    }
    double* out = partial + static_cast<int64_t>(blockIdx.y) * 2 * d * MP;
    for (int a = 0; a < d; ++a) {
        double vp = accP[a * 128 + tid], vg = accG[a * 128 + tid];
        if (accumulate) {
            vp += out[a * MP + j];
            vg += out[(d + a) * MP + j];
        }
        out[a * MP + j] = vp;
        out[(d + a) * MP + j] = vg;
    }
}

__global__ void __launch_bounds__(128)
backproj_diag_finish_kernel(Params P, const double* __restrict__ partial, int nslab, double* __restrict__ dP,
                            double* __restrict__ full) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = P.d, m = P.m, MP = P.MP;
    if (j >= m) return;
    for (int a = 0; a < d; ++a) {
        double sp = 0.0, sg = 0.0;
        for (int s = 0; s < nslab; ++s) {
            sp += partial[(static_cast<int64_t>(s) * 2 * d + a) * MP + j];
            sg += partial[(static_cast<int64_t>(s) * 2 * d + d + a) * MP + j];
        }
        dP[a * m + j] = sp;
        full[a * m + j] = -P.Gt[a * MP + j] * sg;
    }
}

int64_t backproj_partial_doubles(const Params& P, int nslab, int has_psi, int has_nan) {
    if (mode_is_cov(P.mode)) return has_psi ? static_cast<int64_t>(P.npat) * nslab * (1 + P.d + P.d * P.d) * P.MP : 0;
    return (has_psi || has_nan) ? static_cast<int64_t>(nslab) * 2 * P.d * P.MP : 0;
}

int backproj_diag_generic(const Params& P, const RowData& R, int64_t r0, int64_t r1, const double* dPhi, int64_t ld,
                          double* partial, int nslab, int accumulate, cudaStream_t st, int64_t* launches) {
    const size_t smem = sizeof(double) * (2 * P.d * 128 + 2 * 32 * P.d);
    if (smem > 48 * 1024) {
        GPZ_CUDA(cudaFuncSetAttribute(backproj_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    }
    int64_t rps = round_up(ceil_div(r1 - r0 > 0 ? r1 - r0 : 1, nslab), 32);
    dim3 grid(static_cast<unsigned>(P.MP / 128), static_cast<unsigned>(nslab));
    backproj_diag_kernel<<<grid, 128, smem, st>>>(P, R.X, R.Psi, R.n, r0, r1, rps, dPhi, ld, partial, accumulate);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

int backproj_diag_generic_finish(const Params& P, const double* partial, int nslab, double* dP, double* dG,
                                 double* scratch, cudaStream_t st, int64_t* launches) {
    backproj_diag_finish_kernel<<<static_cast<unsigned>(ceil_div(P.m, 128)), 128, 0, st>>>(P, partial, nslab, dP, scratch);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return mode_reduce(P, scratch, dG, st, launches);
}

// ------------------------------------------------------------------------------------------------
// cov modes + Psi: partial[group][slab][1 + d + d*d][MP]  (s0, u, Q)  with, per (i,j) on the observed dims o of the row's pattern:
//   iPS = (Sigma_j(o,o) + Psi_i(o,o))^{-1}, z = iPS Delta_o',  u += dPHI z,  Q += dPHI (z z' - iPS),  s0 += dPHI   (GPz.m:166-184)
// (o,o) blocks are embedded in d x d (identity on the missing dims while inverting, zero in the accumulators).
// ------------------------------------------------------------------------------------------------
template <int DMAX>
__global__ void __launch_bounds__(128)
backproj_cov_psi_kernel(Params P, const double* __restrict__ X, const double* __restrict__ Psi, int64_t n, int64_t r0,
                        int64_t r1, int64_t rows_per_slab, int pat, const double* __restrict__ dPhi, int64_t ld, int64_t row_off,
                        double* __restrict__ partial, int accumulate) {
    const int d = P.d, MP = P.MP, m = P.m;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= MP) return;
    const int64_t sb = r0 + static_cast<int64_t>(blockIdx.y) * rows_per_slab;
    int64_t se = sb + rows_per_slab;
    if (se > r1) se = r1;
    const int nacc = 1 + d + d * d;
    double* out = partial + static_cast<int64_t>(blockIdx.y) * nacc * MP + j;
    if (!accumulate)
        for (int e = 0; e < nacc; ++e) out[static_cast<int64_t>(e) * MP] = 0.0;
    if (j >= m) return;
    const unsigned char* ob = P.obs + pat * d;
    double S[DMAX * DMAX], dl[DMAX], z[DMAX];
    LocalMat Sm{S, d};
    for (int64_t gi = sb; gi < se; ++gi) {
        const double dphi = dPhi[(gi - row_off) * ld + j];
        const double* psi = Psi + gi * d * d;
        for (int a = 0; a < d; ++a) {
            for (int b = 0; b <= a; ++b)
                Sm(a, b) = (ob[a] && ob[b]) ? psi[a + b * d] + P.Sj[(static_cast<int64_t>(a) * d + b) * MP + j] : (a == b ? 1.0 : 0.0);
            dl[a] = ob[a] ? X[a * n + gi] - P.Pt[a * MP + j] : 0.0;
        }
        double hl;
        if (!spd_inv(Sm, d, &hl)) {
            out[0] = nan("");
            continue;
        }
        for (int a = 0; a < d; ++a) {
            double s = 0.0;
            for (int b = 0; b < d; ++b) s += Sm(a, b) * dl[b];
            z[a] = s;
        }
        out[0] += dphi;
        for (int a = 0; a < d; ++a) {
            if (!ob[a]) continue;
            out[static_cast<int64_t>(1 + a) * MP] += dphi * z[a];
            for (int b = 0; b < d; ++b)
                if (ob[b]) out[static_cast<int64_t>(1 + d + a * d + b) * MP] += dphi * (z[a] * z[b] - Sm(a, b));
        }
    }
}

// per basis and pattern (o observed, u missing):  dSoo = 1/2 (Sigma_oo^-1 s0 + Q);  diSoo = -Sigma_oo dSoo Sigma_oo;
//   dGo = 2 (Gamma(:,o) - Gamma(:,u) G) diSoo,  dGamma(:,o) += dGo,  dGamma(:,u) -= dGo G',  dP(o) += u      (GPz.m:172-181)
// with G = iSigma(u,u)^-1 iSigma(u,o) (P.Gg) and Sigma_oo^-1 the marginal precision (P.Mg).  Without missing dims this is
// dGamma_j = -2 Gamma_j Sigma_j B Sigma_j.  scratch: [3][d*d][MP] + full gradient [g_full]
__global__ void __launch_bounds__(128)
backproj_cov_psi_finish_kernel(Params P, int pat, const double* __restrict__ partial, int nslab, int accumulate, double* __restrict__ dP,
                               double* __restrict__ work, double* __restrict__ full) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = P.d, m = P.m, MP = P.MP, dp = P.dp;
    if (j >= m) return;
    const unsigned char* ob = P.obs + pat * d;
    const double* Mg = P.Mg + static_cast<int64_t>(pat) * d * d * MP;
    const double* Gg = P.Gg + static_cast<int64_t>(pat) * d * d * MP;
    const int nacc = 1 + d + d * d;
    StridedMat B{work + j, MP, d};
    StridedMat T{work + static_cast<int64_t>(d) * d * MP + j, MP, d};
    StridedMat E{work + 2LL * d * d * MP + j, MP, d};
    double s0 = 0.0;
    for (int s = 0; s < nslab; ++s) s0 += partial[(static_cast<int64_t>(s) * nacc) * MP + j];
    for (int a = 0; a < d; ++a) {
        double u = 0.0;
        for (int s = 0; s < nslab; ++s) u += partial[(static_cast<int64_t>(s) * nacc + 1 + a) * MP + j];
        dP[a * m + j] = (accumulate ? dP[a * m + j] : 0.0) + u;
        for (int b = 0; b < d; ++b) {
            double q = 0.0;
            for (int s = 0; s < nslab; ++s) q += partial[(static_cast<int64_t>(s) * nacc + 1 + d + a * d + b) * MP + j];
            B(a, b) = 0.5 * (Mg[(static_cast<int64_t>(a) * d + b) * MP + j] * s0 + q);       // dSoo (zero outside (o,o))
        }
    }
    // T = dSoo * Sigma_oo ;  B = -Sigma_oo * T = diSoo   (Sigma_oo = Sigma_j restricted to the observed dims)
    for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) {
            double s = 0.0;
            if (ob[b])
                for (int c = 0; c < d; ++c)
                    if (ob[c]) s += B(a, c) * P.Sj[(static_cast<int64_t>(c) * d + b) * MP + j];
            T(a, b) = s;
        }
    for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) {
            double s = 0.0;
            if (ob[a])
                for (int c = 0; c < d; ++c)
                    if (ob[c]) s += P.Sj[(static_cast<int64_t>(a) * d + c) * MP + j] * T(c, b);
            B(a, b) = -s;
        }
    // E(c,a) = dGo(c,a) = 2 sum_{b in o} (Gamma(c,b) - sum_{e in u} Gamma(c,e) G(e,b)) diSoo(b,a),  a in o
    for (int c = 0; c < d; ++c)
        for (int a = 0; a < d; ++a) {
            double s = 0.0;
            if (ob[a])
                for (int b = 0; b < d; ++b) {
                    if (!ob[b]) continue;
                    double ge = P.Gam[(static_cast<int64_t>(c) * dp + b) * MP + j];
                    for (int e = 0; e < d; ++e)
                        if (!ob[e]) ge -= P.Gam[(static_cast<int64_t>(c) * dp + e) * MP + j] * Gg[(static_cast<int64_t>(e) * d + b) * MP + j];
                    s += ge * B(b, a);
                }
            E(c, a) = 2.0 * s;
        }
    for (int a = 0; a < d; ++a)
        for (int c = 0; c < d; ++c) {
            double v;
            if (ob[a]) v = E(c, a);
            else {                                   // missing dim e = a: - sum_{b in o} dGo(c,b) G(e,b)
                v = 0.0;
                for (int b = 0; b < d; ++b)
                    if (ob[b]) v -= E(c, b) * Gg[(static_cast<int64_t>(a) * d + b) * MP + j];
            }
            const int64_t o = c + a * d + static_cast<int64_t>(d) * d * j;
            full[o] = (accumulate ? full[o] : 0.0) + v;
        }
}

static void group_range(const RowData& R, size_t g, int64_t r0, int64_t r1, int64_t* s0, int64_t* s1, int* pat) {
    *s0 = r0;
    *s1 = r1;
    *pat = 0;
    if (!R.g_pat.empty()) {
        *s0 = R.g_r0[g] > r0 ? R.g_r0[g] : r0;
        *s1 = R.g_r1[g] < r1 ? R.g_r1[g] : r1;
        *pat = R.g_pat[g];
    }
}

// partial: [groups][nslab][1 + d + d*d][MP]; rows r0..r1 of this chunk, dPhi row 0 = row r0
int backproj_cov_psi(const Params& P, const RowData& R, int64_t r0, int64_t r1, const double* dPhi, int64_t ld,
                     double* partial, int nslab, int accumulate, cudaStream_t st, int64_t* launches) {
    const size_t ng = R.g_pat.empty() ? 1 : R.g_pat.size();
    const int64_t gstride = static_cast<int64_t>(nslab) * (1 + P.d + P.d * P.d) * P.MP;
    for (size_t g = 0; g < ng; ++g) {
        int64_t s0, s1;
        int pat;
        group_range(R, g, r0, r1, &s0, &s1, &pat);
        if (s1 < s0) s1 = s0;                                    // empty intersection: still zero / keep the accumulators
        int64_t rps = ceil_div(s1 - s0 > 0 ? s1 - s0 : 1, nslab);
        dim3 grid(static_cast<unsigned>(ceil_div(P.MP, 128)), static_cast<unsigned>(nslab));
        double* pg = partial + static_cast<int64_t>(g) * gstride;
        if (P.d <= 8)
            backproj_cov_psi_kernel<8><<<grid, 128, 0, st>>>(P, R.X, R.Psi, R.n, s0, s1, rps, pat, dPhi, ld, r0, pg, accumulate);
        else if (P.d <= 16)
            backproj_cov_psi_kernel<16><<<grid, 128, 0, st>>>(P, R.X, R.Psi, R.n, s0, s1, rps, pat, dPhi, ld, r0, pg, accumulate);
        else
            backproj_cov_psi_kernel<32><<<grid, 128, 0, st>>>(P, R.X, R.Psi, R.n, s0, s1, rps, pat, dPhi, ld, r0, pg, accumulate);
        GPZ_KERNEL_CHECK();
        ++*launches;
    }
    return GPZ_OK;
}

int backproj_cov_psi_finish(const Params& P, const RowData& R, const double* partial, int nslab, double* dP, double* dG, double* scratch,
                            cudaStream_t st, int64_t* launches) {
    // scratch layout: work [3*d*d*MP] | full [d*d*m]
    double* work = scratch;
    double* full = scratch + 3LL * P.d * P.d * P.MP;
    const size_t ng = R.g_pat.empty() ? 1 : R.g_pat.size();
    const int64_t gstride = static_cast<int64_t>(nslab) * (1 + P.d + P.d * P.d) * P.MP;
    for (size_t g = 0; g < ng; ++g) {
        const int pat = R.g_pat.empty() ? 0 : R.g_pat[g];
        backproj_cov_psi_finish_kernel<<<static_cast<unsigned>(ceil_div(P.m, 128)), 128, 0, st>>>(
            P, pat, partial + static_cast<int64_t>(g) * gstride, nslab, g > 0, dP, work, full);
        GPZ_KERNEL_CHECK();
        ++*launches;
    }
    return mode_reduce(P, full, dG, st, launches);
}

}  // namespace gpz
