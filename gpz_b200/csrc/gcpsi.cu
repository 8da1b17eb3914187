// Global-covariance mode with per-sample input noise (method GC + Psi; BASELINE config 5).
//
// The reference (GPz/getPHI.m:80-88, GPz/GPz.m:166-184) loops over every (sample i, basis j) pair with a d x d inverse of
// Psi_i + Sigma_j: n m d^3.  With ONE covariance for all bases (GC) that matrix, S_i = Psi_i + Sigma, depends on the row
// only, so everything reduces to one d x d factorisation per row plus GEMMs:
//   forward   ln PHI_ij = -1/2 (x_i-p_j)' M_i (x_i-p_j) + 1/2 ln|Sigma| - 1/2 ln|S_i|,  M_i = S_i^-1
//             = F_i . W_j  with row features  F_i = [ -1/2 x'M x + 1/2 ln|Sigma| - 1/2 ln|S_i|,  z_i = M_i x_i,  M_i(a,b) a<=b ]
//             and basis columns               W_j = [ 1,  p_j,  -1/2 p_a^2 (a=b) / -p_a p_b (a<b) ]          K = 1 + d + d(d+1)/2
//             -> the same tensor-core kernel as the no-Psi path (PHI = exp(F W), gemm.cu), F rebuilt per evaluation;
//   backward  dP_j = sum_i dPHI_ij (x_i-p_j)' M_i = (dPHI' Z)_j - p_j' mat((dPHI' vech M)_j)      = the moment GEMM dPHI' F
//             dGamma = -2 Gamma Sigma B Sigma,  B = 1/2 (s Sigma^-1 + sum_i [M_i Q_i M_i - a_i M_i]),
//             Q_i = sum_j dPHI_ij (x_i-p_j)(x_i-p_j)' = a_i x x' - x u' - u x' + V_i  with  [a_i, u_i, vech V_i] = (dPHI G)_i,
//             G_j = [1, p_j, p_a p_b (a<=b)]  -> one row-tile GEMM dPHI G and a per-row kernel (2 d^3 per row).
// Cost 2 n m K per GEMM instead of n m d^3 (d = 32: K = 561 against d^3 = 32768).
#include "internal.cuh"
#include "smallmat.cuh"

namespace gpz {

constexpr int GC_BP_WARPS = 12;      // warps per SM of gc_rows_backproj_kernel (its partial-sum slots)

int gc_feature_width(int d) { return static_cast<int>(round_up(1 + d + d * (d + 1) / 2, 32)); }
static int gc_gwidth(int d) { return static_cast<int>(round_up(gc_feature_width(d), TILE)); }      // N of the row-tile GEMM dPHI G

// warp per row, everything in registers: lane r holds row r of S = Psi_i + Sigma (lower triangle), the Cholesky factor is formed
// right-looking with the pivot and the column moving by shuffles (no division: the pivot's reciprocal square root also scales
// the inverse), then lane b builds column b of W = L^-1 by forward substitution (L_rq broadcast from lane r), column b of
// M = W'W, z_b = (M x)_b, and the features are written.  The first form of this kernel kept S in shared memory with three warp
// barriers per column and a division, a square root and a log per pivot: ~10 ms for 250 000 rows at d = 32
// (profiles/r02u_cfg5_launches_partial.md).  ln|S| comes from the running product of the pivots (mantissa and exponent kept
// apart, one log per row).  Dims >= d are padded with the identity.
template <int DMAX>
__global__ void __launch_bounds__(128)
gc_features_kernel(Params P, const double* __restrict__ X, const double* __restrict__ Psi, int64_t n, int64_t r0, int64_t r1, int KQ,
                   double* __restrict__ F) {
    extern __shared__ double gcf_sm[];                     // Sigma of basis 0 (GC shares it): [DMAX][DMAX + 1]
    constexpr int LD = DMAX + 1;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int d = P.d, MP = P.MP, r = lane;
    for (int e = threadIdx.x; e < DMAX * DMAX; e += blockDim.x) {
        const int a = e / DMAX, b = e - a * DMAX;
        gcf_sm[a * LD + b] = (a < d && b < d) ? P.Sj[(static_cast<int64_t>(a) * d + b) * MP] : 0.0;
    }
    __syncthreads();
    const int nv = d * (d + 1) / 2;
    const double half_lndS = 0.5 * P.lndS[0];
    for (int64_t i = r0 + static_cast<int64_t>(blockIdx.x) * wpb + warp; i < r1; i += static_cast<int64_t>(gridDim.x) * wpb) {
        const double* psi = Psi + i * d * d;
        double* f = F + (i - r0) * KQ;
        double a[DMAX];
#pragma unroll
        for (int q = 0; q < DMAX; ++q)
            a[q] = (r < d && q <= r) ? psi[r + q * d] + gcf_sm[r * LD + q] : ((q == r) ? 1.0 : 0.0);
        const double xv = (r < d) ? X[r * n + i] : 0.0;
        // ---- Cholesky (lower), lane = row
        double pm = 1.0, ilself = 0.0;
        int pe = 0;
        bool ok = true;
#pragma unroll
        for (int c = 0; c < DMAX; ++c) {
            const double piv = __shfl_sync(FULL, a[c], c);
            if (c < d) {
                if (!(piv >= 2.2250738585072014e-308) || !(piv < 1.7e308)) ok = false;
                const int ph = __double2hiint(piv);                                   // piv = mant * 2^ex, mant in [1, 2)
                pe += ((ph >> 20) & 0x7ff) - 1023;
                pm *= __hiloint2double((ph & 0x800fffff) | 0x3ff00000, __double2loint(piv));
                const int mh = __double2hiint(pm);                                    // pm back into [1, 2)
                pe += ((mh >> 20) & 0x7ff) - 1023;
                pm = __hiloint2double((mh & 0x800fffff) | 0x3ff00000, __double2loint(pm));
            }
            const double rs = rsqrt(piv);
            const double mine = (r == c) ? piv * rs : a[c] * rs;                      // L[r][c]
            a[c] = mine;
            if (r == c) ilself = rs;                                                  // 1 / L[r][r]
#pragma unroll
            for (int q = c + 1; q < DMAX; ++q) {
                const double bq = __shfl_sync(FULL, mine, q);                         // L[q][c]
                a[q] = fma(-mine, bq, a[q]);
            }
        }
        if (!ok) {                                                                    // uniform: every lane saw the same pivots
            for (int c = lane; c < KQ; c += 32) f[c] = nan("");
            continue;
        }
        // ---- W = L^-1, lane = column b: x[rr] = W[rr][b]
        double x[DMAX];
#pragma unroll
        for (int rr = 0; rr < DMAX; ++rr) {
            double sacc = (rr == r) ? 1.0 : 0.0;
#pragma unroll
            for (int q = 0; q < rr; ++q) {
                const double lrq = __shfl_sync(FULL, a[q], rr);                       // L[rr][q]
                sacc = fma(-lrq, x[q], sacc);
            }
            const double ilr = __shfl_sync(FULL, ilself, rr);
            x[rr] = (rr >= r) ? sacc * ilr : 0.0;
        }
        // ---- M = W'W, lane = column b: mcol[aa] = M[aa][b] = sum_{c >= aa} W[c][aa] W[c][b]
        double mcol[DMAX];
#pragma unroll
        for (int aa = 0; aa < DMAX; ++aa) {
            double sacc = 0.0;
#pragma unroll
            for (int c = aa; c < DMAX; ++c) {
                const double wca = __shfl_sync(FULL, x[c], aa);
                sacc = fma(wca, x[c], sacc);
            }
            mcol[aa] = sacc;
        }
        // ---- z = M x (M symmetric: z_b = sum_a M[a][b] x_a), q = x'z
        double z = 0.0;
#pragma unroll
        for (int aa = 0; aa < DMAX; ++aa) {
            const double xa = __shfl_sync(FULL, xv, aa);
            z = fma(mcol[aa], xa, z);
        }
        double qv = (r < d) ? z * xv : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) qv += __shfl_xor_sync(FULL, qv, o);
        if (r < d) f[1 + r] = z;
        if (lane == 0) f[0] = -0.5 * qv + half_lndS - 0.5 * (log(pm) + pe * 0.69314718055994530942);
        int off = 0;                                       // packed upper triangle: row aa, columns aa .. d-1
#pragma unroll
        for (int aa = 0; aa < DMAX; ++aa) {
            if (aa < d) {
                if (r >= aa && r < d) f[1 + d + off + r - aa] = mcol[aa];
                off += d - aa;
            }
        }
        for (int e = nv + lane; e < KQ - 1 - d; e += 32) f[1 + d + e] = 0.0;
    }
}

// W [KQ][MP] (forward coefficients) and G [MP][KQ] (plain monomials of p_j for the back-projection GEMM)
__global__ void gc_basis_kernel(Params P, int KQ, int KN, double* __restrict__ W, double* __restrict__ G) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = P.d, MP = P.MP;
    if (j >= MP) return;
    const bool in = j < P.m;
    W[j] = in ? 1.0 : 0.0;
    G[static_cast<int64_t>(j) * KN] = in ? 1.0 : 0.0;
    int idx = 1 + d;
    for (int a = 0; a < d; ++a) {
        const double pa = in ? P.Pt[a * MP + j] : 0.0;
        W[static_cast<int64_t>(1 + a) * MP + j] = pa;
        G[static_cast<int64_t>(j) * KN + 1 + a] = pa;
        for (int b = a; b < d; ++b, ++idx) {
            const double pb = in ? P.Pt[b * MP + j] : 0.0;
            W[static_cast<int64_t>(idx) * MP + j] = (a == b) ? -0.5 * pa * pa : -pa * pb;
            G[static_cast<int64_t>(j) * KN + idx] = pa * pb;
        }
    }
    for (int e = idx; e < KQ; ++e) W[static_cast<int64_t>(e) * MP + j] = 0.0;
    for (int e = idx; e < KN; ++e) G[static_cast<int64_t>(j) * KN + e] = 0.0;
}

// features of rows [r0, r1) into R.gcF (row r0 at index 0) and the basis tables R.gcW / R.gcG; phi.cu then runs PHI = exp(F W)
int gc_features(const Params& P, const RowData& R, int64_t r0, int64_t r1, cudaStream_t st, int64_t* launches) {
    const int KQ = gc_feature_width(P.d);
    const int64_t rows = r1 - r0;
    if (rows > R.gc_chunk) {
        set_error("gc_features: chunk of %lld rows exceeds the feature buffer (%lld)", static_cast<long long>(rows),
                  static_cast<long long>(R.gc_chunk));
        return GPZ_ERR_USAGE;
    }
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int64_t want = ceil_div(rows, 4);
    const unsigned nb = static_cast<unsigned>(want < 4LL * sms ? want : 4LL * sms);
    if (P.d <= 8) gc_features_kernel<8><<<nb, 128, sizeof(double) * 8 * 9, st>>>(P, R.X, R.Psi, R.n, r0, r1, KQ, R.gcF);
    else if (P.d <= 16) gc_features_kernel<16><<<nb, 128, sizeof(double) * 16 * 17, st>>>(P, R.X, R.Psi, R.n, r0, r1, KQ, R.gcF);
    else gc_features_kernel<32><<<nb, 128, sizeof(double) * 32 * 33, st>>>(P, R.X, R.Psi, R.n, r0, r1, KQ, R.gcF);
    GPZ_KERNEL_CHECK();
    gc_basis_kernel<<<static_cast<unsigned>(ceil_div(P.MP, 128)), 128, 0, st>>>(P, KQ, gc_gwidth(P.d), R.gcW, R.gcG);
    GPZ_KERNEL_CHECK();
    *launches += 2;
    return GPZ_OK;
}

// one warp per row at a time, lane = matrix column b; per-warp accumulators  acc[a] (column b = lane) of
//   sum_i [ a_i z z' - z y' - y z' + M V M - a_i M ],  y = M u        and (lane 0) s = sum_i a_i
// smem per warp: M (d x d), V then T = M V (d x d), u, z.   partial: [warps_total][d*d + 1]
template <int DMAX>
__global__ void __launch_bounds__(128)
gc_rows_backproj_kernel(int d, int KQ, int KN, int64_t rows, const double* __restrict__ F, const double* __restrict__ G1,
                        double* __restrict__ partial, int accumulate) {
    extern __shared__ double gc_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + warp, nw = gridDim.x * wpb;
    double* Ms = gc_sm + static_cast<int64_t>(warp) * (2 * DMAX * DMAX + 2 * DMAX);
    double* Vs = Ms + DMAX * DMAX;
    double* us = Vs + DMAX * DMAX;
    double* zs = us + DMAX;
    double acc[DMAX];
#pragma unroll
    for (int a = 0; a < DMAX; ++a) acc[a] = 0.0;
    double ssum = 0.0;
    const int nv = d * (d + 1) / 2;
    for (int e = lane; e < 2 * DMAX * DMAX; e += 32) Ms[e] = 0.0;      // M and V: rows / columns >= d stay zero (the products run to DMAX)
    for (int64_t i = gw; i < rows; i += nw) {
        const double* f = F + i * KQ;
        const double* g = G1 + i * KN;
        __syncwarp();
        for (int e = lane; e < nv; e += 32) {          // unpack the packed symmetric matrices
            int a = 0, rem = e;
            while (rem >= d - a) { rem -= d - a; ++a; }
            const int b = a + rem;
            const double mv = f[1 + d + e], vv = g[1 + d + e];
            Ms[a * DMAX + b] = Ms[b * DMAX + a] = mv;
            Vs[a * DMAX + b] = Vs[b * DMAX + a] = vv;
        }
        if (lane < d) {
            us[lane] = g[1 + lane];
            zs[lane] = f[1 + lane];
        }
        const double ai = g[0];
        __syncwarp();
        // Both d x d x d products keep the lane's operand COLUMN in registers and read the other operand as warp-wide broadcasts
        // (two values per 16-byte load): one shared-memory instruction per two FMAs.  The first form read both operands from
        // shared memory for every FMA -- 4 096 LDS.64 per row, i.e. 1 MB of shared-memory traffic per row, which is what bounded
        // the kernel (10.9 ms for 250 000 rows at d = 32, profiles/r02u_cfg5_launches_partial.md).
        double tcol[DMAX];
        double yb = 0.0;
        if (lane < d) {
            double vcol[DMAX];
#pragma unroll
            for (int c = 0; c < DMAX; ++c) vcol[c] = (c < d) ? Vs[c * DMAX + lane] : 0.0;
            for (int c = 0; c < d; ++c) yb += Ms[lane * DMAX + c] * us[c];          // y_b, b = lane
#pragma unroll 2
            for (int a = 0; a < DMAX; ++a) {                                         // T[a][b] = sum_c M[a][c] V[c][b]
                double s = 0.0;
                if (a < d) {
                    const double2* mrow = reinterpret_cast<const double2*>(Ms + a * DMAX);
#pragma unroll
                    for (int c2 = 0; c2 < DMAX / 2; ++c2) {
                        const double2 mv = mrow[c2];                                 // broadcast; columns >= d of M are 0
                        s = fma(mv.x, vcol[2 * c2], s);
                        s = fma(mv.y, vcol[2 * c2 + 1], s);
                    }
                }
                tcol[a] = s;
            }
        }
        __syncwarp();
        if (lane < d)
            for (int a = 0; a < d; ++a) Vs[a * DMAX + lane] = tcol[a];               // V <- T
        // y as a shared vector for the z y' term
        __syncwarp();
        if (lane < d) us[lane] = yb;                                                 // u <- y
        __syncwarp();
        if (lane < d) {
            const double zb = zs[lane];
            double mcol[DMAX];
#pragma unroll
            for (int c = 0; c < DMAX; ++c) mcol[c] = (c < d) ? Ms[c * DMAX + lane] : 0.0;
#pragma unroll 2
            for (int a = 0; a < DMAX; ++a) {
                if (a < d) {
                    double r = 0.0;
                    const double2* trow = reinterpret_cast<const double2*>(Vs + a * DMAX);
#pragma unroll
                    for (int c2 = 0; c2 < DMAX / 2; ++c2) {
                        const double2 tv = trow[c2];                                 // (T M)[a][b]; columns >= d of T are 0
                        r = fma(tv.x, mcol[2 * c2], r);
                        r = fma(tv.y, mcol[2 * c2 + 1], r);
                    }
                    acc[a] += ai * zs[a] * zb - zs[a] * yb - us[a] * zb + r - ai * mcol[a];
                }
            }
        }
        ssum += ai;
    }
    double* out = partial + static_cast<int64_t>(gw) * (d * d + 1);
    if (lane < d)
        for (int a = 0; a < d; ++a) out[a * d + lane] = (accumulate ? out[a * d + lane] : 0.0) + acc[a];
    if (lane == 0) out[d * d] = (accumulate ? out[d * d] : 0.0) + ssum;
}

// single block: B = 1/2 (s A + sum of the partials); dGamma = -2 Gamma Sigma B Sigma  (GC: one d x d matrix, theta layout
// dG[c + a d]); then thread per basis: dP_j = R2[j][1..d] - mat(R2[j][1+d..]) p_j
__global__ void __launch_bounds__(256)
gc_finish_kernel(Params P, int KQ, const double* __restrict__ partial, int nw, const double* __restrict__ R2, double* __restrict__ dP,
                 double* __restrict__ dG, double* __restrict__ work) {
    const int d = P.d, MP = P.MP, m = P.m, dp = P.dp;
    const int tid = threadIdx.x;
    double* B = work;                 // [d*d]
    double* T = work + d * d;         // [d*d]
    // partial: the per-warp partials already summed over the warps by gc_partial_sum_kernel -> [d*d + 1]
    const double s_tot = partial[d * d];
    for (int e = tid; e < d * d; e += 256) {
        const int a = e / d, b = e % d;
        B[e] = 0.5 * (s_tot * P.Aj[(static_cast<int64_t>(a) * d + b) * MP] + partial[e]);
    }
    __syncthreads();
    for (int e = tid; e < d * d; e += 256) {          // T = B Sigma
        const int a = e / d, b = e % d;
        double s = 0.0;
        for (int c = 0; c < d; ++c) s += B[a * d + c] * P.Sj[(static_cast<int64_t>(c) * d + b) * MP];
        T[e] = s;
    }
    __syncthreads();
    for (int e = tid; e < d * d; e += 256) {          // B = Sigma T
        const int a = e / d, b = e % d;
        double s = 0.0;
        for (int c = 0; c < d; ++c) s += P.Sj[(static_cast<int64_t>(a) * d + c) * MP] * T[c * d + b];
        B[e] = s;
    }
    __syncthreads();
    for (int e = tid; e < d * d; e += 256) {          // dGamma(c,a) = -2 sum_b Gamma(c,b) B(b,a)
        const int a = e / d, c = e % d;
        double s = 0.0;
        for (int b = 0; b < d; ++b) s += P.Gam[(static_cast<int64_t>(c) * dp + b) * MP] * B[b * d + a];
        dG[c + a * d] = -2.0 * s;
    }
}

// out[e] = sum over the nw per-warp partials of entry e, in warp order per thread slice and a fixed-order block sum
__global__ void __launch_bounds__(256)
gc_partial_sum_kernel(const double* __restrict__ partial, int nw, int stride, double* __restrict__ out) {
    __shared__ double sh[8];
    const int e = blockIdx.x;
    double v = 0.0;
    for (int w = threadIdx.x; w < nw; w += 256) v += partial[static_cast<int64_t>(w) * stride + e];
    v = block_sum<256>(v, sh);
    if (threadIdx.x == 0) out[e] = v;
}

// dP_j = R2[j][1..d] - mat(R2[j][1+d..]) p_j : one thread per (basis, dimension) (it was a loop over the bases inside the single
// block above: 2.7 ms of a 105 ms evaluation at m = 2000, d = 32)
__global__ void __launch_bounds__(256)
gc_dp_kernel(Params P, int KQ, const double* __restrict__ R2, double* __restrict__ dP) {
    const int d = P.d, MP = P.MP, m = P.m;
    const int64_t e = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (e >= static_cast<int64_t>(m) * d) return;
    const int j = static_cast<int>(e % m), a = static_cast<int>(e / m);
    const double* r = R2 + static_cast<int64_t>(j) * KQ;
    double v = r[1 + a];
    for (int b = 0; b < d; ++b) {
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        const int idx = 1 + d + lo * d - lo * (lo - 1) / 2 + (hi - lo);
        v -= r[idx] * P.Pt[b * MP + j];
    }
    dP[a * m + j] = v;
}

int64_t gc_backproj_ws_doubles(const Params& P, int64_t chunk_rows, int nsplit, int sm_count) {
    const int KQ = gc_feature_width(P.d);
    const int64_t nw = static_cast<int64_t>(sm_count) * GC_BP_WARPS;
    return chunk_rows * gc_gwidth(P.d) /*G1*/ + static_cast<int64_t>(nsplit) * (P.MP / TILE) * (KQ / 32) * TILE * 32 /*atb partial*/ +
           static_cast<int64_t>(P.MP) * KQ /*R2*/ + nw * (P.d * P.d + 1) + 3LL * P.d * P.d + 8;
}

// one chunk of rows: dPhi row 0 = row r0; R.gcF holds this chunk's features (row r0 at index 0)
int gc_backproj(const Params& P, const RowData& R, int64_t r0, int64_t r1, const double* dPhi, int64_t ld, double* ws, int nsplit,
                int sm_count, int accumulate, int last, cudaStream_t st, int64_t* launches) {
    const int KQ = gc_feature_width(P.d), d = P.d;
    const int64_t rows = r1 - r0;
    double* G1 = ws;
    double* atbp = G1 + R.gc_chunk * gc_gwidth(P.d);
    double* R2 = atbp + static_cast<int64_t>(nsplit) * (P.MP / TILE) * (KQ / 32) * TILE * 32;
    double* partial = R2 + static_cast<int64_t>(P.MP) * KQ;
    int rc;
    const int KN = gc_gwidth(d);
    if (R.gc_digits > 0) {
        // both products as error-free digit GEMMs.  Row digits of dPHI first: their exponents also bound the moment GEMM's operand
        const int kvalid = 1 + d + d * (d + 1) / 2, K128 = static_cast<int>(round_up(KQ, 128));
        if ((rc = ozaki_row_digits(dPhi, ld, P.m, P.MP, rows, R.gc_digits, R.gcA8, R.gcEa, R.flag, st, launches))) return rc;
        // moment GEMM  R2 = dPHI' F  (contraction over the rows, MN-major operands like the Gram)
        if ((rc = ozaki_row_digits(R.gcF, KQ, kvalid, K128, rows, R.gc_digits, R.gcF8, R.gcEaF, R.flag, st, launches))) return rc;
        if ((rc = ozaki_moment_gemm(dPhi, ld, P.m, P.MP, rows, R.gcEa, R.gcF8, R.gcEaF, K128, KQ, R.gc_digits, R.gcX8, accumulate, R2, KQ,
                                    R.gcMws, st, launches))) return rc;
    } else if ((rc = atb_general(dPhi, ld, P.MP, R.gcF, KQ, KQ, R.gc_ones, 0, rows, nsplit, atbp, accumulate, last, R2, st, launches))) return rc;
    // G1 = dPHI G   (rows x KQ)
    if (R.gc_digits > 0) {               // K = m: the T-GEMM's engine, results stored directly
        const int kvalid = 1 + d + d * (d + 1) / 2;
        if ((rc = ozaki_transpose_digits(R.gcG, KN, P.m, kvalid, P.MP, KN, R.gc_digits, R.gcGD8, R.gcEbG, R.flag, st, launches))) return rc;
        if ((rc = ozmma_gemm_rows(R.gcA8, R.gcEa, rows, R.gcGD8, R.gcEbG, kvalid, P.MP, R.gc_digits, G1, KN, st, launches))) return rc;
    } else if ((rc = gemm_rows(dPhi, ld, static_cast<int>(round_up(P.m, KSTEP)), R.gcG, KN, rows, G1, st, launches))) return rc;
    // 4 warps per block, 3 blocks per SM: at d = 32 a warp needs 16.9 KB of shared memory, so 12 warps per SM fit where one
    // 8-warp block did (r02u launch list: 11 ms of a 105 ms evaluation in this kernel, latency-bound)
    const int nblk = sm_count * (GC_BP_WARPS / 4);
    const size_t smem = sizeof(double) * 4 * (2 * 32 * 32 + 2 * 32);
    if (d <= 8) {
        gc_rows_backproj_kernel<8><<<nblk, 128, sizeof(double) * 4 * (2 * 8 * 8 + 2 * 8), st>>>(d, KQ, KN, rows, R.gcF, G1, partial, accumulate);
    } else if (d <= 16) {
        gc_rows_backproj_kernel<16><<<nblk, 128, sizeof(double) * 4 * (2 * 16 * 16 + 2 * 16), st>>>(d, KQ, KN, rows, R.gcF, G1, partial, accumulate);
    } else {
        static PerDeviceOnce once;
        if (once.need()) {
            GPZ_CUDA(cudaFuncSetAttribute(gc_rows_backproj_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        }
        gc_rows_backproj_kernel<32><<<nblk, 128, smem, st>>>(d, KQ, KN, rows, R.gcF, G1, partial, accumulate);
    }
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

int gc_backproj_finish(const Params& P, const RowData& R, double* ws, int nsplit, int sm_count, double* dP, double* dG, cudaStream_t st,
                       int64_t* launches) {
    const int KQ = gc_feature_width(P.d);
    double* G1 = ws;
    double* atbp = G1 + R.gc_chunk * gc_gwidth(P.d);
    double* R2 = atbp + static_cast<int64_t>(nsplit) * (P.MP / TILE) * (KQ / 32) * TILE * 32;
    double* partial = R2 + static_cast<int64_t>(P.MP) * KQ;
    double* work = partial + static_cast<int64_t>(sm_count) * GC_BP_WARPS * (P.d * P.d + 1);
    double* sums = work + 2LL * P.d * P.d;
    gc_partial_sum_kernel<<<P.d * P.d + 1, 256, 0, st>>>(partial, sm_count * GC_BP_WARPS, P.d * P.d + 1, sums);
    ++*launches;
    gc_finish_kernel<<<1, 256, 0, st>>>(P, KQ, sums, sm_count * GC_BP_WARPS, R2, dP, dG, work);
    gc_dp_kernel<<<static_cast<unsigned>(ceil_div(static_cast<int64_t>(P.m) * P.d, 256)), 256, 0, st>>>(P, KQ, R2, dP);
    GPZ_KERNEL_CHECK();
    *launches += 2;
    return GPZ_OK;
}

}  // namespace gpz
