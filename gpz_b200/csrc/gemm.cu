// fp64 tensor-core (DMMA.8x8x4) GEMM kernels of the GPz hot path.
//
//   atb_kernel    C[j,c] = sum_i wgt_i A[i,j] B[i,c]     K = rows (split-K, deterministic partials)
//                 - SYRK form (B == A == PHI): the weighted Gram PHI' diag(omega*beta) PHI, GPz/GPz.m:63-65
//                 - general form (A = dPHI, B = row features): back-projection contractions, GPz.m:133-213
//   tgemm_kernel  T = PHI * iSigma with a fused epilogue: H += rw_i * PHI_ij * T_ij and the row
//                 partial sums nu_i = sum_j PHI_ij T_ij            GPz/GPz.m:69,72 (predictDiag.m:66-68)
//   sgemm_kernel  small strided GEMM used by the blocked Cholesky / inverse (solve.cu)
//
// Operand tiles are staged global -> shared with cp.async in a 4-stage pipeline; the smem strides
// (132 / 20 doubles) make every DMMA fragment load conflict-free.  All reductions have a fixed order,
// so repeated evaluations at the same theta are bit-identical.
#include <cuda.h>

#include "internal.cuh"

namespace gpz {

constexpr int STAGES = 4;
int g_phi_persist = 2;     // PHI = exp(F W) through the persistent column-stationary kernel where it applies ("phi_persist" option): 2 staggered M-groups (default), 1 CTA-synchronous, 0 off
int g_moment_warps = 0;     // warps per CTA of the back-projection moment GEMM ("moment_warps" option: 0 = by shape, 8 or 16)
int g_gemm_warps = 0;       // 0 = defaults (Gram / T-GEMM 8 warps, PHI build 16 warps: its exp epilogue likes more warps);
                            // 8 / 16 force all three (gpz_set_option "gemm_warps").  8 vs 16 differ by <4 % either way across boxes.

// ------------------------------------------------------------------------------------------------
// C = A' diag(w) B over a row range, 128 x TN output tile per CTA, split over rows
// ------------------------------------------------------------------------------------------------
template <int WARPS_M, int WARPS_N, bool SYRK>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, 1)
atb_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
           const double* __restrict__ wgt, int64_t row0, int64_t row1, int64_t rows_per_split,
           int ntiles_n, double* __restrict__ partial, int accumulate) {
    constexpr int NTHR = WARPS_M * WARPS_N * 32;
    constexpr int TN = WARPS_N * 32;
    constexpr int LDB = TN + 4;
    constexpr int WM = TILE / WARPS_M;   // warp tile rows
    constexpr int MT = WM / 8;           // 8-row DMMA tiles per warp
    constexpr int NT = 4;                // 8-col DMMA tiles per warp (warp tile is WM x 32)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);                    // [STAGES][KSTEP][LDT]
    double* Bs = As + STAGES * KSTEP * LDT;                              // [STAGES][KSTEP][LDB]
    double* Ws = Bs + STAGES * KSTEP * LDB;                              // [STAGES][KSTEP]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / WARPS_N) * WM;
    const int wn0 = (warp % WARPS_N) * 32;

    int jt, ct;
    const int tile = blockIdx.x;
    if (SYRK) {
        jt = static_cast<int>((sqrtf(8.0f * tile + 1.0f) - 1.0f) * 0.5f);
        while ((jt + 1) * (jt + 2) / 2 <= tile) ++jt;
        while (jt * (jt + 1) / 2 > tile) --jt;
        ct = tile - jt * (jt + 1) / 2;
    } else {
        jt = tile / ntiles_n;
        ct = tile % ntiles_n;
    }
    const bool alias = SYRK && (jt == ct);     // diagonal Gram tile: both operands are the same tile
    const int64_t a_col0 = static_cast<int64_t>(jt) * TILE;
    const int64_t b_col0 = static_cast<int64_t>(ct) * TN;

    const int64_t rbeg = row0 + static_cast<int64_t>(blockIdx.y) * rows_per_split;
    int64_t rend = rbeg + rows_per_split;
    if (rend > row1) rend = row1;
    const int64_t nrows = rend > rbeg ? rend - rbeg : 0;
    const int nk = static_cast<int>((nrows + KSTEP - 1) / KSTEP);

    auto load_stage = [&](int kt, int stage) {
        const int64_t rb = rbeg + static_cast<int64_t>(kt) * KSTEP;
        double* as = As + stage * KSTEP * LDT;
        double* bs = Bs + stage * KSTEP * LDB;
#pragma unroll
        for (int q = 0; q < (KSTEP * TILE / 2) / NTHR; ++q) {
            const int c = tid + q * NTHR;
            const int r = c >> 6, cc = c & 63;
            const int64_t gr = rb + r;
            const bool ok = gr < rend;
            const double* src = A + (ok ? gr : row0) * lda + a_col0 + cc * 2;
            cp_async16(as + r * LDT + cc * 2, src, ok ? 16 : 0);
        }
        if (!alias) {
            constexpr int BCH = KSTEP * TN / 2;      // 16-byte chunks in the B tile
#pragma unroll
            for (int q = 0; q < (BCH + NTHR - 1) / NTHR; ++q) {
                const int c = tid + q * NTHR;
                if (c < BCH) {
                    const int r = c / (TN / 2), cc = c % (TN / 2);
                    const int64_t gr = rb + r;
                    const bool ok = gr < rend;
                    const double* src = B + (ok ? gr : row0) * ldb + b_col0 + cc * 2;
                    cp_async16(bs + r * LDB + cc * 2, src, ok ? 16 : 0);
                }
            }
        }
        if (tid < KSTEP) {
            const int64_t gr = rb + tid;
            const bool ok = gr < rend;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(Ws + stage * KSTEP + tid)),
                         "l"(wgt + (ok ? gr : row0)), "r"(ok ? 8 : 0));
        }
    };

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = kt + STAGES - 1;
            if (nx < nk) load_stage(nx, nx % STAGES);
            cp_async_commit();
        }
        const int stage = kt % STAGES;
        const double* as = As + stage * KSTEP * LDT;
        const double* bs = alias ? as : (Bs + stage * KSTEP * LDB);
        const int ldb_s = alias ? LDT : LDB;
        const double* wsm = Ws + stage * KSTEP;
#pragma unroll
        for (int kk = 0; kk < KSTEP / 4; ++kk) {
            const int kr = kk * 4 + t;
            const double wv = wsm[kr];
            double bf[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = bs[kr * ldb_s + wn0 + j * 8 + g] * wv;
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const double af = as[kr * LDT + wm0 + i * 8 + g];
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
            }
        }
    }
    cp_async_wait<0>();

    double* out = partial + (static_cast<int64_t>(blockIdx.y) * gridDim.x + tile) * (TILE * TN);
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int r = wm0 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int c = wn0 + j * 8 + 2 * t;
            double2* p = reinterpret_cast<double2*>(out + r * TN + c);
            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
            if (accumulate) {
                const double2 o = *p;
                v.x += o.x;
                v.y += o.y;
            }
            *p = v;
        }
    }
}

// sum the split-K partials in fixed order and scatter the tiles into C (row-major, ldc)
template <bool SYRK>
__global__ void atb_reduce_kernel(const double* __restrict__ partial, int nsplit, int ntiles, int ntiles_n, int TN,
                                  double* __restrict__ C, int64_t ldc) {
    const int tile = blockIdx.x;
    int jt, ct;
    if (SYRK) {
        jt = static_cast<int>((sqrtf(8.0f * tile + 1.0f) - 1.0f) * 0.5f);
        while ((jt + 1) * (jt + 2) / 2 <= tile) ++jt;
        while (jt * (jt + 1) / 2 > tile) --jt;
        ct = tile - jt * (jt + 1) / 2;
    } else {
        jt = tile / ntiles_n;
        ct = tile % ntiles_n;
    }
    // blockIdx.y splits the tile's elements, so that few-tile problems (small m) still fill the machine
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < TILE * TN; e += blockDim.x * gridDim.y) {
        double s = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) s += partial[(static_cast<int64_t>(sp) * ntiles + tile) * (TILE * TN) + e];
        const int r = e / TN, c = e % TN;
        const int64_t gr = static_cast<int64_t>(jt) * TILE + r, gc = static_cast<int64_t>(ct) * TN + c;
        C[gr * ldc + gc] = s;
        if (SYRK && jt != ct) C[gc * ldc + gr] = s;
    }
}

static dim3 atb_reduce_grid(int ntiles, int TN) {
    const int per_tile = static_cast<int>(ceil_div(static_cast<int64_t>(TILE) * TN, 256));
    const int want = static_cast<int>(ceil_div(592, ntiles));
    return dim3(ntiles, want < per_tile ? want : per_tile);
}

template <int WARPS_M, int WARPS_N, bool SYRK>
static int launch_atb(const double* A, int64_t lda, const double* B, int64_t ldb, const double* wgt, int64_t row0,
                      int64_t row1, int ntiles, int ntiles_n, int nsplit, double* partial, int accumulate,
                      cudaStream_t st) {
    constexpr int TN = WARPS_N * 32;
    const size_t smem = sizeof(double) * (STAGES * KSTEP * LDT + STAGES * KSTEP * (TN + 4) + STAGES * KSTEP);
    static PerDeviceOnce once;
    if (once.need()) {
        GPZ_CUDA(cudaFuncSetAttribute(atb_kernel<WARPS_M, WARPS_N, SYRK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    }
    int64_t rps = ceil_div(row1 - row0, nsplit);
    rps = round_up(rps > 0 ? rps : 1, KSTEP);
    dim3 grid(ntiles, nsplit);
    atb_kernel<WARPS_M, WARPS_N, SYRK><<<grid, WARPS_M * WARPS_N * 32, smem, st>>>(A, lda, B, ldb, wgt, row0, row1, rps, ntiles_n,
                                                                                  partial, accumulate);
    GPZ_KERNEL_CHECK();
    return GPZ_OK;
}

// Gram: S(MP x MP, both triangles) = PHI' diag(w) PHI over rows [row0,row1).
// partial must hold nsplit * ntri * 128*128 doubles; the finish step sums the split partials (fixed order).
int gram_syrk_main(const double* Phi, int64_t ld, int MP, const double* wgt, int64_t row0, int64_t row1, int nsplit,
                   double* partial, int accumulate, cudaStream_t st, int64_t* launches) {
    const int T = MP / TILE;
    const int ntri = T * (T + 1) / 2;
    int rc = g_gemm_warps == 16 ? launch_atb<4, 4, true>(Phi, ld, Phi, ld, wgt, row0, row1, ntri, T, nsplit, partial, accumulate, st)
                                : launch_atb<2, 4, true>(Phi, ld, Phi, ld, wgt, row0, row1, ntri, T, nsplit, partial, accumulate, st);
    if (rc) return rc;
    ++*launches;
    return GPZ_OK;
}

int gram_syrk_finish(const double* partial, int nsplit, int MP, int reduce, double* S, cudaStream_t st, int64_t* launches) {
    if (!reduce) return GPZ_OK;
    const int T = MP / TILE;
    const int ntri = T * (T + 1) / 2;
    atb_reduce_kernel<true><<<atb_reduce_grid(ntri, TILE), 256, 0, st>>>(partial, nsplit, ntri, T, TILE, S, MP);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

int gram_nsplit(int MP, int sm_count) {
    const int T = MP / TILE;
    const int ntri = T * (T + 1) / 2;
    int ns = sm_count / ntri;
    return ns < 1 ? 1 : ns;
}

// R(MP x QP) = A' B over rows [row0,row1); QP multiple of 32.
int atb_general(const double* A, int64_t lda, int MP, const double* B, int64_t ldb, int QP, const double* wgt,
                int64_t row0, int64_t row1, int nsplit, double* partial, int accumulate, int reduce, double* R,
                cudaStream_t st, int64_t* launches) {
    const int tm = MP / TILE, tn = QP / 32;
    int rc = launch_atb<8, 1, false>(A, lda, B, ldb, wgt, row0, row1, tm * tn, tn, nsplit, partial, accumulate, st);
    if (rc) return rc;
    ++*launches;
    if (reduce) {
        atb_reduce_kernel<false><<<atb_reduce_grid(tm * tn, 32), 256, 0, st>>>(partial, nsplit, tm * tn, tn, 32, R, QP);
        GPZ_KERNEL_CHECK();
        ++*launches;
    }
    return GPZ_OK;
}

// ------------------------------------------------------------------------------------------------
// row-tile GEMM  C(128 x 128 tile) = A[rows][K] * B[K][cols]  with two fused epilogues
//   EPI_T   (A = PHI, B = iSigma):  h = PHI_ij * T_ij ; nu partial = sum_j h ; H (+)= rw_i * h      GPz.m:69,72
//   EPI_PHI (A = row features F, B = per-basis coefficients W):  PHI_ij = exp(F_i . W_j)             getPHI.m:60-113
//           (the quadratic form -1/2 (x-p)' A_j (x-p) expanded in the monomials of x), PHI stored row-major,
//           plus up to two fused row-dot partials  sum_j PHI_ij vec_q[j]  (ln-noise head getPHI.m:124, PHI*w)
// ------------------------------------------------------------------------------------------------
struct TEpi {
    const double* Phi;      // [rows][ld] (same matrix as A)
    const double* rw;       // row weights omega*beta (may be null)
    double* H;              // may be null
    int accumulate;
    double* nupart;         // [ntn][nu_ld]
    int64_t nu_ld;
    int aug_col;            // >= 0: column that carries y in PHI and w in B; its T value is PHI_i.w (GPz.m:77)
    double* pred;           // [rows] receives T[:, aug_col]
};
struct PEpi {
    int m;
    double* Phi;            // [rows][MP] (may be null)
    int ndot;
    const double* vec[2];
    double* part[2];        // [ntn][part_ld]
    int64_t part_ld;
    const double* ycol;     // non-null: PHI[:, m] := y (spare padded column), so the Gram also yields PHI'(w y) (GPz.m:70)
    int plain;              // != 0: store the product itself (no exp): C = A B, used by the GC+Psi back-projection (gcpsi.cu)
};

template <int EPI, int WARPS_M>
__global__ void __launch_bounds__(WARPS_M * 128, 1)
tgemm_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int MP, int nk, int klast, int64_t n,
             TEpi te, PEpi pe) {
    constexpr int NTHR = WARPS_M * 128;
    constexpr int WM = TILE / WARPS_M, MT = WM / 8, NT = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);          // [STAGES][TILE][LDK]
    double* Bs = As + STAGES * TILE * LDK;                     // [STAGES][KSTEP][LDT]
    __shared__ double red[2][4][TILE];
    __shared__ double exp_sm[32];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 2) * WM;
    const int wn0 = (warp & 3) * 32;
    const int ntn = MP / TILE;
    if (EPI == 1) exp_tab_stage(exp_sm);              // visible after the first __syncthreads of the K loop
    const int ct = blockIdx.x % ntn;
    const int64_t rt = blockIdx.x / ntn;
    const int64_t i0 = rt * TILE;
    const int64_t j0 = static_cast<int64_t>(ct) * TILE;

    auto load_stage = [&](int kt, int stage) {
        const int k0 = kt * KSTEP;
        double* as = As + stage * TILE * LDK;
        double* bs = Bs + stage * KSTEP * LDT;
#pragma unroll
        for (int q = 0; q < (TILE * KSTEP / 2) / NTHR; ++q) {
            const int c = tid + q * NTHR;
            const int r = c >> 3, cc = c & 7;
            const int64_t gr = i0 + r;
            const bool ok = gr < n;
            cp_async16(as + r * LDK + cc * 2, A + (ok ? gr : 0) * lda + k0 + cc * 2, ok ? 16 : 0);
        }
#pragma unroll
        for (int q = 0; q < (KSTEP * TILE / 2) / NTHR; ++q) {
            const int c = tid + q * NTHR;
            const int r = c >> 6, cc = c & 63;
            cp_async16(bs + r * LDT + cc * 2, B + static_cast<int64_t>(k0 + r) * MP + j0 + cc * 2, 16);
        }
    };

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = kt + STAGES - 1;
            if (nx < nk) load_stage(nx, nx % STAGES);
            cp_async_commit();
        }
        const int stage = kt % STAGES;
        const double* as = As + stage * TILE * LDK;
        const double* bs = Bs + stage * KSTEP * LDT;
        // the last stage only runs the 4-deep K steps that hold non-zero operand columns (PHI build: q = 21 or 66 features,
        // not a multiple of 16)
        const int kend = (kt == nk - 1) ? klast : KSTEP / 4;
#pragma unroll
        for (int kk = 0; kk < KSTEP / 4; ++kk) {
            if (kk >= kend) break;
            const int kc = kk * 4 + t;
            double bf[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = bs[kc * LDT + wn0 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const double af = as[(wm0 + i * 8 + g) * LDK + kc];
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
            }
        }
    }
    cp_async_wait<0>();

    double rs0[MT], rs1[MT];
    if (EPI == 0) {
        // h = PHI_ij * T_ij ; nu partial over this tile's 128 columns ; H (+)= rw_i * h
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int64_t gi = i0 + wm0 + i * 8 + g;
            const bool ok = gi < n;
            double s = 0.0;
            const double wrow = (te.rw != nullptr && ok) ? te.rw[gi] : 1.0;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int64_t gj = j0 + wn0 + j * 8 + 2 * t;
                if (ok) {
                    const double2 ph = *reinterpret_cast<const double2*>(te.Phi + gi * lda + gj);
                    double h0 = ph.x * acc[i][j][0], h1 = ph.y * acc[i][j][1];
                    if (te.aug_col >= 0) {
                        if (gj == te.aug_col) { te.pred[gi] = acc[i][j][0]; h0 = 0.0; }
                        if (gj + 1 == te.aug_col) { te.pred[gi] = acc[i][j][1]; h1 = 0.0; }
                    }
                    s += h0 + h1;
                    if (te.H != nullptr) {
                        double2* hp = reinterpret_cast<double2*>(te.H + gi * lda + gj);
                        double2 v = make_double2(wrow * h0, wrow * h1);
                        if (te.accumulate) {
                            const double2 o = *hp;
                            v.x += o.x;
                            v.y += o.y;
                        }
                        *hp = v;
                    }
                }
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            rs0[i] = s;
            rs1[i] = 0.0;
        }
    } else {
        double2 v0[NT], v1[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int64_t gj = j0 + wn0 + j * 8 + 2 * t;
            v0[j] = pe.ndot > 0 ? *reinterpret_cast<const double2*>(pe.vec[0] + gj) : make_double2(0.0, 0.0);
            v1[j] = pe.ndot > 1 ? *reinterpret_cast<const double2*>(pe.vec[1] + gj) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int64_t gi = i0 + wm0 + i * 8 + g;
            const bool ok = gi < n;
            double s0 = 0.0, s1 = 0.0;
            double ex[2 * NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                ex[2 * j] = acc[i][j][0];
                ex[2 * j + 1] = acc[i][j][1];
            }
            if (!pe.plain) exp_tab_vec(ex, exp_sm);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int64_t gj = j0 + wn0 + j * 8 + 2 * t;
                double p0 = (gj < pe.m) ? ex[2 * j] : 0.0;
                double p1 = (gj + 1 < pe.m) ? ex[2 * j + 1] : 0.0;
                if (pe.ycol != nullptr && ok) {
                    if (gj == pe.m) p0 = pe.ycol[gi];
                    if (gj + 1 == pe.m) p1 = pe.ycol[gi];
                }
                if (ok && pe.Phi != nullptr) *reinterpret_cast<double2*>(pe.Phi + gi * MP + gj) = make_double2(p0, p1);
                s0 = fma(p0, v0[j].x, fma(p1, v0[j].y, s0));
                s1 = fma(p0, v1[j].x, fma(p1, v1[j].y, s1));
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            rs0[i] = s0;
            rs1[i] = s1;
        }
    }
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            red[0][warp & 3][wm0 + i * 8 + g] = rs0[i];
            if (EPI == 1) red[1][warp & 3][wm0 + i * 8 + g] = rs1[i];
        }
    }
    __syncthreads();
    if (tid < TILE) {
        const int64_t gi = i0 + tid;
        if (gi < n) {
            if (EPI == 0) {
                te.nupart[static_cast<int64_t>(ct) * te.nu_ld + gi] = red[0][0][tid] + red[0][1][tid] + red[0][2][tid] + red[0][3][tid];
            } else {
                if (pe.ndot > 0) pe.part[0][static_cast<int64_t>(ct) * pe.part_ld + gi] = red[0][0][tid] + red[0][1][tid] + red[0][2][tid] + red[0][3][tid];
                if (pe.ndot > 1) pe.part[1][static_cast<int64_t>(ct) * pe.part_ld + gi] = red[1][0][tid] + red[1][1][tid] + red[1][2][tid] + red[1][3][tid];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// PHI = exp(F W), persistent form for short K (the monomial features of d <= 10: K <= 68 after rounding to the DMMA k step).
// The kernel above gives every 128 x 128 tile its own CTA: operand loads start cold for each tile and its exp/store epilogue
// overlaps nothing (one CTA per SM), so the fp64 pipe idles 35-40 % of the time (profiles/r01h, r02e: DMMA 52-57 % + FP64
// 11-12 % active).  Here a CTA keeps ONE column tile of W in shared memory for its whole life and walks down the row tiles
// of that column: the [128][K] row-feature tile of the next row tile arrives by TMA (one cp.async.bulk.tensor per tile,
// double-buffered, completion on an mbarrier) while the current tile is multiplied and exponentiated, and the K loop has
// no barrier inside (the whole K extent is resident).  Shared-memory strides LDK = 4 (mod 8) and 132 keep the DMMA fragment
// loads conflict-free.  Same arithmetic, same order as tgemm_kernel<1>: bit-identical results.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pp_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pp_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pp_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (!ok && clock64() - t0 > 20000000000LL) __trap();      // a protocol bug fails the launch instead of hanging the GPU
    }
}
__device__ __forceinline__ void pp_tma_2d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}

// exp + store + row-dot partials of one warp's 32 x 32 accumulator block (both persistent kernels).  NDOT row dots are compiled in
// (the training sweep needs one, PHI v; validation / predict two), and EDGE = false is the interior tile: no row / column masks, no
// y column.  One exp_tab_vec call per 8 accumulators keeps their fp64 chains interleaved.
template <int NDOT, bool EDGE, int MT, int NT>
__device__ __forceinline__ void phi_epilogue(const double (&acc)[MT][NT][2], const PEpi& pe, int MP, int64_t n, int64_t i0, int64_t j0,
                                             int wm0, int wn0, int g, int t, const double2 (&v0)[NT], const double2 (&v1)[NT],
                                             const double* __restrict__ exp_sm, double (&rs0)[MT], double (&rs1)[MT]) {
    double* prow = pe.Phi != nullptr ? pe.Phi + (i0 + wm0 + g) * MP + j0 + wn0 + 2 * t : nullptr;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int64_t gi = i0 + wm0 + i * 8 + g;
        const bool ok = !EDGE || gi < n;
        double ex[2 * NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            ex[2 * j] = acc[i][j][0];
            ex[2 * j + 1] = acc[i][j][1];
        }
        exp_tab_vec(ex, exp_sm);
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            double p0 = ex[2 * j], p1 = ex[2 * j + 1];
            if (EDGE) {
                const int64_t gj = j0 + wn0 + j * 8 + 2 * t;
                if (gj >= pe.m) p0 = 0.0;
                if (gj + 1 >= pe.m) p1 = 0.0;
                if (pe.ycol != nullptr && ok) {
                    if (gj == pe.m) p0 = pe.ycol[gi];
                    if (gj + 1 == pe.m) p1 = pe.ycol[gi];
                }
            }
            if (ok && prow != nullptr) *reinterpret_cast<double2*>(prow + static_cast<int64_t>(i) * 8 * MP + j * 8) = make_double2(p0, p1);
            if (NDOT > 0) s0 = fma(p0, v0[j].x, fma(p1, v0[j].y, s0));
            if (NDOT > 1) s1 = fma(p0, v1[j].x, fma(p1, v1[j].y, s1));
        }
        if (NDOT > 0) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        }
        if (NDOT > 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        }
        rs0[i] = s0;
        rs1[i] = s1;
    }
}

struct PhiPersist {
    const double* W;        // [K][MP]
    int MP, K4, LDK, ntn;
    int64_t n, tiles_m;
    PEpi pe;
};

template <int NDOT>
__global__ void __launch_bounds__(512, 1)
phi_persist_kernel(const __grid_constant__ CUtensorMap mapF, const PhiPersist a) {
    constexpr int WARPS_M = 4, WM = TILE / WARPS_M, MT = WM / 8, NT = 4;
    extern __shared__ __align__(128) unsigned char pp_smem[];
    double* Ws = reinterpret_cast<double*>(pp_smem);                       // [K4][LDT]
    double* As = Ws + a.K4 * LDT;                                           // [2][TILE][LDK]
    __shared__ double red[2][4][TILE];
    __shared__ double exp_sm[32];
    __shared__ __align__(8) unsigned long long bars[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 2) * WM, wn0 = (warp & 3) * 32;
    const int ct = blockIdx.x % a.ntn;
    const int64_t step = gridDim.x / a.ntn;
    const int64_t j0 = static_cast<int64_t>(ct) * TILE;
    const int LDK = a.LDK;
    const uint32_t tile_bytes = static_cast<uint32_t>(TILE * LDK * sizeof(double));
    const PEpi& pe = a.pe;

    for (int c = tid; c < a.K4 * (TILE / 2); c += 512) {                    // this CTA's column tile of W, once
        const int r = c >> 6, cc = c & 63;
        cp_async16(Ws + r * LDT + cc * 2, a.W + static_cast<int64_t>(r) * a.MP + j0 + cc * 2, 16);
    }
    cp_async_commit();
    exp_tab_stage(exp_sm);
    if (tid == 0) {
        pp_mbar_init(smem_u32(&bars[0]), 1);
        pp_mbar_init(smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double2 v0[NT], v1[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int64_t gj = j0 + wn0 + j * 8 + 2 * t;
        v0[j] = NDOT > 0 ? *reinterpret_cast<const double2*>(pe.vec[0] + gj) : make_double2(0.0, 0.0);
        v1[j] = NDOT > 1 ? *reinterpret_cast<const double2*>(pe.vec[1] + gj) : make_double2(0.0, 0.0);
    }
    cp_async_wait<0>();
    __syncthreads();
    int64_t rt = blockIdx.x / a.ntn;
    if (tid == 0 && rt < a.tiles_m) {
        pp_mbar_expect(smem_u32(&bars[0]), tile_bytes);
        pp_tma_2d(&mapF, smem_u32(As), smem_u32(&bars[0]), 0, static_cast<int>(rt * TILE));
    }
    for (uint32_t it = 0; rt < a.tiles_m; rt += step, ++it) {
        const uint32_t buf = it & 1u;
        if (tid == 0 && rt + step < a.tiles_m) {                             // next row tile into the other buffer (freed by the
            pp_mbar_expect(smem_u32(&bars[buf ^ 1u]), tile_bytes);           // barrier that ended the previous iteration)
            pp_tma_2d(&mapF, smem_u32(As + (buf ^ 1u) * TILE * LDK), smem_u32(&bars[buf ^ 1u]), 0, static_cast<int>((rt + step) * TILE));
        }
        pp_mbar_wait(smem_u32(&bars[buf]), (it >> 1) & 1u);
        const double* as = As + buf * TILE * LDK;
        double acc[MT][NT][2];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kk = 0; kk < a.K4 / 4; ++kk) {
            const int kc = kk * 4 + t;
            double bf[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = Ws[kc * LDT + wn0 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const double af = as[(wm0 + i * 8 + g) * LDK + kc];
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
            }
        }
        const int64_t i0 = rt * TILE;
        double rs0[MT], rs1[MT];
        if (j0 + TILE <= pe.m && i0 + TILE <= a.n)
            phi_epilogue<NDOT, false>(acc, pe, a.MP, a.n, i0, j0, wm0, wn0, g, t, v0, v1, exp_sm, rs0, rs1);
        else
            phi_epilogue<NDOT, true>(acc, pe, a.MP, a.n, i0, j0, wm0, wn0, g, t, v0, v1, exp_sm, rs0, rs1);
        if (NDOT > 0) {
            if (t == 0) {
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    red[0][warp & 3][wm0 + i * 8 + g] = rs0[i];
                    if (NDOT > 1) red[1][warp & 3][wm0 + i * 8 + g] = rs1[i];
                }
            }
            __syncthreads();
            if (tid < TILE) {
                const int64_t gi = i0 + tid;
                if (gi < a.n) {
                    pe.part[0][static_cast<int64_t>(ct) * pe.part_ld + gi] = red[0][0][tid] + red[0][1][tid] + red[0][2][tid] + red[0][3][tid];
                    if (NDOT > 1)
                        pe.part[1][static_cast<int64_t>(ct) * pe.part_ld + gi] = red[1][0][tid] + red[1][1][tid] + red[1][2][tid] + red[1][3][tid];
                }
            }
        }
        __syncthreads();                                                     // everyone is done with As[buf] and red
    }
}

// ------------------------------------------------------------------------------------------------
// Staggered form of the kernel above (g_phi_persist == 2, the default).  ncu on phi_persist_kernel (profiles/r02z_aux.md): DMMA
// pipe 58 % active, FP64 pipe 12 %, top stall math-pipe throttle -- the CTA-wide barrier at the end of every row tile keeps
// all 16 warps in phase, so the DMMA pipe is saturated during the K loops and idle during the exp/store epilogues.  Here no
// barrier spans the CTA inside the tile loop:
//   * the four M-groups (4 warps each, one per SM sub-partition) start one K-loop apart (a chain of named barriers in the
//     first iteration) and stay that way: while one group multiplies, the other three exponentiate and store;
//   * shared-memory row-feature buffers are released per warp (mbarrier empty[b], 16 arrivals) and refilled by TMA from
//     thread 0 after ITS K loop, one tile ahead;
//   * the row-dot partials of the 4 N-warps of an M-group are combined behind a 128-thread named barrier, double-buffered.
// Same arithmetic in the same order: bit-identical to phi_persist_kernel and tgemm_kernel<1>.
// Measured, interleaved A/B on one box each (profiles/r02z_stagger_ab*.log, r02z_digits.log).  With the per-element exp epilogue
// this variant was SLOWER (8.42 vs 7.85 ms at the headline shape): the exp/store warps do not fill the DMMA pipe's idle time,
// because DFMA issue starves while other warps of the sub-partition stream DMMAs (the tcgen05 kernel shows the same),
// so the phases serialise whatever the warp schedule and the extra barriers only cost.  With the grouped epilogue (phi_epilogue)
// the two are level at the headline shape (7.55 vs 7.72, 7.53 vs 7.37 ms on two boxes) and this one is ~5 % ahead at config 3
// (1.94 vs 2.04 ms), where the epilogue is the larger share -- hence the default.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pp_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void pp_named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void pp_named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int NDOT>
__global__ void __launch_bounds__(512, 1)
phi_stagger_kernel(const __grid_constant__ CUtensorMap mapF, const PhiPersist a) {
    constexpr int WARPS_M = 4, WM = TILE / WARPS_M, MT = WM / 8, NT = 4;
    extern __shared__ __align__(128) unsigned char pp_smem[];
    double* Ws = reinterpret_cast<double*>(pp_smem);                       // [K4][LDT]
    double* As = Ws + a.K4 * LDT;                                           // [2][TILE][LDK]
    __shared__ double red[2][2][4][TILE];                                   // [tile parity][dot][N-warp][row]
    __shared__ double exp_sm[32];
    __shared__ __align__(8) unsigned long long bars[4];                     // full[2] (TMA landed), empty[2] (16 warps done reading)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int mg = warp >> 2, nw = warp & 3;
    const int wm0 = mg * WM, wn0 = nw * 32;
    const int ct = blockIdx.x % a.ntn;
    const int64_t step = gridDim.x / a.ntn;
    const int64_t j0 = static_cast<int64_t>(ct) * TILE;
    const int LDK = a.LDK;
    const uint32_t tile_bytes = static_cast<uint32_t>(TILE * LDK * sizeof(double));
    const PEpi& pe = a.pe;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[2]);

    for (int c = tid; c < a.K4 * (TILE / 2); c += 512) {                    // this CTA's column tile of W, once
        const int r = c >> 6, cc = c & 63;
        cp_async16(Ws + r * LDT + cc * 2, a.W + static_cast<int64_t>(r) * a.MP + j0 + cc * 2, 16);
    }
    cp_async_commit();
    exp_tab_stage(exp_sm);
    if (tid == 0) {
        pp_mbar_init(full0, 1);
        pp_mbar_init(full0 + 8, 1);
        pp_mbar_init(empty0, 16);
        pp_mbar_init(empty0 + 8, 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double2 v0[NT], v1[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int64_t gj = j0 + wn0 + j * 8 + 2 * t;
        v0[j] = NDOT > 0 ? *reinterpret_cast<const double2*>(pe.vec[0] + gj) : make_double2(0.0, 0.0);
        v1[j] = NDOT > 1 ? *reinterpret_cast<const double2*>(pe.vec[1] + gj) : make_double2(0.0, 0.0);
    }
    cp_async_wait<0>();
    __syncthreads();
    int64_t rt = blockIdx.x / a.ntn;
    if (tid == 0 && rt < a.tiles_m) {
        pp_mbar_expect(full0, tile_bytes);
        pp_tma_2d(&mapF, smem_u32(As), full0, 0, static_cast<int>(rt * TILE));
    }
    for (uint32_t it = 0; rt < a.tiles_m; rt += step, ++it) {
        const uint32_t buf = it & 1u;
        if (it == 0 && mg > 0) pp_named_sync(4 + mg, 256);                   // start one K loop after the M-group before
        pp_mbar_wait(full0 + 8 * buf, (it >> 1) & 1u);
        const double* as = As + buf * TILE * LDK;
        double acc[MT][NT][2];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kk = 0; kk < a.K4 / 4; ++kk) {
            const int kc = kk * 4 + t;
            double bf[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = Ws[kc * LDT + wn0 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const double af = as[(wm0 + i * 8 + g) * LDK + kc];
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
            }
        }
        __syncwarp();
        if (lane == 0) pp_mbar_arrive(empty0 + 8 * buf);                     // this warp no longer reads As[buf]
        if (it == 0 && mg < WARPS_M - 1) pp_named_arrive(4 + mg + 1, 256);
        if (tid == 0 && rt + step < a.tiles_m) {                             // next row tile into the other buffer: every warp has
            if (it >= 1) pp_mbar_wait(empty0 + 8 * (buf ^ 1u), ((it - 1) >> 1) & 1u);   // left the K loop of the tile before this one
            pp_mbar_expect(full0 + 8 * (buf ^ 1u), tile_bytes);
            pp_tma_2d(&mapF, smem_u32(As + (buf ^ 1u) * TILE * LDK), full0 + 8 * (buf ^ 1u), 0, static_cast<int>((rt + step) * TILE));
        }
        const int64_t i0 = rt * TILE;
        double rs0[MT], rs1[MT];
        if (j0 + TILE <= pe.m && i0 + TILE <= a.n)
            phi_epilogue<NDOT, false>(acc, pe, a.MP, a.n, i0, j0, wm0, wn0, g, t, v0, v1, exp_sm, rs0, rs1);
        else
            phi_epilogue<NDOT, true>(acc, pe, a.MP, a.n, i0, j0, wm0, wn0, g, t, v0, v1, exp_sm, rs0, rs1);
        if (NDOT > 0) {
            double (*rd)[4][TILE] = red[buf];
            if (t == 0) {
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    rd[0][nw][wm0 + i * 8 + g] = rs0[i];
                    if (NDOT > 1) rd[1][nw][wm0 + i * 8 + g] = rs1[i];
                }
            }
            pp_named_sync(1 + mg, 128);                                      // the 4 N-warps of this M-group; red[buf] is written
            if (nw == 0) {                                                   // again two tiles on, behind the next such barrier
                const int row = wm0 + lane;
                const int64_t gi = i0 + row;
                if (gi < a.n) {
                    pe.part[0][static_cast<int64_t>(ct) * pe.part_ld + gi] = rd[0][0][row] + rd[0][1][row] + rd[0][2][row] + rd[0][3][row];
                    if (NDOT > 1)
                        pe.part[1][static_cast<int64_t>(ct) * pe.part_ld + gi] = rd[1][0][row] + rd[1][1][row] + rd[1][2][row] + rd[1][3][row];
                }
            }
        }
    }
}

static int launch_phi_persist(const double* F, int64_t ldf, int kvalid, const double* W, int MP, int64_t n, const PEpi& pe, cudaStream_t st,
                              bool* done) {
    *done = false;
    const int K4 = static_cast<int>(round_up(kvalid, 4));
    const int LDK = (K4 % 8 == 4) ? K4 : K4 + 4;
    const size_t smem = sizeof(double) * (static_cast<size_t>(K4) * LDT + 2 * static_cast<size_t>(TILE) * LDK);
    const bool stagger = g_phi_persist >= 2;                                 // 8 KB more static shared memory (double-buffered partials)
    if (!ozmma_available() || LDK > ldf || smem > (stagger ? 210 : 218) * 1024 || n >= (1LL << 31) - TILE ||
        (reinterpret_cast<uintptr_t>(F) & 15) != 0)
        return GPZ_OK;                                                       // the tile-per-CTA kernel takes it
    static PerDeviceOnce once;
    static int sms_dev[128];
    int dev = 0;
    cudaGetDevice(&dev);
    if (once.need()) {
        // dynamic + 8.4 KB (16.7 KB staggered) static = the 227 KB a CTA may use
        GPZ_CUDA(cudaFuncSetAttribute(phi_persist_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024));
        GPZ_CUDA(cudaFuncSetAttribute(phi_persist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024));
        GPZ_CUDA(cudaFuncSetAttribute(phi_persist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024));
        GPZ_CUDA(cudaFuncSetAttribute(phi_stagger_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        GPZ_CUDA(cudaFuncSetAttribute(phi_stagger_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        GPZ_CUDA(cudaFuncSetAttribute(phi_stagger_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        sms_dev[dev & 127] = sms;
    }
    alignas(64) CUtensorMap map;
    int rc;
    if ((rc = tensor_map_2d_f64(&map, F, ldf, n, ldf, LDK, TILE))) return rc;
    PhiPersist a;
    a.W = W;
    a.MP = MP;
    a.K4 = K4;
    a.LDK = LDK;
    a.ntn = MP / TILE;
    a.n = n;
    a.tiles_m = ceil_div(n, TILE);
    a.pe = pe;
    int64_t per = sms_dev[dev & 127] / a.ntn;
    if (per < 1) per = 1;
    if (per > a.tiles_m) per = a.tiles_m;
    const unsigned grid = static_cast<unsigned>(per * a.ntn);
    const int nd = pe.ndot < 0 ? 0 : (pe.ndot > 2 ? 2 : pe.ndot);
    if (stagger) {
        if (nd == 0) phi_stagger_kernel<0><<<grid, 512, smem, st>>>(map, a);
        else if (nd == 1) phi_stagger_kernel<1><<<grid, 512, smem, st>>>(map, a);
        else phi_stagger_kernel<2><<<grid, 512, smem, st>>>(map, a);
    } else {
        if (nd == 0) phi_persist_kernel<0><<<grid, 512, smem, st>>>(map, a);
        else if (nd == 1) phi_persist_kernel<1><<<grid, 512, smem, st>>>(map, a);
        else phi_persist_kernel<2><<<grid, 512, smem, st>>>(map, a);
    }
    GPZ_KERNEL_CHECK();
    *done = true;
    return GPZ_OK;
}

template <int EPI, int WARPS_M>
static int launch_tgemm(const double* A, int64_t lda, const double* B, int MP, int nk, int klast, int64_t n, const TEpi& te,
                        const PEpi& pe, cudaStream_t st) {
    const size_t smem = sizeof(double) * (STAGES * TILE * LDK + STAGES * KSTEP * LDT);
    static PerDeviceOnce once;
    if (once.need()) {
        GPZ_CUDA(cudaFuncSetAttribute(tgemm_kernel<EPI, WARPS_M>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    }
    const int64_t nblk = ceil_div(n, TILE) * (MP / TILE);
    if (nblk > 2147483647LL) {
        set_error("tgemm: grid too large");
        return GPZ_ERR_USAGE;
    }
    tgemm_kernel<EPI, WARPS_M><<<static_cast<unsigned>(nblk), WARPS_M * 128, smem, st>>>(A, lda, B, MP, nk, klast, n, te, pe);
    GPZ_KERNEL_CHECK();
    return GPZ_OK;
}

int tgemm(const double* Phi, int64_t ld, const double* Sinv, int MP, int m, int64_t n, const double* rw, double* H,
          int accumulate, double* nupart, int64_t nu_ld, double* pred_aug, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return GPZ_OK;
    const int nk = static_cast<int>(round_up(m, KSTEP) / KSTEP);
    TEpi te{Phi, rw, H, accumulate, nupart, nu_ld, pred_aug != nullptr ? m : -1, pred_aug};
    PEpi pe{};
    int rc = g_gemm_warps == 16 ? launch_tgemm<0, 4>(Phi, ld, Sinv, MP, nk, KSTEP / 4, n, te, pe, st) : launch_tgemm<0, 2>(Phi, ld, Sinv, MP, nk, KSTEP / 4, n, te, pe, st);
    if (!rc) ++*launches;
    return rc;
}

// PHI = exp(F W): F [n][ldf] row features, W [kq][MP] coefficients (kq % 16 == 0); only the first kvalid of the kq operand
// columns / rows are non-zero, so the K loop stops at the 4-deep step that holds column kvalid - 1
int phi_gemm(const double* F, int64_t ldf, int kq, int kvalid, const double* W, int MP, int m, int64_t n, double* Phi, int ndot,
             const double* vec0, const double* vec1, double* part0, double* part1, int64_t part_ld, const double* ycol,
             cudaStream_t st, int64_t* launches) {
    if (n <= 0) return GPZ_OK;
    TEpi te{};
    PEpi pe{m, Phi, ndot, {vec0, vec1}, {part0, part1}, part_ld, ycol};
    if (kvalid < 1 || kvalid > kq) kvalid = kq;
    if (g_gemm_warps == 0 && g_phi_persist) {
        bool done = false;
        const int rcp = launch_phi_persist(F, ldf, kvalid, W, MP, n, pe, st, &done);
        if (rcp) return rcp;
        if (done) {
            ++*launches;
            return GPZ_OK;
        }
    }
    const int nk = static_cast<int>(ceil_div(kvalid, KSTEP));
    const int klast = static_cast<int>(ceil_div(kvalid - (nk - 1) * KSTEP, 4));
    int rc = g_gemm_warps != 8 ? launch_tgemm<1, 4>(F, ldf, W, MP, nk, klast, n, te, pe, st)
                                : launch_tgemm<1, 2>(F, ldf, W, MP, nk, klast, n, te, pe, st);
    if (!rc) ++*launches;
    return rc;
}

// C[n][N] = A[n][K] B[K][N]  (N a multiple of 128, K a multiple of 16; A rows of stride lda): the PHI kernel without the exp
int gemm_rows(const double* A, int64_t lda, int K, const double* B, int N, int64_t n, double* C, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return GPZ_OK;
    if (N % TILE != 0 || K % KSTEP != 0) {
        set_error("gemm_rows: N=%d must be a multiple of %d and K=%d of %d", N, TILE, K, KSTEP);
        return GPZ_ERR_USAGE;
    }
    TEpi te{};
    PEpi pe{N, C, 0, {nullptr, nullptr}, {nullptr, nullptr}, 0, nullptr, 1};
    int rc = g_gemm_warps != 8 ? launch_tgemm<1, 4>(A, lda, B, N, K / KSTEP, KSTEP / 4, n, te, pe, st) : launch_tgemm<1, 2>(A, lda, B, N, K / KSTEP, KSTEP / 4, n, te, pe, st);
    if (!rc) ++*launches;
    return rc;
}

// ------------------------------------------------------------------------------------------------
// fused back-projection GEMM:  R[j][c] = sum_i dPHI_ij F_ic  with dPHI formed on the fly,
//   dPHI_ij = -H_ij + PHI_ij (cw_i w_j + dbeta_i v_j)                         GPz.m:72,90,106,113
// and the column sums  q_j = sum_i PHI_ij (-cw_i),  dv_j = sum_i PHI_ij dbeta_i   GPz.m:89,104
// The A operand goes global -> registers -> (transform) -> shared; F tiles use cp.async.
// CTA tile: 128 bases x TN = 8*NT feature columns (all of QP), split over rows.
// ------------------------------------------------------------------------------------------------
constexpr int DSTAGES = 3;

// WARPS = 8: 16 rows per stage, each warp 16 bases x TN; WARPS = 16: 32 rows per stage, each warp 8 bases x TN -- half as many CTA
// barriers per row and four warps per sub-partition instead of two ("moment_warps" option).
template <int NT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
atb_dphi_kernel(const double* __restrict__ Phi, const double* __restrict__ H, int64_t ld, const double* __restrict__ F,
                int64_t ldf, const double* __restrict__ cw, const double* __restrict__ dbeta,
                const double* __restrict__ w, const double* __restrict__ v, int64_t row0, int64_t row1,
                int64_t rows_per_split, double* __restrict__ partial, double* __restrict__ colp, int MP, int accumulate,
                int col_accumulate) {
    constexpr int TN = 8 * NT;
    constexpr int LDB = TN + 4;
    constexpr int THREADS = WARPS * 32, KS = (THREADS / 64) * 4, MT = 16 / WARPS;      // rows per stage; 8-base blocks per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);           // [DSTAGES][KS][LDT]
    double* Bs = As + DSTAGES * KS * LDT;                       // [DSTAGES][KS][LDB]
    __shared__ double csum[THREADS / 64][4][64];                // [row group][q0,q1,v0,v1][column pair]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = warp * (8 * MT);
    const int jt = blockIdx.x;
    const int64_t a_col0 = static_cast<int64_t>(jt) * TILE;
    const int64_t rbeg = row0 + static_cast<int64_t>(blockIdx.y) * rows_per_split;
    int64_t rend = rbeg + rows_per_split;
    if (rend > row1) rend = row1;
    const int64_t nrows = rend > rbeg ? rend - rbeg : 0;
    const int nk = static_cast<int>((nrows + KS - 1) / KS);

    // this thread's fixed column pair and row group in the A tile
    const int cc = tid & 63, rg = tid >> 6;
    const double2 wj = *reinterpret_cast<const double2*>(w + a_col0 + cc * 2);
    const double2 vj = *reinterpret_cast<const double2*>(v + a_col0 + cc * 2);
    double sq0 = 0.0, sq1 = 0.0, sv0 = 0.0, sv1 = 0.0;

    double2 rp[4], rh[4];
    double rc_[4], rd_[4];
    auto load_a_regs = [&](int kt) {
        const int64_t rb = rbeg + static_cast<int64_t>(kt) * KS;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t gr = rb + rg + (THREADS / 64) * q;
            if (gr < rend) {
                rp[q] = *reinterpret_cast<const double2*>(Phi + gr * ld + a_col0 + cc * 2);
                rh[q] = *reinterpret_cast<const double2*>(H + gr * ld + a_col0 + cc * 2);
                rc_[q] = __ldg(cw + gr);
                rd_[q] = __ldg(dbeta + gr);
            } else {
                rp[q] = rh[q] = make_double2(0.0, 0.0);
                rc_[q] = rd_[q] = 0.0;
            }
        }
    };
    auto store_a_smem = [&](int stage) {
        double* as = As + stage * KS * LDT;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double c0 = fma(rc_[q], wj.x, rd_[q] * vj.x), c1 = fma(rc_[q], wj.y, rd_[q] * vj.y);
            const double2 dphi = make_double2(fma(rp[q].x, c0, -rh[q].x), fma(rp[q].y, c1, -rh[q].y));
            *reinterpret_cast<double2*>(as + (rg + (THREADS / 64) * q) * LDT + cc * 2) = dphi;
            sq0 = fma(rp[q].x, -rc_[q], sq0);
            sq1 = fma(rp[q].y, -rc_[q], sq1);
            sv0 = fma(rp[q].x, rd_[q], sv0);
            sv1 = fma(rp[q].y, rd_[q], sv1);
        }
    };
    auto load_b = [&](int kt, int stage) {
        const int64_t rb = rbeg + static_cast<int64_t>(kt) * KS;
        double* bs = Bs + stage * KS * LDB;
        constexpr int BCH = KS * TN / 2;
#pragma unroll
        for (int q = 0; q < (BCH + THREADS - 1) / THREADS; ++q) {
            const int c = tid + q * THREADS;
            if (c < BCH) {
                const int r = c / (TN / 2), c2 = c % (TN / 2);
                const int64_t gr = rb + r;
                const bool ok = gr < rend;
                cp_async16(bs + r * LDB + c2 * 2, F + (ok ? gr : row0) * ldf + c2 * 2, ok ? 16 : 0);
            }
        }
    };

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < DSTAGES - 1; ++s) {
        if (s < nk) {
            load_a_regs(s);
            store_a_smem(s);
            load_b(s, s);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<DSTAGES - 2>();
        __syncthreads();
        const int nx = kt + DSTAGES - 1;
        const bool more = nx < nk;
        if (more) {
            load_a_regs(nx);
            load_b(nx, nx % DSTAGES);
        }
        cp_async_commit();
        const int stage = kt % DSTAGES;
        const double* as = As + stage * KS * LDT;
        const double* bs = Bs + stage * KS * LDB;
#pragma unroll
        for (int kk = 0; kk < KS / 4; ++kk) {
            const int kr = kk * 4 + t;
            double bf[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = bs[kr * LDB + j * 8 + g];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const double af = as[kr * LDT + wm0 + i * 8 + g];
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
            }
        }
        if (more) store_a_smem(nx % DSTAGES);
    }
    cp_async_wait<0>();

    double* out = partial + (static_cast<int64_t>(blockIdx.y) * gridDim.x + jt) * (TILE * TN);
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int r = wm0 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int c = j * 8 + 2 * t;
            double2* p = reinterpret_cast<double2*>(out + r * TN + c);
            double2 val = make_double2(acc[i][j][0], acc[i][j][1]);
            if (accumulate) {
                const double2 o = *p;
                val.x += o.x;
                val.y += o.y;
            }
            *p = val;
        }
    }
    // column sums: fixed-order sum over the 4 row groups
    csum[rg][0][cc] = sq0;
    csum[rg][1][cc] = sq1;
    csum[rg][2][cc] = sv0;
    csum[rg][3][cc] = sv1;
    __syncthreads();
    if (tid < 128) {
        const int c2 = tid >> 1, e = tid & 1;
        double q = csum[0][e][c2] + csum[1][e][c2] + csum[2][e][c2] + csum[3][e][c2];
        double dv = csum[0][2 + e][c2] + csum[1][2 + e][c2] + csum[2][2 + e][c2] + csum[3][2 + e][c2];
        if (THREADS / 64 == 8) {
            q += csum[4][e][c2] + csum[5][e][c2] + csum[6][e][c2] + csum[7][e][c2];
            dv += csum[4][2 + e][c2] + csum[5][2 + e][c2] + csum[6][2 + e][c2] + csum[7][2 + e][c2];
        }
        double* cp = colp + static_cast<int64_t>(blockIdx.y) * 2 * MP + a_col0 + tid;
        cp[0] = (col_accumulate ? cp[0] : 0.0) + q;
        cp[MP] = (col_accumulate ? cp[MP] : 0.0) + dv;
    }
}

template <int NT, int WARPS>
static int launch_atb_dphi_w(const double* Phi, const double* H, int64_t ld, const double* F, int64_t ldf, const double* cw,
                           const double* dbeta, const double* w, const double* v, int64_t row0, int64_t row1, int nsplit,
                           double* partial, double* colp, int MP, int accumulate, int col_accumulate, cudaStream_t st) {
    constexpr int TN = 8 * NT, KS = (WARPS / 2) * 4;
    const size_t smem = sizeof(double) * (DSTAGES * KS * LDT + DSTAGES * KS * (TN + 4));
    static PerDeviceOnce once;
    if (once.need()) {
        GPZ_CUDA(cudaFuncSetAttribute(atb_dphi_kernel<NT, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    }
    int64_t rps = ceil_div(row1 - row0, nsplit);
    rps = round_up(rps > 0 ? rps : 1, KS);
    dim3 grid(MP / TILE, nsplit);
    atb_dphi_kernel<NT, WARPS><<<grid, WARPS * 32, smem, st>>>(Phi, H, ld, F, ldf, cw, dbeta, w, v, row0, row1, rps, partial, colp, MP,
                                                             accumulate, col_accumulate);
    GPZ_KERNEL_CHECK();
    return GPZ_OK;
}

template <int NT>
static int launch_atb_dphi(const double* Phi, const double* H, int64_t ld, const double* F, int64_t ldf, const double* cw,
                           const double* dbeta, const double* w, const double* v, int64_t row0, int64_t row1, int nsplit,
                           double* partial, double* colp, int MP, int accumulate, int col_accumulate, cudaStream_t st) {
    // 16 warps pay for narrow feature tiles on many rows (config 3, VD: NT = 3, 1.87 -> 1.54 ms) and cost 35 % at NT = 9, where each
    // warp's B fragment loads are no longer shared by two row blocks (profiles/r02zf_moment_warps.log)
    const bool wide = g_moment_warps == 16 || (g_moment_warps == 0 && NT <= 4 && row1 - row0 >= 200000);
    if (wide)
        return launch_atb_dphi_w<NT, 16>(Phi, H, ld, F, ldf, cw, dbeta, w, v, row0, row1, nsplit, partial, colp, MP, accumulate, col_accumulate, st);
    return launch_atb_dphi_w<NT, 8>(Phi, H, ld, F, ldf, cw, dbeta, w, v, row0, row1, nsplit, partial, colp, MP, accumulate, col_accumulate, st);
}

// F rows have stride QP (multiple of 32, <= 128); only the first q feature columns are multiplied, rounded up to the
// narrowest instantiated tile (24, 32, 48, 64, 72, 96 or 128 columns).
// partial: nsplit * (MP/128) * 128 * QP doubles; colp: nsplit * 2 * MP doubles.
int atb_dphi(const double* Phi, const double* H, int64_t ld, int MP, const double* F, int QP, int q, const double* cw,
             const double* dbeta, const double* w, const double* v, int64_t row0, int64_t row1, int nsplit, double* partial,
             double* colp, int accumulate, int col_accumulate, int reduce, double* R, cudaStream_t st, int64_t* launches) {
    int rc, TN;
#define GPZ_DPHI(NT_)                                                                                                      \
    TN = 8 * NT_;                                                                                                         \
    rc = launch_atb_dphi<NT_>(Phi, H, ld, F, QP, cw, dbeta, w, v, row0, row1, nsplit, partial, colp, MP, accumulate, col_accumulate, st)
    if (QP > 128 || q > QP) {
        set_error("atb_dphi: unsupported feature width %d (q=%d)", QP, q);
        return GPZ_ERR_USAGE;
    }
    if (q <= 24) { GPZ_DPHI(3); }
    else if (q <= 32) { GPZ_DPHI(4); }
    else if (q <= 48 && QP >= 48) { GPZ_DPHI(6); }
    else if (q <= 64) { GPZ_DPHI(8); }
    else if (q <= 72 && QP >= 72) { GPZ_DPHI(9); }
    else if (q <= 96) { GPZ_DPHI(12); }
    else { GPZ_DPHI(16); }
#undef GPZ_DPHI
    if (rc) return rc;
    ++*launches;
    if (reduce) {
        const int tm = MP / TILE;
        atb_reduce_kernel<false><<<atb_reduce_grid(tm, TN), 256, 0, st>>>(partial, nsplit, tm, 1, TN, R, QP);
        GPZ_KERNEL_CHECK();
        ++*launches;
    }
    return GPZ_OK;
}

// ------------------------------------------------------------------------------------------------
// small strided GEMM: C(M x N, row-major ldc) = beta*C + alpha * sum_k A(i,k) B(k,j)
//   A(i,k) = A[i*sAi + k*sAk],  B(k,j) = B[k*sBk + j*sBj];  64x64 tile, 4 warps, DMMA
//   lower_only: skip tiles strictly above the diagonal
// ------------------------------------------------------------------------------------------------
//   batched over blockIdx.z: operand pointers advance by bsA/bsB/bsC per batch; with mlim > 0 the batch z works on
//   M_z = min(M, mlim - z*mstep) rows (and K_z = M_z when k_follows_m), which lets one launch cover the ragged last pair of
//   the recursive triangular inverse
struct SgemmBatch {
    int64_t bsA, bsB, bsC;
    int mlim, mstep, k_follows_m;
};

// TS = 64: 64 x 64 tile per CTA (4 warps x 32 x 32); TS = 32: 32 x 32 tile (4 warps x 16 x 16) for skinny products whose 64-tiles
// would leave most SMs idle (a CTA is DMMA-bound at TS*TS*16 MACs per K step whatever else happens).
// KS = K elements per step.  The operands go global -> registers -> shared one step ahead, so every step exposes whatever of the
// global-load latency its own DMMAs do not cover; with 16 the ncu launch list (profiles/r02z_solve_launches.csv) showed 0.8 us per
// step for the 32-tile and 2 us per step for K = 64 products -- latency, not arithmetic.  KS = 32 (64-tile) / 64 (32-tile) moves
// 2-4x the bytes per latency.
template <int TS, int KS>
__global__ void __launch_bounds__(128)
sgemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, int64_t sAi, int64_t sAk,
             const double* __restrict__ B, int64_t sBk, int64_t sBj, double beta, double* __restrict__ C, int64_t ldc,
             int lower_only, SgemmBatch bt) {
    constexpr int LA = KS + 4, LB = TS + 4, WT = TS / 2, MI = WT / 8, NQ = TS * KS / 128;
    __shared__ double As[TS][LA];
    __shared__ double Bs[KS][LB];
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (lower_only && tj > ti) return;
    if (lower_only == 2 && ti == 0 && tj == 0) return;      // look-ahead Cholesky: the next diagonal block is updated by its own kernel
    if (bt.mlim > 0) {
        const int mz = bt.mlim - static_cast<int>(blockIdx.z) * bt.mstep;
        if (mz < M) M = mz;
        if (bt.k_follows_m) K = M;
        if (M <= 0 || ti * TS >= M) return;
    }
    A += blockIdx.z * bt.bsA;
    B += blockIdx.z * bt.bsB;
    C += blockIdx.z * bt.bsC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 1) * WT, wn0 = (warp & 1) * WT;
    const int i0 = ti * TS, j0 = tj * TS;

    double acc[MI][MI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < MI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // software pipeline: the next K tile is fetched into registers while the current one is multiplied
    const bool a_kfast = (sAk == 1), b_jfast = (sBj == 1);
    double ra[NQ], rb[NQ];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int e = tid + q * 128;
            int r, c;
            if (a_kfast) { r = e / KS; c = e % KS; } else { c = e / TS; r = e % TS; }
            const int gi = i0 + r, gk = k0 + c;
            ra[q] = (gi < M && gk < K) ? A[gi * sAi + gk * sAk] : 0.0;
            int rr, cc;
            if (b_jfast) { rr = e / TS; cc = e % TS; } else { cc = e / KS; rr = e % KS; }
            const int gk2 = k0 + rr, gj = j0 + cc;
            rb[q] = (gk2 < K && gj < N) ? B[gk2 * sBk + gj * sBj] : 0.0;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int e = tid + q * 128;
            int r, c;
            if (a_kfast) { r = e / KS; c = e % KS; } else { c = e / TS; r = e % TS; }
            As[r][c] = ra[q];
            int rr, cc;
            if (b_jfast) { rr = e / TS; cc = e % TS; } else { cc = e / KS; rr = e % KS; }
            Bs[rr][cc] = rb[q];
        }
    };
    if (K > 0) fetch(0);
    for (int k0 = 0; k0 < K; k0 += KS) {
        stash();
        __syncthreads();
        if (k0 + KS < K) fetch(k0 + KS);
#pragma unroll
        for (int kk = 0; kk < KS / 4; ++kk) {
            const int kc = kk * 4 + t;
            double bf[MI];
#pragma unroll
            for (int j = 0; j < MI; ++j) bf[j] = Bs[kc][wn0 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const double af = As[wm0 + i * 8 + g][kc];
#pragma unroll
                for (int j = 0; j < MI; ++j) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int gi = i0 + wm0 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < MI; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = j0 + wn0 + j * 8 + 2 * t + e;
                if (gi < M && gj < N) {
                    double* p = C + gi * ldc + gj;
                    const double v = alpha * acc[i][j][e];
                    *p = (beta == 0.0) ? v : (beta * (*p) + v);
                }
            }
        }
    }
}

int sgemm(int M, int N, int K, double alpha, const double* A, int64_t sAi, int64_t sAk, const double* B, int64_t sBk,
          int64_t sBj, double beta, double* C, int64_t ldc, int lower_only, cudaStream_t st, int64_t* launches) {
    if (M <= 0 || N <= 0) return GPZ_OK;
    // few 64-tiles and a long K: 32-tiles give 4x the CTAs (only without the tile-skipping options, whose tile index means 64)
    if (!lower_only && K >= 256 && ceil_div(N, 64) * ceil_div(M, 64) <= 64) {
        dim3 grid32(static_cast<unsigned>(ceil_div(N, 32)), static_cast<unsigned>(ceil_div(M, 32)));
        sgemm_kernel<32, 64><<<grid32, 128, 0, st>>>(M, N, K, alpha, A, sAi, sAk, B, sBk, sBj, beta, C, ldc, 0, SgemmBatch{0, 0, 0, 0, 0, 0});
        GPZ_KERNEL_CHECK();
        ++*launches;
        return GPZ_OK;
    }
    dim3 grid(static_cast<unsigned>(ceil_div(N, 64)), static_cast<unsigned>(ceil_div(M, 64)));
    sgemm_kernel<64, 32><<<grid, 128, 0, st>>>(M, N, K, alpha, A, sAi, sAk, B, sBk, sBj, beta, C, ldc, lower_only, SgemmBatch{0, 0, 0, 0, 0, 0});
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

// batched variant (see SgemmBatch): `batches` problems, pointers advancing by bsA/bsB/bsC
int sgemm_batched(int M, int N, int K, double alpha, const double* A, int64_t sAi, int64_t sAk, int64_t bsA, const double* B,
                  int64_t sBk, int64_t sBj, int64_t bsB, double beta, double* C, int64_t ldc, int64_t bsC, int batches, int mlim,
                  int mstep, int k_follows_m, cudaStream_t st, int64_t* launches) {
    if (M <= 0 || N <= 0 || batches <= 0) return GPZ_OK;
    dim3 grid(static_cast<unsigned>(ceil_div(N, 64)), static_cast<unsigned>(ceil_div(M, 64)), static_cast<unsigned>(batches));
    sgemm_kernel<64, 32><<<grid, 128, 0, st>>>(M, N, K, alpha, A, sAi, sAk, B, sBk, sBj, beta, C, ldc, 0,
                                           SgemmBatch{bsA, bsB, bsC, mlim, mstep, k_follows_m});
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

}  // namespace gpz
