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
#include "internal.cuh"

namespace gpz {

constexpr int STAGES = 4;

// ------------------------------------------------------------------------------------------------
// C = A' diag(w) B over a row range, 128 x TN output tile per CTA, split over rows
// ------------------------------------------------------------------------------------------------
template <int WARPS_M, int WARPS_N, bool SYRK>
__global__ void __launch_bounds__(256, 1)
atb_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
           const double* __restrict__ wgt, int64_t row0, int64_t row1, int64_t rows_per_split,
           int ntiles_n, double* __restrict__ partial, int accumulate) {
    static_assert(WARPS_M * WARPS_N == 8, "8 warps");
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
        for (int q = 0; q < (KSTEP * TILE / 2) / 256; ++q) {
            const int c = tid + q * 256;
            const int r = c >> 6, cc = c & 63;
            const int64_t gr = rb + r;
            const bool ok = gr < rend;
            const double* src = A + (ok ? gr : row0) * lda + a_col0 + cc * 2;
            cp_async16(as + r * LDT + cc * 2, src, ok ? 16 : 0);
        }
        if (!alias) {
            constexpr int BCH = KSTEP * TN / 2;      // 16-byte chunks in the B tile
#pragma unroll
            for (int q = 0; q < (BCH + 255) / 256; ++q) {
                const int c = tid + q * 256;
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
template <int TN, bool SYRK>
__global__ void atb_reduce_kernel(const double* __restrict__ partial, int nsplit, int ntiles, int ntiles_n,
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
    for (int e = threadIdx.x; e < TILE * TN; e += blockDim.x) {
        double s = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) s += partial[(static_cast<int64_t>(sp) * ntiles + tile) * (TILE * TN) + e];
        const int r = e / TN, c = e % TN;
        const int64_t gr = static_cast<int64_t>(jt) * TILE + r, gc = static_cast<int64_t>(ct) * TN + c;
        C[gr * ldc + gc] = s;
        if (SYRK && jt != ct) C[gc * ldc + gr] = s;
    }
}

template <int WARPS_M, int WARPS_N, bool SYRK>
static int launch_atb(const double* A, int64_t lda, const double* B, int64_t ldb, const double* wgt, int64_t row0,
                      int64_t row1, int ntiles, int ntiles_n, int nsplit, double* partial, int accumulate,
                      cudaStream_t st) {
    constexpr int TN = WARPS_N * 32;
    const size_t smem = sizeof(double) * (STAGES * KSTEP * LDT + STAGES * KSTEP * (TN + 4) + STAGES * KSTEP);
    static bool configured = false;
    if (!configured) {
        GPZ_CUDA(cudaFuncSetAttribute(atb_kernel<WARPS_M, WARPS_N, SYRK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
        configured = true;
    }
    int64_t rps = ceil_div(row1 - row0, nsplit);
    rps = round_up(rps > 0 ? rps : 1, KSTEP);
    dim3 grid(ntiles, nsplit);
    atb_kernel<WARPS_M, WARPS_N, SYRK><<<grid, 256, smem, st>>>(A, lda, B, ldb, wgt, row0, row1, rps, ntiles_n, partial,
                                                               accumulate);
    GPZ_KERNEL_CHECK();
    return GPZ_OK;
}

// Gram: S(MP x MP, both triangles) = PHI' diag(w) PHI over rows [row0,row1).
// partial must hold nsplit * ntri * 128*128 doubles; the finish step sums the split partials (fixed order).
int gram_syrk_main(const double* Phi, int64_t ld, int MP, const double* wgt, int64_t row0, int64_t row1, int nsplit,
                   double* partial, int accumulate, cudaStream_t st, int64_t* launches) {
    const int T = MP / TILE;
    const int ntri = T * (T + 1) / 2;
    int rc = launch_atb<2, 4, true>(Phi, ld, Phi, ld, wgt, row0, row1, ntri, T, nsplit, partial, accumulate, st);
    if (rc) return rc;
    ++*launches;
    return GPZ_OK;
}

int gram_syrk_finish(const double* partial, int nsplit, int MP, int reduce, double* S, cudaStream_t st, int64_t* launches) {
    if (!reduce) return GPZ_OK;
    const int T = MP / TILE;
    const int ntri = T * (T + 1) / 2;
    atb_reduce_kernel<TILE, true><<<ntri, 256, 0, st>>>(partial, nsplit, ntri, T, S, MP);
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
        atb_reduce_kernel<32, false><<<tm * tn, 256, 0, st>>>(partial, nsplit, tm * tn, tn, R, QP);
        GPZ_KERNEL_CHECK();
        ++*launches;
    }
    return GPZ_OK;
}

// ------------------------------------------------------------------------------------------------
// T = PHI * iSigma with fused epilogue
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
tgemm_kernel(const double* __restrict__ Phi, int64_t ld, const double* __restrict__ Sinv, int MP, int nk, int64_t n,
             const double* __restrict__ rw, double* __restrict__ H, int accumulate, double* __restrict__ nupart,
             int64_t nu_ld) {
    constexpr int WM = 64, MT = 8, NT = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);          // [STAGES][TILE][LDK]
    double* Bs = As + STAGES * TILE * LDK;                     // [STAGES][KSTEP][LDT]
    __shared__ double red[4][TILE];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 2) * WM;
    const int wn0 = (warp & 3) * 32;
    const int ntn = MP / TILE;
    const int ct = blockIdx.x % ntn;
    const int64_t rt = blockIdx.x / ntn;
    const int64_t i0 = rt * TILE;
    const int64_t j0 = static_cast<int64_t>(ct) * TILE;

    auto load_stage = [&](int kt, int stage) {
        const int k0 = kt * KSTEP;
        double* as = As + stage * TILE * LDK;
        double* bs = Bs + stage * KSTEP * LDT;
#pragma unroll
        for (int q = 0; q < (TILE * KSTEP / 2) / 256; ++q) {
            const int c = tid + q * 256;
            const int r = c >> 3, cc = c & 7;
            const int64_t gr = i0 + r;
            const bool ok = gr < n;
            cp_async16(as + r * LDK + cc * 2, Phi + (ok ? gr : 0) * ld + k0 + cc * 2, ok ? 16 : 0);
        }
#pragma unroll
        for (int q = 0; q < (KSTEP * TILE / 2) / 256; ++q) {
            const int c = tid + q * 256;
            const int r = c >> 6, cc = c & 63;
            cp_async16(bs + r * LDT + cc * 2, Sinv + static_cast<int64_t>(k0 + r) * MP + j0 + cc * 2, 16);
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
#pragma unroll
        for (int kk = 0; kk < KSTEP / 4; ++kk) {
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

    // epilogue: h = PHI_ij * T_ij ; nu partial over this tile's 128 columns ; H (+)= rw_i * h
    double rs[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int64_t gi = i0 + wm0 + i * 8 + g;
        const bool ok = gi < n;
        double s = 0.0;
        const double wrow = (rw != nullptr && ok) ? rw[gi] : 1.0;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int64_t gj = j0 + wn0 + j * 8 + 2 * t;
            if (ok) {
                const double2 ph = *reinterpret_cast<const double2*>(Phi + gi * ld + gj);
                const double h0 = ph.x * acc[i][j][0], h1 = ph.y * acc[i][j][1];
                s += h0 + h1;
                if (H != nullptr) {
                    double2* hp = reinterpret_cast<double2*>(H + gi * ld + gj);
                    double2 v = make_double2(wrow * h0, wrow * h1);
                    if (accumulate) {
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
        rs[i] = s;
    }
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < MT; ++i) red[warp & 3][wm0 + i * 8 + g] = rs[i];
    }
    __syncthreads();
    if (tid < TILE) {
        const int64_t gi = i0 + tid;
        if (gi < n) nupart[static_cast<int64_t>(ct) * nu_ld + gi] = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
    }
}

int tgemm(const double* Phi, int64_t ld, const double* Sinv, int MP, int m, int64_t n, const double* rw, double* H,
          int accumulate, double* nupart, int64_t nu_ld, cudaStream_t st, int64_t* launches) {
    const size_t smem = sizeof(double) * (STAGES * TILE * LDK + STAGES * KSTEP * LDT);
    static bool configured = false;
    if (!configured) {
        GPZ_CUDA(cudaFuncSetAttribute(tgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured = true;
    }
    if (n <= 0) return GPZ_OK;
    const int nk = static_cast<int>(round_up(m, KSTEP) / KSTEP);
    const int64_t nblk = ceil_div(n, TILE) * (MP / TILE);
    if (nblk > 2147483647LL) {
        set_error("tgemm: grid too large");
        return GPZ_ERR_USAGE;
    }
    tgemm_kernel<<<static_cast<unsigned>(nblk), 256, smem, st>>>(Phi, ld, Sinv, MP, nk, n, rw, H, accumulate, nupart, nu_ld);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

// ------------------------------------------------------------------------------------------------
// small strided GEMM: C(M x N, row-major ldc) = beta*C + alpha * sum_k A(i,k) B(k,j)
//   A(i,k) = A[i*sAi + k*sAk],  B(k,j) = B[k*sBk + j*sBj];  64x64 tile, 4 warps, DMMA
//   lower_only: skip tiles strictly above the diagonal
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
sgemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, int64_t sAi, int64_t sAk,
             const double* __restrict__ B, int64_t sBk, int64_t sBj, double beta, double* __restrict__ C, int64_t ldc,
             int lower_only) {
    constexpr int TS = 64, LA = KSTEP + 4, LB = TS + 4;
    __shared__ double As[TS][LA];
    __shared__ double Bs[KSTEP][LB];
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (lower_only && tj > ti) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 1) * 32, wn0 = (warp & 1) * 32;
    const int i0 = ti * TS, j0 = tj * TS;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int k0 = 0; k0 < K; k0 += KSTEP) {
        // A tile 64 x 16
        for (int e = tid; e < TS * KSTEP; e += 128) {
            int r, c;
            if (sAk == 1) { r = e / KSTEP; c = e % KSTEP; } else { c = e / TS; r = e % TS; }
            const int gi = i0 + r, gk = k0 + c;
            As[r][c] = (gi < M && gk < K) ? A[gi * sAi + gk * sAk] : 0.0;
        }
        for (int e = tid; e < TS * KSTEP; e += 128) {
            int r, c;
            if (sBj == 1) { r = e / TS; c = e % TS; } else { c = e / KSTEP; r = e % KSTEP; }
            const int gk = k0 + r, gj = j0 + c;
            Bs[r][c] = (gk < K && gj < N) ? B[gk * sBk + gj * sBj] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KSTEP / 4; ++kk) {
            const int kc = kk * 4 + t;
            double bf[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[kc][wn0 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double af = As[wm0 + i * 8 + g][kc];
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = i0 + wm0 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
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
    dim3 grid(static_cast<unsigned>(ceil_div(N, 64)), static_cast<unsigned>(ceil_div(M, 64)));
    sgemm_kernel<<<grid, 128, 0, st>>>(M, N, K, alpha, A, sAi, sAk, B, sBk, sBj, beta, C, ldc, lower_only);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

}  // namespace gpz
