// m x m SPD solve of the hot path: SIGMA^{-1} and ln det SIGMA (GPz/inv_logdet.m:1-15, called at
// GPz/GPz.m:67).  The reference pseudo-inverts by SVD; SIGMA = PHI' W PHI + diag(alpha) is SPD by
// construction (alpha > 0), so this is a blocked right-looking Cholesky (64-wide panels: diagonal
// block factored and inverted by one CTA in shared memory, panel solve and trailing update as DMMA
// GEMMs), a blocked triangular inverse and one W'W product.  A non-positive pivot raises a device
// flag; the caller turns that into NaN outputs (SURVEY.md 8b "error convention", H3).
#include "internal.cuh"

namespace gpz {

constexpr int NB = 64;

// warp-synchronous Cholesky of the 32 x 32 block at (o,o) of the smem tile A (lower, in place); lane = row.
// (A register/shuffle variant was measured slower: 2.7 vs 2.2 ms for the whole m=1000 solve.)
// returns false (uniformly) on a non-positive pivot; *ld2 accumulates 2*sum(log L_cc)
__device__ __forceinline__ bool chol32(double (*A)[NB + 1], int o, double* ld2) {
    const int lane = threadIdx.x & 31;
    for (int c = 0; c < 32; ++c) {
        const double piv = A[o + c][o + c];
        if (!(piv > 0.0)) return false;
        const double l = sqrt(piv);
        *ld2 += 2.0 * log(l);
        __syncwarp();
        if (lane == c) A[o + c][o + c] = l;
        else if (lane > c) A[o + lane][o + c] /= l;
        __syncwarp();
        if (lane > c) {
            const double arc = A[o + lane][o + c];
            for (int q = c + 1; q <= lane; ++q) A[o + lane][o + q] -= arc * A[o + q][o + c];
        }
        __syncwarp();
    }
    return true;
}

// W(o..o+32, o..o+32) = inverse of the lower-triangular 32 x 32 block of A at (o,o); lane = column (forward substitution
// by rows, no cross-lane traffic: L[r][q] is a broadcast read)
__device__ __forceinline__ void trinv32(double (*A)[NB + 1], double (*W)[NB + 1], int o) {
    const int c = threadIdx.x & 31;
    double x[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < r; ++q) s -= A[o + r][o + q] * x[q];
        x[r] = (r >= c) ? s / A[o + r][o + r] : 0.0;
        W[o + r][o + c] = x[r];
    }
}

// C(32x32 at (ro,co) of Cm) = alpha * sum_k X[xr+i][xc+k] * Y(k,j) + beta*C, Y(k,j) = Yt ? Ym[yr+j][yc+k] : Ym[yr+k][yc+j];
// all 128 threads, 8 outputs each
__device__ __forceinline__ void mm32(double (*Cm)[NB + 1], int ro, int co, double alpha, double (*X)[NB + 1], int xr, int xc,
                                     double (*Ym)[NB + 1], int yr, int yc, bool Yt, double beta, bool lower_only) {
    for (int e = threadIdx.x; e < 1024; e += 128) {
        const int i = e >> 5, j = e & 31;
        if (lower_only && j > i) continue;
        double s = 0.0;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) s += X[xr + i][xc + k] * (Yt ? Ym[yr + j][yc + k] : Ym[yr + k][yc + j]);
        Cm[ro + i][co + j] = alpha * s + (beta != 0.0 ? beta * Cm[ro + i][co + j] : 0.0);
    }
}

// factor the nb x nb diagonal block at (k0,k0) in place (lower), invert the factor into Linv (row-major 64 x 64, zero
// upper triangle), accumulate logdet.  The 64 x 64 block is handled as 2 x 2 blocks of 32: warp-level Cholesky and
// triangular inverse on the diagonal blocks, 32^3 products by the whole CTA in between.  Rows/cols >= nb are padded
// with the identity.
__global__ void __launch_bounds__(128)
potf2_trti_kernel(double* __restrict__ S, int64_t ld, int k0, int nb, double* __restrict__ Linv,
                  double* __restrict__ logdet, int* __restrict__ flag, int first) {
    extern __shared__ double sm_potf[];
    double (*A)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm_potf);
    double (*W)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm_potf + NB * (NB + 1));
    double (*T)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm_potf + 2 * NB * (NB + 1));
    __shared__ int bad;
    __shared__ double ldsum;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { bad = 0; ldsum = 0.0; }
    for (int e = tid; e < NB * NB; e += 128) {
        const int r = e / NB, c = e % NB;
        A[r][c] = (r < nb && c < nb && c <= r) ? S[static_cast<int64_t>(k0 + r) * ld + k0 + c] : (r == c ? 1.0 : 0.0);
        W[r][c] = 0.0;
    }
    __syncthreads();
    double ld2 = 0.0;
    if (warp == 0) {                              // L11, W11
        const bool ok = chol32(A, 0, &ld2);
        if (!ok && tid == 0) bad = 1;
        if (ok) trinv32(A, W, 0);
    }
    __syncthreads();
    if (!bad) {
        mm32(T, 32, 0, 1.0, A, 32, 0, W, 0, 0, true, 0.0, false);       // T21 = A21 * W11'  (= L21)
        __syncthreads();
        for (int e = tid; e < 1024; e += 128) A[32 + (e >> 5)][e & 31] = T[32 + (e >> 5)][e & 31];
        __syncthreads();
        mm32(A, 32, 32, -1.0, A, 32, 0, A, 32, 0, true, 1.0, true);     // A22 -= L21 L21'  (lower)
        __syncthreads();
        if (warp == 0) {                          // L22, W22
            const bool ok = chol32(A, 32, &ld2);
            if (!ok && tid == 0) bad = 1;
            if (ok) trinv32(A, W, 32);
        }
        __syncthreads();
    }
    if (bad) {
        if (tid == 0) *flag = 1;
        return;
    }
    mm32(T, 32, 0, 1.0, A, 32, 0, W, 0, 0, false, 0.0, false);          // T21 = L21 * W11
    __syncthreads();
    mm32(W, 32, 0, -1.0, W, 32, 32, T, 32, 0, false, 0.0, false);       // W21 = -W22 * T21
    __syncthreads();
    if (tid == 0) ldsum = ld2;                    // both chol32 calls ran on warp 0: lane 0 holds the full sum
    __syncthreads();
    for (int e = tid; e < NB * NB; e += 128) {
        const int r = e / NB, c = e % NB;
        Linv[e] = (r < nb && c < nb && c <= r) ? W[r][c] : 0.0;
        if (r < nb && c < nb && c <= r) S[static_cast<int64_t>(k0 + r) * ld + k0 + c] = A[r][c];
    }
    if (tid == 0) *logdet = (first ? 0.0 : *logdet) + ldsum;
}

__global__ void zero_kernel(double* p, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0;
}

// copy the 64x64 inverse diagonal blocks into W
__global__ void place_diag_kernel(const double* __restrict__ Linv, double* __restrict__ W, int64_t ld, int m) {
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
        const int r = e / NB, c = e % NB;
        const int gr = b * NB + r, gc = b * NB + c;
        if (gr < m && gc < m) W[static_cast<int64_t>(gr) * ld + gc] = Linv[static_cast<int64_t>(b) * NB * NB + e];
    }
}

int solve_ws_alloc(SolveWs& ws, int MP) {
    const int nblk = MP / NB;
    GPZ_CUDA(cudaMalloc(&ws.W, sizeof(double) * MP * MP));
    GPZ_CUDA(cudaMalloc(&ws.Linv, sizeof(double) * nblk * NB * NB));
    GPZ_CUDA(cudaMalloc(&ws.tmp, sizeof(double) * NB * MP));
    GPZ_CUDA(cudaMalloc(&ws.flag, sizeof(int)));
    return GPZ_OK;
}

void solve_ws_free(SolveWs& ws) {
    cudaFree(ws.W);
    cudaFree(ws.Linv);
    cudaFree(ws.tmp);
    cudaFree(ws.flag);
    ws = SolveWs();
}

int spd_inverse(double* S, int m, int MP, double* Sinv, double* d_logdet, SolveWs& ws, cudaStream_t st,
                int64_t* launches) {
    const int nblk = static_cast<int>(ceil_div(m, NB));
    const int64_t ld = MP;
    int rc;
    constexpr size_t kPotfSmem = sizeof(double) * 3 * NB * (NB + 1);
    static bool configured = false;
    if (!configured) {
        GPZ_CUDA(cudaFuncSetAttribute(potf2_trti_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kPotfSmem)));
        configured = true;
    }
    // ---- blocked Cholesky: S(lower) <- L
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB;
        const int nb = (m - k0 < NB) ? (m - k0) : NB;
        double* Lk = ws.Linv + static_cast<int64_t>(kb) * NB * NB;
        potf2_trti_kernel<<<1, 128, kPotfSmem, st>>>(S, ld, k0, nb, Lk, d_logdet, ws.flag, kb == 0);
        GPZ_KERNEL_CHECK();
        ++*launches;
        const int rem = m - k0 - nb;
        if (rem > 0) {
            double* panel = S + static_cast<int64_t>(k0 + nb) * ld + k0;
            // L_ik = A_ik * Linv_kk'   (B(k,j) = Linv[j][k])
            rc = sgemm(rem, nb, nb, 1.0, panel, ld, 1, Lk, 1, NB, 0.0, panel, ld, 0, st, launches);
            if (rc) return rc;
            // trailing: A_ij -= L_ik L_jk'  (lower tiles only)
            double* trail = S + static_cast<int64_t>(k0 + nb) * ld + (k0 + nb);
            rc = sgemm(rem, rem, nb, -1.0, panel, ld, 1, panel, 1, ld, 1.0, trail, ld, 1, st, launches);
            if (rc) return rc;
        }
    }
    // ---- W = L^{-1} (lower triangular) by recursive doubling: the diagonal 64-blocks are already inverted; at block size b
    //      every pair of neighbouring inverted blocks [1 | 2] is merged with W21 = -W22 * (L21 * W11), all pairs in one
    //      batched launch (the last pair may be ragged: rows of block 2 clipped to m).  Sinv serves as the scratch T.
    zero_kernel<<<static_cast<unsigned>(ceil_div(static_cast<int64_t>(MP) * MP, 256)), 256, 0, st>>>(
        ws.W, static_cast<int64_t>(MP) * MP);
    GPZ_KERNEL_CHECK();
    ++*launches;
    place_diag_kernel<<<nblk, 256, 0, st>>>(ws.Linv, ws.W, ld, m);
    GPZ_KERNEL_CHECK();
    ++*launches;
    for (int b = NB; b < m; b *= 2) {
        const int npairs = static_cast<int>(ceil_div(m - b, 2 * b));        // pairs whose block 2 is non-empty
        const int64_t bs = static_cast<int64_t>(2 * b) * ld + 2 * b;         // pointer step from pair to pair
        const double* L21 = S + static_cast<int64_t>(b) * ld;
        const double* W11 = ws.W;
        const double* W22 = ws.W + static_cast<int64_t>(b) * ld + b;
        double* T21 = Sinv + static_cast<int64_t>(b) * ld;
        double* W21 = ws.W + static_cast<int64_t>(b) * ld;
        // T21 = L21 * W11          (rows of block 2: min(b, m - (2pb + b)))
        rc = sgemm_batched(b, b, b, 1.0, L21, ld, 1, bs, W11, ld, 1, bs, 0.0, T21, ld, bs, npairs, m - b, 2 * b, 0, st, launches);
        if (rc) return rc;
        // W21 = -W22 * T21         (K = rows of block 2)
        rc = sgemm_batched(b, b, b, -1.0, W22, ld, 1, bs, T21, ld, 1, bs, 0.0, W21, ld, bs, npairs, m - b, 2 * b, 1, st, launches);
        if (rc) return rc;
    }
    // ---- Sinv = W' W (full symmetric): A(i,k) = W[k][i], B(k,j) = W[k][j]
    zero_kernel<<<static_cast<unsigned>(ceil_div(static_cast<int64_t>(MP) * MP, 256)), 256, 0, st>>>(
        Sinv, static_cast<int64_t>(MP) * MP);
    GPZ_KERNEL_CHECK();
    ++*launches;
    rc = sgemm(m, m, m, 1.0, ws.W, 1, ld, ws.W, ld, 1, 0.0, Sinv, ld, 0, st, launches);
    return rc;
}

}  // namespace gpz
