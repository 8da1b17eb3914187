// m x m SPD solve of the hot path: SIGMA^{-1} and ln det SIGMA (GPz/inv_logdet.m:1-15, called at
// GPz/GPz.m:67).  The reference pseudo-inverts by SVD; SIGMA = PHI' W PHI + diag(alpha) is SPD by
// construction (alpha > 0), so this is a blocked right-looking Cholesky (64-wide panels: diagonal
// block factored and inverted by one CTA in shared memory, panel solve and trailing update as DMMA
// GEMMs), a blocked triangular inverse (recursive doubling) and one W'W product.  A non-positive pivot raises a device
// flag; the caller turns that into NaN outputs (SURVEY.md 8b "error convention", H3).
#include "internal.cuh"

namespace gpz {

constexpr int NB = 64;

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
// all threads of the CTA
__device__ __forceinline__ void mm32(double (*Cm)[NB + 1], int ro, int co, double alpha, double (*X)[NB + 1], int xr, int xc,
                                     double (*Ym)[NB + 1], int yr, int yc, bool Yt, double beta, bool lower_only) {
    for (int e = threadIdx.x; e < 1024; e += blockDim.x) {
        const int i = e >> 5, j = e & 31;
        if (lower_only && j > i) continue;
        double s = 0.0;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) s += X[xr + i][xc + k] * (Yt ? Ym[yr + j][yc + k] : Ym[yr + k][yc + j]);
        Cm[ro + i][co + j] = alpha * s + (beta != 0.0 ? beta * Cm[ro + i][co + j] : 0.0);
    }
}

// right-looking Cholesky of the 64 x 64 smem tile A (lower, in place) by the whole CTA (256 threads).  Thread (r, q0) keeps
// the entries A[r][q], q = q0 (mod 4), q <= r, in REGISTERS for the whole factorisation; per column c the owners publish the
// (unscaled) column through a double-buffered smem vector, every thread derives the pivot's reciprocal square root itself
// (no broadcast, no fp64 division or log in the serial chain) and applies the rank-1 update to its registers.
// One barrier per column.  returns false (uniformly) on a non-positive / non-finite pivot.
__device__ __forceinline__ bool chol64(double (*A)[NB + 1], double (*colbuf)[NB]) {
    const int tid = threadIdx.x;
    const int r = tid >> 2, q0 = tid & 3;
    double reg[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int q = q0 + 4 * k;
        reg[k] = (q <= r) ? A[r][q] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) {                         // fully unrolled: register indices and the update range are static
        double* col = colbuf[c & 1];
        if ((c & 3) == q0 && r >= c) {                     // owner of (r, c): publish the finished, unscaled column entry
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (k == (c >> 2)) col[r] = reg[k];
        }
        __syncthreads();
        const double piv = col[c];
        if (!(piv > 0.0) || !(piv < 1.7e308)) return false;
        const double rl = rsqrt(piv);
        if (r >= c) {
            const double arc = col[r] * rl;                // L[r][c]
            if ((c & 3) == q0) A[r][c] = (r == c) ? piv * rl : arc;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int q = q0 + 4 * k;
                if (q > c && q <= r) reg[k] -= arc * (col[q] * rl);
            }
        }
    }
    __syncthreads();
    return true;
}
// (Tried in round 2 and dropped: the same factorisation by ONE warp with rows in registers and the column values exchanged by
// shuffles -- no block barrier at all, but 105 us per 64 x 64 block against 51 us for this version in the ncu launch list of an
// m = 1000 evaluation, gpurun_out r02k: a single warp cannot hide the shuffle -> FMA latency chain.)

// factor the nb x nb diagonal block at (k0,k0) in place (lower), invert the factor into Linv (row-major 64 x 64, zero
// upper triangle), accumulate logdet.  Cholesky of the 64 x 64 block by the whole CTA; the inverse as 2 x 2 blocks of 32:
// two warps invert the diagonal blocks concurrently, W21 = -W22 L21 W11 by the whole CTA.  Rows/cols >= nb are padded
// with the identity.
__global__ void __launch_bounds__(256)
potf2_trti_kernel(double* __restrict__ S, int64_t ld, int k0, int nb, double* __restrict__ Linv,
                  double* __restrict__ logdet, int* __restrict__ flag, int first) {
    extern __shared__ double sm_potf[];
    double (*A)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm_potf);
    double (*W)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm_potf + NB * (NB + 1));
    double (*T)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm_potf + 2 * NB * (NB + 1));
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < NB * NB; e += 256) {
        const int r = e / NB, c = e % NB;
        A[r][c] = (r < nb && c < nb && c <= r) ? S[static_cast<int64_t>(k0 + r) * ld + k0 + c] : (r == c ? 1.0 : 0.0);
        W[r][c] = 0.0;
    }
    __syncthreads();
    __shared__ double colbuf[2][NB];
    if (!chol64(A, colbuf)) {
        if (tid == 0) *flag = 1;
        return;
    }
    __shared__ double ldsh[2];
    if (warp < 2) {                               // 2 * sum_c log L_cc, fixed order
        double v = 2.0 * log(A[tid][tid]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) ldsh[warp] = v;
    }
    if (warp == 0) trinv32(A, W, 0);              // W11
    else if (warp == 1) trinv32(A, W, 32);        // W22
    __syncthreads();
    mm32(T, 32, 0, 1.0, A, 32, 0, W, 0, 0, false, 0.0, false);          // T21 = L21 * W11
    __syncthreads();
    mm32(W, 32, 0, -1.0, W, 32, 32, T, 32, 0, false, 0.0, false);       // W21 = -W22 * T21
    __syncthreads();
    for (int e = tid; e < NB * NB; e += 256) {
        const int r = e / NB, c = e % NB;
        Linv[e] = (r < nb && c < nb && c <= r) ? W[r][c] : 0.0;
        if (r < nb && c < nb && c <= r) S[static_cast<int64_t>(k0 + r) * ld + k0 + c] = A[r][c];
    }
    if (tid == 0) *logdet = (first ? 0.0 : *logdet) + (ldsh[0] + ldsh[1]);
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
    static PerDeviceOnce once;
    if (once.need()) {
        GPZ_CUDA(cudaFuncSetAttribute(potf2_trti_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kPotfSmem)));
    }
    // ---- blocked Cholesky: S(lower) <- L
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB;
        const int nb = (m - k0 < NB) ? (m - k0) : NB;
        double* Lk = ws.Linv + static_cast<int64_t>(kb) * NB * NB;
        potf2_trti_kernel<<<1, 256, kPotfSmem, st>>>(S, ld, k0, nb, Lk, d_logdet, ws.flag, kb == 0);
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
