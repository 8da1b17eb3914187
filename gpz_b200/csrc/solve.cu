// m x m SPD solve of the hot path: SIGMA^{-1} and ln det SIGMA (GPz/inv_logdet.m:1-15, called at
// GPz/GPz.m:67).  The reference pseudo-inverts by SVD; SIGMA = PHI' W PHI + diag(alpha) is SPD by
// construction (alpha > 0), so this is a blocked right-looking Cholesky (64-wide panels: diagonal
// block factored and inverted by one CTA in shared memory, panel solve and trailing update as DMMA
// GEMMs), a blocked triangular inverse (recursive doubling) and one W'W product.  A non-positive pivot raises a device
// flag; the caller turns that into NaN outputs (SURVEY.md 8b "error convention", H3).
#include <algorithm>
#include <vector>

#include "internal.cuh"

namespace gpz {

constexpr int NB = 64;
int g_solve_lookahead = 2;      // "solve_lookahead" option (process-wide): 0 in-stream, 1 look-ahead factorisation, 2 + incremental inverse (default)

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
// upper triangle), accumulate logdet.  Rows/cols >= nb are padded with the identity.
//
// This kernel is the serial bottleneck of the blocked solve (16 dependent launches at m = 1000; with 8 GPUs the solve is
// 15 % of an evaluation), so it is organised to have as few dependent steps as possible: the 64 x 64 block is factored
// in four 16-column steps.  Per step ONE warp factors the 16 x 16 diagonal sub-block with its rows in registers (pivot and
// column values move by shuffles: no block barrier inside) and inverts it by forward substitution with the pivots'
// reciprocal square roots it already has (no division anywhere); then all 256 threads form the sub-panel
// L21 = A21 L11^-T and apply the rank-16 update to the trailing sub-matrix.  The inverse of the whole factor follows from
// the four inverted diagonal sub-blocks by block forward substitution (three levels).  ~20 block barriers in total; the
// column-at-a-time version this replaces (chol64 above, kept for reference) needs 64 for the factorisation alone.
constexpr int SB = 16;

__device__ __forceinline__ void mm16(double (*Cm)[NB + 1], int ro, int co, double alpha, double (*X)[NB + 1], int xr, int xc,
                                     double (*Ym)[NB + 1], int yr, int yc, bool Yt, int kdepth) {
    // C(16 x 16 at (ro,co)) = alpha * X(16 x kdepth at (xr,xc)) * Y, Y(k,j) = Yt ? Ym[yr+j][yc+k] : Ym[yr+k][yc+j]; one thread per entry
    const int i = threadIdx.x >> 4, j = threadIdx.x & 15;
    double s0 = 0.0;
    for (int k = 0; k < kdepth; ++k) s0 = fma(X[xr + i][xc + k], Yt ? Ym[yr + j][yc + k] : Ym[yr + k][yc + j], s0);
    Cm[ro + i][co + j] = alpha * s0;
}

//
// Look-ahead form (LinvPrev != nullptr): the block row left of this diagonal block, A_{k,k-1}, still holds the values BEFORE
// panel k-1 was applied (the panel solve and trailing update of panel k-1 run concurrently on a side stream and leave this
// diagonal block alone).  The kernel first forms L_{k,k-1} = A_{k,k-1} Linv_{k-1}' itself and applies its rank-64 update to
// the diagonal block in shared memory, so that it depends only on the PREVIOUS diagonal block's kernel and on the trailing
// update of panel k-2.  L goes to Lout (== S in the in-place form).
__global__ void __launch_bounds__(256)
potf2_trti_kernel(const double* __restrict__ S, double* __restrict__ Lout, int64_t ld, int k0, int nb, double* __restrict__ Linv,
                  const double* __restrict__ LinvPrev, double* __restrict__ logdet, int* __restrict__ flag, int first,
                  double* __restrict__ Wfull) {
    extern __shared__ double sm_potf[];
    double (*A)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm_potf);
    double (*W)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm_potf + NB * (NB + 1));
    double (*T)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm_potf + 2 * NB * (NB + 1));
    __shared__ int ok_sm;
    __shared__ double ldsh[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // the three 64 x 64 blocks arrive by cp.async (all 48 copies of a thread in flight at once): as a load -> store loop the
    // prologue paid one L2 round trip per 256 elements, ~10 us of the kernel (profiles/r02z_solve_launches.csv)
    for (int e = tid; e < NB * NB; e += 256) {
        const int r = e / NB, c = e % NB;
        if (r < nb && c < nb && c <= r) cp_async8(&A[r][c], S + static_cast<int64_t>(k0 + r) * ld + k0 + c);
        else A[r][c] = (r == c) ? 1.0 : 0.0;
        if (LinvPrev != nullptr) {
            cp_async8(&W[r][c], LinvPrev + e);
            if (r < nb) cp_async8(&T[r][c], S + static_cast<int64_t>(k0 + r) * ld + (k0 - NB) + c);
            else T[r][c] = 0.0;
        } else {
            W[r][c] = 0.0;
        }
    }
    cp_async_commit();
    if (tid == 0) ok_sm = 1;
    cp_async_wait<0>();
    __syncthreads();
    if (LinvPrev != nullptr) {
        // Both 64 x 64 x 64 products on the DMMA pipe (8 warps x 8 rows x 64 columns each): as scalar FMA loops on 256 threads
        // they cost ~18 us of the kernel's ~50, as much as the look-ahead saves by not waiting for the panel update.
        const int g = lane >> 2, t = lane & 3, r0w = warp * 8;
        {   // L_{k,k-1} = A_{k,k-1} Linv_{k-1}'  (Linv is stored with zeros above its diagonal): rows r0w.. of T, in place
            double acc[8][2];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
#pragma unroll 4
            for (int kk = 0; kk < NB / 4; ++kk) {
                const double af = T[r0w + g][4 * kk + t];
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[j][0], acc[j][1], af, W[8 * j + g][4 * kk + t]);
            }
            __syncwarp();                      // a warp reads only its own rows of T
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                T[r0w + g][8 * j + 2 * t] = acc[j][0];
                T[r0w + g][8 * j + 2 * t + 1] = acc[j][1];
            }
        }
        __syncthreads();
        {   // A_kk -= L_{k,k-1} L_{k,k-1}'  (lower triangle)
            double acc[8][2];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
#pragma unroll 4
            for (int kk = 0; kk < NB / 4; ++kk) {
                const double af = T[r0w + g][4 * kk + t];
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[j][0], acc[j][1], af, T[8 * j + g][4 * kk + t]);
            }
            const int row = r0w + g;
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * j + 2 * t + e;
                    if (col <= row && row < nb) A[row][col] -= acc[j][e];
                }
        }
        for (int e = tid; e < NB * NB; e += 256) W[e / NB][e % NB] = 0.0;
        __syncthreads();
    }
    for (int jb = 0; jb < NB / SB; ++jb) {
        const int o = jb * SB;
        if (warp == 0) {
            // ---- 16 x 16 diagonal sub-block: Cholesky with lane = row (lanes 16..31 shadow lanes 0..15), then its inverse
            const int r = lane & 15;
            double a[SB], rd[SB];
#pragma unroll
            for (int q = 0; q < SB; ++q) a[q] = (q <= r) ? A[o + r][o + q] : 0.0;
            bool ok = true;
#pragma unroll
            for (int c = 0; c < SB; ++c) {
                const double piv = __shfl_sync(0xffffffffu, a[c], c);
                if (!(piv > 0.0) || !(piv < 1.7e308)) ok = false;
                const double rl = rsqrt(piv);
                rd[c] = rl;                                            // 1 / L[c][c]
                const double l = a[c] * rl;
                a[c] = (r == c) ? piv * rl : l;
#pragma unroll
                for (int q = c + 1; q < SB; ++q) {
                    const double bq = __shfl_sync(0xffffffffu, l, q);  // L[q][c]
                    a[q] = fma(-l, bq, a[q]);
                }
            }
            __syncwarp();                          // lanes 16..31 read the rows that lanes 0..15 now overwrite
            if (lane < SB) {
#pragma unroll
                for (int q = 0; q < SB; ++q)
                    if (q <= r) A[o + r][o + q] = a[q];
            }
            if (!ok && lane == 0) ok_sm = 0;
            __syncwarp();
            // inverse of the lower-triangular sub-block: lane = column, x_r = (delta_rc - sum_{q<r} L[r][q] x_q) / L[r][r]
            const int c = lane & 15;
            double x[SB];
#pragma unroll
            for (int rr = 0; rr < SB; ++rr) {
                double s0 = (rr == c) ? 1.0 : 0.0;
#pragma unroll
                for (int q = 0; q < rr; ++q) s0 = fma(-A[o + rr][o + q], x[q], s0);
                x[rr] = (rr >= c) ? s0 * rd[rr] : 0.0;
                if (lane < SB) W[o + rr][o + c] = x[rr];
            }
        }
        __syncthreads();
        if (!ok_sm) {
            if (tid == 0) *flag = 1;
            return;
        }
        const int rem = NB - o - SB;                                   // rows below this sub-block
        // ---- sub-panel L21 = A21 W11'  (W11 = L11^-1, lower): L21[r][c] = sum_{q<=c} A21[r][q] W11[c][q]  -> T, then back
        for (int e = tid; e < rem * SB; e += 256) {
            const int r = o + SB + e / SB, c = e % SB;
            double s0 = 0.0;
            for (int q = 0; q <= c; ++q) s0 = fma(A[r][o + q], W[o + c][o + q], s0);
            T[r][c] = s0;
        }
        __syncthreads();
        for (int e = tid; e < rem * SB; e += 256) {
            const int r = o + SB + e / SB, c = e % SB;
            A[r][o + c] = T[r][c];
        }
        // ---- trailing update A22 -= L21 L21' (lower triangle), from T
        for (int e = tid; e < rem * rem; e += 256) {
            const int i = e / rem, j = e % rem;
            if (j > i) continue;
            double s0 = 0.0;
#pragma unroll
            for (int c = 0; c < SB; ++c) s0 = fma(T[o + SB + i][c], T[o + SB + j][c], s0);
            A[o + SB + i][o + SB + j] -= s0;
        }
        __syncthreads();
    }
    if (warp < 2) {                               // 2 * sum_c log L_cc, fixed order
        double v = 2.0 * log(A[tid][tid]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((tid & 31) == 0) ldsh[warp] = v;
    }
    // ---- W = L^-1 by block forward substitution over the 4 x 4 grid of 16-blocks: W_ij = -W_ii (sum_{k=j}^{i-1} L_ik W_kj),
    //      blocks with the same i - j are independent
    for (int dist = 1; dist < NB / SB; ++dist) {
        for (int j = 0; j + dist < NB / SB; ++j) {
            const int i = j + dist;
            mm16(T, i * SB, j * SB, 1.0, A, i * SB, j * SB, W, j * SB, j * SB, false, dist * SB);      // T_ij = L_i,[j..i-1] W_[j..i-1],j
        }
        __syncthreads();
        for (int j = 0; j + dist < NB / SB; ++j) {
            const int i = j + dist;
            mm16(W, i * SB, j * SB, -1.0, W, i * SB, i * SB, T, i * SB, j * SB, false, SB);            // W_ij = -W_ii T_ij
        }
        __syncthreads();
    }
    for (int e = tid; e < NB * NB; e += 256) {
        const int r = e / NB, c = e % NB;
        const double wv = (r < nb && c < nb && c <= r) ? W[r][c] : 0.0;
        Linv[e] = wv;
        if (Wfull != nullptr && r < nb && c < nb) Wfull[static_cast<int64_t>(k0 + r) * ld + k0 + c] = wv;   // diagonal block of L^-1 in place
        if (r < nb && c < nb && c <= r) Lout[static_cast<int64_t>(k0 + r) * ld + k0 + c] = A[r][c];
    }
    if (tid == 0) *logdet = (first ? 0.0 : *logdet) + (ldsh[0] + ldsh[1]);
}

__global__ void zero_kernel(double* p, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0;
}

// copy the 64x64 inverse diagonal blocks into W
__global__ void place_diag_kernel(const double* __restrict__ Linv, double* __restrict__ W, int64_t ld, int m, int b0) {
    const int b = b0 + blockIdx.x;
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
    // look-ahead factorisation: L is written to its own matrix, the panel GEMMs run on a side stream
    GPZ_CUDA(cudaMalloc(&ws.Lbuf, sizeof(double) * MP * MP));
    GPZ_CUDA(cudaStreamCreateWithFlags(&ws.side, cudaStreamNonBlocking));
    GPZ_CUDA(cudaStreamCreateWithFlags(&ws.inv, cudaStreamNonBlocking));
    GPZ_CUDA(cudaStreamCreateWithFlags(&ws.acc, cudaStreamNonBlocking));
    for (int i = 0; i < SolveWs::MAXBLK; ++i) {
        GPZ_CUDA(cudaEventCreateWithFlags(&ws.evP[i], cudaEventDisableTiming));
        GPZ_CUDA(cudaEventCreateWithFlags(&ws.evT[i], cudaEventDisableTiming));
        GPZ_CUDA(cudaEventCreateWithFlags(&ws.evL[i], cudaEventDisableTiming));
        GPZ_CUDA(cudaEventCreateWithFlags(&ws.evW[i], cudaEventDisableTiming));
    }
    GPZ_CUDA(cudaEventCreateWithFlags(&ws.evF, cudaEventDisableTiming));
    GPZ_CUDA(cudaEventCreateWithFlags(&ws.evI, cudaEventDisableTiming));
    return GPZ_OK;
}

void solve_ws_free(SolveWs& ws) {
    cudaFree(ws.W);
    cudaFree(ws.Linv);
    cudaFree(ws.tmp);
    cudaFree(ws.flag);
    cudaFree(ws.Lbuf);
    if (ws.side) cudaStreamDestroy(ws.side);
    if (ws.inv) cudaStreamDestroy(ws.inv);
    if (ws.acc) cudaStreamDestroy(ws.acc);
    for (int i = 0; i < SolveWs::MAXBLK; ++i) {
        if (ws.evP[i]) cudaEventDestroy(ws.evP[i]);
        if (ws.evT[i]) cudaEventDestroy(ws.evT[i]);
        if (ws.evL[i]) cudaEventDestroy(ws.evL[i]);
        if (ws.evW[i]) cudaEventDestroy(ws.evW[i]);
    }
    if (ws.evF) cudaEventDestroy(ws.evF);
    if (ws.evI) cudaEventDestroy(ws.evI);
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
    // ---- blocked Cholesky.  Look-ahead form: the diagonal-block kernels form ONE dependent chain on `st` (each applies the
    //      previous panel's update to its own block itself, see potf2_trti_kernel), the panel solves and trailing updates form a
    //      second chain on the side stream; the two only meet through events: diagonal block k needs trailing update k-2,
    //      panel solve k needs diagonal block k.  L is written to ws.Lbuf (the panel solve is out of place so that the next
    //      diagonal kernel can still read A_{k+1,k}), S keeps being updated as the trailing matrix.
    const bool la = g_solve_lookahead && ws.side != nullptr && nblk >= 3 && nblk <= SolveWs::MAXBLK;
    // Incremental inverse (solve_lookahead = 2): W = L^-1 and Sinv = W'W do not have to wait for the whole factor.  Block row k of W
    // needs only block row k of L (complete after panel solve k-1), the inverted diagonal block k and the rows of W above it:
    //     W_k,0:k = -Linv_kk (L_k,0:k W_0:k,0:k),   W_kk = Linv_kk,   Sinv(0:k+1, 0:k+1) += W_k' W_k
    // so a third chain on ws.inv follows the diagonal chain one step behind, and what is left after the last diagonal block is one
    // block row (~50 us) instead of the 4-level recursive-doubling inverse plus an m^3 product (~0.3 ms at m = 1000).
    const bool inc = la && g_solve_lookahead >= 2 && ws.inv != nullptr;
    double* Lm = la ? ws.Lbuf : S;
    if (inc) {
        GPZ_CUDA(cudaMemsetAsync(ws.W, 0, sizeof(double) * static_cast<size_t>(MP) * MP, st));   // the diagonal kernels write into it
        GPZ_CUDA(cudaEventRecord(ws.evF, st));
        GPZ_CUDA(cudaStreamWaitEvent(ws.inv, ws.evF, 0));
        GPZ_CUDA(cudaStreamWaitEvent(ws.acc, ws.evF, 0));
        GPZ_CUDA(cudaMemsetAsync(Sinv, 0, sizeof(double) * static_cast<size_t>(MP) * MP, ws.acc));
    }
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB;
        const int nb = (m - k0 < NB) ? (m - k0) : NB;
        double* Lk = ws.Linv + static_cast<int64_t>(kb) * NB * NB;
        if (la && kb >= 2) GPZ_CUDA(cudaStreamWaitEvent(st, ws.evT[kb - 2], 0));
        potf2_trti_kernel<<<1, 256, kPotfSmem, st>>>(S, Lm, ld, k0, nb, Lk, (la && kb > 0) ? Lk - NB * NB : nullptr, d_logdet, ws.flag,
                                                     kb == 0, inc ? ws.W : nullptr);
        GPZ_KERNEL_CHECK();
        ++*launches;
        const int rem = m - k0 - nb;
        if (rem > 0) {
            const double* panelA = S + static_cast<int64_t>(k0 + nb) * ld + k0;
            double* panelL = Lm + static_cast<int64_t>(k0 + nb) * ld + k0;
            double* trail = S + static_cast<int64_t>(k0 + nb) * ld + (k0 + nb);
            cudaStream_t sg = st;
            if (la) {
                GPZ_CUDA(cudaEventRecord(ws.evP[kb], st));
                GPZ_CUDA(cudaStreamWaitEvent(ws.side, ws.evP[kb], 0));
                sg = ws.side;
            }
            // L_ik = A_ik * Linv_kk'   (B(k,j) = Linv[j][k])
            rc = sgemm(rem, nb, nb, 1.0, panelA, ld, 1, Lk, 1, NB, 0.0, panelL, ld, 0, sg, launches);
            if (rc) return rc;
            if (inc) GPZ_CUDA(cudaEventRecord(ws.evL[kb], ws.side));
            // trailing: A_ij -= L_ik L_jk'  (lower tiles only; look-ahead: not the next diagonal block, its kernel does that)
            rc = sgemm(rem, rem, nb, -1.0, panelL, ld, 1, panelL, 1, ld, 1.0, trail, ld, la ? 2 : 1, sg, launches);
            if (rc) return rc;
            if (la) GPZ_CUDA(cudaEventRecord(ws.evT[kb], ws.side));
        }
        if (inc) {
            // block row kb of W and its rank-nb contribution to Sinv, on the third chain
            if (rem <= 0) GPZ_CUDA(cudaEventRecord(ws.evP[kb], st));           // (recorded above when there is a panel below)
            GPZ_CUDA(cudaStreamWaitEvent(ws.inv, ws.evP[kb], 0));                // Linv_kk
            if (kb > 0) GPZ_CUDA(cudaStreamWaitEvent(ws.inv, ws.evL[kb - 1], 0));   // L_kb,0:kb
            double* Wk = ws.W + static_cast<int64_t>(k0) * ld;
            if (kb > 0) {
                // T = L_k,0:k0 * W_0:k0,0:k0   (nb x k0)
                rc = sgemm(nb, k0, k0, 1.0, Lm + static_cast<int64_t>(k0) * ld, ld, 1, ws.W, ld, 1, 0.0, ws.tmp, ld, 0, ws.inv, launches);
                if (rc) return rc;
                // W_k,0:k0 = -Linv_kk * T
                rc = sgemm(nb, k0, nb, -1.0, Lk, NB, 1, ws.tmp, ld, 1, 0.0, Wk, ld, 0, ws.inv, launches);
                if (rc) return rc;
            }
            // Sinv(0:k0+nb, 0:k0+nb) += W_k' W_k :  A(i,q) = W[k0+q][i], B(q,j) = W[k0+q][j]; on its own chain (ws.acc) so that the
            // next block row of W does not wait for it
            GPZ_CUDA(cudaEventRecord(ws.evW[kb], ws.inv));
            GPZ_CUDA(cudaStreamWaitEvent(ws.acc, ws.evW[kb], 0));
            rc = sgemm(k0 + nb, k0 + nb, nb, 1.0, Wk, 1, ld, Wk, ld, 1, 1.0, Sinv, ld, 0, ws.acc, launches);
            if (rc) return rc;
        }
    }
    if (inc) {                                       // join the side chains (ws.acc is behind ws.inv by construction); Sinv is complete
        GPZ_CUDA(cudaStreamWaitEvent(st, ws.evT[nblk - 2], 0));
        GPZ_CUDA(cudaEventRecord(ws.evI, ws.acc));
        GPZ_CUDA(cudaStreamWaitEvent(st, ws.evI, 0));
        return GPZ_OK;
    }
    if (la) {                                        // join: everything the side stream did precedes the inverse below
        const int last = nblk - 2;                   // the last panel with rows below it
        GPZ_CUDA(cudaStreamWaitEvent(st, ws.evT[last], 0));
    }
    // ---- W = L^{-1} (lower triangular) by recursive doubling: the diagonal 64-blocks are already inverted; at block size b
    //      every pair of neighbouring inverted blocks [1 | 2] is merged with W21 = -W22 * (L21 * W11), all pairs in one
    //      batched launch (the last pair may be ragged: rows of block 2 clipped to m).  Sinv serves as the scratch T.
    zero_kernel<<<static_cast<unsigned>(ceil_div(static_cast<int64_t>(MP) * MP, 256)), 256, 0, st>>>(
        ws.W, static_cast<int64_t>(MP) * MP);
    GPZ_KERNEL_CHECK();
    ++*launches;
    place_diag_kernel<<<nblk, 256, 0, st>>>(ws.Linv, ws.W, ld, m, 0);
    GPZ_KERNEL_CHECK();
    ++*launches;
    for (int b = NB; b < m; b *= 2) {
        const int npairs = static_cast<int>(ceil_div(m - b, 2 * b));        // pairs whose block 2 is non-empty
        const int64_t bs = static_cast<int64_t>(2 * b) * ld + 2 * b;         // pointer step from pair to pair
        const double* L21 = Lm + static_cast<int64_t>(b) * ld;
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

// ------------------------------------------------------------------------------------------------
// Pseudo-inverse and log-determinant with the reference's SVD semantics (GPz/inv_logdet.m:3-15):
//     [U,S,V] = svd(X); tol = m * eps(max s); keep s > tol; Xi = V diag(1/s) U'; logdet = sum(log s) over the kept ones.
// Used by gpz_inv_logdet when the Cholesky route fails or the factor says cond(X) is beyond 1 / (m eps) -- the regime where the
// reference truncates.  One-sided (Hestenes) Jacobi on the ROWS of G = V X, V orthogonal: rotations make the rows of G mutually
// orthogonal, then X = V' diag(s) W' with s_j = |g_j|, w_j = g_j / s_j, and pinv(X) = sum_j g_j v_j' / s_j^2 = G' diag(1/s^2) V.
// Round-robin ordering: M/2 disjoint row pairs per step (one CTA each), M - 1 steps per sweep; all reductions in fixed order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
jacobi_step_kernel(double* __restrict__ G, double* __restrict__ V, int64_t ld, int m, int M, int step, int* __restrict__ rotations) {
    __shared__ double sh[8];
    __shared__ double cs[2];
    const int i = blockIdx.x;
    int p, q;
    if (i == 0) {
        p = M - 1;
        q = step;
    } else {
        p = (step + i) % (M - 1);
        q = (step - i + (M - 1)) % (M - 1);
    }
    if (p >= m || q >= m) return;                      // padding row of an odd m
    if (p > q) {
        const int t = p;
        p = q;
        q = t;
    }
    double* gp = G + static_cast<int64_t>(p) * ld;
    double* gq = G + static_cast<int64_t>(q) * ld;
    double a = 0.0, b = 0.0, c = 0.0;
    for (int j = threadIdx.x; j < m; j += 256) {
        const double x = gp[j], y = gq[j];
        a = fma(x, x, a);
        b = fma(y, y, b);
        c = fma(x, y, c);
    }
    a = block_sum<256>(a, sh);
    b = block_sum<256>(b, sh);
    c = block_sum<256>(c, sh);
    if (threadIdx.x == 0) {
        double cc = 1.0, ss = 0.0;
        const double lim = 4.0 * 2.220446049250313e-16 * sqrt(a) * sqrt(b);
        if (fabs(c) > lim && a > 0.0 && b > 0.0) {
            const double zeta = (b - a) / (2.0 * c);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            cc = 1.0 / sqrt(1.0 + t * t);
            ss = cc * t;
            atomicAdd(rotations, 1);
        }
        cs[0] = cc;
        cs[1] = ss;
    }
    __syncthreads();
    const double cc = cs[0], ss = cs[1];
    if (ss == 0.0) return;
    double* vp = V + static_cast<int64_t>(p) * ld;
    double* vq = V + static_cast<int64_t>(q) * ld;
    for (int j = threadIdx.x; j < m; j += 256) {
        const double x = gp[j], y = gq[j];
        gp[j] = cc * x - ss * y;
        gq[j] = ss * x + cc * y;
        const double u = vp[j], w = vq[j];
        vp[j] = cc * u - ss * w;
        vq[j] = ss * u + cc * w;
    }
}

__global__ void identity_kernel(double* __restrict__ V, int64_t ld, int m) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<int64_t>(m) * ld) return;
    const int r = static_cast<int>(e / ld), c = static_cast<int>(e % ld);
    V[e] = (r == c && c < m) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(256) row_norm2_kernel(const double* __restrict__ G, int64_t ld, int m, double* __restrict__ n2) {
    __shared__ double sh[8];
    const double* g = G + static_cast<int64_t>(blockIdx.x) * ld;
    double a = 0.0;
    for (int j = threadIdx.x; j < m; j += 256) a = fma(g[j], g[j], a);
    a = block_sum<256>(a, sh);
    if (threadIdx.x == 0) n2[blockIdx.x] = a;
}

__global__ void scale_rows_kernel(double* __restrict__ G, int64_t ld, int m, const double* __restrict__ f) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<int64_t>(m) * ld) return;
    G[e] *= f[e / ld];
}

// G: [MP][MP] holding X (m x m valid, row-major) on entry -- destroyed; V, Xi: [MP][MP] scratch / result; *h_logdet on the host
int svd_pinv_logdet(double* G, int m, int MP, double* V, double* Xi, double* h_logdet, int* d_counter, double* d_vec, cudaStream_t st,
                    int64_t* launches) {
    const int64_t ld = MP;
    const int M = (m + 1) / 2 * 2;
    identity_kernel<<<static_cast<unsigned>(ceil_div(static_cast<int64_t>(m) * ld, 256)), 256, 0, st>>>(V, ld, m);
    GPZ_KERNEL_CHECK();
    int rot = 1;
    for (int sweep = 0; sweep < 40 && rot > 0 && M > 1; ++sweep) {
        GPZ_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(int), st));
        for (int step = 0; step < M - 1; ++step) {
            jacobi_step_kernel<<<M / 2, 256, 0, st>>>(G, V, ld, m, M, step, d_counter);
            ++*launches;
        }
        GPZ_KERNEL_CHECK();
        GPZ_CUDA(cudaMemcpyAsync(&rot, d_counter, sizeof(int), cudaMemcpyDeviceToHost, st));
        GPZ_CUDA(cudaStreamSynchronize(st));
    }
    row_norm2_kernel<<<m, 256, 0, st>>>(G, ld, m, d_vec);
    GPZ_KERNEL_CHECK();
    std::vector<double> n2(static_cast<size_t>(m));
    GPZ_CUDA(cudaMemcpyAsync(n2.data(), d_vec, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    GPZ_CUDA(cudaStreamSynchronize(st));
    double smax = 0.0;
    for (double v : n2) smax = fmax(smax, sqrt(v));
    int ex = 0;
    frexp(smax, &ex);                                             // smax in [2^(ex-1), 2^ex): eps(smax) = 2^(ex-1-52)
    const double tol = static_cast<double>(m) * (smax > 0.0 ? ldexp(1.0, ex - 53) : 0.0);      // inv_logdet.m:7
    double ldet = 0.0;
    std::vector<double> f(static_cast<size_t>(m), 0.0);
    std::vector<double> kept;
    for (int j = 0; j < m; ++j) {
        const double sj = sqrt(n2[j]);
        if (sj > tol) {
            f[j] = 1.0 / n2[j];
            kept.push_back(sj);
        }
    }
    std::sort(kept.begin(), kept.end(), [](double a, double b) { return a > b; });             // svd order, as sum(log(s)) adds them
    for (double sj : kept) ldet += log(sj);
    *h_logdet = ldet;
    GPZ_CUDA(cudaMemcpyAsync(d_vec, f.data(), sizeof(double) * m, cudaMemcpyHostToDevice, st));
    scale_rows_kernel<<<static_cast<unsigned>(ceil_div(static_cast<int64_t>(m) * ld, 256)), 256, 0, st>>>(G, ld, m, d_vec);
    GPZ_KERNEL_CHECK();
    GPZ_CUDA(cudaMemsetAsync(Xi, 0, sizeof(double) * MP * MP, st));
    // Xi = G' V  (A(i,k) = G[k][i], B(k,j) = V[k][j])
    return sgemm(m, m, m, 1.0, G, 1, ld, V, ld, 1, 0.0, Xi, ld, 0, st, launches);
}

// (max L_ii / min L_ii)^2 <= cond(X) for the Cholesky factor L: cheap lower bound used to decide whether the reference's
// SVD truncation would have been active
__global__ void __launch_bounds__(256) diag_minmax_kernel(const double* __restrict__ L, int64_t ld, int m, double* __restrict__ out) {
    __shared__ double smin[256], smax[256];
    double lo = 1.7e308, hi = 0.0;
    for (int j = threadIdx.x; j < m; j += 256) {
        const double v = fabs(L[static_cast<int64_t>(j) * ld + j]);
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    smin[threadIdx.x] = lo;
    smax[threadIdx.x] = hi;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < 256; ++t) {
            lo = fmin(lo, smin[t]);
            hi = fmax(hi, smax[t]);
        }
        out[0] = lo;
        out[1] = hi;
    }
}

int chol_diag_minmax(const SolveWs& ws, const double* S, int m, int MP, double* d_out2, cudaStream_t st) {
    const int nblk = static_cast<int>(ceil_div(m, NB));
    const bool la = g_solve_lookahead && ws.side != nullptr && nblk >= 3 && nblk <= SolveWs::MAXBLK;
    diag_minmax_kernel<<<1, 256, 0, st>>>(la ? ws.Lbuf : S, MP, m, d_out2);
    GPZ_KERNEL_CHECK();
    return GPZ_OK;
}

}  // namespace gpz
