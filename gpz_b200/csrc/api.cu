// C ABI of libgpz_b200 (include/gpz_b200.h): dataset context, the two-sweep NLML+gradient
// evaluation (GPz/GPz.m:1-263), the fit exit (GPz.m:84-87), getPHI, predict, inv_logdet, Dxy and the
// NCCL row-sharding plumbing.
//
// One evaluation =
//   sweep 1   PHI build (+ln-noise row-dot)  ->  row weights  ->  Gram PHI'W PHI and PHI'W y   [allreduce #1]
//   solve     SIGMA = Gram + diag(alpha): blocked Cholesky, SIGMA^-1, logdet, w, dw/dalpha     (replicated)
//   sweep 2   T = PHI SIGMA^-1 with fused nu / H epilogue -> row gradients -> dPHI -> back-projection
//             onto P, Gamma; column sums for dv, dlnAlpha; train/valid statistics               [allreduce #2]
//   finish    assemble nlogML, grad (scaled by -1/(n k)), the four statistics
// Everything is enqueued on one stream; the host only copies theta in and (f, g, stats) out.
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "internal.cuh"

namespace gpz {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen (libnccl.so.2 is only needed when gpz_comm_init is used)
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.h) return GPZ_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("cannot load libnccl.so.2: %s", dlerror());
        return GPZ_ERR_NCCL;
    }
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        set_error("libnccl.so.2 lacks required symbols");
        return GPZ_ERR_NCCL;
    }
    g_nccl.h = h;
    return GPZ_OK;
}

// timing events are host-visible only outside stream capture (inside, an event record would become a capture-internal
// dependency): a captured evaluation keeps the phase times of the last un-captured one
#define GPZ_EVREC(ev)                                        \
    do {                                                     \
        if (!c->capturing) GPZ_CUDA(cudaEventRecord(ev, st)); \
    } while (0)

#define GPZ_NCCL(call)                                                                                   \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) {                                                                        \
            gpz::set_error("%s:%d NCCL error: %s", __FILE__, __LINE__,                                    \
                           gpz::g_nccl.GetErrorString ? gpz::g_nccl.GetErrorString(r__) : "?");          \
            return GPZ_ERR_NCCL;                                                                         \
        }                                                                                                \
    } while (0)

// A synchronous cudaMemcpy from pageable host memory returns once the data is STAGED, not once it has landed in device
// memory (CUDA "API synchronization behavior"), and the context streams are non-blocking: order the upload explicitly
// before any kernel on those streams can read it.  Only used at context creation / predict set-up, never per evaluation.
static inline cudaError_t h2d(void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return e;
}

// ------------------------------------------------------------------------------------------------
// O(n) row kernels and small assembly kernels
// ------------------------------------------------------------------------------------------------
constexpr int RB = 256;        // rows per block of the row kernels
constexpr int KMAX = 4;        // fused multi-output limit of the element-wise kernels

// sweep 1 rows: lnbi = b + PHI v ; beta ; ob = omega*beta ; yw = ob*y ; partial sums
//   part[blk][0..k-1] = sum omega*lnbi_o,  [k] = sum omega,  [k+1] = row count
__global__ void __launch_bounds__(RB)
rows1_kernel(Params P, const double* __restrict__ Y, const double* __restrict__ omega, int64_t n, int64_t r0, int64_t r1,
             double* __restrict__ lnbi /*in: PHI v (het) ; out: ln beta^-1*/, double* __restrict__ beta,
             double* __restrict__ ob, double* __restrict__ yw /*[n][32]*/, double* __restrict__ part, int nsc) {
    __shared__ double sh[RB / 32];
    const int64_t i = r0 + static_cast<int64_t>(blockIdx.x) * RB + threadIdx.x;
    const bool live = i < r1;
    const double om = live ? omega[i] : 0.0;
    const int64_t blk = r0 / RB + blockIdx.x;
    for (int o = 0; o < P.k; ++o) {
        double ln = 0.0;
        if (live) {
            ln = P.bk[o] + (P.het ? lnbi[o * n + i] : 0.0);
            const double be = exp(-ln);
            lnbi[o * n + i] = ln;
            beta[o * n + i] = be;
            ob[o * n + i] = om * be;
            yw[i * 32 + o] = om * be * Y[o * n + i];
        }
        const double s = block_sum<RB>(om * ln, sh);
        if (threadIdx.x == 0) part[blk * nsc + o] = s;
    }
    if (live)
        for (int o = P.k; o < 32; ++o) yw[i * 32 + o] = 0.0;
    const double so = block_sum<RB>(om, sh);
    const double sc = block_sum<RB>(live ? 1.0 : 0.0, sh);
    if (threadIdx.x == 0) {
        part[blk * nsc + P.k] = so;
        part[blk * nsc + P.k + 1] = sc;
    }
}

// sweep 2 rows for output o: nu = sum of tile partials ; delta ; cw = -omega beta delta ; dbeta
//   part[blk][o] = sum ob*delta^2, [k+o] = sum dbeta, [2k] += sum omega delta^2, [2k+1] += sum omega(-.5 beta delta^2 - .5 lnbi)
__global__ void __launch_bounds__(RB)
rows2_kernel(Params P, int o, const double* __restrict__ Y, const double* __restrict__ omega, int64_t n, int64_t r0,
             int64_t r1, const double* __restrict__ pred, const double* __restrict__ nupart, int ntn,
             const double* __restrict__ lnbi, const double* __restrict__ beta, const double* __restrict__ ob,
             double* __restrict__ nu, double* __restrict__ cw, double* __restrict__ dbeta, double* __restrict__ part,
             int nsc) {
    __shared__ double sh[RB / 32];
    const int64_t i = r0 + static_cast<int64_t>(blockIdx.x) * RB + threadIdx.x;
    const bool live = i < r1;
    const int64_t blk = r0 / RB + blockIdx.x;
    const int k = P.k;
    double a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
    if (live) {
        double nv = 0.0;
        if (nupart != nullptr)
            for (int t = 0; t < ntn; ++t) nv += nupart[static_cast<int64_t>(t) * n + i];
        const double om = omega[i];
        const double be = beta[o * n + i];
        const double dl = pred[o * n + i] - Y[o * n + i];
        const double obd = ob[o * n + i] * dl;
        const double db = -0.5 * om * (1.0 - be * (dl * dl + nv));          // GPz.m:93
        nu[o * n + i] = nv;
        cw[o * n + i] = -obd;
        dbeta[o * n + i] = db;
        a1 = obd * dl;
        a2 = db;
        a3 = om * dl * dl;
        a4 = om * (-0.5 * be * dl * dl - 0.5 * lnbi[o * n + i]);
    }
    a1 = block_sum<RB>(a1, sh);
    a2 = block_sum<RB>(a2, sh);
    a3 = block_sum<RB>(a3, sh);
    a4 = block_sum<RB>(a4, sh);
    if (threadIdx.x == 0) {
        double* p = part + blk * nsc;
        p[o] = a1;
        p[k + o] = a2;
        p[2 * k] = (o == 0 ? 0.0 : p[2 * k]) + a3;
        p[2 * k + 1] = (o == 0 ? 0.0 : p[2 * k + 1]) + a4;
    }
}

// validation rows: lnbi_v = b + PHI_v v, pred_v = PHI_v w  ->  part[blk][0] += sum omega delta^2, [1] += LL sum, [2] = count
__global__ void __launch_bounds__(RB)
rowsv_kernel(Params P, int o, const double* __restrict__ Y, const double* __restrict__ omega, int64_t n, int64_t r0,
             int64_t r1, const double* __restrict__ dotv, const double* __restrict__ pred, double* __restrict__ part) {
    __shared__ double sh[RB / 32];
    const int64_t i = r0 + static_cast<int64_t>(blockIdx.x) * RB + threadIdx.x;
    const bool live = i < r1;
    const int64_t blk = r0 / RB + blockIdx.x;
    double a1 = 0.0, a2 = 0.0;
    if (live) {
        const double om = omega[i];
        const double ln = P.bk[o] + (P.het ? dotv[o * n + i] : 0.0);
        const double be = exp(-ln);
        const double dl = pred[o * n + i] - Y[o * n + i];
        a1 = om * dl * dl;
        a2 = om * (-0.5 * be * dl * dl - 0.5 * ln);
    }
    a1 = block_sum<RB>(a1, sh);
    a2 = block_sum<RB>(a2, sh);
    const double c = block_sum<RB>(live ? 1.0 : 0.0, sh);
    if (threadIdx.x == 0) {
        double* p = part + blk * 3;
        p[0] = (o == 0 ? 0.0 : p[0]) + a1;
        p[1] = (o == 0 ? 0.0 : p[1]) + a2;
        p[2] = c;
    }
}

// out[s] = sum_b part[b][s]  (fixed order)
__global__ void __launch_bounds__(256)
reduce_parts_kernel(const double* __restrict__ part, int64_t nblk, int nsc, double* __restrict__ out) {
    __shared__ double sh[8];
    const int s = blockIdx.x;
    double v = 0.0;
    for (int64_t b = threadIdx.x; b < nblk; b += 256) v += part[b * nsc + s];
    v = block_sum<256>(v, sh);
    if (threadIdx.x == 0) out[s] = v;
}

// dPHI_ij = -H_ij + PHI_ij * sum_o (cw_io w_jo + dbeta_io v_jo)     (GPz.m:72,90,106,113), in place over H
// column partial sums: colp[slab][o][MP] = sum_i PHI_ij * (-cw_io)  (= PHI'(omega beta delta), GPz.m:89)
//                      colp[slab][k+o][MP] = sum_i PHI_ij * dbeta_io (GPz.m:104)
__global__ void __launch_bounds__(128)
dphi_kernel(Params P, const double* __restrict__ Phi, double* __restrict__ H, int64_t ld, int64_t n, int64_t r0,
            int64_t r1, int64_t rows_per_slab, const double* __restrict__ cw, const double* __restrict__ dbeta,
            const double* __restrict__ w, double* __restrict__ colp, int accumulate) {
    const int j = blockIdx.x * 128 + threadIdx.x;
    const int k = P.k, MP = P.MP;
    const int64_t sb = r0 + static_cast<int64_t>(blockIdx.y) * rows_per_slab;
    int64_t se = sb + rows_per_slab;
    if (se > r1) se = r1;
    double wj[KMAX], vj[KMAX], sq[KMAX], sv[KMAX];
#pragma unroll
    for (int o = 0; o < KMAX; ++o) {
        wj[o] = (o < k) ? w[o * MP + j] : 0.0;
        vj[o] = (o < k) ? P.v[o * MP + j] : 0.0;
        sq[o] = sv[o] = 0.0;
    }
    for (int64_t ib = sb; ib < se; ib += 4) {
        double ph[4], hh[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {              // issue the loads of 4 rows before using any of them
            const int64_t i = ib + u;
            const int64_t off = (i - r0) * ld + j;
            ph[u] = (i < se) ? Phi[off] : 0.0;
            hh[u] = (i < se) ? H[off] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = ib + u;
            if (i >= se) break;
            double c = 0.0;
#pragma unroll
            for (int o = 0; o < KMAX; ++o) {
                if (o < k) {
                    const double cwi = __ldg(cw + o * n + i), dbi = __ldg(dbeta + o * n + i);
                    c = fma(cwi, wj[o], fma(dbi, vj[o], c));
                    sq[o] = fma(ph[u], -cwi, sq[o]);
                    sv[o] = fma(ph[u], dbi, sv[o]);
                }
            }
            H[(i - r0) * ld + j] = fma(ph[u], c, -hh[u]);
        }
    }
    double* out = colp + static_cast<int64_t>(blockIdx.y) * 2 * k * MP;
#pragma unroll
    for (int o = 0; o < KMAX; ++o) {
        if (o < k) {
            double a = sq[o], b = sv[o];
            if (accumulate) {
                a += out[o * MP + j];
                b += out[(k + o) * MP + j];
            }
            out[o * MP + j] = a;
            out[(k + o) * MP + j] = b;
        }
    }
}

// colsum[c][j] = sum_slab colp[slab][c][j]
__global__ void colsum_reduce_kernel(const double* __restrict__ colp, int nslab, int nc, int MP, double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nc * MP) return;
    double s = 0.0;
    for (int sl = 0; sl < nslab; ++sl) s += colp[static_cast<int64_t>(sl) * nc * MP + e];
    out[e] = s;
}

// one block per column: mean of the non-NaN entries (fixed summation order), then x -= mean in place; shift[a] = mean
__global__ void __launch_bounds__(1024) colmean_shift_kernel(double* __restrict__ X, int64_t n, double* __restrict__ shift) {
    __shared__ double sh[32];
    __shared__ double mean;
    double* col = X + static_cast<int64_t>(blockIdx.x) * n;
    double s = 0.0, c = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        const double v = col[i];
        if (v == v) s += v, c += 1.0;
    }
    const double S = gpz::block_sum<1024>(s, sh);
    const double Cn = gpz::block_sum<1024>(c, sh);
    if (threadIdx.x == 0) {
        mean = Cn > 0.0 ? S / Cn : 0.0;
        shift[blockIdx.x] = mean;
    }
    __syncthreads();
    const double mu = mean;
    for (int64_t i = threadIdx.x; i < n; i += 1024) col[i] -= mu;
}

__global__ void nan_if_flag_kernel(double* __restrict__ v, int64_t n, const int* __restrict__ flag) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n && *flag != 0) v[i] = nan("");
}

__global__ void add_diag_kernel(double* __restrict__ S, int MP, int m, const double* __restrict__ alpha) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < MP) S[static_cast<int64_t>(j) * MP + j] += (j < m) ? alpha[j] : 1.0;
}

// y[j] = scale * sum_l Sinv[j][l] * (x[l*xs] * (xmul ? xmul[l] : 1))      warp per row
__global__ void __launch_bounds__(256)
symv_kernel(const double* __restrict__ Sinv, int MP, int m, const double* __restrict__ x, int xs,
            const double* __restrict__ xmul, double scale, double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= MP) return;
    double s = 0.0;
    if (j < m)
        for (int l = lane; l < m; l += 32) s += Sinv[static_cast<int64_t>(j) * MP + l] * x[static_cast<int64_t>(l) * xs] * (xmul ? xmul[l] : 1.0);
    s = warp_sum(s);
    if (lane == 0) y[j] = (j < m) ? scale * s : 0.0;
}

// iSigma[l][m] := w[l] (l < m): the spare column of the T-GEMM's B operand, so that T[:, m] = PHI*w
__global__ void set_aug_col_kernel(double* __restrict__ Sinv, int MP, int m, const double* __restrict__ w) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < m) Sinv[static_cast<int64_t>(l) * MP + m] = w[l];
}

// out[0] = max_i |v_i|  (single CTA, order-independent)
__global__ void __launch_bounds__(1024)
max_abs_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
    __shared__ double sh[32];
    double mx = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) mx = fmax(mx, fabs(v[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 32; ++q) mx = fmax(mx, sh[q]);
        out[0] = mx;
    }
}

struct FinishArgs {
    Params P;
    const double* theta;
    const double* w;        // [k][MP]
    const double* dwda;     // [k][MP]
    const double* Sinv;     // [k][MP][MP]
    const double* logdet;   // [k]
    const double* scal1;    // [k]: sum omega lnbi_o ; [k]: sum omega ; [k+1]: n rows (global)
    const double* red2;     // dP | dG | q[k][MP] | dvraw[k][MP] | scal2
    const int* flag;
    int has_valid;
    double* out;            // nlogML | grad[p] | stats[4]
};

// grad(dP, dGamma) = -red2 / (n k)   (GPz.m:227-234), NaN when the solve flagged a bad pivot
__global__ void __launch_bounds__(256)
grad_scale_kernel(const double* __restrict__ red2, int64_t len, const double* __restrict__ scal1, int k, const int* __restrict__ flag,
                  double* __restrict__ grad) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (e >= len) return;
    const double sgn = -1.0 / (scal1[k + 1] * k);
    grad[e] = (*flag != 0) ? nan("") : sgn * red2[e];
}

// single CTA: assemble nlogML (GPz.m:81-82,103,110,233), the small gradients (GPz.m:89,104-105), pack and
// scale (GPz.m:227-234) and the statistics (GPz.m:236-259)
__global__ void __launch_bounds__(256)
finish_kernel(FinishArgs A) {
    __shared__ double sh[8];
    __shared__ double s_total;
    const Params& P = A.P;
    const int m = P.m, k = P.k, MP = P.MP, d = P.d;
    const int tid = threadIdx.x;
    const int64_t md = static_cast<int64_t>(m) * d;
    const double* dP = A.red2;
    const double* dG = dP + md;
    const double* q = dG + P.g_dim;
    const double* dvraw = q + static_cast<int64_t>(k) * MP;
    const double* sc2 = dvraw + static_cast<int64_t>(k) * MP;
    const double nrow = A.scal1[k + 1];
    const double nk = nrow * k;
    double* grad = A.out + 1;
    const bool bad = (*A.flag != 0);

    double tot = 0.0;
    for (int o = 0; o < k; ++o) {
        double s = 0.0;
        for (int j = tid; j < m; j += 256) {
            const double al = P.alpha[o * MP + j], wv = A.w[o * MP + j];
            s += -0.5 * al * wv * wv + 0.5 * A.theta[P.oA + static_cast<int64_t>(o) * m + j];
            if (P.het) {
                const double vv = P.v[o * MP + j], ta = P.tau[o * MP + j];
                s += -0.5 * vv * vv * ta + 0.5 * A.theta[P.oT + static_cast<int64_t>(o) * m + j];
            }
        }
        s = block_sum<256>(s, sh);
        if (tid == 0) {
            s += -0.5 * sc2[o] - 0.5 * A.logdet[o] - 0.5 * A.scal1[o];
            if (P.het) s += -0.5 * m * k * 1.8378770664093454836;     // m*k on every column (GPz.m:103)
            tot += s;
        }
    }
    if (tid == 0) {
        tot -= 0.5 * 1.8378770664093454836 * A.scal1[k];               // GPz.m:110
        s_total = tot;
    }
    __syncthreads();
    const double nan_ = nan("");
    const double sgn = -1.0 / nk;
    if (tid == 0) A.out[0] = bad ? nan_ : -s_total / nk;
    // (the m d + g_dim entries of dP, dGamma are scaled by grad_scale_kernel: one CTA copying 1.1e5 entries was 0.2 ms)
    for (int e = tid; e < m * k; e += 256) {
        const int o = e / m, j = e % m;
        const double al = P.alpha[o * MP + j], wv = A.w[o * MP + j], dw = A.dwda[o * MP + j];
        const double sii = A.Sinv[(static_cast<int64_t>(o) * MP + j) * MP + j];
        const double dla = -0.5 * sii * al - q[o * MP + j] * dw - al * wv * dw - 0.5 * al * wv * wv + 0.5;   // GPz.m:73,89
        grad[P.oA + e] = bad ? nan_ : sgn * dla;
        if (P.het) {
            const double vv = P.v[o * MP + j], ta = P.tau[o * MP + j];
            grad[P.oV + e] = bad ? nan_ : sgn * (dvraw[o * MP + j] - vv * ta);      // GPz.m:104
            grad[P.oT + e] = bad ? nan_ : sgn * (-0.5 * ta * vv * vv + 0.5);        // GPz.m:105
        }
    }
    if (tid < k) grad[P.oB + tid] = bad ? nan_ : sgn * sc2[k + tid];               // db, GPz.m:94
    if (tid == 0) {
        double* st = A.out + 1 + P.p;
        st[0] = bad ? nan_ : sqrt(sc2[2 * k] / nk);                                 // trainRMSE  GPz.m:236
        st[1] = bad ? nan_ : sc2[2 * k + 1] / nk - 0.5 * 1.8378770664093454836;     // trainLL    GPz.m:237
        if (A.has_valid) {
            const double nv = sc2[2 * k + 4] * k;
            st[2] = bad ? nan_ : sqrt(sc2[2 * k + 2] / nv);                         // validRMSE  GPz.m:258
            st[3] = bad ? nan_ : sc2[2 * k + 3] / nv - 0.5 * 1.8378770664093454836; // validLL    GPz.m:259
        } else {
            st[2] = nan_;
            st[3] = nan_;
        }
    }
}

// fit exit: nl[o] = -1/2 sum ob delta^2 - 1/2 sum alpha w^2 + 1/2 sum lnAlpha - 1/2 logdet - 1/2 sum omega lnbi  (GPz.m:81-82)
__global__ void __launch_bounds__(256)
fit_nl_kernel(Params P, const double* theta, const double* w, const double* logdet, const double* scal1,
              const double* sc2, const int* flag, double* out) {
    __shared__ double sh[8];
    for (int o = 0; o < P.k; ++o) {
        double s = 0.0;
        for (int j = threadIdx.x; j < P.m; j += 256) {
            const double al = P.alpha[o * P.MP + j], wv = w[o * P.MP + j];
            s += -0.5 * al * wv * wv + 0.5 * theta[P.oA + static_cast<int64_t>(o) * P.m + j];
        }
        s = block_sum<256>(s, sh);
        if (threadIdx.x == 0) out[o] = (*flag) ? nan("") : s - 0.5 * sc2[o] - 0.5 * logdet[o] - 0.5 * scal1[o];
    }
}

__global__ void nan_fill_kernel(double* p, int64_t n, const int* flag) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n && *flag) p[i] = nan("");
}

__global__ void sum_cols_kernel(const double* __restrict__ nupart, int ntn, int64_t stride, int64_t count, double* __restrict__ nu) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (int t = 0; t < ntn; ++t) s += nupart[static_cast<int64_t>(t) * stride + i];
    nu[i] = s;
}

__global__ void exp_rows_kernel(const double* __restrict__ dotv, const double* __restrict__ bk, int het, int k, int64_t n,
                                double* __restrict__ elns, double* __restrict__ beta_i) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int o = 0; o < k; ++o) {
        const double e = bk[o] + (het ? dotv[o * n + i] : 0.0);
        elns[o * n + i] = e;
        beta_i[o * n + i] = exp(e);
    }
}

}  // namespace gpz

using namespace gpz;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct gpz_ctx {
    Params P{};
    int device = 0;
    int sm_count = 148;
    cudaStream_t st = nullptr;
    RowData tr, va;
    int has_psi = 0;
    int64_t launches = 0;
    // NCCL
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // options
    int64_t opt_chunk_rows = 0;     // 0 = auto
    int opt_tensor_phi = -1;        // 2: PHI = exp(F W) with F W on the int8 tensor cores; 1: on the DMMA pipe; 0: direct-difference kernels
    int opt_fused_bp = 1;           // 1: fused dPHI + back-projection GEMM; 0: materialise dPHI first
    int opt_aug = 1;                // 1: spare-column trick (see aug)
    int opt_ozaki = -1;             // >0: T-GEMM through the int8 tensor cores with this many base-256 digits (ozaki.cu);
                                    // -1: default = 8 when the tcgen05 int8 GEMM was built in, else 0 (fp64 DMMA)
    void* oz_ws = nullptr;
    int8_t *oz_D8 = nullptr, *oz_F8 = nullptr;     // base-256 digits of PHI (plain, row-scaled) and of w .* PHI (ozaki.cu)
    double* oz_ea = nullptr;        // [rows] row scales of D8
    int opt_ozaki_gs = -1;          // digits used by the Gram (<= opt_ozaki)
    int opt_gc_fast = 1;            // GC + Psi through per-row factorisations + GEMMs (gcpsi.cu); 0: generic per-(i,j) kernels
    int opt_gc_int8 = 1;            // ... with its two K-major GEMMs (F W, dPHI G) on the int8 tensor cores when K > 128 (d >= 15)
    bool gc_fast = false;
    int gc_ns = 1;
    double* gc_ws = nullptr;
    int opt_ozaki_gram = -1;        // Gram through the int8 tensor cores too (-1: follows ozaki_slices > 0, k == 1)
    void* ozg_ws = nullptr;
    double* d_scal = nullptr;       // [0] max row weight of this eval, [1] max |y| (constant)
    std::vector<double> h_shift;    // constant subtracted from X at upload
    double* Wc_alloc = nullptr;
    double* dot_scratch = nullptr;
    int dphi_slabs = 1, fused_ns = 1;
    bool aug = false;               // PHI[:, m] carries y and iSigma[:, m] carries w: r and PHI*w come out of the two big GEMMs
    // workspaces (allocated at first use)
    bool ws_ready = false;
    int64_t chunk_rows = 0;
    bool resident = true;
    double *d_theta = nullptr, *d_out = nullptr;
    double *Phi = nullptr, *H = nullptr;
    double *lnbi = nullptr, *beta = nullptr, *ob = nullptr, *pred = nullptr, *nu = nullptr, *cw = nullptr, *dbeta = nullptr;
    double *nupart = nullptr, *yw = nullptr, *ones = nullptr;
    double *dotv_va = nullptr, *pred_va = nullptr;
    double *gram_partial = nullptr, *atb_partial = nullptr;
    int gram_ns = 1, atb_ns = 1, nslab = 1;
    double *red1 = nullptr, *red2 = nullptr;      // allreduce payloads
    int64_t red1_len = 0, red2_len = 0;
    double *S = nullptr, *Rvec = nullptr, *scal1 = nullptr;
    double *Sinv = nullptr, *w = nullptr, *dwda = nullptr, *logdet = nullptr;
    double *part1 = nullptr, *part2 = nullptr, *partv = nullptr;
    double *colp = nullptr, *bp_partial = nullptr, *Fbuf = nullptr, *Rm = nullptr, *scratch = nullptr;
    int QP = 32;
    SolveWs sws;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t kev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // + [4,5] around the int8 level GEMMs of row chunk 0   // [0,1] around the first Gram launch, [2,3] around the first T-GEMM launch
    cudaEvent_t kt[12] = {};   // pairs around single kernels of row chunk 0 (gpz_kernel_timing): PHI build, digits, Gram GEMM, moment GEMM
    bool kt_valid[6] = {};
    double* h_out = nullptr;   // pinned
    double* h_theta = nullptr; // pinned
    std::vector<void*> allocs;
    // launch-bound problems: the whole evaluation replayed as one CUDA graph (see eval_device)
    int opt_graph = -1;             // -1 auto (small single-GPU problems), 0 off, 1 on
    bool capturing = false;
    cudaGraphExec_t g_exec = nullptr;
    const double* g_theta = nullptr;  // pointer pair the graph was captured for
    double* g_out = nullptr;
    const double* seen_theta = nullptr;
    double* seen_out = nullptr;
    int64_t g_launches = 0;
    int64_t graph_replays = 0;
    bool last_was_replay = false;
    cudaEvent_t gev[2] = {nullptr, nullptr};       // around the last graph launch
};

namespace {

int parse_mode(const char* s) {
    static const char* names[6] = {"GL", "VL", "GD", "VD", "GC", "VC"};
    for (int i = 0; i < 6; ++i)
        if (s[0] == names[i][0] && s[1] == names[i][1]) return i;
    return -1;
}

int64_t g_dim_of(int mode, int m, int d) {
    switch (mode) {
        case GL: return 1;
        case VL: return m;
        case GD: return d;
        case VD: return static_cast<int64_t>(m) * d;
        case GC: return static_cast<int64_t>(d) * d;
        default: return static_cast<int64_t>(d) * d * m;
    }
}

int fill_params(Params& P, const gpz_model* model) {
    if (!model) {
        set_error("model is NULL");
        return GPZ_ERR_USAGE;
    }
    const int mode = parse_mode(model->method);
    if (mode < 0 || model->d < 1 || model->m < 1 || model->k < 1) {
        set_error("bad model (method '%c%c', d=%d, m=%d, k=%d)", model->method[0], model->method[1], model->d, model->m, model->k);
        return GPZ_ERR_USAGE;
    }
    if (model->k > KMAX) {
        set_error("k=%d outputs: at most %d supported", model->k, KMAX);
        return GPZ_ERR_USAGE;
    }
    if (model->d > 32 && mode_is_cov(mode)) {
        set_error("covariance modes support d <= 32 (got %d)", model->d);
        return GPZ_ERR_USAGE;
    }
    P.d = model->d;
    P.dp = static_cast<int>(round_up(model->d, 2));
    P.k = model->k;
    P.m = model->m;
    P.MP = static_cast<int>(round_up(model->m, TILE));
    P.mode = mode;
    P.het = model->heteroscedastic ? 1 : 0;
    P.g_dim = g_dim_of(mode, P.m, P.d);
    const int64_t md = static_cast<int64_t>(P.m) * P.d, mk = static_cast<int64_t>(P.m) * P.k;
    P.oG = md;
    P.oA = md + P.g_dim;
    P.oB = P.oA + mk;
    P.oV = P.oB + P.k;
    P.oT = P.oV + mk;
    P.p = P.oB + P.k + (P.het ? 2 * mk : 0);
    P.q = mode_is_cov(mode) ? 1 + P.d + P.d * (P.d + 1) / 2 : 1 + 2 * P.d;
    P.KQ = static_cast<int>(round_up(P.q, KSTEP));
    P.QP = static_cast<int>(round_up(P.q, 32));
    P.Wc = nullptr;
    P.xshift = nullptr;
    P.npat = 1;
    P.obs = nullptr;
    P.Mg = P.Gg = P.lndM = nullptr;
    return GPZ_OK;
}

template <class T>
int dev_alloc(std::vector<void*>& list, T** p, int64_t count) {
    *p = nullptr;
    if (count <= 0) count = 1;
    GPZ_CUDA(cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * static_cast<size_t>(count)));
    list.push_back(*p);
    return GPZ_OK;
}

int alloc_params(Params& P, std::vector<void*>& list, int need_sigma) {
    const int64_t MP = P.MP, d = P.d;
    int rc;
    if ((rc = dev_alloc(list, &P.Pt, d * MP))) return rc;
    if ((rc = dev_alloc(list, &P.Ct, d * MP))) return rc;
    if (!mode_is_cov(P.mode)) {
        if ((rc = dev_alloc(list, &P.Gt, d * MP))) return rc;
        P.Gam = P.Aj = P.Sj = P.lndS = nullptr;
    } else {
        P.Gt = nullptr;
        if ((rc = dev_alloc(list, &P.Gam, d * P.dp * MP))) return rc;
        if ((rc = dev_alloc(list, &P.Aj, d * d * MP))) return rc;
        if (need_sigma) {
            if ((rc = dev_alloc(list, &P.Sj, d * d * MP))) return rc;
            if ((rc = dev_alloc(list, &P.lndS, MP))) return rc;
        } else {
            P.Sj = P.lndS = nullptr;
        }
    }
    if ((rc = dev_alloc(list, &P.alpha, P.k * MP))) return rc;
    if ((rc = dev_alloc(list, &P.v, P.k * MP))) return rc;
    if ((rc = dev_alloc(list, &P.tau, P.k * MP))) return rc;
    if ((rc = dev_alloc(list, &P.bk, 32))) return rc;
    if ((rc = dev_alloc(list, &P.Wc, static_cast<int64_t>(P.npat) * P.KQ * MP))) return rc;
    if (mode_is_cov(P.mode)) {
        if ((rc = dev_alloc(list, &P.Mg, static_cast<int64_t>(P.npat) * d * d * MP))) return rc;
        if ((rc = dev_alloc(list, &P.Gg, static_cast<int64_t>(P.npat) * d * d * MP))) return rc;
        if ((rc = dev_alloc(list, &P.lndM, static_cast<int64_t>(P.npat) * MP))) return rc;
        if ((rc = dev_alloc(list, &P.obs, static_cast<int64_t>(P.npat) * d))) return rc;
        GPZ_CUDA(cudaMemset(P.obs, 1, static_cast<size_t>(P.npat) * d));
    }
    if ((rc = dev_alloc(list, &P.xshift, d))) return rc;
    GPZ_CUDA(cudaMemset(P.xshift, 0, sizeof(double) * d));
    return GPZ_OK;
}

int check_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); libgpz_b200 has no CPU fallback", cudaGetErrorString(e));
        return GPZ_ERR_NODEVICE;
    }
    if (device < 0 || device >= count) {
        set_error("device %d out of range (%d devices)", device, count);
        return GPZ_ERR_USAGE;
    }
    GPZ_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GPZ_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libgpz_b200 is built for sm_100a only", device, prop.major, prop.minor);
        return GPZ_ERR_NODEVICE;
    }
    return GPZ_OK;
}

// gather selected rows of a column-major n_all x cols host matrix into a [cols][n] host buffer
void gather_cols(const double* src, int64_t n_all, int cols, const std::vector<int64_t>& idx, std::vector<double>& dst) {
    const int64_t n = static_cast<int64_t>(idx.size());
    dst.resize(static_cast<size_t>(n) * cols);
    for (int c = 0; c < cols; ++c) {
        const double* s = src + static_cast<int64_t>(c) * n_all;
        double* o = dst.data() + static_cast<int64_t>(c) * n;
        for (int64_t i = 0; i < n; ++i) o[i] = s[idx[i]];
    }
}

int upload_rows(gpz_ctx* c, RowData& R, const std::vector<int64_t>& idx, int64_t n_all, const double* X, const double* Y,
                const double* Psi, const double* omega, bool zero_fill_nan) {
    const Params& P = c->P;
    const int64_t n = static_cast<int64_t>(idx.size());
    R.n = n;
    int rc;
    std::vector<double> buf;
    if ((rc = dev_alloc(c->allocs, &R.X, n * P.d))) return rc;
    gather_cols(X, n_all, P.d, idx, buf);
    R.has_nan = 0;
    for (double v : buf)
        if (v != v) { R.has_nan = 1; break; }
    for (int a = 0; a < P.d; ++a) {          // only x - p enters the path: store X relative to a fixed shift
        const double sh = c->h_shift[a];
        double* col = buf.data() + static_cast<int64_t>(a) * n;
        for (int64_t i = 0; i < n; ++i) col[i] -= sh;
    }
    if (zero_fill_nan) {                     // pattern-grouped covariance modes: missing dims drop out of the monomials
        for (double& v : buf)
            if (v != v) v = 0.0;
        R.has_nan = 0;
    }
    GPZ_CUDA(h2d(R.X, buf.data(), sizeof(double) * buf.size()));
    if ((rc = dev_alloc(c->allocs, &R.Y, n * P.k))) return rc;
    if (Y) {
        gather_cols(Y, n_all, P.k, idx, buf);
        GPZ_CUDA(h2d(R.Y, buf.data(), sizeof(double) * buf.size()));
    } else {
        GPZ_CUDA(cudaMemset(R.Y, 0, sizeof(double) * (n * P.k > 0 ? n * P.k : 1)));
    }
    if ((rc = dev_alloc(c->allocs, &R.omega, n))) return rc;
    buf.assign(static_cast<size_t>(n), 1.0);
    if (omega)
        for (int64_t i = 0; i < n; ++i) buf[i] = omega[idx[i]];
    GPZ_CUDA(h2d(R.omega, buf.data(), sizeof(double) * n));
    R.Psi = nullptr;
    if (Psi) {
        if (mode_is_cov(P.mode)) {
            const int64_t dd = static_cast<int64_t>(P.d) * P.d;
            buf.resize(static_cast<size_t>(n * dd));
            for (int64_t i = 0; i < n; ++i) {
                memcpy(buf.data() + i * dd, Psi + idx[i] * dd, sizeof(double) * dd);
                if (zero_fill_nan)               // only Psi(o,o) enters (getPHI.m:84): rows / columns of the missing dims may hold anything
                    for (int a = 0; a < P.d; ++a) {
                        const double xv = X[static_cast<int64_t>(a) * n_all + idx[i]];
                        if (xv == xv) continue;
                        for (int b = 0; b < P.d; ++b) buf[i * dd + a + b * P.d] = buf[i * dd + b + a * P.d] = 0.0;
                    }
            }
            if ((rc = dev_alloc(c->allocs, &R.Psi, n * dd))) return rc;
        } else {
            gather_cols(Psi, n_all, P.d, idx, buf);
            if ((rc = dev_alloc(c->allocs, &R.Psi, n * P.d))) return rc;
        }
        GPZ_CUDA(h2d(R.Psi, buf.data(), sizeof(double) * buf.size()));
    }
    return GPZ_OK;
}

int ensure_workspace(gpz_ctx* c) {
    if (c->ws_ready) return GPZ_OK;
    Params& P = c->P;
    const int64_t n = c->tr.n, nv = c->va.n, MP = P.MP, k = P.k;
    int rc;
    auto A = [&](double** p, int64_t cnt) { return dev_alloc(c->allocs, p, cnt); };
    // row chunking: keep PHI and H (2 x rows x MP doubles) within ~55% of the free memory
    size_t free_b = 0, total_b = 0;
    GPZ_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // default GEMM engine: the int8 digit GEMMs (ozmma.cu), except for a single 128-wide basis tile on few rows, where half
    // of every 256-row UMMA tile is padding and the launch is latency-bound: there the fp64 DMMA kernels win
    // (tools/small_latency.py: m=100, 2.4e3..1.6e5 rows: 0.38/0.41/1.03 ms against 0.43/0.48/1.19 ms per evaluation;
    // m>=500 is faster on the int8 path from a few thousand rows on)
    if (c->opt_ozaki < 0) c->opt_ozaki = (ozmma_available() && !(MP <= 128 && n < 250000)) ? 7 : 0;
    const double bytes_per_elt = 16.0 + (c->opt_ozaki > 0 ? 2.0 * c->opt_ozaki : 0.0);     // PHI, H (+ two digit sets)
    int64_t max_rows = static_cast<int64_t>(0.55 * static_cast<double>(free_b) / (bytes_per_elt * MP));
    max_rows = max_rows / 1024 * 1024;
    if (max_rows < 1024) max_rows = 1024;
    int64_t need = n > nv ? n : nv;
    if (need < 1) need = 1;
    c->chunk_rows = need < max_rows ? need : max_rows;
    if (c->opt_chunk_rows > 0) {
        c->chunk_rows = round_up(c->opt_chunk_rows, 1024);
        if (c->chunk_rows > round_up(need, 1024)) c->chunk_rows = round_up(need, 1024);
    }
    c->resident = c->chunk_rows >= n;
    if ((rc = A(&c->Phi, c->chunk_rows * MP))) return rc;
    if ((rc = A(&c->H, c->chunk_rows * MP))) return rc;
    if ((rc = A(&c->d_theta, P.p))) return rc;
    if ((rc = A(&c->d_out, P.p + 5))) return rc;
    const int64_t nn = n > 0 ? n : 1;
    if ((rc = A(&c->lnbi, k * nn))) return rc;
    if ((rc = A(&c->beta, k * nn))) return rc;
    if ((rc = A(&c->ob, k * nn))) return rc;
    if ((rc = A(&c->pred, k * nn))) return rc;
    if ((rc = A(&c->nu, k * nn))) return rc;
    if ((rc = A(&c->cw, k * nn))) return rc;
    if ((rc = A(&c->dbeta, k * nn))) return rc;
    if ((rc = A(&c->nupart, (MP / TILE) * nn))) return rc;
    if ((rc = A(&c->yw, 32 * nn))) return rc;
    {
        const int64_t no = need;
        if ((rc = A(&c->ones, no))) return rc;
        std::vector<double> one(static_cast<size_t>(no), 1.0);
        GPZ_CUDA(h2d(c->ones, one.data(), sizeof(double) * no));
        c->tr.gc_ones = c->va.gc_ones = c->ones;
    }
    if ((rc = A(&c->dotv_va, k * (nv > 0 ? nv : 1)))) return rc;
    if ((rc = A(&c->pred_va, k * (nv > 0 ? nv : 1)))) return rc;
    // GEMM partial buffers
    const int T = static_cast<int>(MP / TILE);
    const int ntri = T * (T + 1) / 2;
    c->gram_ns = gram_nsplit(static_cast<int>(MP), c->sm_count);
    if ((rc = A(&c->gram_partial, static_cast<int64_t>(k) * c->gram_ns * ntri * TILE * TILE))) return rc;   // per output: row chunks accumulate
    c->QP = P.QP;
    const int tn = c->QP / 32;
    c->atb_ns = c->sm_count / (T * tn) > 0 ? c->sm_count / (T * tn) : 1;
    {
        const int ns1 = c->sm_count / T > 0 ? c->sm_count / T : 1;     // the PHI'(ob*y) product has QP = 32
        c->fused_ns = ns1;
        int64_t a = static_cast<int64_t>(c->atb_ns) * T * tn * TILE * 32;
        const int64_t b = static_cast<int64_t>(ns1) * T * TILE * 32;
        const int64_t f = static_cast<int64_t>(c->fused_ns) * T * TILE * c->QP;
        if (b > a) a = b;
        if (f > a) a = f;
        if ((rc = A(&c->atb_partial, a))) return rc;
    }
    c->nslab = (2 * c->sm_count) / T > 0 ? (2 * c->sm_count) / T : 1;
    c->dphi_slabs = 16 * c->nslab;
    // (capping these split factors for few rows was tried: 4 splits instead of 148 at 2 400 rows makes the Gram 2.4x
    // slower -- the serial K loop per CTA costs more than adding up 148 partial tiles, tools/small_latency.py)
    if ((rc = A(&c->dot_scratch, 2 * (MP / TILE) * c->chunk_rows))) return rc;
    // allreduce payloads
    c->red1_len = k * MP * MP + MP * 32 + (k + 2);
    if ((rc = A(&c->red1, c->red1_len))) return rc;
    c->S = c->red1;
    c->Rvec = c->S + k * MP * MP;
    c->scal1 = c->Rvec + MP * 32;
    c->red2_len = static_cast<int64_t>(P.m) * P.d + P.g_dim + 2 * k * MP + (2 * k + 5);
    if ((rc = A(&c->red2, c->red2_len))) return rc;
    if ((rc = A(&c->Sinv, k * MP * MP))) return rc;
    if ((rc = A(&c->w, k * MP))) return rc;
    if ((rc = A(&c->dwda, k * MP))) return rc;
    if ((rc = A(&c->logdet, 32))) return rc;
    const int64_t nb1 = ceil_div(nn, RB) + 1, nbv = ceil_div(nv > 0 ? nv : 1, RB) + 1;
    if ((rc = A(&c->part1, nb1 * (k + 2)))) return rc;
    if ((rc = A(&c->part2, nb1 * (2 * k + 2)))) return rc;
    if ((rc = A(&c->partv, nbv * 3))) return rc;
    {
        const int64_t sl = c->dphi_slabs > c->fused_ns ? c->dphi_slabs : c->fused_ns;
        if ((rc = A(&c->colp, sl * 2 * k * MP))) return rc;
    }
    const int64_t bpd = backproj_partial_doubles(P, c->nslab, c->has_psi, c->tr.has_nan);
    if ((rc = A(&c->bp_partial, bpd))) return rc;
    // GC + Psi without missing inputs: per-row factorisation + GEMMs instead of the per-(i,j) kernels (gcpsi.cu)
    c->gc_fast = P.mode == GC && c->has_psi && c->opt_gc_fast && c->tr.g_pat.size() <= 1 && !c->tr.has_nan &&
                 (c->tr.g_pat.empty() || c->tr.g_pat[0] == 0) && c->va.g_pat.size() <= 1 && (c->va.g_pat.empty() || c->va.g_pat[0] == 0);
    if (c->gc_fast) {
        const int KQ = gc_feature_width(P.d);
        for (RowData* R : {&c->tr, &c->va}) {
            if (R->n <= 0) continue;
            R->gc_chunk = R->n < c->chunk_rows ? R->n : c->chunk_rows;
            if ((rc = A(&R->gcF, R->gc_chunk * KQ)) || (rc = A(&R->gcW, static_cast<int64_t>(KQ) * MP)) || (rc = A(&R->gcG, MP * round_up(KQ, TILE)))) return rc;
        }
        c->gc_ns = c->sm_count / (static_cast<int>(MP / TILE) * (KQ / 32)) > 0 ? c->sm_count / (static_cast<int>(MP / TILE) * (KQ / 32)) : 1;
        if ((rc = A(&c->gc_ws, gc_backproj_ws_doubles(P, c->tr.gc_chunk > 0 ? c->tr.gc_chunk : 1, c->gc_ns, c->sm_count)))) return rc;
    }
    const bool fast_bp = !c->has_psi && !c->tr.has_nan;
    if (fast_bp) {
        if ((rc = A(&c->Rm, MP * c->QP))) return rc;
        // monomial row features are dataset constants: built once, resident
        if ((rc = A(&c->tr.F, (n > 0 ? n : 1) * c->QP))) return rc;
        if ((rc = build_features(P, c->tr.X, n, 0, n, c->tr.F, c->st, &c->launches))) return rc;
    }
    if (!c->has_psi && !c->va.has_nan && nv > 0) {
        if ((rc = A(&c->va.F, nv * c->QP))) return rc;
        if ((rc = build_features(P, c->va.X, nv, 0, nv, c->va.F, c->st, &c->launches))) return rc;
    }
    c->Wc_alloc = P.Wc;
    if (!c->opt_tensor_phi) P.Wc = nullptr;
    // measured (n=1e6): the int8 variant is epilogue-instruction bound (7 level folds + exp per element: VD m=500 4.6 ms, VC m=1000
    // 10.9 ms) and does not beat the DMMA kernel (2.9 / 10.0 ms), so the DMMA path stays the default; tensor_phi=2 selects it
    if (c->opt_tensor_phi < 0) c->opt_tensor_phi = 1;
    if (c->opt_tensor_phi == 2 && P.Wc != nullptr && P.KQ <= 128 && ozmma_available()) {
        // PHI = exp(F W) with the quadratic forms on the int8 tensor cores: the digits of the row features are dataset constants
        for (RowData* R : {&c->tr, &c->va}) {
            if (R->F == nullptr || R->n <= 0) continue;
            double* tmp = nullptr;
            if ((rc = A(&tmp, oz_padded_rows(R->n) * 7 * 128 / 8 + 1))) return rc;
            R->FD8 = reinterpret_cast<int8_t*>(tmp);
            if ((rc = A(&R->eaF, oz_padded_rows(R->n)))) return rc;
            if ((rc = A(&tmp, MP * 7 * 128 / 8 + 1))) return rc;
            R->WD8 = reinterpret_cast<int8_t*>(tmp);
            if ((rc = A(&R->ebW, MP))) return rc;
            R->phi_digits = 7;
            if ((rc = ozaki_feature_digits(R->F, c->QP, P.q, R->n, 7, R->FD8, R->eaF, nullptr, c->st, &c->launches))) return rc;
        }
    }
    c->aug = fast_bp && P.Wc != nullptr && k == 1 && P.m < MP && c->opt_aug;
    if (c->aug) c->tr.ycol = c->tr.Y;
    {
        const int64_t dd = static_cast<int64_t>(P.d) * P.d;
        const int64_t sc = 3 * dd * MP + dd * P.m + static_cast<int64_t>(P.m) * P.d + 64;
        if ((rc = A(&c->scratch, sc))) return rc;
    }
    if (c->opt_ozaki > 0) {
        if (!ozmma_available()) {
            set_error("ozaki_slices: the driver does not export cuTensorMapEncodeTiled (needed by the tcgen05 digit GEMM)");
            return GPZ_ERR_USAGE;
        }
        const int64_t rows = n < c->chunk_rows ? (n > 0 ? n : 1) : c->chunk_rows;
        double* tmp = nullptr;
        if ((rc = A(&tmp, oz_workspace_bytes(static_cast<int>(MP), c->opt_ozaki) / 8 + 1))) return rc;
        c->oz_ws = tmp;
        if ((rc = A(&tmp, oz_digit_bytes(static_cast<int>(MP), c->opt_ozaki, rows) / 8 + 1))) return rc;
        c->oz_D8 = reinterpret_cast<int8_t*>(tmp);
        if ((rc = A(&c->oz_ea, oz_padded_rows(rows)))) return rc;
    }
    if (c->opt_ozaki_gram < 0) c->opt_ozaki_gram = (c->opt_ozaki > 0 && k == 1) ? 1 : 0;
    if (c->opt_ozaki_gram > 0 && (c->opt_ozaki <= 0 || k != 1)) c->opt_ozaki_gram = 0;
    // measured on a Gram-like product of real PHI values (tools/gram_digits_accuracy.py, K = 16384, long-double reference):
    // max error / max|S| = 9.6e-17 with 7 digits, 8.8e-15 with 6, 1.7e-12 with 5; fp64 BLAS 7.5e-16.  cond(SIGMA) ~ 1e8-1e9
    // amplifies the error of S into the gradient, so the Gram keeps all 7 digits (6 would save ~3.5 ms of 64 at the headline size
    // and still match the fp64 DMMA path, but not the 1e-5 tolerance with margin); "ozaki_gram_slices" overrides
    if (c->opt_ozaki_gs < 0) c->opt_ozaki_gs = c->opt_ozaki;
    if (c->opt_ozaki_gs > c->opt_ozaki) c->opt_ozaki_gs = c->opt_ozaki;
    if ((rc = A(&c->d_scal, 8))) return rc;
    if (c->opt_ozaki_gram > 0) {
        double* tmp = nullptr;
        const int64_t rows = n < c->chunk_rows ? (n > 0 ? n : 1) : c->chunk_rows;
        if ((rc = A(&tmp, oz_gram_workspace_bytes(static_cast<int>(MP), rows) / 8 + 1))) return rc;
        c->ozg_ws = tmp;
        if ((rc = A(&tmp, oz_digit_bytes(static_cast<int>(MP), c->opt_ozaki, rows) / 8 + 1))) return rc;
        c->oz_F8 = reinterpret_cast<int8_t*>(tmp);
        max_abs_kernel<<<1, 1024, 0, c->st>>>(c->tr.Y, n, c->d_scal + 1);
        GPZ_KERNEL_CHECK();
    }
    if (c->gc_fast && c->opt_gc_int8 && c->opt_ozaki > 0 && gc_feature_width(P.d) > 128) {
        // digit GEMMs of the GC + Psi path: the left operands' digits live in the PHI digit buffer (free at those points of the
        // evaluation: before the digits of PHI are made / after the T-GEMM consumed them)
        const int KQ = gc_feature_width(P.d), K128 = static_cast<int>(round_up(KQ, 128)), KN = static_cast<int>(round_up(KQ, TILE));
        double* t1 = nullptr;
        double* t2 = nullptr;
        double *e1 = nullptr, *e2 = nullptr;
        if ((rc = A(&t1, MP * 7 * K128 / 8 + 1)) || (rc = A(&t2, static_cast<int64_t>(KN) * 7 * MP / 8 + 1)) || (rc = A(&e1, MP)) || (rc = A(&e2, KN))) return rc;
        int8_t* a8 = c->oz_D8;
        if (K128 > MP) {                      // few bases: the feature digits are wider than a PHI row's
            const int64_t rows = n < c->chunk_rows ? (n > 0 ? n : 1) : c->chunk_rows;
            double* t3 = nullptr;
            if ((rc = A(&t3, oz_padded_rows(rows) * 7 * K128 / 8 + 1))) return rc;
            a8 = reinterpret_cast<int8_t*>(t3);
        }
        const int64_t crow = n < c->chunk_rows ? (n > 0 ? n : 1) : c->chunk_rows;
        double *t4 = nullptr, *e3 = nullptr, *t5 = nullptr;
        int8_t* x8 = c->oz_F8;
        if (x8 == nullptr) {                  // Gram not on the digit kernel: no second PHI digit buffer to borrow
            double* t6 = nullptr;
            if ((rc = A(&t6, oz_digit_bytes(static_cast<int>(MP), 7, crow) / 8 + 1))) return rc;
            x8 = reinterpret_cast<int8_t*>(t6);
        }
        if ((rc = A(&t4, oz_padded_rows(crow) * 7 * K128 / 8 + 1)) || (rc = A(&e3, oz_padded_rows(crow))) ||
            (rc = A(&t5, oz_moment_workspace_bytes(static_cast<int>(MP), KQ, crow) / 8 + 1))) return rc;
        for (RowData* R : {&c->tr, &c->va}) {
            if (R->n > n) continue;           // the digit buffer is sized for the training rows of a chunk: a larger validation set keeps the DMMA GEMMs
            R->gc_digits = 7;
            R->gcF8 = reinterpret_cast<int8_t*>(t4);
            R->gcEaF = e3;
            R->gcX8 = x8;                     // the second PHI digit buffer: free after the Gram
            R->gcMws = t5;
            R->gcA8 = a8;
            R->gcEa = c->oz_ea;
            R->gcWD8 = reinterpret_cast<int8_t*>(t1);
            R->gcEbW = e1;
            R->gcGD8 = reinterpret_cast<int8_t*>(t2);
            R->gcEbG = e2;
        }
    }
    if ((rc = solve_ws_alloc(c->sws, static_cast<int>(MP)))) return rc;
    c->tr.flag = c->va.flag = c->sws.flag;          // non-finite coefficients seen by the int8 PHI build raise the same failure flag
    for (auto& e : c->ev) GPZ_CUDA(cudaEventCreate(&e));
    for (auto& e : c->kev) GPZ_CUDA(cudaEventCreate(&e));
    for (auto& e : c->gev) GPZ_CUDA(cudaEventCreate(&e));
    for (auto& e : c->kt) GPZ_CUDA(cudaEventCreate(&e));
    GPZ_CUDA(cudaMallocHost(&c->h_out, sizeof(double) * (P.p + 5)));
    GPZ_CUDA(cudaMallocHost(&c->h_theta, sizeof(double) * P.p));
    c->ws_ready = true;
    return GPZ_OK;
}

int allreduce(gpz_ctx* c, double* buf, int64_t len) {
    if (c->world <= 1 || c->comm == nullptr) return GPZ_OK;
    GPZ_NCCL(g_nccl.AllReduce(buf, buf, static_cast<size_t>(len), ncclDouble, ncclSum, c->comm, c->st));
    return GPZ_OK;
}

// sweep 1 + solve.  Leaves Sinv, w, dwda, logdet, scal1 (global) on the device.
int forward_and_solve(gpz_ctx* c, const double* d_theta) {
    Params& P = c->P;
    cudaStream_t st = c->st;
    const int64_t n = c->tr.n, MP = P.MP;
    const int k = P.k;
    int rc;
    GPZ_CUDA(cudaMemsetAsync(c->sws.flag, 0, sizeof(int), st));
    GPZ_EVREC(c->ev[0]);
    if ((rc = prep_params(d_theta, P, c->has_psi, st, &c->launches))) return rc;
    const int T = static_cast<int>(MP / TILE);
    const int ns1 = c->sm_count / T > 0 ? c->sm_count / T : 1;
    int nchunks = 0;
    for (int64_t r0 = 0; r0 < n || (n == 0 && r0 == 0); r0 += c->chunk_rows) {
        const int64_t r1 = (r0 + c->chunk_rows < n) ? r0 + c->chunk_rows : n;
        const bool last = r1 >= n;
        double* phi = c->resident ? c->Phi + r0 * MP : c->Phi;
        if (n > 0) {
            DotSpec ds{0, {nullptr, nullptr}, {nullptr, nullptr}};
            if (P.het && k == 1) ds = DotSpec{1, {P.v, nullptr}, {c->lnbi, nullptr}};
            const bool kt_on = r0 == 0 && !c->capturing;
            if (kt_on) GPZ_CUDA(cudaEventRecord(c->kt[0], st));
            if ((rc = phi_build(P, c->tr, r0, r1, phi, ds, c->dot_scratch, st, &c->launches))) return rc;
            if (kt_on) {
                GPZ_CUDA(cudaEventRecord(c->kt[1], st));
                c->kt_valid[0] = true;
            }
            if (P.het && k > 1)
                for (int o = 0; o < k; ++o)
                    if ((rc = rowdot(phi, MP, P.m, r1 - r0, DotSpec{1, {P.v + o * MP, nullptr}, {c->lnbi + o * n + r0, nullptr}}, st, &c->launches))) return rc;
            rows1_kernel<<<static_cast<unsigned>(ceil_div(r1 - r0, RB)), RB, 0, st>>>(P, c->tr.Y, c->tr.omega, n, r0, r1, c->lnbi, c->beta,
                                                                                    c->ob, c->yw, c->part1, k + 2);
            GPZ_KERNEL_CHECK();
            ++c->launches;
        }
        if (nchunks == 0) GPZ_EVREC(c->ev[1]);
        for (int o = 0; o < k; ++o) {
            // rows of this chunk are [0, r1-r0) of phi; the weights are indexed by absolute row
            const bool timed = (nchunks == 0 && o == 0);
            if (timed) GPZ_EVREC(c->kev[0]);
            if (c->opt_ozaki_gram > 0) {
                if (nchunks == 0) {
                    max_abs_kernel<<<1, 1024, 0, st>>>(c->ob, n, c->d_scal);     // max row weight over ALL rows of this rank
                    GPZ_KERNEL_CHECK();
                    ++c->launches;
                }
                const bool kt_on = nchunks == 0 && !c->capturing;
                if (kt_on) GPZ_CUDA(cudaEventRecord(c->kt[2], st));
                if ((rc = ozaki_digits(phi, MP, static_cast<int>(MP), P.m, r1 - r0, c->opt_ozaki, c->ob + r0, c->d_scal, c->aug ? 1 : 0,
                                       c->oz_D8, c->oz_F8, c->oz_ea, c->sws.flag, st, &c->launches))) return rc;
                if (kt_on) {
                    GPZ_CUDA(cudaEventRecord(c->kt[3], st));
                    GPZ_CUDA(cudaEventRecord(c->kt[4], st));
                }
                if ((rc = ozaki_gram(c->oz_F8, c->oz_D8, static_cast<int>(MP), P.m, r1 - r0, c->opt_ozaki, c->opt_ozaki_gs, c->d_scal,
                                     c->aug ? 1 : 0, nchunks > 0, c->S, c->ozg_ws, c->sws.flag, st, &c->launches))) return rc;
                if (kt_on) {
                    GPZ_CUDA(cudaEventRecord(c->kt[5], st));
                    c->kt_valid[1] = c->kt_valid[2] = true;
                }
                if (timed) GPZ_EVREC(c->kev[1]);
                continue;
            }
            double* gp = c->gram_partial + static_cast<int64_t>(o) * c->gram_ns * (T * (T + 1) / 2) * TILE * TILE;
            if ((rc = gram_syrk_main(phi, MP, static_cast<int>(MP), c->ob + o * n + r0, 0, r1 - r0, c->gram_ns, gp, nchunks > 0, st,
                                     &c->launches))) return rc;
            if (timed) GPZ_EVREC(c->kev[1]);
            if ((rc = gram_syrk_finish(gp, c->gram_ns, static_cast<int>(MP), last, c->S + static_cast<int64_t>(o) * MP * MP, st,
                                       &c->launches))) return rc;
        }
        if (!c->aug)
            if ((rc = atb_general(phi, MP, static_cast<int>(MP), c->yw + r0 * 32, 32, 32, c->ones, 0, r1 - r0, ns1, c->atb_partial,
                                  nchunks > 0, last, c->Rvec, st, &c->launches))) return rc;
        ++nchunks;
        if (last) break;
    }
    {
        const int64_t nb = ceil_div(n > 0 ? n : 1, RB);
        if (n == 0) GPZ_CUDA(cudaMemsetAsync(c->part1, 0, sizeof(double) * (k + 2), st));
        reduce_parts_kernel<<<k + 2, 256, 0, st>>>(c->part1, n > 0 ? nb : 1, k + 2, c->scal1);
        GPZ_KERNEL_CHECK();
        ++c->launches;
    }
    if ((rc = allreduce(c, c->red1, c->red1_len))) return rc;
    GPZ_EVREC(c->ev[2]);
    for (int o = 0; o < k; ++o) {
        double* S = c->S + static_cast<int64_t>(o) * MP * MP;
        double* Si = c->Sinv + static_cast<int64_t>(o) * MP * MP;
        add_diag_kernel<<<static_cast<unsigned>(ceil_div(MP, 256)), 256, 0, st>>>(S, static_cast<int>(MP), P.m, P.alpha + o * MP);
        GPZ_KERNEL_CHECK();
        ++c->launches;
        if ((rc = spd_inverse(S, P.m, static_cast<int>(MP), Si, c->logdet + o, c->sws, st, &c->launches))) return rc;
        const unsigned gb = static_cast<unsigned>(ceil_div(MP, 8));
        if (c->aug)     // r = PHI'(omega beta y) is row m of the Gram (the spare column of PHI carries y)
            symv_kernel<<<gb, 256, 0, st>>>(Si, static_cast<int>(MP), P.m, S + static_cast<int64_t>(P.m) * MP, 1, nullptr, 1.0, c->w + o * MP);
        else
            symv_kernel<<<gb, 256, 0, st>>>(Si, static_cast<int>(MP), P.m, c->Rvec + o, 32, nullptr, 1.0, c->w + o * MP);
        GPZ_KERNEL_CHECK();
        symv_kernel<<<gb, 256, 0, st>>>(Si, static_cast<int>(MP), P.m, c->w + o * MP, 1, P.alpha + o * MP, -1.0, c->dwda + o * MP);
        GPZ_KERNEL_CHECK();
        c->launches += 2;
        if (c->aug) {
            set_aug_col_kernel<<<static_cast<unsigned>(ceil_div(P.m, 256)), 256, 0, st>>>(Si, static_cast<int>(MP), P.m, c->w + o * MP);
            GPZ_KERNEL_CHECK();
            ++c->launches;
        }
    }
    GPZ_EVREC(c->ev[3]);
    return GPZ_OK;
}

// pred = PHI w for the training rows of one chunk (PHI resident or rebuilt)
int chunk_phi_and_pred(gpz_ctx* c, int64_t r0, int64_t r1, double** phi_out, bool need_pred) {
    Params& P = c->P;
    const int64_t n = c->tr.n, MP = P.MP;
    int rc;
    double* phi = c->resident ? c->Phi + r0 * MP : c->Phi;
    if (!need_pred) {
        if (!c->resident)
            if ((rc = phi_build(P, c->tr, r0, r1, phi, DotSpec{0, {nullptr, nullptr}, {nullptr, nullptr}}, c->dot_scratch, c->st, &c->launches))) return rc;
        *phi_out = phi;
        return GPZ_OK;
    }
    if (!c->resident) {
        DotSpec ds{1, {c->w, nullptr}, {c->pred, nullptr}};
        if (P.k > 1) ds.n = 0;
        if ((rc = phi_build(P, c->tr, r0, r1, phi, ds, c->dot_scratch, c->st, &c->launches))) return rc;
        if (P.k > 1)
            for (int o = 0; o < P.k; ++o)
                if ((rc = rowdot(phi, MP, P.m, r1 - r0, DotSpec{1, {c->w + o * MP, nullptr}, {c->pred + o * n + r0, nullptr}}, c->st, &c->launches))) return rc;
    } else {
        for (int o = 0; o < P.k; ++o)
            if ((rc = rowdot(phi, MP, P.m, r1 - r0, DotSpec{1, {c->w + o * MP, nullptr}, {c->pred + o * n + r0, nullptr}}, c->st, &c->launches))) return rc;
    }
    *phi_out = phi;
    return GPZ_OK;
}

int eval_device_enqueue(gpz_ctx* c, const double* d_theta, double* d_out) {
    Params& P = c->P;
    cudaStream_t st = c->st;
    int rc;
    if (c->tr.n == 0) {                  // before anything collective is enqueued
        set_error("no training rows on this rank");
        return GPZ_ERR_USAGE;
    }
    if ((rc = ensure_workspace(c))) return rc;
    if ((rc = forward_and_solve(c, d_theta))) return rc;
    const int64_t n = c->tr.n, nv = c->va.n, MP = P.MP;
    const int k = P.k;
    const int ntn = static_cast<int>(MP / TILE);
    const int64_t md = static_cast<int64_t>(P.m) * P.d;
    double* dP = c->red2;
    double* dG = dP + md;
    double* qcol = dG + P.g_dim;
    double* sc2 = qcol + 2LL * k * MP;
    const bool fast_bp = !c->has_psi && !c->tr.has_nan;
    const int T = static_cast<int>(MP / TILE);
    int nchunks = 0, colp_slabs = 1, pieces = 0;      // pieces: (row chunk, pattern group) parts of the moment GEMM done so far
    bool ev4 = false;
    for (int64_t r0 = 0; r0 < n; r0 += c->chunk_rows) {
        const int64_t r1 = (r0 + c->chunk_rows < n) ? r0 + c->chunk_rows : n;
        const bool last = r1 >= n;
        const int64_t rows = r1 - r0;
        double* phi = nullptr;
        if ((rc = chunk_phi_and_pred(c, r0, r1, &phi, !c->aug))) return rc;
        for (int o = 0; o < k; ++o) {
            const bool timed = (nchunks == 0 && o == 0);
            if (timed) GPZ_EVREC(c->kev[2]);
            if (c->opt_ozaki > 0) {
                // the plain digits of this chunk are still there when PHI is resident and the Gram went through them
                if (o == 0 && !(c->resident && c->opt_ozaki_gram > 0))
                    if ((rc = ozaki_digits(phi, MP, static_cast<int>(MP), P.m, rows, c->opt_ozaki, nullptr, c->d_scal, 0, c->oz_D8, nullptr,
                                           c->oz_ea, c->sws.flag, st, &c->launches))) return rc;
                if ((rc = ozaki_tgemm(phi, MP, c->oz_D8, c->oz_ea, c->Sinv + static_cast<int64_t>(o) * MP * MP, static_cast<int>(MP), P.m,
                                      rows, c->opt_ozaki, c->ob + o * n + r0, c->H, o > 0, c->nupart + r0, n,
                                      c->aug ? c->w + o * MP : nullptr, c->aug ? c->pred + r0 : nullptr, c->oz_ws, st,
                                      timed && !c->capturing ? c->kev[4] : nullptr, timed && !c->capturing ? c->kev[5] : nullptr,
                                      &c->launches))) return rc;
            } else {
                if ((rc = tgemm(phi, MP, c->Sinv + static_cast<int64_t>(o) * MP * MP, static_cast<int>(MP), P.m, rows, c->ob + o * n + r0,
                                c->H, o > 0, c->nupart + r0, n, c->aug ? c->pred + r0 : nullptr, st, &c->launches))) return rc;
            }
            if (timed) GPZ_EVREC(c->kev[3]);
            rows2_kernel<<<static_cast<unsigned>(ceil_div(rows, RB)), RB, 0, st>>>(P, o, c->tr.Y, c->tr.omega, n, r0, r1, c->pred,
                                                                                 c->nupart, ntn, c->lnbi, c->beta,
                                                                                 c->ob, c->nu, c->cw, c->dbeta, c->part2, 2 * k + 2);
            GPZ_KERNEL_CHECK();
            ++c->launches;
        }
        if (!ev4) {
            GPZ_EVREC(c->ev[4]);
            ev4 = true;
        }
        const bool fused = fast_bp && c->opt_fused_bp && k == 1 && c->QP <= 128;
        const size_t ng = c->tr.g_pat.empty() ? 1 : c->tr.g_pat.size();
        if (!fused) {
            const int64_t rps = ceil_div(rows, c->dphi_slabs);
            dim3 grid(static_cast<unsigned>(T), static_cast<unsigned>(c->dphi_slabs));
            dphi_kernel<<<grid, 128, 0, st>>>(P, phi, c->H, MP, n, r0, r1, rps, c->cw, c->dbeta, c->w, c->colp, nchunks > 0);
            GPZ_KERNEL_CHECK();
            ++c->launches;
            colp_slabs = c->dphi_slabs;
        } else {
            colp_slabs = c->fused_ns;
        }
        if (fast_bp) {
            // one moment GEMM per missing-input pattern group (a single group when the data have no NaN)
            for (size_t g = 0; g < ng; ++g) {
                int64_t s0 = r0, s1 = r1;
                if (ng > 1) {                  // the part of this pattern group inside the row chunk
                    s0 = c->tr.g_r0[g] > r0 ? c->tr.g_r0[g] : r0;
                    s1 = c->tr.g_r1[g] < r1 ? c->tr.g_r1[g] : r1;
                    if (s1 <= s0) continue;
                }
                const int acc = ng > 1 ? 0 : (nchunks > 0);
                const int red = ng > 1 ? 1 : (last ? 1 : 0);
                const bool kt_on = nchunks == 0 && g == 0 && !c->capturing;
                if (kt_on) GPZ_CUDA(cudaEventRecord(c->kt[6], st));
                if (fused) {
                    if ((rc = atb_dphi(phi + (s0 - r0) * MP, c->H + (s0 - r0) * MP, MP, static_cast<int>(MP), c->tr.F + s0 * c->QP, c->QP, P.q,
                                       c->cw + s0, c->dbeta + s0, c->w, P.v, 0, s1 - s0, c->fused_ns, c->atb_partial, c->colp, acc,
                                       pieces > 0, red, c->Rm, st, &c->launches))) return rc;
                } else {
                    if ((rc = atb_general(c->H + (s0 - r0) * MP, MP, static_cast<int>(MP), c->tr.F + s0 * c->QP, c->QP, c->QP, c->ones, 0,
                                          s1 - s0, c->atb_ns, c->atb_partial, acc, red, c->Rm, st, &c->launches))) return rc;
                }
                if (kt_on) {
                    GPZ_CUDA(cudaEventRecord(c->kt[7], st));
                    c->kt_valid[3] = true;
                }
                if (ng > 1)
                    if ((rc = moments_to_grad(P, c->tr.g_pat[g], c->Rm, c->QP, dP, c->scratch, pieces > 0, st, &c->launches))) return rc;
                ++pieces;
            }
        } else if (!mode_is_cov(P.mode)) {
            if ((rc = backproj_diag_generic(P, c->tr, r0, r1, c->H, MP, c->bp_partial, c->nslab, nchunks > 0, st, &c->launches))) return rc;
        } else if (c->gc_fast) {
            if ((rc = gc_backproj(P, c->tr, r0, r1, c->H, MP, c->gc_ws, c->gc_ns, c->sm_count, nchunks > 0, last, st, &c->launches))) return rc;
        } else {
            if ((rc = backproj_cov_psi(P, c->tr, r0, r1, c->H, MP, c->bp_partial, c->nslab, nchunks > 0, st, &c->launches))) return rc;
        }
        ++nchunks;
    }
    if (!ev4) GPZ_EVREC(c->ev[4]);
    if (fast_bp) {
        if (c->tr.g_pat.size() <= 1)
            rc = moments_to_grad(P, c->tr.g_pat.empty() ? 0 : c->tr.g_pat[0], c->Rm, c->QP, dP, c->scratch, 0, st, &c->launches);
        if (!rc) rc = mode_reduce(P, c->scratch, dG, st, &c->launches);
    }
    else if (!mode_is_cov(P.mode)) rc = backproj_diag_generic_finish(P, c->bp_partial, c->nslab, dP, dG, c->scratch, st, &c->launches);
    else if (c->gc_fast) rc = gc_backproj_finish(P, c->tr, c->gc_ws, c->gc_ns, c->sm_count, dP, dG, st, &c->launches);
    else rc = backproj_cov_psi_finish(P, c->tr, c->bp_partial, c->nslab, dP, dG, c->scratch, st, &c->launches);
    if (rc) return rc;
    colsum_reduce_kernel<<<static_cast<unsigned>(ceil_div(2LL * k * MP, 256)), 256, 0, st>>>(c->colp, colp_slabs, 2 * k, static_cast<int>(MP), qcol);
    GPZ_KERNEL_CHECK();
    reduce_parts_kernel<<<2 * k + 2, 256, 0, st>>>(c->part2, ceil_div(n, RB), 2 * k + 2, sc2);
    GPZ_KERNEL_CHECK();
    c->launches += 2;
    // validation statistics (GPz.m:239-259): second PHI build on the validation rows
    if (nv > 0) {
        for (int64_t r0 = 0; r0 < nv; r0 += c->chunk_rows) {
            const int64_t r1 = (r0 + c->chunk_rows < nv) ? r0 + c->chunk_rows : nv;
            const bool need_store = (k > 1) || (mode_is_cov(P.mode) && c->has_psi);
            DotSpec ds{2, {P.v, c->w}, {c->dotv_va, c->pred_va}};
            if (k > 1) ds.n = 0;
            if ((rc = phi_build(P, c->va, r0, r1, need_store ? c->Phi : nullptr, ds, c->dot_scratch, st, &c->launches))) return rc;
            if (k > 1)
                for (int o = 0; o < k; ++o)
                    if ((rc = rowdot(c->Phi, MP, P.m, r1 - r0, DotSpec{2, {P.v + o * MP, c->w + o * MP}, {c->dotv_va + o * nv + r0, c->pred_va + o * nv + r0}}, st, &c->launches))) return rc;
            for (int o = 0; o < k; ++o) {
                rowsv_kernel<<<static_cast<unsigned>(ceil_div(r1 - r0, RB)), RB, 0, st>>>(P, o, c->va.Y, c->va.omega, nv, r0, r1, c->dotv_va,
                                                                                        c->pred_va, c->partv);
                GPZ_KERNEL_CHECK();
                ++c->launches;
            }
        }
        reduce_parts_kernel<<<3, 256, 0, st>>>(c->partv, ceil_div(nv, RB), 3, sc2 + 2 * k + 2);
        GPZ_KERNEL_CHECK();
        ++c->launches;
    } else {
        GPZ_CUDA(cudaMemsetAsync(sc2 + 2 * k + 2, 0, sizeof(double) * 3, st));
    }
    if ((rc = allreduce(c, c->red2, c->red2_len))) return rc;
    FinishArgs fa{P, d_theta, c->w, c->dwda, c->Sinv, c->logdet, c->scal1, c->red2, c->sws.flag,
                  (nv > 0 || c->world > 1) ? 1 : 0, d_out};
    finish_kernel<<<1, 256, 0, st>>>(fa);
    grad_scale_kernel<<<static_cast<unsigned>(ceil_div(md + P.g_dim, 256)), 256, 0, st>>>(c->red2, md + P.g_dim, c->scal1, k, c->sws.flag,
                                                                                         d_out + 1);
    ++c->launches;
    GPZ_KERNEL_CHECK();
    ++c->launches;
    GPZ_EVREC(c->ev[5]);
    return GPZ_OK;
}

// Launch-bound regime (the reference's demo sizes: ~35 kernels of a few microseconds each): the evaluation has no
// host-side decision inside, so it is captured once into a CUDA graph and replayed with one launch.  The graph is tied
// to the (theta, out) pointer pair it was captured for; capture happens the second time the same pair is seen (the first
// run does the lazy one-time setup un-captured and provides the phase timings gpz_last_timing reports).
bool graph_wanted(const gpz_ctx* c) {
    if (c->opt_graph == 0 || c->comm != nullptr) return false;
    if (c->opt_graph > 0) return true;
    const double work = static_cast<double>(c->tr.n + c->va.n) * c->P.MP * c->P.MP;
    return work < 2e10 && c->tr.n <= c->chunk_rows && c->va.n <= c->chunk_rows;   // a few hundred microseconds of kernels, one row chunk
}

void graph_drop(gpz_ctx* c) {
    if (c->g_exec) cudaGraphExecDestroy(c->g_exec);
    c->g_exec = nullptr;
    c->g_theta = nullptr;
    c->g_out = nullptr;
}

int eval_device(gpz_ctx* c, const double* d_theta, double* d_out) {
    int rc;
    if ((rc = ensure_workspace(c))) return rc;
    if (!graph_wanted(c)) return eval_device_enqueue(c, d_theta, d_out);
    if (c->g_exec && c->g_theta == d_theta && c->g_out == d_out) {
        GPZ_CUDA(cudaEventRecord(c->gev[0], c->st));
        GPZ_CUDA(cudaGraphLaunch(c->g_exec, c->st));
        GPZ_CUDA(cudaEventRecord(c->gev[1], c->st));
        c->launches += c->g_launches;
        ++c->graph_replays;
        c->last_was_replay = true;
        return GPZ_OK;
    }
    c->last_was_replay = false;
    if (c->seen_theta != d_theta || c->seen_out != d_out) {        // first sight of this pointer pair: plain run
        c->seen_theta = d_theta;
        c->seen_out = d_out;
        return eval_device_enqueue(c, d_theta, d_out);
    }
    graph_drop(c);
    cudaGraph_t graph = nullptr;
    const int64_t l0 = c->launches;
    GPZ_CUDA(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
    c->capturing = true;
    rc = eval_device_enqueue(c, d_theta, d_out);
    c->capturing = false;
    const cudaError_t e = cudaStreamEndCapture(c->st, &graph);
    c->g_launches = c->launches - l0;
    c->launches = l0;
    if (rc || e != cudaSuccess || !graph) {                        // not capturable here: fall back to plain launches for good
        if (getenv("GPZ_B200_DEBUG"))
            fprintf(stderr, "gpz_b200: graph capture failed (rc %d, %s; %s)\n", rc, cudaGetErrorString(e), gpz_last_error());
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        c->opt_graph = 0;
        return rc ? rc : eval_device_enqueue(c, d_theta, d_out);
    }
    const cudaError_t ei = cudaGraphInstantiate(&c->g_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) {
        cudaGetLastError();
        c->g_exec = nullptr;
        c->opt_graph = 0;
        return eval_device_enqueue(c, d_theta, d_out);
    }
    c->g_theta = d_theta;
    c->g_out = d_out;
    GPZ_CUDA(cudaEventRecord(c->gev[0], c->st));
    GPZ_CUDA(cudaGraphLaunch(c->g_exec, c->st));
    GPZ_CUDA(cudaEventRecord(c->gev[1], c->st));
    c->launches += c->g_launches;
    ++c->graph_replays;
    c->last_was_replay = true;
    return GPZ_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// exported C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* gpz_last_error(void) { return gpz::g_err; }
int gpz_version(void) { return 100; }

int64_t gpz_theta_len(const gpz_model* model) {
    Params P{};
    if (fill_params(P, model)) return -1;
    return P.p;
}

int64_t gpz_g_dim(const gpz_model* model) {
    Params P{};
    if (fill_params(P, model)) return -1;
    return P.g_dim;
}

int gpz_create(gpz_ctx** out, const gpz_model* model, int64_t n_all, const double* X, const double* Y, const double* Psi,
               const double* omega, const uint8_t* training, const uint8_t* validation, int device) {
    if (!out || !X || n_all < 0) {
        set_error("gpz_create: bad arguments");
        return GPZ_ERR_USAGE;
    }
    *out = nullptr;
    Params P{};
    int rc;
    if ((rc = fill_params(P, model))) return rc;
    if ((rc = check_device(device))) return rc;
    gpz_ctx* c = new gpz_ctx();
    c->P = P;
    c->device = device;
    c->has_psi = Psi != nullptr;
    auto fail = [&](int code) {
        gpz_destroy(c);
        return code;
    };
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        return fail(GPZ_ERR_CUDA);
    }
    std::vector<int64_t> itr, iva;
    for (int64_t i = 0; i < n_all; ++i) {
        if (!training || training[i]) itr.push_back(i);
        if (validation && validation[i]) iva.push_back(i);
    }
    c->h_shift.assign(static_cast<size_t>(P.d), 0.0);
    for (int a = 0; a < P.d; ++a) {          // column means of the finite training entries
        double s = 0.0;
        int64_t cnt = 0;
        const double* col = X + static_cast<int64_t>(a) * n_all;
        for (int64_t i : itr) {
            const double v = col[i];
            if (v == v) { s += v; ++cnt; }
        }
        c->h_shift[a] = cnt > 0 ? s / static_cast<double>(cnt) : 0.0;
    }
    // covariance modes with missing inputs: group rows by NaN pattern (getPHI.m:43-54), patterns numbered in order
    // of first appearance over training then validation rows
    bool grouped = false;
    std::vector<std::vector<unsigned char>> pats;
    std::vector<int> pat_tr, pat_va;
    if (mode_is_cov(P.mode)) {
        std::map<std::string, int> seen;
        auto scan = [&](const std::vector<int64_t>& idx, std::vector<int>& out_pat) {
            out_pat.resize(idx.size());
            std::string key(static_cast<size_t>(P.d), '1');
            for (size_t r = 0; r < idx.size(); ++r) {
                bool any = false;
                for (int a = 0; a < P.d; ++a) {
                    const double v = X[static_cast<int64_t>(a) * n_all + idx[r]];
                    key[a] = (v == v) ? '1' : '0';
                    any = any || (v != v);
                }
                grouped = grouped || any;
                auto it = seen.find(key);
                if (it == seen.end()) {
                    it = seen.emplace(key, static_cast<int>(pats.size())).first;
                    std::vector<unsigned char> ob(static_cast<size_t>(P.d));
                    for (int a = 0; a < P.d; ++a) ob[a] = key[a] == '1';
                    pats.push_back(ob);
                }
                out_pat[r] = it->second;
            }
        };
        scan(itr, pat_tr);
        scan(iva, pat_va);
        if (grouped && pats.size() > 4096) {
            set_error("too many distinct missing-input patterns (%zu)", pats.size());
            return fail(GPZ_ERR_USAGE);
        }
    }
    if (grouped) c->P.npat = static_cast<int>(pats.size());
    if ((rc = alloc_params(c->P, c->allocs, c->has_psi))) return fail(rc);
    if (h2d(c->P.xshift, c->h_shift.data(), sizeof(double) * P.d) != cudaSuccess) {
        set_error("gpz_create: upload of the shift failed");
        return fail(GPZ_ERR_CUDA);
    }
    auto sort_groups = [&](std::vector<int64_t>& idx, const std::vector<int>& pat, RowData& R) {
        if (!grouped) return;
        std::vector<int64_t> order(idx.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return pat[a] < pat[b]; });
        std::vector<int64_t> sorted(idx.size());
        for (size_t r = 0; r < order.size(); ++r) sorted[r] = idx[order[r]];
        R.perm = order;
        size_t r = 0;
        while (r < order.size()) {
            size_t e = r;
            while (e < order.size() && pat[order[e]] == pat[order[r]]) ++e;
            R.g_r0.push_back(static_cast<int64_t>(r));
            R.g_r1.push_back(static_cast<int64_t>(e));
            R.g_pat.push_back(pat[order[r]]);
            r = e;
        }
        idx.swap(sorted);
    };
    sort_groups(itr, pat_tr, c->tr);
    sort_groups(iva, pat_va, c->va);
    if (grouped) {
        std::vector<unsigned char> flat;
        for (auto& ob : pats) flat.insert(flat.end(), ob.begin(), ob.end());
        if (h2d(c->P.obs, flat.data(), flat.size()) != cudaSuccess) {
            set_error("gpz_create: upload of the patterns failed");
            return fail(GPZ_ERR_CUDA);
        }
    }
    if ((rc = upload_rows(c, c->tr, itr, n_all, X, Y, Psi, omega, grouped))) return fail(rc);
    if ((rc = upload_rows(c, c->va, iva, n_all, X, Y, Psi, omega, grouped))) return fail(rc);
    *out = c;
    return GPZ_OK;
}

void gpz_destroy(gpz_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (void* p : c->allocs) cudaFree(p);
    solve_ws_free(c->sws);
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->kev)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->gev)
        if (e) cudaEventDestroy(e);
    if (c->g_exec) cudaGraphExecDestroy(c->g_exec);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->h_theta) cudaFreeHost(c->h_theta);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

int gpz_comm_unique_id(char id[128]) {
    int rc;
    if ((rc = nccl_load())) return rc;
    ncclUniqueId u;
    GPZ_NCCL(g_nccl.GetUniqueId(&u));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id, &u, 128);
    return GPZ_OK;
}

int gpz_comm_init(gpz_ctx* c, int rank, int world, const char id[128]) {
    if (!c || world < 1 || rank < 0 || rank >= world) {
        set_error("gpz_comm_init: bad arguments");
        return GPZ_ERR_USAGE;
    }
    if (world > 1 && c->tr.n == 0) {     // such a rank would leave the others blocked inside an allreduce
        set_error("gpz_comm_init: rank %d holds no training rows; every rank of a sharded run needs at least one", rank);
        return GPZ_ERR_USAGE;
    }
    c->rank = rank;
    c->world = world;
    if (world == 1) return GPZ_OK;
    int rc;
    if ((rc = nccl_load())) return rc;
    GPZ_CUDA(cudaSetDevice(c->device));
    ncclUniqueId u;
    memcpy(&u, id, 128);
    GPZ_NCCL(g_nccl.CommInitRank(&c->comm, world, u, rank));
    return GPZ_OK;
}

int gpz_eval_dev(gpz_ctx* c, const double* d_theta, double* d_out) {
    if (!c || !d_theta || !d_out) {
        set_error("gpz_eval_dev: NULL argument");
        return GPZ_ERR_USAGE;
    }
    GPZ_CUDA(cudaSetDevice(c->device));
    return eval_device(c, d_theta, d_out);
}

int gpz_eval(gpz_ctx* c, const double* theta, double* nlogML, double* grad, double stats[4]) {
    if (!c || !theta) {
        set_error("gpz_eval: NULL argument");
        return GPZ_ERR_USAGE;
    }
    GPZ_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure_workspace(c))) return rc;
    const int64_t p = c->P.p;
    memcpy(c->h_theta, theta, sizeof(double) * p);
    GPZ_CUDA(cudaMemcpyAsync(c->d_theta, c->h_theta, sizeof(double) * p, cudaMemcpyHostToDevice, c->st));
    if ((rc = eval_device(c, c->d_theta, c->d_out))) return rc;
    GPZ_CUDA(cudaMemcpyAsync(c->h_out, c->d_out, sizeof(double) * (p + 5), cudaMemcpyDeviceToHost, c->st));
    GPZ_CUDA(cudaStreamSynchronize(c->st));
    if (nlogML) *nlogML = c->h_out[0];
    if (grad) memcpy(grad, c->h_out + 1, sizeof(double) * p);
    if (stats) memcpy(stats, c->h_out + 1 + p, sizeof(double) * 4);
    return GPZ_OK;
}

static int train_objective(void* user, const double* d_x, double* d_out, void* /*stream*/) {
    return eval_device(static_cast<gpz_ctx*>(user), d_x, d_out);
}

int gpz_train(gpz_ctx* c, const gpz_train_options* opt, double* theta, double* best_theta, double* best_valid,
              gpz_train_callback cb, void* user, gpz_train_result* res) {
    if (!c || !theta || !best_theta || !best_valid) {
        set_error("gpz_train: NULL argument");
        return GPZ_ERR_USAGE;
    }
    GPZ_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure_workspace(c))) return rc;
    gpz_train_options o;
    if (opt)
        o = *opt;
    else
        gpz_train_default_options(&o);
    if (o.training_only < 0) o.training_only = c->va.n == 0 ? 1 : 0;
    return gpz::lbfgs_train(c->P.p, train_objective, c, &o, theta, best_theta, best_valid, cb, user, res, c->st, &c->launches);
}

int gpz_fit(gpz_ctx* c, const double* theta, double* nlogML_k, double* w, double* iSigma_w) {
    if (!c || !theta) {
        set_error("gpz_fit: NULL argument");
        return GPZ_ERR_USAGE;
    }
    GPZ_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure_workspace(c))) return rc;
    Params& P = c->P;
    const int64_t p = P.p, MP = P.MP, n = c->tr.n;
    const int k = P.k;
    cudaStream_t st = c->st;
    memcpy(c->h_theta, theta, sizeof(double) * p);
    GPZ_CUDA(cudaMemcpyAsync(c->d_theta, c->h_theta, sizeof(double) * p, cudaMemcpyHostToDevice, st));
    if ((rc = forward_and_solve(c, c->d_theta))) return rc;
    const int64_t m2 = static_cast<int64_t>(P.m) * P.m;
    nan_fill_kernel<<<static_cast<unsigned>(ceil_div(k * MP * MP, 256)), 256, 0, st>>>(c->Sinv, k * MP * MP, c->sws.flag);
    nan_fill_kernel<<<static_cast<unsigned>(ceil_div(k * MP, 256)), 256, 0, st>>>(c->w, k * MP, c->sws.flag);
    GPZ_KERNEL_CHECK();
    if (nlogML_k) {
        double* sc2 = c->red2 + static_cast<int64_t>(P.m) * P.d + P.g_dim + 2LL * k * MP;
        for (int64_t r0 = 0; r0 < n; r0 += c->chunk_rows) {
            const int64_t r1 = (r0 + c->chunk_rows < n) ? r0 + c->chunk_rows : n;
            double* phi = nullptr;
            if ((rc = chunk_phi_and_pred(c, r0, r1, &phi, true))) return rc;
            for (int o = 0; o < k; ++o) {
                rows2_kernel<<<static_cast<unsigned>(ceil_div(r1 - r0, RB)), RB, 0, st>>>(P, o, c->tr.Y, c->tr.omega, n, r0, r1, c->pred, nullptr,
                                                                                        0, c->lnbi, c->beta, c->ob, c->nu, c->cw, c->dbeta,
                                                                                        c->part2, 2 * k + 2);
                GPZ_KERNEL_CHECK();
                ++c->launches;
            }
        }
        reduce_parts_kernel<<<2 * k + 2, 256, 0, st>>>(c->part2, ceil_div(n > 0 ? n : 1, RB), 2 * k + 2, sc2);
        GPZ_KERNEL_CHECK();
        if ((rc = allreduce(c, sc2, 2 * k + 2))) return rc;
        fit_nl_kernel<<<1, 256, 0, st>>>(P, c->d_theta, c->w, c->logdet, c->scal1, sc2, c->sws.flag, c->d_out);
        GPZ_KERNEL_CHECK();
        c->launches += 2;
        GPZ_CUDA(cudaMemcpyAsync(c->h_out, c->d_out, sizeof(double) * k, cudaMemcpyDeviceToHost, st));
    }
    if (w)
        GPZ_CUDA(cudaMemcpy2DAsync(w, sizeof(double) * P.m, c->w, sizeof(double) * MP, sizeof(double) * P.m, k, cudaMemcpyDeviceToHost, st));
    if (iSigma_w)
        for (int o = 0; o < k; ++o)
            GPZ_CUDA(cudaMemcpy2DAsync(iSigma_w + o * m2, sizeof(double) * P.m, c->Sinv + static_cast<int64_t>(o) * MP * MP, sizeof(double) * MP,
                                       sizeof(double) * P.m, P.m, cudaMemcpyDeviceToHost, st));
    GPZ_CUDA(cudaStreamSynchronize(st));
    if (nlogML_k) memcpy(nlogML_k, c->h_out, sizeof(double) * k);
    return GPZ_OK;
}

int64_t gpz_rows(const gpz_ctx* c, int which) { return c ? (which ? c->va.n : c->tr.n) : -1; }

int gpz_phi(gpz_ctx* c, const double* theta, int which, double* PHI, double* lnBeta_i, double* N) {
    if (!c || !theta) {
        set_error("gpz_phi: NULL argument");
        return GPZ_ERR_USAGE;
    }
    GPZ_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure_workspace(c))) return rc;
    Params& P = c->P;
    RowData& R = which ? c->va : c->tr;
    const int64_t n = R.n, MP = P.MP;
    cudaStream_t st = c->st;
    if (n == 0) return GPZ_OK;
    GPZ_CUDA(cudaMemcpyAsync(c->d_theta, theta, sizeof(double) * P.p, cudaMemcpyHostToDevice, st));
    if ((rc = prep_params(c->d_theta, P, c->has_psi, st, &c->launches))) return rc;
    double *dotv = nullptr, *colmaj = nullptr;
    struct Scratch {                     // freed on every exit path
        double*& a;
        double*& b;
        ~Scratch() {
            if (a) cudaFree(a);
            if (b) cudaFree(b);
        }
    } scratch_guard{dotv, colmaj};
    GPZ_CUDA(cudaMalloc(&dotv, sizeof(double) * n * P.k));
    if (PHI || N) GPZ_CUDA(cudaMalloc(&colmaj, sizeof(double) * c->chunk_rows * P.m));
    std::vector<double> hbuf;
    for (int64_t r0 = 0; r0 < n; r0 += c->chunk_rows) {
        const int64_t r1 = (r0 + c->chunk_rows < n) ? r0 + c->chunk_rows : n;
        if ((rc = phi_build(P, R, r0, r1, c->Phi, DotSpec{0, {nullptr, nullptr}, {nullptr, nullptr}}, c->dot_scratch, st, &c->launches))) return rc;
        if (P.het)
            for (int o = 0; o < P.k; ++o)
                if ((rc = rowdot(c->Phi, MP, P.m, r1 - r0, DotSpec{1, {P.v + o * MP, nullptr}, {dotv + o * n + r0, nullptr}}, st, &c->launches))) return rc;
        if (PHI) {
            if ((rc = transpose_out(c->Phi, MP, r1 - r0, P.m, colmaj, st))) return rc;
            // column-major (r1-r0) x m chunk -> rows r0..r1 of the n x m host matrix
            GPZ_CUDA(cudaMemcpy2DAsync(PHI + r0, sizeof(double) * n, colmaj, sizeof(double) * (r1 - r0), sizeof(double) * (r1 - r0), P.m,
                                       cudaMemcpyDeviceToHost, st));
            GPZ_CUDA(cudaStreamSynchronize(st));
        }
        if (N) {
            if ((rc = phi_to_density(P, R, r0, r1, c->Phi, c->H, st, &c->launches))) return rc;
            if ((rc = transpose_out(c->H, MP, r1 - r0, P.m, colmaj, st))) return rc;
            GPZ_CUDA(cudaMemcpy2DAsync(N + r0, sizeof(double) * n, colmaj, sizeof(double) * (r1 - r0), sizeof(double) * (r1 - r0), P.m,
                                       cudaMemcpyDeviceToHost, st));
            GPZ_CUDA(cudaStreamSynchronize(st));
        }
    }
    for (double* M : {PHI, N}) {
        if (M && !R.perm.empty()) {          // rows are stored sorted by missing-input pattern: restore selection order
            std::vector<double> tmp(static_cast<size_t>(n));
            for (int j = 0; j < P.m; ++j) {
                double* col = M + static_cast<int64_t>(j) * n;
                for (int64_t i = 0; i < n; ++i) tmp[static_cast<size_t>(R.perm[i])] = col[i];
                memcpy(col, tmp.data(), sizeof(double) * n);
            }
        }
    }
    if (lnBeta_i) {
        hbuf.resize(static_cast<size_t>(n * P.k));
        std::vector<double> hb(32);
        if (P.het) GPZ_CUDA(cudaMemcpyAsync(hbuf.data(), dotv, sizeof(double) * n * P.k, cudaMemcpyDeviceToHost, st));
        GPZ_CUDA(cudaStreamSynchronize(st));
        for (int o = 0; o < P.k; ++o) {
            const double b = theta[P.oB + o];
            for (int64_t i = 0; i < n; ++i) {
                const int64_t dst = R.perm.empty() ? i : R.perm[i];
                lnBeta_i[o * n + dst] = b + (P.het ? hbuf[o * n + i] : 0.0);
            }
        }
    }
    GPZ_CUDA(cudaStreamSynchronize(st));
    return GPZ_OK;
}

// priors over the bases by EM on the normalised densities (GPz/getPrior.m:1-22), training rows
__global__ void prior_inv_kernel(const double* __restrict__ s, int64_t n, double* __restrict__ yw) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    yw[i * 32] = 1.0 / s[i];
    for (int c = 1; c < 32; ++c) yw[i * 32 + c] = 0.0;
}
// prior_j <- prior_j * t_j / n ; out[0] = |old-new|^2, out[1] = |old+new|^2      (getPrior.m:12-17)
__global__ void __launch_bounds__(256)
prior_update_kernel(double* __restrict__ prior, const double* __restrict__ R32, const double* __restrict__ nrows, int m,
                    double* __restrict__ out) {
    __shared__ double sh[8];
    double a = 0.0, b = 0.0;
    for (int j = threadIdx.x; j < m; j += 256) {
        const double o = prior[j];
        const double nw = o * R32[static_cast<int64_t>(j) * 32] / nrows[0];
        prior[j] = nw;
        a += (o - nw) * (o - nw);
        b += (o + nw) * (o + nw);
    }
    a = gpz::block_sum<256>(a, sh);
    b = gpz::block_sum<256>(b, sh);
    if (threadIdx.x == 0) {
        out[0] = a;
        out[1] = b;
    }
}

int gpz_get_prior(gpz_ctx* c, const double* theta, double* prior) {
    if (!c || !theta || !prior) {
        set_error("gpz_get_prior: NULL argument");
        return GPZ_ERR_USAGE;
    }
    GPZ_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure_workspace(c))) return rc;
    Params& P = c->P;
    const int64_t n = c->tr.n, MP = P.MP;
    cudaStream_t st = c->st;
    GPZ_CUDA(cudaMemcpyAsync(c->d_theta, theta, sizeof(double) * P.p, cudaMemcpyHostToDevice, st));
    if ((rc = prep_params(c->d_theta, P, c->has_psi, st, &c->launches))) return rc;
    // N (getPHI's 4th output) of the rows [r0, r1) into the H buffer.  PHI resident: all rows once, before the EM loop; otherwise
    // rebuilt per row chunk in every iteration -- which is what the reference does for ALL rows (getPrior.m:10 calls getPHI
    // inside the loop although N does not depend on the prior)
    auto build_N = [&](int64_t r0, int64_t r1) -> int {
        double* phi = c->resident ? c->Phi + r0 * MP : c->Phi;
        double* Nn = c->resident ? c->H + r0 * MP : c->H;
        int e = phi_build(P, c->tr, r0, r1, phi, DotSpec{0, {nullptr, nullptr}, {nullptr, nullptr}}, c->dot_scratch, st, &c->launches);
        if (!e) e = phi_to_density(P, c->tr, r0, r1, phi, Nn, st, &c->launches);
        return e;
    };
    if (c->resident && n > 0)
        if ((rc = build_N(0, n))) return rc;
    double* d_prior = c->dwda;             // [MP] scratch vectors of the eval path are free here
    double* d_s = c->pred;
    double* d_R = c->Rvec;                 // [MP][32]
    double* d_cnt = c->scal1;              // [0] = global row count
    std::vector<double> h(static_cast<size_t>(MP), 0.0);
    for (int j = 0; j < P.m; ++j) h[j] = 1.0 / P.m;                                               // getPrior.m:5
    GPZ_CUDA(cudaMemcpyAsync(d_prior, h.data(), sizeof(double) * MP, cudaMemcpyHostToDevice, st));
    const double nd = static_cast<double>(n);
    GPZ_CUDA(cudaMemcpyAsync(d_cnt, &nd, sizeof(double), cudaMemcpyHostToDevice, st));
    if ((rc = allreduce(c, d_cnt, 1))) return rc;
    const int T = static_cast<int>(MP / TILE);
    const int ns1 = c->sm_count / T > 0 ? c->sm_count / T : 1;
    for (int iter = 0; iter < 100; ++iter) {                                                      // getPrior.m:7
        if (n == 0) GPZ_CUDA(cudaMemsetAsync(d_R, 0, sizeof(double) * MP * 32, st));
        int nchunks = 0;
        for (int64_t r0 = 0; r0 < n; r0 += c->chunk_rows, ++nchunks) {
            const int64_t r1 = (r0 + c->chunk_rows < n) ? r0 + c->chunk_rows : n;
            const int64_t rows = r1 - r0;
            if (!c->resident)
                if ((rc = build_N(r0, r1))) return rc;
            const double* Nn = c->resident ? c->H + r0 * MP : c->H;
            if ((rc = rowdot(Nn, MP, P.m, rows, DotSpec{1, {d_prior, nullptr}, {d_s + r0, nullptr}}, st, &c->launches))) return rc;
            prior_inv_kernel<<<static_cast<unsigned>(ceil_div(rows, 256)), 256, 0, st>>>(d_s + r0, rows, c->yw + r0 * 32);
            GPZ_KERNEL_CHECK();
            ++c->launches;
            if ((rc = atb_general(Nn, MP, static_cast<int>(MP), c->yw + r0 * 32, 32, 32, c->ones, 0, rows, ns1, c->atb_partial, nchunks > 0,
                                  r1 >= n, d_R, st, &c->launches))) return rc;
        }
        if ((rc = allreduce(c, d_R, MP * 32))) return rc;
        prior_update_kernel<<<1, 256, 0, st>>>(d_prior, d_R, d_cnt, P.m, c->d_out);
        GPZ_KERNEL_CHECK();
        ++c->launches;
        double nr[2];
        GPZ_CUDA(cudaMemcpyAsync(nr, c->d_out, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
        GPZ_CUDA(cudaStreamSynchronize(st));
        if (!(sqrt(nr[0]) / sqrt(nr[1]) >= 1e-10)) break;                                         // getPrior.m:17 (also stops on NaN)
    }
    GPZ_CUDA(cudaMemcpyAsync(prior, d_prior, sizeof(double) * P.m, cudaMemcpyDeviceToHost, st));
    GPZ_CUDA(cudaStreamSynchronize(st));
    return GPZ_OK;
}

// one group of rows sharing a missing-input pattern (predictDiag.m:127-295, predictCov.m:134-336); host buffers in/out;
// Psig: diagonal modes [n x d] column-major, covariance modes [n][d*d]
static int predict_missing_group(const gpz_model* model, const double* theta, const double* w, const double* iSigma_w,
                                 int64_t n, const double* Xg, const double* Psig, const double* priors,
                                 const std::vector<unsigned char>& ob, double* mu, double* nu, double* beta_i, double* gamma,
                                 double* PHI, int device) {
    Params P{};
    int rc;
    if ((rc = fill_params(P, model))) return rc;
    if ((rc = check_device(device))) return rc;
    std::vector<void*> allocs;
    cudaStream_t st = nullptr;
    int64_t launches = 0;
    auto cleanup = [&](int code) {
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
        for (void* p : allocs) cudaFree(p);
        return code;
    };
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        return GPZ_ERR_CUDA;
    }
    const int64_t MP = P.MP;
    const int k = P.k, d = P.d;
    const bool cov = mode_is_cov(P.mode);
    const int64_t psi_w = cov ? static_cast<int64_t>(P.d) * P.d : P.d;
    if ((rc = alloc_params(P, allocs, cov ? 1 : 0))) return cleanup(rc);
    double *d_theta, *d_w, *d_Sinv, *d_prior, *d_X, *d_Psi = nullptr, *d_out, *d_Phi, *d_col = nullptr;
    unsigned char* d_ob;
    if ((rc = dev_alloc(allocs, &d_theta, P.p)) || (rc = dev_alloc(allocs, &d_w, k * MP)) || (rc = dev_alloc(allocs, &d_Sinv, k * MP * MP)) ||
        (rc = dev_alloc(allocs, &d_prior, MP)) || (rc = dev_alloc(allocs, &d_X, n * d)) || (rc = dev_alloc(allocs, &d_out, 4 * k * n)) ||
        (rc = dev_alloc(allocs, &d_Phi, n * MP)) || (rc = dev_alloc(allocs, &d_ob, d)))
        return cleanup(rc);
    if (Psig && (rc = dev_alloc(allocs, &d_Psi, n * psi_w))) return cleanup(rc);
    if (PHI && (rc = dev_alloc(allocs, &d_col, n * P.m))) return cleanup(rc);
    std::vector<double> hx(Xg, Xg + n * d);
    for (double& v : hx)
        if (v != v) v = 0.0;                      // missing dims are never read (ob mask); keep the buffer NaN-free
    bool ok = h2d(d_theta, theta, sizeof(double) * P.p) == cudaSuccess;
    ok = ok && cudaMemset(d_w, 0, sizeof(double) * k * MP) == cudaSuccess && cudaMemset(d_Sinv, 0, sizeof(double) * k * MP * MP) == cudaSuccess;
    ok = ok && cudaMemset(d_prior, 0, sizeof(double) * MP) == cudaSuccess;
    ok = ok && cudaMemcpy2D(d_w, sizeof(double) * MP, w, sizeof(double) * P.m, sizeof(double) * P.m, k, cudaMemcpyHostToDevice) == cudaSuccess;
    for (int o = 0; o < k && ok; ++o)
        ok = cudaMemcpy2D(d_Sinv + static_cast<int64_t>(o) * MP * MP, sizeof(double) * MP, iSigma_w + static_cast<int64_t>(o) * P.m * P.m,
                          sizeof(double) * P.m, sizeof(double) * P.m, P.m, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && h2d(d_prior, priors, sizeof(double) * P.m) == cudaSuccess;
    ok = ok && h2d(d_X, hx.data(), sizeof(double) * n * d) == cudaSuccess;
    ok = ok && h2d(d_ob, ob.data(), d) == cudaSuccess;
    if (Psig) ok = ok && h2d(d_Psi, Psig, sizeof(double) * n * psi_w) == cudaSuccess;
    if (!ok) {
        set_error("predict_missing_group: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return cleanup(GPZ_ERR_CUDA);
    }
    if ((rc = prep_params(d_theta, P, cov ? 1 : 0, st, &launches))) return cleanup(rc);
    double *d_mu = d_out, *d_nu = d_out + k * n, *d_be = d_out + 2 * k * n, *d_ga = d_out + 3 * k * n;
    rc = cov ? predict_missing_cov(P, d_X, d_Psi, n, d_ob, d_prior, d_w, d_Sinv, d_mu, d_nu, d_be, d_ga, d_Phi, st, &launches)
             : predict_missing_diag(P, d_X, d_Psi, n, d_ob, d_prior, d_w, d_Sinv, d_mu, d_nu, d_be, d_ga, d_Phi, st, &launches);
    if (rc) return cleanup(rc);
    if (PHI) {
        if ((rc = transpose_out(d_Phi, MP, n, P.m, d_col, st))) return cleanup(rc);
        ok = cudaMemcpyAsync(PHI, d_col, sizeof(double) * n * P.m, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    }
    ok = ok && cudaMemcpyAsync(mu, d_mu, sizeof(double) * k * n, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(nu, d_nu, sizeof(double) * k * n, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(beta_i, d_be, sizeof(double) * k * n, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(gamma, d_ga, sizeof(double) * k * n, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) {
        set_error("predict_missing_group: %s", cudaGetErrorString(cudaGetLastError()));
        return cleanup(GPZ_ERR_CUDA);
    }
    return cleanup(GPZ_OK);
}

int gpz_predict(const gpz_model* model, const double* theta, const double* w, const double* iSigma_w, int64_t n,
                const double* Xz, const double* Psi, const double* priors, double* mu, double* nu, double* beta_i,
                double* gamma, double* PHI, int device) {
    if (!theta || !w || !iSigma_w || !Xz || !mu || !nu || !beta_i || !gamma || n < 0) {
        set_error("gpz_predict: NULL argument");
        return GPZ_ERR_USAGE;
    }
    Params P{};
    int rc;
    if ((rc = fill_params(P, model))) return rc;
    if ((rc = check_device(device))) return rc;
    {   // rows with missing inputs: group by NaN pattern and dispatch per group (predict.m:45-69)
        bool any_nan = false;
        for (int64_t i = 0; i < n * P.d && !any_nan; ++i) any_nan = Xz[i] != Xz[i];
        if (any_nan) {
            if (!priors) {
                set_error("gpz_predict: rows with missing inputs need the basis priors (model.best.priors, getPrior.m)");
                return GPZ_ERR_USAGE;
            }
            std::map<std::string, std::vector<int64_t>> groups;
            std::vector<std::string> order;
            std::string key(static_cast<size_t>(P.d), '1');
            for (int64_t i = 0; i < n; ++i) {
                for (int a = 0; a < P.d; ++a) key[a] = (Xz[static_cast<int64_t>(a) * n + i] == Xz[static_cast<int64_t>(a) * n + i]) ? '1' : '0';
                auto it = groups.find(key);
                if (it == groups.end()) {
                    order.push_back(key);
                    it = groups.emplace(key, std::vector<int64_t>()).first;
                }
                it->second.push_back(i);
            }
            const int k = P.k;
            for (const std::string& pk : order) {
                const std::vector<int64_t>& rows = groups[pk];
                const int64_t ng = static_cast<int64_t>(rows.size());
                std::vector<double> Xg(static_cast<size_t>(ng) * P.d), Pg, o_mu(static_cast<size_t>(ng) * k), o_nu(o_mu.size()),
                    o_be(o_mu.size()), o_ga(o_mu.size()), o_phi(PHI ? static_cast<size_t>(ng) * P.m : 0);
                for (int a = 0; a < P.d; ++a)
                    for (int64_t r = 0; r < ng; ++r) Xg[static_cast<size_t>(a) * ng + r] = Xz[static_cast<int64_t>(a) * n + rows[r]];
                if (Psi && mode_is_cov(P.mode)) {         // d x d x n: one contiguous d*d block per row
                    const size_t dd = static_cast<size_t>(P.d) * P.d;
                    Pg.resize(static_cast<size_t>(ng) * dd);
                    for (int64_t r = 0; r < ng; ++r) {
                        memcpy(Pg.data() + r * dd, Psi + rows[r] * dd, sizeof(double) * dd);
                    }
                    for (int64_t r = 0; r < ng; ++r)         // only Psi(o,o) enters: whatever sits in the missing rows / columns is dropped
                        for (int a = 0; a < P.d; ++a)
                            if (pk[a] != '1')
                                for (int b = 0; b < P.d; ++b) Pg[r * dd + a + b * P.d] = Pg[r * dd + b + a * P.d] = 0.0;
                } else if (Psi) {
                    Pg.resize(static_cast<size_t>(ng) * P.d);
                    for (int a = 0; a < P.d; ++a)
                        for (int64_t r = 0; r < ng; ++r) Pg[static_cast<size_t>(a) * ng + r] = Psi[static_cast<int64_t>(a) * n + rows[r]];
                }
                bool full = true;
                std::vector<unsigned char> ob(static_cast<size_t>(P.d));
                for (int a = 0; a < P.d; ++a) {
                    ob[a] = pk[a] == '1';
                    full = full && ob[a];
                }
                if (full)
                    rc = gpz_predict(model, theta, w, iSigma_w, ng, Xg.data(), Psi ? Pg.data() : nullptr, priors, o_mu.data(), o_nu.data(),
                                     o_be.data(), o_ga.data(), PHI ? o_phi.data() : nullptr, device);
                else
                    rc = predict_missing_group(model, theta, w, iSigma_w, ng, Xg.data(), Psi ? Pg.data() : nullptr, priors, ob, o_mu.data(),
                                               o_nu.data(), o_be.data(), o_ga.data(), PHI ? o_phi.data() : nullptr, device);
                if (rc) return rc;
                for (int o = 0; o < k; ++o)
                    for (int64_t r = 0; r < ng; ++r) {
                        mu[o * n + rows[r]] = o_mu[o * ng + r];
                        nu[o * n + rows[r]] = o_nu[o * ng + r];
                        beta_i[o * n + rows[r]] = o_be[o * ng + r];
                        gamma[o * n + rows[r]] = o_ga[o * ng + r];
                    }
                if (PHI)
                    for (int j = 0; j < P.m; ++j)
                        for (int64_t r = 0; r < ng; ++r) PHI[static_cast<int64_t>(j) * n + rows[r]] = o_phi[static_cast<size_t>(j) * ng + r];
            }
            return GPZ_OK;
        }
    }
    if (n == 0) return GPZ_OK;
    const bool cov_psi = Psi && mode_is_cov(P.mode);
    std::vector<void*> allocs;
    cudaStream_t st = nullptr;
    int64_t launches = 0;
    auto cleanup = [&]() {
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
        for (void* p : allocs) cudaFree(p);
    };
#define PR(call)             \
    do {                     \
        int r__ = (call);    \
        if (r__) {           \
            cleanup();       \
            return r__;      \
        }                    \
    } while (0)
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        return GPZ_ERR_CUDA;
    }
    const int64_t MP = P.MP;
    const int k = P.k;
    const bool dbg = getenv("GPZ_B200_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tdbg = now();
    auto lap = [&](const char* what, bool sync = true) {
        if (!dbg) return;
        if (sync) cudaStreamSynchronize(st);
        const double t1 = now();
        fprintf(stderr, "gpz_predict: %-18s %8.2f ms\n", what, t1 - tdbg);
        tdbg = t1;
    };
    PR(alloc_params(P, allocs, cov_psi ? 1 : 0));
    double *d_theta, *d_w, *d_Sinv, *d_X, *d_Psi = nullptr, *d_Phi, *d_dotv, *d_mu, *d_nupart, *d_nu, *d_elns, *d_beta, *d_gamma, *d_col = nullptr;
    PR(dev_alloc(allocs, &d_theta, P.p));
    PR(dev_alloc(allocs, &d_w, k * MP));
    PR(dev_alloc(allocs, &d_Sinv, k * MP * MP));
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // nu = rowsum((PHI iSigma) .* PHI) (predictDiag.m:66 / predictCov.m:61) is the same n m^2 product as T in the objective:
    // same engine choice (int8 digit GEMM unless one basis tile on few rows, see ensure_workspace)
    const int oz_s = (!Psi && ozmma_available() && !(MP <= 128 && n < 250000)) ? 7 : 0;
    int64_t chunk = static_cast<int64_t>(0.4 * static_cast<double>(free_b) / (8.0 * (MP + (PHI ? P.m : 0)) + static_cast<double>(oz_s) * MP));
    chunk = chunk / 1024 * 1024;
    // predict is a stateless call: its buffers are allocated and freed every time, and multi-GB cudaMalloc / cudaFree cost
    // far more than the kernels (1e6 rows x 1000 bases: 45 ms of kernels).  Row chunks of at most ~2 GB of PHI keep every
    // launch large enough to fill the machine and the allocations small.
    {
        const char* env = getenv("GPZ_B200_PREDICT_CHUNK");
        int64_t cap = env ? atoll(env) : (int64_t{1} << 31) / (8 * MP);
        cap = cap / 1024 * 1024;
        if (cap < 1024) cap = 1024;
        if (chunk > cap) chunk = cap;
    }
    if (chunk < 1024) chunk = 1024;
    if (chunk > n) chunk = n;
    PR(dev_alloc(allocs, &d_X, n * P.d));
    if (Psi) PR(dev_alloc(allocs, &d_Psi, n * P.d * (cov_psi ? P.d : 1)));
    PR(dev_alloc(allocs, &d_Phi, chunk * MP));
    if (PHI) PR(dev_alloc(allocs, &d_col, chunk * P.m));
    PR(dev_alloc(allocs, &d_dotv, k * n));
    PR(dev_alloc(allocs, &d_mu, k * n));
    PR(dev_alloc(allocs, &d_nupart, (MP / TILE) * n));
    PR(dev_alloc(allocs, &d_nu, k * n));
    PR(dev_alloc(allocs, &d_elns, k * n));
    PR(dev_alloc(allocs, &d_beta, k * n));
    PR(dev_alloc(allocs, &d_gamma, k * n));
    double *oz_D8 = nullptr, *oz_ea = nullptr, *oz_ws = nullptr, *oz_flag = nullptr;
    if (oz_s > 0) {
        PR(dev_alloc(allocs, &oz_D8, oz_digit_bytes(static_cast<int>(MP), oz_s, chunk) / 8 + 1));
        PR(dev_alloc(allocs, &oz_ea, oz_padded_rows(chunk)));
        PR(dev_alloc(allocs, &oz_ws, oz_workspace_bytes(static_cast<int>(MP), oz_s) / 8 + 1));
        PR(dev_alloc(allocs, &oz_flag, 1));
    }
    auto cuda_ok = [&](cudaError_t e, const char* what) -> int {
        if (e == cudaSuccess) return GPZ_OK;
        set_error("gpz_predict: %s: %s", what, cudaGetErrorString(e));
        return (int)GPZ_ERR_CUDA;
    };
    PR(cuda_ok(cudaMemcpyAsync(d_theta, theta, sizeof(double) * P.p, cudaMemcpyHostToDevice, st), "H2D theta"));
    PR(cuda_ok(cudaMemsetAsync(d_w, 0, sizeof(double) * k * MP, st), "memset"));
    PR(cuda_ok(cudaMemsetAsync(d_Sinv, 0, sizeof(double) * k * MP * MP, st), "memset"));
    PR(cuda_ok(cudaMemcpy2DAsync(d_w, sizeof(double) * MP, w, sizeof(double) * P.m, sizeof(double) * P.m, k, cudaMemcpyHostToDevice, st), "H2D w"));
    for (int o = 0; o < k; ++o)
        PR(cuda_ok(cudaMemcpy2DAsync(d_Sinv + static_cast<int64_t>(o) * MP * MP, sizeof(double) * MP, iSigma_w + static_cast<int64_t>(o) * P.m * P.m,
                                     sizeof(double) * P.m, sizeof(double) * P.m, P.m, cudaMemcpyHostToDevice, st), "H2D iSigma_w"));
    // store X relative to its column means (only x - p matters; prep_params shifts P by the same constant): raw upload,
    // NaN-aware column means and the subtraction on the device (the host pass over n x d cost more than all the kernels)
    PR(cuda_ok(h2d(d_X, Xz, sizeof(double) * n * P.d), "H2D X"));
    colmean_shift_kernel<<<P.d, 1024, 0, st>>>(d_X, n, P.xshift);
    PR(cuda_ok(cudaGetLastError(), "colmean_shift_kernel"));
    ++launches;
    if (Psi) PR(cuda_ok(cudaMemcpyAsync(d_Psi, Psi, sizeof(double) * n * P.d * (cov_psi ? P.d : 1), cudaMemcpyHostToDevice, st), "H2D Psi"));
    lap("alloc + upload");
    PR(prep_params(d_theta, P, cov_psi ? 1 : 0, st, &launches));
    RowData R;
    R.n = n;
    R.X = d_X;
    R.Psi = d_Psi;
    double* d_scratch = nullptr;
    PR(dev_alloc(allocs, &d_scratch, 2 * (MP / TILE) * chunk));
    if (!Psi) {
        PR(dev_alloc(allocs, &R.F, n * P.QP));
        PR(build_features(P, d_X, n, 0, n, R.F, st, &launches));
    }
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t r1 = (r0 + chunk < n) ? r0 + chunk : n;
        DotSpec ds{2, {P.v, d_w}, {d_dotv, d_mu}};
        if (k > 1) ds.n = 0;
        PR(phi_build(P, R, r0, r1, d_Phi, ds, d_scratch, st, &launches));
        if (k > 1)
            for (int o = 0; o < k; ++o)
                PR(rowdot(d_Phi, MP, P.m, r1 - r0, DotSpec{2, {P.v + o * MP, d_w + o * MP}, {d_dotv + o * n + r0, d_mu + o * n + r0}}, st, &launches));
        if (!Psi) {
            if (oz_s > 0) {
                if (r0 == 0) PR(cuda_ok(cudaMemsetAsync(oz_flag, 0, sizeof(double), st), "memset"));
                PR(ozaki_digits(d_Phi, MP, static_cast<int>(MP), P.m, r1 - r0, oz_s, nullptr, nullptr, 0, reinterpret_cast<int8_t*>(oz_D8),
                                nullptr, oz_ea, reinterpret_cast<int*>(oz_flag), st, &launches));
            }
            for (int o = 0; o < k; ++o) {
                if (oz_s > 0)
                    PR(ozaki_tgemm(d_Phi, MP, reinterpret_cast<const int8_t*>(oz_D8), oz_ea, d_Sinv + static_cast<int64_t>(o) * MP * MP,
                                   static_cast<int>(MP), P.m, r1 - r0, oz_s, nullptr, nullptr, 0, d_nupart + r0, n, nullptr, nullptr, oz_ws, st,
                                   nullptr, nullptr, &launches));
                else
                    PR(tgemm(d_Phi, MP, d_Sinv + static_cast<int64_t>(o) * MP * MP, static_cast<int>(MP), P.m, r1 - r0, nullptr, nullptr, 0,
                             d_nupart + r0, n, nullptr, st, &launches));
                sum_cols_kernel<<<static_cast<unsigned>(ceil_div(r1 - r0, 256)), 256, 0, st>>>(d_nupart + r0, static_cast<int>(MP / TILE), n, r1 - r0, d_nu + o * n + r0);
            }
        }
        if (PHI) {
            PR(transpose_out(d_Phi, MP, r1 - r0, P.m, d_col, st));
            PR(cuda_ok(cudaMemcpy2DAsync(PHI + r0, sizeof(double) * n, d_col, sizeof(double) * (r1 - r0), sizeof(double) * (r1 - r0), P.m,
                                         cudaMemcpyDeviceToHost, st), "D2H PHI"));
            PR(cuda_ok(cudaStreamSynchronize(st), "sync"));
        }
    }
    if (oz_s > 0) {                  // non-finite PHI has no digits: the affected answer is NaN, as the fp64 path would give
        nan_if_flag_kernel<<<static_cast<unsigned>(ceil_div(k * n, 256)), 256, 0, st>>>(d_nu, k * n, reinterpret_cast<const int*>(oz_flag));
        ++launches;
    }
    lap("row chunks");
    exp_rows_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, st>>>(d_dotv, P.bk, P.het, k, n, d_elns, d_beta);
    PR(cuda_ok(cudaMemsetAsync(d_gamma, 0, sizeof(double) * k * n, st), "memset"));
    if (cov_psi) PR(predict_noisy_cov(P, R, d_w, d_Sinv, d_elns, d_mu, d_nu, d_beta, d_gamma, st, &launches));
    else if (Psi) PR(predict_noisy_diag(P, R, d_w, d_Sinv, d_elns, d_mu, d_nu, d_beta, d_gamma, st, &launches));
    PR(cuda_ok(cudaMemcpyAsync(mu, d_mu, sizeof(double) * k * n, cudaMemcpyDeviceToHost, st), "D2H"));
    PR(cuda_ok(cudaMemcpyAsync(nu, d_nu, sizeof(double) * k * n, cudaMemcpyDeviceToHost, st), "D2H"));
    PR(cuda_ok(cudaMemcpyAsync(beta_i, d_beta, sizeof(double) * k * n, cudaMemcpyDeviceToHost, st), "D2H"));
    PR(cuda_ok(cudaMemcpyAsync(gamma, d_gamma, sizeof(double) * k * n, cudaMemcpyDeviceToHost, st), "D2H"));
    PR(cuda_ok(cudaStreamSynchronize(st), "sync"));
    PR(cuda_ok(cudaGetLastError(), "kernel"));
#undef PR
    lap("tail + download");
    cleanup();
    lap("free", false);
    return GPZ_OK;
}

int gpz_inv_logdet(int32_t m, const double* X, double* Xi, double* logdet, int device) {
    if (m < 1 || !X || !Xi) {
        set_error("gpz_inv_logdet: bad arguments");
        return GPZ_ERR_USAGE;
    }
    int rc;
    if ((rc = check_device(device))) return rc;
    const int64_t MP = round_up(m, TILE);
    double *S = nullptr, *Si = nullptr, *ld = nullptr;
    SolveWs ws;
    cudaStream_t st = nullptr;
    int64_t launches = 0;
    struct Cleanup {                      // every exit path below releases what was allocated so far
        double*& a;
        double*& b;
        double*& c;
        SolveWs& w;
        cudaStream_t& s;
        ~Cleanup() {
            if (a) cudaFree(a);
            if (b) cudaFree(b);
            if (c) cudaFree(c);
            solve_ws_free(w);
            if (s) cudaStreamDestroy(s);
        }
    } cleanup{S, Si, ld, ws, st};
    GPZ_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    GPZ_CUDA(cudaMalloc(&S, sizeof(double) * MP * MP));
    GPZ_CUDA(cudaMalloc(&Si, sizeof(double) * MP * MP));
    GPZ_CUDA(cudaMalloc(&ld, sizeof(double)));
    if ((rc = solve_ws_alloc(ws, static_cast<int>(MP)))) return rc;
    GPZ_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * MP * MP, st));
    GPZ_CUDA(cudaMemsetAsync(ws.flag, 0, sizeof(int), st));
    GPZ_CUDA(cudaMemcpy2DAsync(S, sizeof(double) * MP, X, sizeof(double) * m, sizeof(double) * m, m, cudaMemcpyHostToDevice, st));
    rc = spd_inverse(S, m, static_cast<int>(MP), Si, ld, ws, st, &launches);
    int flag = 0;
    double hl = 0.0;
    if (!rc) {
        double* mm = nullptr;
        double hmm[2] = {1.0, 1.0};
        GPZ_CUDA(cudaMalloc(&mm, sizeof(double) * (m + 2)));
        if ((rc = chol_diag_minmax(ws, S, m, static_cast<int>(MP), mm, st))) return rc;
        GPZ_CUDA(cudaMemcpyAsync(hmm, mm, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
        GPZ_CUDA(cudaMemcpyAsync(&flag, ws.flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        GPZ_CUDA(cudaStreamSynchronize(st));
        // The reference pseudo-inverts by SVD and drops singular values <= m eps(max s) (inv_logdet.m:7-12).  The Cholesky route is
        // that same inverse while every singular value is kept; when the factorisation fails (not positive definite: rank
        // deficient or indefinite) or the factor's diagonal says cond(X) >= (max L_ii / min L_ii)^2 comes within 1e3 of
        // 1 / (m eps), the truncating SVD itself is run (one-sided Jacobi, solve.cu)
        const double ratio = (hmm[0] > 0.0) ? hmm[1] / hmm[0] : 1e300;
        const bool near_singular = flag || !(ratio * ratio * 1e3 < 1.0 / (static_cast<double>(m) * 2.220446049250313e-16));
        if (near_singular) {
            int* cnt = nullptr;
            GPZ_CUDA(cudaMalloc(&cnt, sizeof(int)));
            GPZ_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * MP * MP, st));
            GPZ_CUDA(cudaMemcpy2DAsync(S, sizeof(double) * MP, X, sizeof(double) * m, sizeof(double) * m, m, cudaMemcpyHostToDevice, st));
            rc = svd_pinv_logdet(S, m, static_cast<int>(MP), ws.W, Si, &hl, cnt, mm, st, &launches);
            cudaFree(cnt);
            if (!rc) {
                GPZ_CUDA(cudaMemcpy2DAsync(Xi, sizeof(double) * m, Si, sizeof(double) * MP, sizeof(double) * m, m, cudaMemcpyDeviceToHost, st));
                GPZ_CUDA(cudaStreamSynchronize(st));
            }
        } else {
            GPZ_CUDA(cudaMemcpy2DAsync(Xi, sizeof(double) * m, Si, sizeof(double) * MP, sizeof(double) * m, m, cudaMemcpyDeviceToHost, st));
            GPZ_CUDA(cudaMemcpyAsync(&hl, ld, sizeof(double), cudaMemcpyDeviceToHost, st));
            GPZ_CUDA(cudaStreamSynchronize(st));
        }
        cudaFree(mm);
        if (!rc && logdet) *logdet = hl;
    }
    return rc;
}

int gpz_dxy(int64_t n, int32_t m, int32_t d, const double* X, const double* Y, double* D, int device) {
    if (n < 0 || m < 0 || d < 1 || !X || !Y || !D) {
        set_error("gpz_dxy: bad arguments");
        return GPZ_ERR_USAGE;
    }
    int rc;
    if ((rc = check_device(device))) return rc;
    if (n == 0 || m == 0) return GPZ_OK;
    double *dX = nullptr, *dY = nullptr, *dD = nullptr;
    GPZ_CUDA(cudaMalloc(&dX, sizeof(double) * n * d));
    GPZ_CUDA(cudaMalloc(&dY, sizeof(double) * m * d));
    GPZ_CUDA(cudaMalloc(&dD, sizeof(double) * n * m));
    GPZ_CUDA(h2d(dX, X, sizeof(double) * n * d));
    GPZ_CUDA(h2d(dY, Y, sizeof(double) * m * d));
    rc = dxy_device(dX, n, dY, m, d, dD, nullptr);
    if (!rc) GPZ_CUDA(cudaMemcpy(D, dD, sizeof(double) * n * m, cudaMemcpyDeviceToHost));
    cudaFree(dX);
    cudaFree(dY);
    cudaFree(dD);
    return rc;
}

int gpz_dxy_colmean(int64_t n, int32_t m, int32_t d, const double* X, const double* Y, double* mean, int device) {
    if (n < 1 || m < 1 || d < 1 || !X || !Y || !mean) {
        set_error("gpz_dxy_colmean: bad arguments");
        return GPZ_ERR_USAGE;
    }
    int rc;
    if ((rc = check_device(device))) return rc;
    double *dX = nullptr, *dY = nullptr, *dP = nullptr, *dM = nullptr;
    GPZ_CUDA(cudaMalloc(&dX, sizeof(double) * n * d));
    GPZ_CUDA(cudaMalloc(&dY, sizeof(double) * m * d));
    GPZ_CUDA(cudaMalloc(&dP, sizeof(double) * m * dxy_colmean_chunks(n)));
    GPZ_CUDA(cudaMalloc(&dM, sizeof(double) * m));
    GPZ_CUDA(h2d(dX, X, sizeof(double) * n * d));
    GPZ_CUDA(h2d(dY, Y, sizeof(double) * m * d));
    rc = dxy_colmean_device(dX, n, dY, m, d, dP, dM, nullptr);
    if (!rc) GPZ_CUDA(cudaMemcpy(mean, dM, sizeof(double) * m, cudaMemcpyDeviceToHost));
    cudaFree(dX);
    cudaFree(dY);
    cudaFree(dP);
    cudaFree(dM);
    return rc;
}

namespace {
__global__ void pad_rows_kernel(const double* __restrict__ src, int64_t ld, int64_t rows, int64_t K, int64_t Kp, double* __restrict__ dst) {
    const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (e >= rows * Kp) return;
    const int64_t r = e / Kp, c = e - r * Kp;
    dst[e] = c < K ? src[r * ld + c] : 0.0;
}
}  // namespace

int gpz_dgemm_nt(int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                 int32_t digits, int device) {
    if (M < 1 || N < 1 || K < 1 || K > 16384 || M > (1 << 30) || N > (1 << 30) || !A || !B || !C || lda < K || ldb < K || ldc < N ||
        digits < 3 || digits > 7) {
        set_error("gpz_dgemm_nt: bad arguments (1 <= K <= 16384, digits 3..7)");
        return GPZ_ERR_USAGE;
    }
    int rc;
    if ((rc = check_device(device))) return rc;
    if (!ozmma_available()) {
        set_error("gpz_dgemm_nt: the driver does not export cuTensorMapEncodeTiled");
        return GPZ_ERR_USAGE;
    }
    const int64_t Kp = round_up(K, 128);
    const int Kpi = static_cast<int>(Kp);
    std::vector<void*> bufs;
    auto D = [&](void** p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) bufs.push_back(*p);
        return e;
    };
    auto cleanup = [&]() { for (void* b : bufs) cudaFree(b); };
    double *dA = nullptr, *dB = nullptr, *pA = nullptr, *pB = nullptr, *dC = nullptr, *ea = nullptr, *eb = nullptr, *partial = nullptr;
    int8_t *A8 = nullptr, *B8 = nullptr;
    int* flag = nullptr;
    const int64_t Mp = oz_padded_rows(M), Np = oz_padded_rows(N);
    const int64_t npart = ozmma_partial_doubles(static_cast<int>(M), static_cast<int>(N), 0, 1, 0);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = D(reinterpret_cast<void**>(&dA), sizeof(double) * M * lda);
    if (e == cudaSuccess) e = D(reinterpret_cast<void**>(&dB), sizeof(double) * N * ldb);
    if (e == cudaSuccess) e = D(reinterpret_cast<void**>(&pA), sizeof(double) * M * Kp);
    if (e == cudaSuccess) e = D(reinterpret_cast<void**>(&pB), sizeof(double) * N * Kp);
    if (e == cudaSuccess) e = D(reinterpret_cast<void**>(&dC), sizeof(double) * M * N);
    if (e == cudaSuccess) e = D(reinterpret_cast<void**>(&ea), sizeof(double) * Mp);
    if (e == cudaSuccess) e = D(reinterpret_cast<void**>(&eb), sizeof(double) * Np);
    if (e == cudaSuccess) e = D(reinterpret_cast<void**>(&partial), sizeof(double) * npart);
    if (e == cudaSuccess) e = D(reinterpret_cast<void**>(&A8), static_cast<size_t>(Mp) * digits * Kp);
    if (e == cudaSuccess) e = D(reinterpret_cast<void**>(&B8), static_cast<size_t>(Np) * digits * Kp);
    if (e == cudaSuccess) e = D(reinterpret_cast<void**>(&flag), sizeof(int));
    if (e == cudaSuccess) e = h2d(dA, A, sizeof(double) * M * lda);
    if (e == cudaSuccess) e = h2d(dB, B, sizeof(double) * N * ldb);
    if (e == cudaSuccess) e = cudaMemset(flag, 0, sizeof(int));
    if (e != cudaSuccess) {
        set_error("gpz_dgemm_nt: CUDA error %s", cudaGetErrorString(e));
        cleanup();
        return GPZ_ERR_CUDA;
    }
    int64_t launches = 0;
    pad_rows_kernel<<<static_cast<unsigned>(ceil_div(M * Kp, 256)), 256>>>(dA, lda, M, K, Kp, pA);
    pad_rows_kernel<<<static_cast<unsigned>(ceil_div(N * Kp, 256)), 256>>>(dB, ldb, N, K, Kp, pB);
    rc = ozaki_digits(pA, Kp, Kpi, static_cast<int>(K), M, digits, nullptr, nullptr, 0, A8, nullptr, ea, flag, nullptr, &launches);
    if (!rc) rc = ozaki_digits(pB, Kp, Kpi, static_cast<int>(K), N, digits, nullptr, nullptr, 0, B8, nullptr, eb, flag, nullptr, &launches);
    if (!rc) {
        // digits addressed {k, digit, row, chunk}; the row scales ea, eb carry the 256^-1 of the digit convention
        const int64_t str[3] = {Kp, static_cast<int64_t>(digits) * Kp, static_cast<int64_t>(digits) * Kp * Mp};
        const int64_t strB[3] = {Kp, static_cast<int64_t>(digits) * Kp, static_cast<int64_t>(digits) * Kp * Np};
        rc = ozmma_gemm_nt(A8, str, static_cast<int>(M), B8, strB, static_cast<int>(N), digits, digits + 1, Kpi, 1, 0, 0, partial, ea, eb,
                           1.0, 0, dC, N, 0, nullptr, &launches);
    }
    int hflag = 0;
    if (!rc) {
        e = cudaMemcpy(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy2D(C, sizeof(double) * ldc, dC, sizeof(double) * N, sizeof(double) * N, M, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            set_error("gpz_dgemm_nt: CUDA error %s", cudaGetErrorString(e));
            rc = GPZ_ERR_CUDA;
        }
    }
    cleanup();
    if (!rc && hflag) {          // non-finite inputs: propagate NaN (error convention)
        for (int64_t i = 0; i < M; ++i)
            for (int64_t j = 0; j < N; ++j) C[i * ldc + j] = nan("");
    }
    return rc;
}

void* gpz_stream(gpz_ctx* c) { return c ? static_cast<void*>(c->st) : nullptr; }

int gpz_sync(gpz_ctx* c) {
    if (!c) return GPZ_ERR_USAGE;
    GPZ_CUDA(cudaSetDevice(c->device));
    GPZ_CUDA(cudaStreamSynchronize(c->st));
    return GPZ_OK;
}

int64_t gpz_launch_count(const gpz_ctx* c) { return c ? c->launches : -1; }

int gpz_last_timing(gpz_ctx* c, double ms[12]) {
    if (!c || !c->ws_ready) {
        set_error("gpz_last_timing: no evaluation yet");
        return GPZ_ERR_USAGE;
    }
    GPZ_CUDA(cudaSetDevice(c->device));
    GPZ_CUDA(cudaStreamSynchronize(c->st));
    float t;
    // ev0 start | ev1 after first PHI build | ev2 after Gram+allreduce | ev3 after solve | ev4 after T-GEMM | ev5 end
    for (int i = 0; i < 5; ++i) {
        GPZ_CUDA(cudaEventElapsedTime(&t, c->ev[i], c->ev[i + 1]));
        ms[i] = t;
    }
    GPZ_CUDA(cudaEventElapsedTime(&t, c->ev[0], c->ev[5]));
    ms[5] = t;
    if (c->last_was_replay) {                                     // phases: last plain run; total: this replay
        GPZ_CUDA(cudaEventElapsedTime(&t, c->gev[0], c->gev[1]));
        ms[5] = t;
    }
    GPZ_CUDA(cudaEventElapsedTime(&t, c->kev[0], c->kev[1]));
    ms[6] = t;
    GPZ_CUDA(cudaEventElapsedTime(&t, c->kev[2], c->kev[3]));
    ms[7] = t;
    ms[8] = ms[9] = ms[10] = ms[11] = 0.0;
    if (c->opt_ozaki > 0) {
        GPZ_CUDA(cudaEventElapsedTime(&t, c->kev[4], c->kev[5]));
        ms[8] = t;                                              // the tcgen05 digit-GEMM launch of T = PHI iSigma (first chunk)
        const double rows = static_cast<double>(c->tr.n < c->chunk_rows ? c->tr.n : c->chunk_rows);
        double levels = 0.0;
        for (int e = 2; e <= c->opt_ozaki + 1; ++e) levels += e - 1;
        ms[9] = 2.0 * rows * c->P.MP * c->P.MP * levels;        // int8 operations those GEMMs executed
    }
    ms[10] = c->opt_ozaki;
    ms[11] = c->opt_ozaki_gram;
    return GPZ_OK;
}

int gpz_kernel_timing(gpz_ctx* c, double ms[4]) {
    if (!c || !c->ws_ready) {
        set_error("gpz_kernel_timing: no evaluation yet");
        return GPZ_ERR_USAGE;
    }
    GPZ_CUDA(cudaSetDevice(c->device));
    GPZ_CUDA(cudaStreamSynchronize(c->st));
    for (int i = 0; i < 4; ++i) {
        ms[i] = -1.0;
        if (!c->kt_valid[i]) continue;
        float t;
        GPZ_CUDA(cudaEventElapsedTime(&t, c->kt[2 * i], c->kt[2 * i + 1]));
        ms[i] = t;
    }
    return GPZ_OK;
}

int64_t gpz_graph_replays(const gpz_ctx* c) { return c ? c->graph_replays : -1; }

int gpz_set_option(gpz_ctx* c, const char* name, double value) {
    if (!c || !name) return GPZ_ERR_USAGE;
    graph_drop(c);                                  // any option may change what an evaluation enqueues
    c->seen_theta = nullptr;
    c->seen_out = nullptr;
    if (strcmp(name, "graph") == 0) {               // CUDA-graph replay of the evaluation: -1 auto (small problems), 0 off, 1 on
        c->opt_graph = value < 0.0 ? -1 : (value != 0.0 ? 1 : 0);
        return GPZ_OK;
    }
    if (strcmp(name, "chunk_rows") == 0) {
        if (c->ws_ready) {
            set_error("chunk_rows must be set before the first evaluation");
            return GPZ_ERR_USAGE;
        }
        c->opt_chunk_rows = static_cast<int64_t>(value);
        return GPZ_OK;
    }
    if (strcmp(name, "gemm_warps") == 0) {          // process-wide: 8 or 16 warps per CTA in the Gram / T-GEMM kernels
        if (value != 0.0 && value != 8.0 && value != 16.0) {
            set_error("gemm_warps must be 0 (defaults), 8 or 16");
            return GPZ_ERR_USAGE;
        }
        g_gemm_warps = static_cast<int>(value);
        return GPZ_OK;
    }
    if (strcmp(name, "solve_lookahead") == 0) {     // process-wide: blocked Cholesky (solve.cu): 2 = look-ahead + incremental inverse (default), 1 = look-ahead, 0 = in-stream
        g_solve_lookahead = value >= 2.0 ? 2 : (value != 0.0 ? 1 : 0);
        return GPZ_OK;
    }
    if (strcmp(name, "moment_warps") == 0) {        // process-wide: warps per CTA of the moment GEMM dPHI'F (gemm.cu): 0 = by shape (default), 8 or 16
        if (value != 0.0 && value != 8.0 && value != 16.0) {
            set_error("moment_warps must be 0 (by shape), 8 or 16");
            return GPZ_ERR_USAGE;
        }
        g_moment_warps = static_cast<int>(value);
        return GPZ_OK;
    }
    if (strcmp(name, "prep_block") == 0) {          // process-wide: CTA size of the per-basis parameter kernels (phi.cu), 32 = default
        if (value != 32.0 && value != 64.0 && value != 128.0) {
            set_error("prep_block must be 32, 64 or 128");
            return GPZ_ERR_USAGE;
        }
        g_prep_block = static_cast<int>(value);
        return GPZ_OK;
    }
    if (strcmp(name, "phi_persist") == 0) {         // process-wide: persistent column-stationary PHI kernel (gemm.cu): 1 = default, 2 = staggered variant, 0 = off
        g_phi_persist = value >= 2.0 ? 2 : (value != 0.0 ? 1 : 0);
        return GPZ_OK;
    }
    if (strcmp(name, "ozaki_int_fold") == 0) {      // process-wide: integer folding of the lowest digit levels (ozmma.cu), 1 = default
        ozmma_set_int_fold(static_cast<int>(value));
        return GPZ_OK;
    }
    if (strcmp(name, "ozaki_prefetch") == 0) {      // process-wide: L2 prefetch of the T-GEMM epilogue's PHI block (ozmma.cu), 1 = default
        ozmma_set_prefetch(static_cast<int>(value));
        return GPZ_OK;
    }
    if (strcmp(name, "ozaki_level_group") == 0) {   // process-wide: schedule of the digit GEMMs (ozmma.cu), 2 = default
        ozmma_set_level_group(static_cast<int>(value));
        return GPZ_OK;
    }
    if (strcmp(name, "tensor_phi") == 0 || strcmp(name, "fused_backproj") == 0) {
        if (c->ws_ready) {
            set_error("%s must be set before the first evaluation", name);
            return GPZ_ERR_USAGE;
        }
        if (name[0] == 't') c->opt_tensor_phi = static_cast<int>(value); else c->opt_fused_bp = value != 0.0;
        return GPZ_OK;
    }
    if (strcmp(name, "ozaki_slices") == 0) {
        if (c->ws_ready) {
            set_error("%s must be set before the first evaluation", name);
            return GPZ_ERR_USAGE;
        }
        if (value != 0.0 && (value < 3.0 || value > 7.0)) {
            set_error("ozaki_slices must be 0 (off) or 3..7 (base-256 digits per operand)");
            return GPZ_ERR_USAGE;
        }
        c->opt_ozaki = static_cast<int>(value);
        return GPZ_OK;
    }
    if (strcmp(name, "gc_int8") == 0) {
        if (c->ws_ready) {
            set_error("%s must be set before the first evaluation", name);
            return GPZ_ERR_USAGE;
        }
        c->opt_gc_int8 = value != 0.0;
        return GPZ_OK;
    }
    if (strcmp(name, "gc_fast") == 0) {
        if (c->ws_ready) {
            set_error("%s must be set before the first evaluation", name);
            return GPZ_ERR_USAGE;
        }
        c->opt_gc_fast = value != 0.0;
        return GPZ_OK;
    }
    if (strcmp(name, "ozaki_gram_slices") == 0) {
        if (c->ws_ready) {
            set_error("%s must be set before the first evaluation", name);
            return GPZ_ERR_USAGE;
        }
        if (value < 2.0 || value > 7.0) {
            set_error("ozaki_gram_slices must be 2..7 (and is capped by ozaki_slices)");
            return GPZ_ERR_USAGE;
        }
        c->opt_ozaki_gs = static_cast<int>(value);
        return GPZ_OK;
    }
    if (strcmp(name, "ozaki_gram") == 0) {
        if (c->ws_ready) {
            set_error("%s must be set before the first evaluation", name);
            return GPZ_ERR_USAGE;
        }
        c->opt_ozaki_gram = value != 0.0;
        return GPZ_OK;
    }
    if (strcmp(name, "spare_column") == 0) {
        if (c->ws_ready) {
            set_error("%s must be set before the first evaluation", name);
            return GPZ_ERR_USAGE;
        }
        c->opt_aug = value != 0.0;
        return GPZ_OK;
    }
    set_error("unknown option '%s'", name);
    return GPZ_ERR_USAGE;
}

}  // extern "C"
