// Error-free fp64 GEMMs on the int8 tensor cores (Ozaki splitting) for the two n x m x m products of an evaluation:
//     T = PHI * iSigma         (GPz/GPz.m:69,72)        S = PHI' diag(w) PHI   (GPz/GPz.m:63-65)
//
// tcgen05.mma has no fp64 kind and the fp64 DMMA pipe tops out at ~36 TFLOP/s.  Here every fp64 operand is scaled by a
// power of two 2^-E into r in (-0.4961, 0.4961) and written as s balanced base-256 digits
//     r = sum_{t=1..s} d_t 256^-t + O(256^-s / 2),    d_t in [-128, 127]
// (round to nearest at 8 s bits, then two's-complement digit extraction with carry).  Digit products are EXACT
// int8 x int8 -> int32 tensor-core GEMMs; all pairs with the same level e = t+u share one TMEM accumulator and the levels
// e = s+1 .. 2 are folded in fp64, smallest first, inside the hand-written tcgen05 kernel of ozmma.cu.  Pairs with
// t+u > s+1 are dropped: they are below 256^-s of (row scale x column scale), i.e. with the default s = 7 below 2^-55.
// s = 7 gives 56-bit fixed point per row/column (28 digit products); s = 6 gives 48 bits (21 products) -- comparable to
// the rounding of an fp64 GEMM with K ~ 1000.
#include "internal.cuh"

namespace gpz {

constexpr int OZ_MAXS = 7;      // 8 s <= 56 bits so that r * 2^(8s) fits an int64

// x < 2^ex / 1.0078125  ->  |x| 2^-(ex+1) < 0.4961: the leading digit stays in [-127, 127] after the carries
__device__ __forceinline__ int oz_exponent(double mx) {
    int ex = 0;
    if (mx > 0.0) frexp(mx * 1.0078125, &ex);
    return ex + 1;
}

// digits of I = sum_t d_t 256^(s-t), least significant first; d[t] for t = 0..s-1 (most significant first)
__device__ __forceinline__ int oz_digit(long long& I) {
    const int d = static_cast<int>(static_cast<signed char>(static_cast<unsigned char>(I & 255)));
    I = (I - d) >> 8;
    return d;
}

// ---- PHI rows -> digits A8[i][t][j] (K = j contiguous).  warp per row; lane handles 4 consecutive columns per step ----
__global__ void __launch_bounds__(256)
oz_slice_rows_kernel(const double* __restrict__ Phi, int64_t ld, int m, int MP, int64_t n, int s, int8_t* __restrict__ A8,
                     double* __restrict__ ea) {
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const double* row = Phi + i * ld;
    double mx = 0.0;
    for (int j = lane * 4; j < MP; j += 128) {
        const double4 v = *reinterpret_cast<const double4*>(row + j);
        if (j < m) mx = fmax(mx, fabs(v.x));
        if (j + 1 < m) mx = fmax(mx, fabs(v.y));
        if (j + 2 < m) mx = fmax(mx, fabs(v.z));
        if (j + 3 < m) mx = fmax(mx, fabs(v.w));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int E = oz_exponent(mx);
    if (lane == 0) ea[i] = ldexp(1.0, E - 8);          // a = ea * sum_t d_t 256^-(t-1)
    const double sc = ldexp(1.0, 8 * s - E);
    int8_t* out = A8 + i * static_cast<int64_t>(s) * MP;
    for (int j = lane * 4; j < MP; j += 128) {
        const double4 v = *reinterpret_cast<const double4*>(row + j);
        long long I[4] = {j < m ? __double2ll_rn(v.x * sc) : 0, j + 1 < m ? __double2ll_rn(v.y * sc) : 0,
                          j + 2 < m ? __double2ll_rn(v.z * sc) : 0, j + 3 < m ? __double2ll_rn(v.w * sc) : 0};
        for (int t = s - 1; t >= 0; --t) {
            char4 q;
            q.x = static_cast<signed char>(oz_digit(I[0]));
            q.y = static_cast<signed char>(oz_digit(I[1]));
            q.z = static_cast<signed char>(oz_digit(I[2]));
            q.w = static_cast<signed char>(oz_digit(I[3]));
            *reinterpret_cast<char4*>(out + static_cast<int64_t>(t) * MP + j) = q;
        }
    }
}

// ---- iSigma columns -> digits B8[j][u][l] (K = l contiguous); column m (aug) is w, every other column j uses the
// symmetric iSigma[j][l].  One block per column j.
__global__ void __launch_bounds__(256)
oz_slice_cols_kernel(const double* __restrict__ Sinv, int MP, int m, const double* __restrict__ waug, int s, int8_t* __restrict__ B8,
                     double* __restrict__ eb) {
    __shared__ double sh[8];
    __shared__ int Esh;
    const int j = blockIdx.x;
    const double* src = (j < m) ? Sinv + static_cast<int64_t>(j) * MP : ((j == m && waug != nullptr) ? waug : nullptr);
    double mx = 0.0;
    if (src != nullptr)
        for (int l = threadIdx.x; l < m; l += 256) mx = fmax(mx, fabs(src[l]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) mx = fmax(mx, sh[q]);
        const int E = oz_exponent(mx);
        Esh = E;
        eb[j] = ldexp(1.0, E - 8);
    }
    __syncthreads();
    const double sc = ldexp(1.0, 8 * s - Esh);
    int8_t* out = B8 + static_cast<int64_t>(j) * s * MP;
    for (int l = threadIdx.x; l < MP; l += 256) {
        long long I = (src != nullptr && l < m) ? __double2ll_rn(src[l] * sc) : 0;
        for (int u = s - 1; u >= 0; --u) out[static_cast<int64_t>(u) * MP + l] = static_cast<int8_t>(oz_digit(I));
    }
}

static int64_t al256(int64_t b) { return (b + 255) / 256 * 256; }

int64_t oz_workspace_bytes(int MP, int s, int64_t chunk_rows) {
    return 2 * al256(chunk_rows * static_cast<int64_t>(s) * MP) + al256(static_cast<int64_t>(MP) * s * MP) + 2 * al256(chunk_rows * 8) +
           al256(static_cast<int64_t>(MP) * 8);
}

// T-GEMM with fused epilogue through the int8 tensor cores.  ws: oz_workspace_bytes(MP, s, chunk_rows) bytes.
// Row chunks are software-pipelined over two streams: the digit extraction of chunk c+1 (HBM bound, stream aux) runs
// under the tcgen05 kernel of chunk c (stream st).  ev: 4 events.  nupart: [MP/128][nu_ld] row-sum partials of PHI .* T.
int ozaki_tgemm(const double* Phi, int64_t ld, const double* Sinv, int MP, int m, int64_t n, int s, int64_t chunk_rows,
                const double* rw, double* H, int accumulate, double* nupart, int64_t nu_ld, const double* waug, double* pred, void* ws,
                cudaStream_t st, cudaStream_t aux, cudaEvent_t* ev, cudaEvent_t tev0, cudaEvent_t tev1, int64_t* launches) {
    if (s < 2 || s > OZ_MAXS) {
        set_error("ozaki_tgemm: digits must be in [2, %d]", OZ_MAXS);
        return GPZ_ERR_USAGE;
    }
    unsigned char* p = static_cast<unsigned char*>(ws);
    auto take = [&](int64_t bytes) {
        unsigned char* r = p;
        p += al256(bytes);
        return r;
    };
    int8_t* A8[2];
    double* ea[2];
    for (int b = 0; b < 2; ++b) A8[b] = reinterpret_cast<int8_t*>(take(chunk_rows * static_cast<int64_t>(s) * MP));
    int8_t* B8 = reinterpret_cast<int8_t*>(take(static_cast<int64_t>(MP) * s * MP));
    for (int b = 0; b < 2; ++b) ea[b] = reinterpret_cast<double*>(take(chunk_rows * 8));
    double* eb = reinterpret_cast<double*>(take(static_cast<int64_t>(MP) * 8));
    cudaEvent_t* evS = ev;          // [2] digits of buffer b ready
    cudaEvent_t* evG = ev + 2;      // [2] kernel reading A8[b] done

    oz_slice_cols_kernel<<<MP, 256, 0, st>>>(Sinv, MP, m, waug, s, B8, eb);
    GPZ_KERNEL_CHECK();
    ++*launches;
    const int nchunks = static_cast<int>(ceil_div(n, chunk_rows));
    auto rows_of = [&](int c) { return (static_cast<int64_t>(c + 1) * chunk_rows <= n) ? chunk_rows : n - static_cast<int64_t>(c) * chunk_rows; };
    auto slice = [&](int c, cudaStream_t sx) {
        const int64_t r0 = static_cast<int64_t>(c) * chunk_rows;
        oz_slice_rows_kernel<<<static_cast<unsigned>(ceil_div(rows_of(c), 8)), 256, 0, sx>>>(Phi + r0 * ld, ld, m, MP, rows_of(c), s,
                                                                                             A8[c & 1], ea[c & 1]);
        ++*launches;
    };
    // everything enqueued so far on st (PHI, iSigma) must be visible to aux
    GPZ_CUDA(cudaEventRecord(evS[1], st));
    GPZ_CUDA(cudaStreamWaitEvent(aux, evS[1], 0));
    slice(0, st);
    GPZ_CUDA(cudaEventRecord(evS[0], st));
    int rc;
    for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1;
        const int64_t r0 = static_cast<int64_t>(c) * chunk_rows;
        const int64_t rows = rows_of(c);
        if (c + 1 < nchunks) {                       // digits of chunk c+1 on aux while the kernel of chunk c runs
            if (c >= 1) GPZ_CUDA(cudaStreamWaitEvent(aux, evG[b ^ 1], 0));      // A8[b^1] was read by the kernel of chunk c-1
            slice(c + 1, aux);
            GPZ_CUDA(cudaEventRecord(evS[b ^ 1], aux));
        }
        GPZ_CUDA(cudaStreamWaitEvent(st, evS[b], 0));
        if (c == 0 && tev0) GPZ_CUDA(cudaEventRecord(tev0, st));
        if ((rc = ozmma_tgemm(A8[b], B8, MP, s, s + 1, rows, ea[b], eb, Phi + r0 * ld, ld, rw != nullptr ? rw + r0 : nullptr,
                              H != nullptr ? H + r0 * ld : nullptr, accumulate, nupart + r0, nu_ld, waug != nullptr ? m : -1,
                              pred != nullptr ? pred + r0 : nullptr, st, launches)))
            return rc;
        if (c == 0 && tev1) GPZ_CUDA(cudaEventRecord(tev1, st));
        GPZ_CUDA(cudaEventRecord(evG[b], st));
    }
    return GPZ_OK;
}

}  // namespace gpz

// ================================================================================================
// Gram  S = PHI' diag(w) PHI  (GPz/GPz.m:63-65) through the int8 tensor cores.
//   A side = (w .* PHI)' , B side = PHI' ; contraction over the rows i.  Rows are cut into chunks of OZG_CH rows
//   (s * OZG_CH * 2^14 < 2^31 keeps every int32 level accumulator exact, and a chunk's digits stay L2-resident while
//   its 28 digit pairs re-stream them); the tcgen05 kernel folds groups of chunks into its fp64 registers and a small
//   kernel adds the group partials in fixed order.  Digits are stored transposed, chunk-major:  F[c][j][t][i] (weighted), G[c][j][t][i] (plain),
//   so that a digit of a 128-row x 128-byte tile is one TMA box.  Only tiles touching the lower triangle are computed
//   and the result is mirrored.  Scales are fixed powers of two: PHI <= 1 -> 4, the weights -> 2^ceil(log2 max w),
//   the spare column (y) -> 4 * 2^ceil(log2 max|y|).
// ================================================================================================
namespace gpz {

constexpr int OZG_CH = 1024;      // rows per K chunk: the digits of ~4 chunks in flight (x all tiles) stay in L2

__device__ __forceinline__ double pow2_ceil(double v) {
    int ex = 0;
    if (v > 0.0) frexp(v, &ex);
    return ldexp(1.0, ex);                      // v < 2^ex
}

// tile: 128 rows (i) x 32 columns (j) of PHI -> transposed digits.  A thread owns 4 consecutive rows of a column and
// emits one char4 per digit, a warp writes 128 contiguous bytes of one (column, digit) row.
__global__ void __launch_bounds__(256)
ozg_slice_kernel(const double* __restrict__ Phi, int64_t ld, int MP, int m, int64_t rows, int s, const double* __restrict__ wgt,
                 const double* __restrict__ scal, int aug, int8_t* __restrict__ F, int8_t* __restrict__ G) {
    __shared__ double tile[128][33];
    __shared__ double wsm[128];
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * 128;
    const int j0 = blockIdx.y * 32;
    const int tid = threadIdx.x;
    for (int e = tid; e < 128 * 32; e += 256) {
        const int r = e >> 5, c = e & 31;
        const int64_t gi = i0 + r;
        tile[r][c] = (gi < rows) ? Phi[gi * ld + j0 + c] : 0.0;
    }
    if (tid < 128) wsm[tid] = (i0 + tid < rows) ? wgt[i0 + tid] : 0.0;
    __syncthreads();
    const double sw = 1.0 / pow2_ceil(scal[0]);                    // weight scale: w * sw < 1
    const double sy = 1.0 / (4.0 * pow2_ceil(scal[1]));            // spare-column scale
    const double two8s = ldexp(1.0, 8 * s);
    const int64_t c = i0 / OZG_CH;                                 // chunk (128 divides OZG_CH)
    const int il = static_cast<int>(i0 % OZG_CH);
    const int ig = (tid & 31) * 4;                                 // 4 consecutive rows per thread
    for (int jj = tid >> 5; jj < 32; jj += 8) {
        const int j = j0 + jj;
        const double sp = (j < m) ? 0.25 : ((j == m && aug) ? sy : 0.0);
        long long IA[4], IB[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double rb = tile[ig + q][jj] * sp;
            const double ra = rb * (wsm[ig + q] * sw);
            IB[q] = __double2ll_rn(rb * two8s);
            IA[q] = __double2ll_rn(ra * two8s);
        }
        int8_t* fo = F + ((c * MP + j) * static_cast<int64_t>(s)) * OZG_CH + il + ig;
        int8_t* go = G + ((c * MP + j) * static_cast<int64_t>(s)) * OZG_CH + il + ig;
        for (int t = s - 1; t >= 0; --t) {
            char4 qa, qb;
            qa.x = static_cast<signed char>(oz_digit(IA[0]));
            qa.y = static_cast<signed char>(oz_digit(IA[1]));
            qa.z = static_cast<signed char>(oz_digit(IA[2]));
            qa.w = static_cast<signed char>(oz_digit(IA[3]));
            qb.x = static_cast<signed char>(oz_digit(IB[0]));
            qb.y = static_cast<signed char>(oz_digit(IB[1]));
            qb.z = static_cast<signed char>(oz_digit(IB[2]));
            qb.w = static_cast<signed char>(oz_digit(IB[3]));
            *reinterpret_cast<char4*>(fo + static_cast<int64_t>(t) * OZG_CH) = qa;
            *reinterpret_cast<char4*>(go + static_cast<int64_t>(t) * OZG_CH) = qb;
        }
    }
}

// sr[j] = (weight scale) * (column scale) / 256, sc[l] = (column scale) / 256: value = 2^E sum_t d_t 256^-t
__global__ void ozg_scales_kernel(const double* __restrict__ scal, int MP, int m, int aug, double* __restrict__ sr, double* __restrict__ sc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= MP) return;
    const double cw = pow2_ceil(scal[0]);
    const double cy = 4.0 * pow2_ceil(scal[1]);
    const double sj = (j < m) ? 4.0 : ((j == m && aug) ? cy : 0.0);
    sr[j] = cw * sj * (1.0 / 256.0);
    sc[j] = sj * (1.0 / 256.0);
}

int64_t oz_gram_workspace_bytes(int MP, int s, int64_t rows) {
    const int64_t nch = ceil_div(rows > 0 ? rows : 1, OZG_CH);
    return 2 * al256(nch * MP * static_cast<int64_t>(s) * OZG_CH) + al256(ozmma_partial_doubles(MP, MP, 1, static_cast<int>(nch), 0) * 8) +
           2 * al256(static_cast<int64_t>(MP) * 8);
}

int ozaki_gram(const double* Phi, int64_t ld, int MP, int m, int64_t rows, int s, const double* wgt, const double* d_scal,
               int aug, int accumulate, double* S, void* ws, cudaStream_t st, cudaStream_t aux, cudaEvent_t* ev, int64_t* launches) {
    if (s < 2 || s > OZ_MAXS) {
        set_error("ozaki_gram: unsupported digit count %d", s);
        return GPZ_ERR_USAGE;
    }
    (void)aux;
    (void)ev;
    const int nch = static_cast<int>(ceil_div(rows > 0 ? rows : 1, OZG_CH));
    unsigned char* p = static_cast<unsigned char*>(ws);
    auto take = [&](int64_t bytes) {
        unsigned char* r = p;
        p += al256(bytes);
        return r;
    };
    const int64_t slab = static_cast<int64_t>(nch) * MP * s * OZG_CH;
    int8_t* F = reinterpret_cast<int8_t*>(take(slab));
    int8_t* G = reinterpret_cast<int8_t*>(take(slab));
    double* partial = reinterpret_cast<double*>(take(ozmma_partial_doubles(MP, MP, 1, nch, 0) * 8));
    double* sr = reinterpret_cast<double*>(take(static_cast<int64_t>(MP) * 8));
    double* sc = reinterpret_cast<double*>(take(static_cast<int64_t>(MP) * 8));
    // digits (the tail of the last chunk is zero-filled by the kernel: rows beyond `rows` read as 0)
    dim3 gs(static_cast<unsigned>(static_cast<int64_t>(nch) * OZG_CH / 128), static_cast<unsigned>(MP / 32));
    ozg_slice_kernel<<<gs, 256, 0, st>>>(Phi, ld, MP, m, rows, s, wgt, d_scal, aug, F, G);
    GPZ_KERNEL_CHECK();
    ozg_scales_kernel<<<static_cast<unsigned>(ceil_div(MP, 256)), 256, 0, st>>>(d_scal, MP, m, aug, sr, sc);
    GPZ_KERNEL_CHECK();
    *launches += 2;
    const int64_t str[3] = {OZG_CH, static_cast<int64_t>(s) * OZG_CH, static_cast<int64_t>(MP) * s * OZG_CH};
    return ozmma_gemm_nt(F, str, MP, G, str, MP, s, s + 1, OZG_CH, nch, 1, partial, sr, sc, 1.0, accumulate, S, MP, 0, st, launches);
}

}  // namespace gpz
