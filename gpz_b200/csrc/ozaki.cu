// Error-free fp64 GEMM on the int8 tensor cores (Ozaki splitting) for T = PHI * iSigma   (GPz/GPz.m:69,72).
//
// tcgen05.mma has no fp64 kind; the fp64 DMMA pipe tops out at ~36 TFLOP/s.  Here each fp64 operand is split into
// s signed 7-bit slices against a per-row (PHI) / per-column (iSigma) power-of-two scale,
//     a = 2^ea * sum_t qa_t 2^(-7t),    b = 2^eb * sum_u qb_u 2^(-7u),      |q| <= 127,
// the slice products are EXACT int8 x int8 -> int32 GEMMs (K <= 9*1024 keeps |acc| < 2^31), all pairs with the same
// level e = t+u are concatenated along K into one GEMM (i8gemm_cutlass.cu: tcgen05 + TMEM + TMA), and the levels are
// summed in fp64 smallest first.  Pairs with t+u > s+1 are dropped (below 2^(-7s) of the row/column scale).
// The combine kernel is fused with the T-GEMM epilogue of the fp64 path: nu_i = sum_j PHI_ij T_ij, H = rw_i PHI .* T,
// and the spare column m delivers PHI*w.
#include "internal.cuh"

namespace gpz {

constexpr int OZ_MAXS = 9;

// ---- PHI rows -> int8 slices.  warp per row; lane handles 4 consecutive columns per step -------------------------
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
    int ex = 0;
    if (mx > 0.0) frexp(mx, &ex);                  // mx = f * 2^ex, f in [0.5, 1)
    const double sc = ldexp(1.0, -ex);
    if (lane == 0) ea[i] = ldexp(1.0, ex);
    int8_t* out = A8 + i * static_cast<int64_t>(s) * MP;
    for (int j = lane * 4; j < MP; j += 128) {
        const double4 v = *reinterpret_cast<const double4*>(row + j);
        const double r[4] = {j < m ? v.x * sc : 0.0, j + 1 < m ? v.y * sc : 0.0, j + 2 < m ? v.z * sc : 0.0, j + 3 < m ? v.w * sc : 0.0};
        // r in (-1, 1): X = trunc(|r| 2^63), slice t = bits [56-7t, 63-7t)  (== the trunc(r*128) recurrence)
        long long X[4];
        int sg[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            X[q] = __double2ll_rz(fabs(r[q]) * 9223372036854775808.0);     // 2^63
            sg[q] = r[q] < 0.0 ? -1 : 1;
        }
        for (int t = 0; t < s; ++t) {
            char4 q;
            const int sh = 56 - 7 * t;
            q.x = static_cast<signed char>(sg[0] * static_cast<int>((X[0] >> sh) & 127));
            q.y = static_cast<signed char>(sg[1] * static_cast<int>((X[1] >> sh) & 127));
            q.z = static_cast<signed char>(sg[2] * static_cast<int>((X[2] >> sh) & 127));
            q.w = static_cast<signed char>(sg[3] * static_cast<int>((X[3] >> sh) & 127));
            *reinterpret_cast<char4*>(out + static_cast<int64_t>(t) * MP + j) = q;
        }
    }
}

// ---- iSigma columns -> per-level concatenated int8 B operands ---------------------------------------------------
// Bcat level e (2..s+1), stored N x K row-major with K = (e-1)*MP:  B_e[j][(t-1)*MP + l] = qb_{e-t}[l][j]
// column scale from max_l |B[l][j]|; column m (aug) is w, every other column j uses the symmetric iSigma[j][l]
__global__ void __launch_bounds__(256)
oz_colmax_kernel(const double* __restrict__ Sinv, int MP, int m, const double* __restrict__ waug, double* __restrict__ eb) {
    __shared__ double sh[8];
    const int j = blockIdx.x;
    double mx = 0.0;
    if (j < m)
        for (int l = threadIdx.x; l < m; l += 256) mx = fmax(mx, fabs(Sinv[static_cast<int64_t>(j) * MP + l]));
    else if (j == m && waug != nullptr)
        for (int l = threadIdx.x; l < m; l += 256) mx = fmax(mx, fabs(waug[l]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) mx = fmax(mx, sh[q]);
        int ex = 0;
        if (mx > 0.0) frexp(mx, &ex);
        eb[j] = ldexp(1.0, ex);
    }
}

struct OzLevels {
    int8_t* B[OZ_MAXS + 2];      // B[e] for e = 2..s+1
};

__global__ void __launch_bounds__(256)
oz_slice_cols_kernel(const double* __restrict__ Sinv, int MP, int m, const double* __restrict__ waug,
                     const double* __restrict__ eb, int s, OzLevels L) {
    const int l = blockIdx.x * 256 + threadIdx.x;      // K index (row of iSigma)
    const int j = blockIdx.y;                          // column
    if (l >= MP) return;
    double v = 0.0;
    if (l < m) {
        if (j < m) v = Sinv[static_cast<int64_t>(j) * MP + l];
        else if (j == m && waug != nullptr) v = waug[l];
    }
    double r = v / eb[j];
    for (int u = 1; u <= s; ++u) {
        const double qd = trunc(r * 128.0);
        r = r * 128.0 - qd;
        const int8_t q = static_cast<int8_t>(qd);
        for (int e = u + 1; e <= s + 1; ++e) {          // pairs (t = e-u, u), t >= 1
            const int t = e - u;
            L.B[e][static_cast<int64_t>(j) * (static_cast<int64_t>(e - 1) * MP) + static_cast<int64_t>(t - 1) * MP + l] = q;
        }
    }
}

// ---- combine the levels in fp64 + the T-GEMM epilogue.  warp per row ---------------------------------------------
struct OzD {
    const int32_t* D[OZ_MAXS + 2];   // D[e], e = 2..s+1, each [rows][MP]
};

__global__ void __launch_bounds__(256)
oz_combine_kernel(OzD Dl, int s, const double* __restrict__ ea, const double* __restrict__ eb, const double* __restrict__ Phi,
                  int64_t ld, int MP, int m, int64_t n, const double* __restrict__ rw, double* __restrict__ H, int accumulate,
                  double* __restrict__ nu, int aug_col, double* __restrict__ pred) {
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const double sa = ea[i];
    const double wrow = rw != nullptr ? rw[i] : 1.0;
    double rs = 0.0;
    for (int j = lane * 2; j < MP; j += 64) {
        double t0 = 0.0, t1 = 0.0;
        double wgt = ldexp(1.0, -7 * (s + 1));
        for (int e = s + 1; e >= 2; --e) {              // smallest level first
            const int2 d = *reinterpret_cast<const int2*>(Dl.D[e] + i * MP + j);
            t0 = fma(static_cast<double>(d.x), wgt, t0);
            t1 = fma(static_cast<double>(d.y), wgt, t1);
            wgt *= 128.0;
        }
        const double2 sb = *reinterpret_cast<const double2*>(eb + j);
        t0 *= sa * sb.x;
        t1 *= sa * sb.y;
        const double2 ph = *reinterpret_cast<const double2*>(Phi + i * ld + j);
        double h0 = ph.x * t0, h1 = ph.y * t1;
        if (aug_col >= 0) {
            if (j == aug_col) { pred[i] = t0; h0 = 0.0; }
            if (j + 1 == aug_col) { pred[i] = t1; h1 = 0.0; }
        }
        rs += h0 + h1;
        if (H != nullptr) {
            double2* hp = reinterpret_cast<double2*>(H + i * ld + j);
            double2 v = make_double2(wrow * h0, wrow * h1);
            if (accumulate) {
                const double2 o = *hp;
                v.x += o.x;
                v.y += o.y;
            }
            *hp = v;
        }
    }
    rs = warp_sum(rs);
    if (lane == 0) nu[i] = rs;
}

static int64_t al256(int64_t b) { return (b + 255) / 256 * 256; }

int64_t oz_workspace_bytes(int MP, int s, int64_t chunk_rows) {
    int64_t b = 0;
    b += 2 * al256(chunk_rows * static_cast<int64_t>(s) * MP);                                  // A8 (double-buffered)
    for (int e = 2; e <= s + 1; ++e) b += 2 * al256(chunk_rows * static_cast<int64_t>(MP) * 4);  // D levels (double-buffered)
    for (int e = 2; e <= s + 1; ++e) b += al256(static_cast<int64_t>(MP) * (e - 1) * MP);        // B levels
    b += 2 * al256(chunk_rows * 8) + al256(static_cast<int64_t>(MP) * 8) + al256(64 << 20);      // scales, CUTLASS workspace
    return b;
}

// T-GEMM with fused epilogue through the int8 tensor cores.  ws: oz_workspace_bytes(MP, s, chunk_rows) bytes.
// Row chunks are software-pipelined over two streams: the int8 GEMMs of chunk c (tensor-pipe bound, stream st) overlap
// the slicing of chunk c+1 and the fp64 combine/epilogue of chunk c-1 (both HBM bound, stream aux).  ev: 6 events.
int ozaki_tgemm(const double* Phi, int64_t ld, const double* Sinv, int MP, int m, int64_t n, int s, int64_t chunk_rows,
                const double* rw, double* H, int accumulate, double* nu, const double* waug, double* pred, void* ws,
                cudaStream_t st, cudaStream_t aux, cudaEvent_t* ev, cudaEvent_t tev0, cudaEvent_t tev1, int64_t* launches) {
    if (s < 2 || s > OZ_MAXS) {
        set_error("ozaki_tgemm: slices must be in [2, %d]", OZ_MAXS);
        return GPZ_ERR_USAGE;
    }
    if (static_cast<int64_t>(s) * MP * 127 * 127 >= 2147483647LL) {
        set_error("ozaki_tgemm: m too large for exact int32 accumulation");
        return GPZ_ERR_USAGE;
    }
    unsigned char* p = static_cast<unsigned char*>(ws);
    auto take = [&](int64_t bytes) {
        unsigned char* r = p;
        p += al256(bytes);
        return r;
    };
    int8_t* A8[2];
    OzD Dl[2] = {};
    int32_t* Dbuf[2][OZ_MAXS + 2] = {};
    double* ea[2];
    for (int b = 0; b < 2; ++b) {
        A8[b] = reinterpret_cast<int8_t*>(take(chunk_rows * static_cast<int64_t>(s) * MP));
        for (int e = 2; e <= s + 1; ++e) {
            Dbuf[b][e] = reinterpret_cast<int32_t*>(take(chunk_rows * static_cast<int64_t>(MP) * 4));
            Dl[b].D[e] = Dbuf[b][e];
        }
        ea[b] = reinterpret_cast<double*>(take(chunk_rows * 8));
    }
    OzLevels L{};
    for (int e = 2; e <= s + 1; ++e) L.B[e] = reinterpret_cast<int8_t*>(take(static_cast<int64_t>(MP) * (e - 1) * MP));
    double* eb = reinterpret_cast<double*>(take(static_cast<int64_t>(MP) * 8));
    void* cws = take(64 << 20);
    cudaEvent_t* evS = ev;          // [2] slices of buffer b ready
    cudaEvent_t* evG = ev + 2;      // [2] GEMMs reading A8[b] / writing D[b] done
    cudaEvent_t* evC = ev + 4;      // [2] combine reading D[b] done

    oz_colmax_kernel<<<MP, 256, 0, st>>>(Sinv, MP, m, waug, eb);
    dim3 g2(static_cast<unsigned>(ceil_div(MP, 256)), static_cast<unsigned>(MP));
    oz_slice_cols_kernel<<<g2, 256, 0, st>>>(Sinv, MP, m, waug, eb, s, L);
    GPZ_KERNEL_CHECK();
    *launches += 2;
    const int nchunks = static_cast<int>(ceil_div(n, chunk_rows));
    auto rows_of = [&](int c) { return (static_cast<int64_t>(c + 1) * chunk_rows <= n) ? chunk_rows : n - static_cast<int64_t>(c) * chunk_rows; };
    auto slice = [&](int c, cudaStream_t sx) {
        const int64_t r0 = static_cast<int64_t>(c) * chunk_rows;
        oz_slice_rows_kernel<<<static_cast<unsigned>(ceil_div(rows_of(c), 8)), 256, 0, sx>>>(Phi + r0 * ld, ld, m, MP, rows_of(c), s,
                                                                                             A8[c & 1], ea[c & 1]);
        ++*launches;
    };
    // everything enqueued so far on st (PHI, iSigma, B levels) must be visible to aux
    GPZ_CUDA(cudaEventRecord(evS[1], st));
    GPZ_CUDA(cudaStreamWaitEvent(aux, evS[1], 0));
    slice(0, st);
    GPZ_CUDA(cudaEventRecord(evS[0], st));
    for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1;
        const int64_t r0 = static_cast<int64_t>(c) * chunk_rows;
        const int64_t rows = rows_of(c);
        if (c + 1 < nchunks) {                       // slice chunk c+1 on aux while the GEMMs of chunk c run
            if (c >= 1) GPZ_CUDA(cudaStreamWaitEvent(aux, evG[b ^ 1], 0));      // A8[b^1] was read by the GEMMs of chunk c-1
            slice(c + 1, aux);
            GPZ_CUDA(cudaEventRecord(evS[b ^ 1], aux));
        }
        GPZ_CUDA(cudaStreamWaitEvent(st, evS[b], 0));
        if (c >= 2) GPZ_CUDA(cudaStreamWaitEvent(st, evC[b], 0));               // D[b] was read by the combine of chunk c-2
        if (c == 0 && tev0) GPZ_CUDA(cudaEventRecord(tev0, st));
        for (int e = 2; e <= s + 1; ++e) {
            const int K = (e - 1) * MP;
            int rc = i8gemm_tn(A8[b], static_cast<int64_t>(s) * MP, L.B[e], K, Dbuf[b][e], MP, static_cast<int>(rows), MP, K, cws, 64 << 20, st);
            if (rc) return rc;
            ++*launches;
        }
        if (c == 0 && tev1) GPZ_CUDA(cudaEventRecord(tev1, st));
        GPZ_CUDA(cudaEventRecord(evG[b], st));
        GPZ_CUDA(cudaStreamWaitEvent(aux, evG[b], 0));
        oz_combine_kernel<<<static_cast<unsigned>(ceil_div(rows, 8)), 256, 0, aux>>>(Dl[b], s, ea[b], eb, Phi + r0 * ld, ld, MP, m, rows,
                                                                                   rw != nullptr ? rw + r0 : nullptr,
                                                                                   H != nullptr ? H + r0 * ld : nullptr, accumulate,
                                                                                   nu + r0, waug != nullptr ? m : -1,
                                                                                   pred != nullptr ? pred + r0 : nullptr);
        GPZ_KERNEL_CHECK();
        ++*launches;
        GPZ_CUDA(cudaEventRecord(evC[b], aux));
    }
    GPZ_CUDA(cudaStreamWaitEvent(st, evC[0], 0));
    GPZ_CUDA(cudaStreamWaitEvent(st, evC[1], 0));
    return GPZ_OK;
}

}  // namespace gpz

// ================================================================================================
// Gram  S = PHI' diag(w) PHI  (GPz/GPz.m:63-65) through the int8 tensor cores.
//   A side = (w .* PHI)' , B side = PHI ; contraction over the rows i.  Rows are cut into chunks of OZG_CH rows
//   (so that (e-1)*OZG_CH*127^2 < 2^31 keeps every int32 accumulator exact) and the chunks are the batch dimension
//   of the GEMM.  Slices are stored transposed, chunk-major:
//     F[c][j][t][i]   forward slice order   (A operand; level e uses slices 0..e-2)
//     R[c][j][s-1-t][i] reversed slice order (B operand; level e uses the suffix starting at slot s-e+1),
//   which makes "all pairs with t+u = e" one GEMM with K = (e-1)*OZG_CH.  Only the block columns on or below the
//   diagonal are computed (256-wide), the result is mirrored.  Scales are fixed powers of two: PHI <= 1 -> 2,
//   the weights -> 2^ceil(log2 max w), the spare column (y) -> 2^ceil(log2 max|y|)+1.
// ================================================================================================
namespace gpz {

constexpr int OZG_CH = 16384;     // rows per chunk: s * OZG_CH * 127^2 < 2^31 for s <= 8 (s = 9 is refused for the Gram)

__device__ __forceinline__ double pow2_ceil(double v) {
    int ex = 0;
    if (v > 0.0) frexp(v, &ex);
    return ldexp(1.0, ex);                      // v < 2^ex
}

// tile: 128 rows (i) x 32 columns (j) of PHI -> transposed int8 slices.  Small footprint on purpose (34 KB smem, few
// registers: 6 CTAs per SM): a thread owns 4 consecutive rows of a column and emits one char4 per slice, a warp writes 128
// contiguous bytes of one (column, slice) row.
__global__ void __launch_bounds__(256)
ozg_slice_kernel(const double* __restrict__ Phi, int64_t ld, int MP, int m, int64_t rows, int s, const double* __restrict__ wgt,
                 const double* __restrict__ scal, int aug, int8_t* __restrict__ F, int8_t* __restrict__ R) {
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
    const double sw = 1.0 / pow2_ceil(scal[0]);                    // weight scale
    const double sy = 1.0 / (2.0 * pow2_ceil(scal[1]));            // spare-column scale
    const int64_t c = i0 / OZG_CH;                                 // chunk (128 divides OZG_CH)
    const int il = static_cast<int>(i0 % OZG_CH);
    const int ig = (tid & 31) * 4;                                 // 4 consecutive rows per thread
    for (int jj = tid >> 5; jj < 32; jj += 8) {
        const int j = j0 + jj;
        const double sp = (j < m) ? 0.5 : ((j == m && aug) ? sy : 0.0);
        long long XA[4], XB[4];
        int sg[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {                              // X = trunc(|r| 2^63): slice t = bits [56-7t, 63-7t)
            const double rb = tile[ig + q][jj] * sp;
            const double ra = rb * (wsm[ig + q] * sw);
            sg[q] = rb < 0.0 ? -1 : 1;
            XB[q] = __double2ll_rz(fabs(rb) * 9223372036854775808.0);
            XA[q] = __double2ll_rz(fabs(ra) * 9223372036854775808.0);
        }
        int8_t* fo = F + ((c * MP + j) * static_cast<int64_t>(s)) * OZG_CH + il + ig;
        int8_t* ro = R + ((c * MP + j) * static_cast<int64_t>(s)) * OZG_CH + il + ig;
        for (int t = 0; t < s; ++t) {
            const int sh = 56 - 7 * t;
            char4 qa, qb;
            qa.x = static_cast<signed char>(sg[0] * static_cast<int>((XA[0] >> sh) & 127));
            qa.y = static_cast<signed char>(sg[1] * static_cast<int>((XA[1] >> sh) & 127));
            qa.z = static_cast<signed char>(sg[2] * static_cast<int>((XA[2] >> sh) & 127));
            qa.w = static_cast<signed char>(sg[3] * static_cast<int>((XA[3] >> sh) & 127));
            qb.x = static_cast<signed char>(sg[0] * static_cast<int>((XB[0] >> sh) & 127));
            qb.y = static_cast<signed char>(sg[1] * static_cast<int>((XB[1] >> sh) & 127));
            qb.z = static_cast<signed char>(sg[2] * static_cast<int>((XB[2] >> sh) & 127));
            qb.w = static_cast<signed char>(sg[3] * static_cast<int>((XB[3] >> sh) & 127));
            *reinterpret_cast<char4*>(fo + static_cast<int64_t>(t) * OZG_CH) = qa;
            *reinterpret_cast<char4*>(ro + static_cast<int64_t>(s - 1 - t) * OZG_CH) = qb;
        }
    }
}

// S[j][l] (+)= sA_j sB_l sum_c sum_e 2^(-7e) D[e][c][j][l]   for block columns on/below the diagonal; mirrored
__global__ void __launch_bounds__(256)
ozg_combine_kernel(OzD Dl, int s, int nchunks, int MP, int m, const double* __restrict__ scal, int aug, int accumulate,
                   double* __restrict__ S) {
    const int l = blockIdx.x * 256 + threadIdx.x;
    const int j = blockIdx.y;
    if (l >= MP) return;
    if ((l >> 8) > (j >> 8)) return;                    // block column above the diagonal block: not computed
    const double cw = pow2_ceil(scal[0]);
    const double cy = 2.0 * pow2_ceil(scal[1]);
    const double sj = (j < m) ? 2.0 : ((j == m && aug) ? cy : 0.0);
    const double sl = (l < m) ? 2.0 : ((l == m && aug) ? cy : 0.0);
    double acc = 0.0;
    double wgt = ldexp(1.0, -7 * (s + 1));
    const int64_t off = static_cast<int64_t>(j) * MP + l;
    for (int e = s + 1; e >= 2; --e) {
        double lev = 0.0;
        const int32_t* D = Dl.D[e];
        for (int c = 0; c < nchunks; ++c) lev += static_cast<double>(D[static_cast<int64_t>(c) * MP * MP + off]);   // exact up to 2^53
        acc = fma(lev, wgt, acc);
        wgt *= 128.0;
    }
    acc *= cw * sj * sl;
    const bool lower = l <= j;
    if (lower) {
        S[off] = (accumulate ? S[off] : 0.0) + acc;
        if (l != j) {
            const int64_t offT = static_cast<int64_t>(l) * MP + j;
            S[offT] = (accumulate ? S[offT] : 0.0) + acc;
        }
    }
}

int64_t oz_gram_workspace_bytes(int MP, int s, int64_t rows) {
    const int64_t nch = ceil_div(rows > 0 ? rows : 1, OZG_CH);
    return 2 * al256(nch * MP * static_cast<int64_t>(s) * OZG_CH) + s * al256(nch * static_cast<int64_t>(MP) * MP * 4) + al256(64 << 20);
}

int ozaki_gram(const double* Phi, int64_t ld, int MP, int m, int64_t rows, int s, const double* wgt, const double* d_scal,
               int aug, int accumulate, double* S, void* ws, cudaStream_t st, cudaStream_t aux, cudaEvent_t* ev, int64_t* launches) {
    if (s < 2 || s > OZ_MAXS || static_cast<int64_t>(s) * OZG_CH * 127 * 127 >= 2147483647LL) {
        set_error("ozaki_gram: unsupported slice count %d", s);
        return GPZ_ERR_USAGE;
    }
    const int nch = static_cast<int>(ceil_div(rows > 0 ? rows : 1, OZG_CH));
    unsigned char* p = static_cast<unsigned char*>(ws);
    auto take = [&](int64_t bytes) {
        unsigned char* r = p;
        p += al256(bytes);
        return r;
    };
    const int64_t slab = static_cast<int64_t>(nch) * MP * s * OZG_CH;
    int8_t* F = reinterpret_cast<int8_t*>(take(slab));
    int8_t* R = reinterpret_cast<int8_t*>(take(slab));
    OzD Dl{};
    int32_t* Dbuf[OZ_MAXS + 2] = {nullptr};
    for (int e = 2; e <= s + 1; ++e) {
        Dbuf[e] = reinterpret_cast<int32_t*>(take(static_cast<int64_t>(nch) * MP * MP * 4));
        Dl.D[e] = Dbuf[e];
    }
    void* cws = take(64 << 20);
    const int64_t rowstride = static_cast<int64_t>(s) * OZG_CH;          // bytes between consecutive j
    const int64_t bstride = static_cast<int64_t>(MP) * rowstride;        // bytes between chunks
    // The chunks are processed in up to 4 groups: the HBM-bound slicing of groups 1.. (stream aux) runs under the int8
    // GEMMs of the earlier groups (stream st).  Every chunk has its own slice and D storage, so a group only needs its
    // "slices ready" event.  ev: >= 4 events.
    const int ngroups = nch >= 8 ? 4 : 1;
    const int cpg = static_cast<int>(ceil_div(nch, ngroups));
    auto slice_group = [&](int gidx, cudaStream_t ss) {
        const int c0 = gidx * cpg;
        const int c1 = (c0 + cpg < nch) ? c0 + cpg : nch;
        if (c1 <= c0) return;
        const int64_t r0 = static_cast<int64_t>(c0) * OZG_CH;
        int64_t rcount = (static_cast<int64_t>(c1) * OZG_CH < rows ? static_cast<int64_t>(c1) * OZG_CH : rows) - r0;
        if (rcount < 0) rcount = 0;
        dim3 gs(static_cast<unsigned>(static_cast<int64_t>(c1 - c0) * OZG_CH / 128), static_cast<unsigned>(MP / 32));
        ozg_slice_kernel<<<gs, 256, 0, ss>>>(Phi + r0 * ld, ld, MP, m, rcount, s, wgt + r0, d_scal, aug,
                                             F + static_cast<int64_t>(c0) * bstride, R + static_cast<int64_t>(c0) * bstride);
        ++*launches;
    };
    if (ngroups > 1) {
        GPZ_CUDA(cudaEventRecord(ev[0], st));                            // PHI / weights / scales are ready
        GPZ_CUDA(cudaStreamWaitEvent(aux, ev[0], 0));
        for (int gidx = 1; gidx < ngroups; ++gidx) {
            slice_group(gidx, aux);
            GPZ_CUDA(cudaEventRecord(ev[gidx], aux));
        }
    }
    slice_group(0, st);
    GPZ_KERNEL_CHECK();
    for (int gidx = 0; gidx < ngroups; ++gidx) {
        const int c0 = gidx * cpg;
        const int c1 = (c0 + cpg < nch) ? c0 + cpg : nch;
        if (c1 <= c0) break;
        if (gidx > 0) GPZ_CUDA(cudaStreamWaitEvent(st, ev[gidx], 0));
        for (int e = 2; e <= s + 1; ++e) {
            const int K = (e - 1) * OZG_CH;
            for (int J = 0; J * 256 < MP; ++J) {
                const int n0 = J * 256;
                const int N = (MP - n0 < 256) ? (MP - n0) : 256;
                const int M = MP - n0;
                int rc = i8gemm_tn_batched(F + static_cast<int64_t>(c0) * bstride + static_cast<int64_t>(n0) * rowstride, rowstride, bstride,
                                           R + static_cast<int64_t>(c0) * bstride + static_cast<int64_t>(n0) * rowstride +
                                               static_cast<int64_t>(s - e + 1) * OZG_CH,
                                           rowstride, bstride,
                                           Dbuf[e] + static_cast<int64_t>(c0) * MP * MP + static_cast<int64_t>(n0) * MP + n0, MP,
                                           static_cast<int64_t>(MP) * MP, M, N, K, c1 - c0, cws, 64 << 20, st);
                if (rc) return rc;
                ++*launches;
            }
        }
    }
    dim3 gc(static_cast<unsigned>(ceil_div(MP, 256)), static_cast<unsigned>(MP));
    ozg_combine_kernel<<<gc, 256, 0, st>>>(Dl, s, nch, MP, m, d_scal, aug, accumulate, S);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

}  // namespace gpz
