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
        double r[4] = {j < m ? v.x * sc : 0.0, j + 1 < m ? v.y * sc : 0.0, j + 2 < m ? v.z * sc : 0.0, j + 3 < m ? v.w * sc : 0.0};
        for (int t = 0; t < s; ++t) {
            char4 q;
            double qd;
            qd = trunc(r[0] * 128.0); r[0] = r[0] * 128.0 - qd; q.x = static_cast<signed char>(qd);
            qd = trunc(r[1] * 128.0); r[1] = r[1] * 128.0 - qd; q.y = static_cast<signed char>(qd);
            qd = trunc(r[2] * 128.0); r[2] = r[2] * 128.0 - qd; q.z = static_cast<signed char>(qd);
            qd = trunc(r[3] * 128.0); r[3] = r[3] * 128.0 - qd; q.w = static_cast<signed char>(qd);
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
    b += al256(chunk_rows * static_cast<int64_t>(s) * MP);                                  // A8
    for (int e = 2; e <= s + 1; ++e) b += al256(chunk_rows * static_cast<int64_t>(MP) * 4);  // D levels
    for (int e = 2; e <= s + 1; ++e) b += al256(static_cast<int64_t>(MP) * (e - 1) * MP);    // B levels
    b += al256(chunk_rows * 8) + al256(static_cast<int64_t>(MP) * 8) + al256(64 << 20);      // scales, CUTLASS workspace
    return b;
}

// T-GEMM with fused epilogue through the int8 tensor cores.  ws: oz_workspace_bytes(MP, s, chunk_rows) bytes.
int ozaki_tgemm(const double* Phi, int64_t ld, const double* Sinv, int MP, int m, int64_t n, int s, int64_t chunk_rows,
                const double* rw, double* H, int accumulate, double* nu, const double* waug, double* pred, void* ws,
                cudaStream_t st, int64_t* launches) {
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
    int8_t* A8 = reinterpret_cast<int8_t*>(take(chunk_rows * static_cast<int64_t>(s) * MP));
    OzD Dl{};
    int32_t* Dbuf[OZ_MAXS + 2] = {nullptr};
    for (int e = 2; e <= s + 1; ++e) {
        Dbuf[e] = reinterpret_cast<int32_t*>(take(chunk_rows * static_cast<int64_t>(MP) * 4));
        Dl.D[e] = Dbuf[e];
    }
    OzLevels L{};
    for (int e = 2; e <= s + 1; ++e) L.B[e] = reinterpret_cast<int8_t*>(take(static_cast<int64_t>(MP) * (e - 1) * MP));
    double* ea = reinterpret_cast<double*>(take(chunk_rows * 8));
    double* eb = reinterpret_cast<double*>(take(static_cast<int64_t>(MP) * 8));
    void* cws = take(64 << 20);

    oz_colmax_kernel<<<MP, 256, 0, st>>>(Sinv, MP, m, waug, eb);
    dim3 g2(static_cast<unsigned>(ceil_div(MP, 256)), static_cast<unsigned>(MP));
    oz_slice_cols_kernel<<<g2, 256, 0, st>>>(Sinv, MP, m, waug, eb, s, L);
    GPZ_KERNEL_CHECK();
    *launches += 2;
    for (int64_t r0 = 0; r0 < n; r0 += chunk_rows) {
        const int64_t rows = (r0 + chunk_rows < n) ? chunk_rows : n - r0;
        oz_slice_rows_kernel<<<static_cast<unsigned>(ceil_div(rows, 8)), 256, 0, st>>>(Phi + r0 * ld, ld, m, MP, rows, s, A8, ea);
        GPZ_KERNEL_CHECK();
        ++*launches;
        for (int e = 2; e <= s + 1; ++e) {
            const int K = (e - 1) * MP;
            int rc = i8gemm_tn(A8, static_cast<int64_t>(s) * MP, L.B[e], K, Dbuf[e], MP, static_cast<int>(rows), MP, K, cws, 64 << 20, st);
            if (rc) return rc;
            ++*launches;
        }
        oz_combine_kernel<<<static_cast<unsigned>(ceil_div(rows, 8)), 256, 0, st>>>(Dl, s, ea, eb, Phi + r0 * ld, ld, MP, m, rows,
                                                                                  rw != nullptr ? rw + r0 : nullptr,
                                                                                  H != nullptr ? H + r0 * ld : nullptr, accumulate,
                                                                                  nu + r0, waug != nullptr ? m : -1,
                                                                                  pred != nullptr ? pred + r0 : nullptr);
        GPZ_KERNEL_CHECK();
        ++*launches;
    }
    return GPZ_OK;
}

}  // namespace gpz
