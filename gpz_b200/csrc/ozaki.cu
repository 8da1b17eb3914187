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

__device__ __forceinline__ double pow2_ceil(double v) {
    int ex = 0;
    if (v > 0.0) frexp(v, &ex);
    return ldexp(1.0, ex);                      // v < 2^ex
}

// All s balanced digits of I at once: adding the bias 0x80 to every byte below the leading one turns the signed-digit
// expansion into the plain unsigned byte expansion (d_k = u_k - 128), and u - 128 as a two's-complement byte is u ^ 0x80;
// the leading byte is the leading digit itself.  Byte k of the result is the digit of weight 256^k.
__device__ __forceinline__ unsigned long long oz_digit_bytes(long long I, unsigned long long bias) {
    return (static_cast<unsigned long long>(I) + bias) ^ bias;
}
// digit k (weight 256^k) of 4 neighbouring elements -> one 32-bit store: a 4 x 4 byte transpose in 8 PRMT
__device__ __forceinline__ void oz_transpose4(unsigned w0, unsigned w1, unsigned w2, unsigned w3, unsigned (&o)[4]) {
    const unsigned a = __byte_perm(w0, w1, 0x5140), b = __byte_perm(w0, w1, 0x7362);
    const unsigned c = __byte_perm(w2, w3, 0x5140), d = __byte_perm(w2, w3, 0x7362);
    o[0] = __byte_perm(a, c, 0x5410);
    o[1] = __byte_perm(a, c, 0x7632);
    o[2] = __byte_perm(b, d, 0x5410);
    o[3] = __byte_perm(b, d, 0x7632);
}
// store the s digit rows of 4 consecutive columns (K[q] from oz_digit_bytes): digit t (0 = leading) is byte s-1-t
__device__ __forceinline__ void oz_store_digits4(const unsigned long long (&K)[4], int s, int8_t* out, int64_t digit_stride) {
    unsigned lo[4], hi[4];
    oz_transpose4(static_cast<unsigned>(K[0]), static_cast<unsigned>(K[1]), static_cast<unsigned>(K[2]), static_cast<unsigned>(K[3]), lo);
    oz_transpose4(static_cast<unsigned>(K[0] >> 32), static_cast<unsigned>(K[1] >> 32), static_cast<unsigned>(K[2] >> 32),
                  static_cast<unsigned>(K[3] >> 32), hi);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k >= s) break;
        *reinterpret_cast<unsigned*>(out + static_cast<int64_t>(s - 1 - k) * digit_stride) = k < 4 ? lo[k] : hi[k - 4];
    }
}

// ---- one pass over the PHI rows -> both digit sets, row-major [i][t][j].  warp per row; lane handles 4 columns per step.
//   D8: plain digits of PHI_ij 2^-E_i (per-row exponent E_i from the row maximum, ea[i] = 2^(E_i - 8)); columns >= m are 0.
//       They are the A operand of T = PHI iSigma (K = j contiguous) AND the B operand of the Gram (MN-major, K = i).
//   F8: digits of w_i 2^E_i PHI_ij 2^-Ef with the fixed exponent 2^Ef = 16 2^ceil(log2 max w) -- the A operand of the Gram:
//       sum_i F_ij D_il = 2^-Ef sum_i w_i PHI_ij PHI_il, the row scale cancels inside the product.  Column m (aug) carries
//       w_i 2^E_i y_i / (2^Ef cy), cy = 2^ceil(log2 max|y|), so that row m of the Gram is PHI'(w y)  (GPz/GPz.m:70).
//   Rows in [n, n_pad) are zero-filled (the Gram contracts over whole 1024-row chunks).
// S > 0: the digit count is a compile-time constant (the store loops and the bias fold); S == 0 reads it from `s`.
// KEEP: rows of <= 1024 columns stay in registers between the two passes (row maximum, then digits); the second pass used to
// re-read the row, and 56 % of that came from HBM again (profiles/r01h: 12.9 GB read for an 8.2 GB input).
// A double4 that lies wholly inside [0, m) takes the mask-free path; the column tail and the y column take the general one.
// ncu before this form (profiles/r02z_aux.md): 94 instructions per element, ALU pipe 60 % / issue 58 % active against DRAM at 60 %.
template <int S>
__device__ __forceinline__ void oz_store_digits4s(const unsigned long long (&K)[4], int s, int8_t* out, int64_t digit_stride) {
    unsigned lo[4], hi[4];
    oz_transpose4(static_cast<unsigned>(K[0]), static_cast<unsigned>(K[1]), static_cast<unsigned>(K[2]), static_cast<unsigned>(K[3]), lo);
    if (S == 0 || S > 4)
        oz_transpose4(static_cast<unsigned>(K[0] >> 32), static_cast<unsigned>(K[1] >> 32), static_cast<unsigned>(K[2] >> 32),
                      static_cast<unsigned>(K[3] >> 32), hi);
    const int ns = S > 0 ? S : s;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (S > 0 ? k >= S : k >= s) break;
        *reinterpret_cast<unsigned*>(out + static_cast<int64_t>(ns - 1 - k) * digit_stride) = k < 4 ? lo[k] : hi[k - 4];
    }
}

template <int S, bool KEEP>
__global__ void __launch_bounds__(256)
oz_digits_kernel(const double* __restrict__ Phi, int64_t ld, int src_cols, int m, int MP, int64_t n, int64_t n_pad, int s_rt,
                 const double* __restrict__ wgt,
                 const double* __restrict__ scal, int aug, int8_t* __restrict__ D8, int8_t* __restrict__ F8, double* __restrict__ ea,
                 int* __restrict__ flag) {
    const int s = S > 0 ? S : s_rt;
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (i >= n_pad) return;
    int8_t* dout = D8 + i * static_cast<int64_t>(s) * MP;
    int8_t* fout = F8 != nullptr ? F8 + i * static_cast<int64_t>(s) * MP : nullptr;
    if (i >= n) {
        const int4 z = make_int4(0, 0, 0, 0);
        for (int b = lane * 16; b < s * MP; b += 512) {
            *reinterpret_cast<int4*>(dout + b) = z;
            if (fout != nullptr) *reinterpret_cast<int4*>(fout + b) = z;
        }
        return;
    }
    const double* row = Phi + i * ld;
    double mx = 0.0;
    bool bad = false;                                   // NaN / Inf cannot be expressed in digits: raise the failure flag instead
    const double4 zero4 = make_double4(0.0, 0.0, 0.0, 0.0);
    constexpr int NK = KEEP ? 8 : 1;
    double4 vb[NK];
    if (KEEP) {
#pragma unroll
        for (int it = 0; it < NK; ++it) {
            const int j = lane * 4 + it * 128;
            vb[it] = (j < MP && j < src_cols) ? *reinterpret_cast<const double4*>(row + j) : zero4;
        }
    }
    const int lim = (aug && fout != nullptr) ? m + 1 : m;
    auto scan = [&](const double4& v, int j) {
        if (j + 4 <= m) {                               // interior: one finite test on the sum of magnitudes, no masks
            const double a0 = fabs(v.x), a1 = fabs(v.y), a2 = fabs(v.z), a3 = fabs(v.w);
            const double m01 = fmax(a0, a1), m23 = fmax(a2, a3);
            const double sum = (a0 + a1) + (a2 + a3), mq = fmax(m01, m23);
            bad |= (sum != sum) || !(mq <= 1.7e308);           // a NaN survives the sum (fmax drops it), an Inf the maximum
            mx = fmax(mx, mq);
            return;
        }
        if (j < lim) bad |= !(fabs(v.x) <= 1.7e308);
        if (j + 1 < lim) bad |= !(fabs(v.y) <= 1.7e308);
        if (j + 2 < lim) bad |= !(fabs(v.z) <= 1.7e308);
        if (j + 3 < lim) bad |= !(fabs(v.w) <= 1.7e308);
        if (j < m) mx = fmax(mx, fabs(v.x));
        if (j + 1 < m) mx = fmax(mx, fabs(v.y));
        if (j + 2 < m) mx = fmax(mx, fabs(v.z));
        if (j + 3 < m) mx = fmax(mx, fabs(v.w));
    };
    if (KEEP) {
#pragma unroll
        for (int it = 0; it < NK; ++it)
            if (lane * 4 + it * 128 < MP) scan(vb[it], lane * 4 + it * 128);
    } else {
        for (int j = lane * 4; j < MP; j += 128) scan(j < src_cols ? *reinterpret_cast<const double4*>(row + j) : zero4, j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (fout != nullptr) bad |= !(fabs(wgt[i]) <= 1.7e308);
    if (__any_sync(0xffffffffu, bad) && lane == 0 && flag != nullptr) atomicExch(flag, 1);
    const int E = oz_exponent(mx);
    if (lane == 0) ea[i] = ldexp(1.0, E - 8);          // PHI_ij = ea_i * sum_t d_t 256^-(t-1)
    const double sc = ldexp(1.0, 8 * s - E);
    const unsigned long long bias = 0x0080808080808080ull >> (8 * (8 - s));      // 0x80 in bytes 0 .. s-2
    double fs = 0.0, fy = 0.0;
    if (fout != nullptr) {
        fs = wgt[i] * ldexp(1.0, 8 * s + E - 4) / pow2_ceil(scal[0]);       // w_i 2^E_i 2^-Ef 2^(8s)
        fy = fs / pow2_ceil(scal[1]);
    }
    auto emit = [&](const double4& v, int j) {
        const double x[4] = {v.x, v.y, v.z, v.w};
        unsigned long long I[4], J[4];
        if (j + 4 <= m) {
#pragma unroll
            for (int q = 0; q < 4; ++q) I[q] = oz_digit_bytes(__double2ll_rn(x[q] * sc), bias);
            oz_store_digits4s<S>(I, s, dout + j, MP);
            if (fout != nullptr) {
#pragma unroll
                for (int q = 0; q < 4; ++q) J[q] = oz_digit_bytes(__double2ll_rn(x[q] * fs), bias);
                oz_store_digits4s<S>(J, s, fout + j, MP);
            }
            return;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool in = j + q < m;
            I[q] = oz_digit_bytes(in ? __double2ll_rn(x[q] * sc) : 0, bias);
            J[q] = oz_digit_bytes(in ? __double2ll_rn(x[q] * fs) : ((aug && j + q == m) ? __double2ll_rn(x[q] * fy) : 0), bias);
        }
        oz_store_digits4s<S>(I, s, dout + j, MP);
        if (fout != nullptr) oz_store_digits4s<S>(J, s, fout + j, MP);
    };
    if (KEEP) {
#pragma unroll
        for (int it = 0; it < NK; ++it)
            if (lane * 4 + it * 128 < MP) emit(vb[it], lane * 4 + it * 128);
    } else {
        for (int j = lane * 4; j < MP; j += 128) emit(j < src_cols ? *reinterpret_cast<const double4*>(row + j) : zero4, j);
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

#ifndef GPZ_OZG_CH
#define GPZ_OZG_CH 1024
#endif
constexpr int OZG_CH = GPZ_OZG_CH;      // rows per K chunk of the Gram: the digits of the ~4 chunks in flight (x all tiles) stay in L2

int64_t oz_padded_rows(int64_t rows) { return round_up(rows > 0 ? rows : 1, OZG_CH); }

// bytes of ONE digit set for `rows` rows (D8 or F8)
int64_t oz_digit_bytes(int MP, int s, int64_t rows) { return al256(oz_padded_rows(rows) * static_cast<int64_t>(s) * MP); }

static void launch_oz_digits(const double* A, int64_t ld, int src_cols, int m, int MP, int64_t rows, int64_t np, int s, const double* wgt,
                             const double* d_scal, int aug, int8_t* D8, int8_t* F8, double* ea, int* flag, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>(ceil_div(np, 8));
    const bool keep = MP <= 1024;
    if (s == 7 && keep) oz_digits_kernel<7, true><<<grid, 256, 0, st>>>(A, ld, src_cols, m, MP, rows, np, s, wgt, d_scal, aug, D8, F8, ea, flag);
    else if (s == 7) oz_digits_kernel<7, false><<<grid, 256, 0, st>>>(A, ld, src_cols, m, MP, rows, np, s, wgt, d_scal, aug, D8, F8, ea, flag);
    else if (keep) oz_digits_kernel<0, true><<<grid, 256, 0, st>>>(A, ld, src_cols, m, MP, rows, np, s, wgt, d_scal, aug, D8, F8, ea, flag);
    else oz_digits_kernel<0, false><<<grid, 256, 0, st>>>(A, ld, src_cols, m, MP, rows, np, s, wgt, d_scal, aug, D8, F8, ea, flag);
}

// PHI rows -> D8 (and F8 when wgt != nullptr) + ea.  d_scal: [0] >= max w, [1] >= max |y| (device).
int ozaki_digits(const double* Phi, int64_t ld, int MP, int m, int64_t rows, int s, const double* wgt, const double* d_scal, int aug,
                 int8_t* D8, int8_t* F8, double* ea, int* flag, cudaStream_t st, int64_t* launches) {
    if (s < 2 || s > OZ_MAXS) {
        set_error("ozaki: digits per operand must be in [2, %d]", OZ_MAXS);
        return GPZ_ERR_USAGE;
    }
    const int64_t np = oz_padded_rows(rows);
    launch_oz_digits(Phi, ld, MP, m, MP, rows, np, s, wgt, d_scal, aug, D8, wgt != nullptr ? F8 : nullptr, ea, flag, st);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

int64_t oz_workspace_bytes(int MP, int s) { return al256(static_cast<int64_t>(MP) * s * MP) + al256(static_cast<int64_t>(MP) * 8); }

// T = PHI iSigma with the fused epilogue (nu partials, H, PHI w) through the int8 tensor cores: one tcgen05 launch over all
// rows.  D8 / ea: from ozaki_digits.  ws: oz_workspace_bytes.  nupart: [MP/128][nu_ld] row-sum partials of PHI .* T.
int ozaki_tgemm(const double* Phi, int64_t ld, const int8_t* D8, const double* ea, const double* Sinv, int MP, int m, int64_t n, int s,
                const double* rw, double* H, int accumulate, double* nupart, int64_t nu_ld, const double* waug, double* pred, void* ws,
                cudaStream_t st, cudaEvent_t tev0, cudaEvent_t tev1, int64_t* launches) {
    if (s < 2 || s > OZ_MAXS) {
        set_error("ozaki_tgemm: digits must be in [2, %d]", OZ_MAXS);
        return GPZ_ERR_USAGE;
    }
    unsigned char* p = static_cast<unsigned char*>(ws);
    int8_t* B8 = reinterpret_cast<int8_t*>(p);
    double* eb = reinterpret_cast<double*>(p + al256(static_cast<int64_t>(MP) * s * MP));
    oz_slice_cols_kernel<<<MP, 256, 0, st>>>(Sinv, MP, m, waug, s, B8, eb);
    GPZ_KERNEL_CHECK();
    ++*launches;
    if (tev0) GPZ_CUDA(cudaEventRecord(tev0, st));
    int rc;
    if ((rc = ozmma_tgemm(D8, B8, MP, s, s + 1, n, ea, eb, Phi, ld, rw, H, accumulate, nupart, nu_ld, waug != nullptr ? m : -1, pred, st,
                          launches)))
        return rc;
    if (tev1) GPZ_CUDA(cudaEventRecord(tev1, st));
    return GPZ_OK;
}

// non-finite inputs were seen while extracting digits: make the Gram (and with it the Cholesky pivot test on every rank of a
// sharded run, after the allreduce) non-finite, so that NaN is returned like on the fp64 path (error convention)
__global__ void ozg_poison_kernel(double* __restrict__ S, const int* __restrict__ flag) {
    if (*flag) S[0] = nan("");
}

// sr[j] = 2^Ef 256^-2 (x cy for the spare column): S_jl = sr_j sum_e 256^-(e-2) acc_e
__global__ void ozg_scales_kernel(const double* __restrict__ scal, int MP, int m, int aug, double* __restrict__ sr) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= MP) return;
    const double f = 16.0 * pow2_ceil(scal[0]) * (1.0 / 65536.0);
    sr[j] = (j < m) ? f : ((j == m && aug) ? f * pow2_ceil(scal[1]) : 0.0);
}

int64_t oz_gram_workspace_bytes(int MP, int64_t rows) {
    const int64_t nch = oz_padded_rows(rows) / OZG_CH;
    return al256(ozmma_partial_doubles(MP, MP, 1, static_cast<int>(nch), 0) * 8) + al256(static_cast<int64_t>(MP) * 8);
}

// S (+)= PHI' diag(w) PHI over `rows` rows: F8' D8 as a digit GEMM contracting over the rows (MN-major operands, K chunks of
// OZG_CH rows so that s * OZG_CH * 2^14 < 2^31 keeps every int32 level accumulator exact).  Only tiles touching the lower
// triangle are computed; the result is mirrored.  gs: digits used by the Gram (<= s, the leading ones of the stored digits).
int ozaki_gram(const int8_t* F8, const int8_t* D8, int MP, int m, int64_t rows, int s, int gs, const double* d_scal, int aug,
               int accumulate, double* S, void* ws, const int* flag, cudaStream_t st, int64_t* launches) {
    if (gs < 2 || gs > s) {
        set_error("ozaki_gram: unsupported digit count %d (stored %d)", gs, s);
        return GPZ_ERR_USAGE;
    }
    const int nch = static_cast<int>(oz_padded_rows(rows) / OZG_CH);
    unsigned char* p = static_cast<unsigned char*>(ws);
    double* partial = reinterpret_cast<double*>(p);
    double* sr = reinterpret_cast<double*>(p + al256(ozmma_partial_doubles(MP, MP, 1, nch, 0) * 8));
    ozg_scales_kernel<<<static_cast<unsigned>(ceil_div(MP, 256)), 256, 0, st>>>(d_scal, MP, m, aug, sr);
    GPZ_KERNEL_CHECK();
    ++*launches;
    // addressed {row index j, digit, k = i within chunk, chunk}
    const int64_t str[3] = {MP, static_cast<int64_t>(s) * MP, static_cast<int64_t>(OZG_CH) * s * MP};
    int rc;
    if ((rc = ozmma_gemm_nt(F8, str, MP, D8, str, MP, gs, gs + 1, OZG_CH, nch, 1, 1, partial, sr, nullptr, 1.0, accumulate, S, MP, 0, st,
                            launches)))
        return rc;
    if (flag != nullptr) {
        ozg_poison_kernel<<<1, 1, 0, st>>>(S, flag);
        GPZ_KERNEL_CHECK();
        ++*launches;
    }
    return GPZ_OK;
}

// ---- general fp64 GEMM with a row-major left operand through the digit kernel: C = A B'  (GC + Psi path, gcpsi.cu) ---------
// digits of the rows of A [rows][lda] (cols valid columns) -> A8 [rows_pad][s][K128] with per-row power-of-two scales ea
int ozaki_row_digits(const double* A, int64_t lda, int cols, int K128, int64_t rows, int s, int8_t* A8, double* ea, int* flag,
                     cudaStream_t st, int64_t* launches) {
    if (K128 % 128 != 0 || cols > K128 || lda % 4 != 0 || s < 2 || s > OZ_MAXS) {
        set_error("ozaki_row_digits: unsupported K=%d cols=%d lda=%lld", K128, cols, static_cast<long long>(lda));
        return GPZ_ERR_USAGE;
    }
    const int64_t np = oz_padded_rows(rows);
    const int src_cols = static_cast<int>(lda < K128 ? lda : K128);
    launch_oz_digits(A, lda, src_cols, cols, K128, rows, np, s, nullptr, nullptr, 0, A8, nullptr, ea, flag, st);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

// digits of the COLUMNS of src [R][lds] (C columns): out [Cpad][s][Rpad] (K = row index contiguous, zero padded), scale[c];
// one block per column
__global__ void __launch_bounds__(256)
oz_transpose_digits_kernel(const double* __restrict__ src, int64_t lds, int R, int C, int Rpad, int s, int8_t* __restrict__ out,
                           double* __restrict__ scale, int* __restrict__ flag) {
    __shared__ double sh[8];
    __shared__ int Esh;
    const int c = blockIdx.x;
    double mx = 0.0;
    bool bad = false;
    if (c < C)
        for (int r = threadIdx.x; r < R; r += 256) {
            const double v = src[static_cast<int64_t>(r) * lds + c];
            bad |= !(fabs(v) <= 1.7e308);
            mx = fmax(mx, fabs(v));
        }
    if (bad && flag != nullptr) atomicExch(flag, 1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) mx = fmax(mx, sh[q]);
        const int E = oz_exponent(mx);
        Esh = E;
        scale[c] = ldexp(1.0, E - 8);
    }
    __syncthreads();
    const double sc = ldexp(1.0, 8 * s - Esh);
    int8_t* o = out + static_cast<int64_t>(c) * s * Rpad;
    for (int r = threadIdx.x; r < Rpad; r += 256) {
        long long I = (c < C && r < R) ? __double2ll_rn(src[static_cast<int64_t>(r) * lds + c] * sc) : 0;
        for (int u = s - 1; u >= 0; --u) o[static_cast<int64_t>(u) * Rpad + r] = static_cast<int8_t>(oz_digit(I));
    }
}

int ozaki_transpose_digits(const double* src, int64_t lds, int R, int C, int Rpad, int Cpad, int s, int8_t* out, double* scale, int* flag,
                           cudaStream_t st, int64_t* launches) {
    oz_transpose_digits_kernel<<<Cpad, 256, 0, st>>>(src, lds, R, C, Rpad, s, out, scale, flag);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

// ---- R (+)= X' F over the rows (K = row index) through the digit kernel, for two DIFFERENT matrices (GC + Psi moment GEMM) ----
// The Gram trick generalised: the B operand carries F_iq 2^-G_i (per-row exponent G_i of F's row, its ordinary row digits), the
// A operand carries X_ij 2^(G_i - Eb) with ONE exponent Eb = max_i (E_i + G_i) (E_i: exponent of X's row), so that the row
// scales cancel inside the product and the contraction over rows is exact up to 2^-56 of 2^Eb.
__global__ void __launch_bounds__(1024) oz_pair_exponent_kernel(const double* __restrict__ eaX, const double* __restrict__ eaF, int64_t n,
                                                                int* __restrict__ Eb) {
    __shared__ int sh[32];
    int mx = -100000;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        const int e = ((__double2hiint(eaX[i]) >> 20) & 0x7ff) + ((__double2hiint(eaF[i]) >> 20) & 0x7ff) - 2046 + 16;   // E_i + G_i (ea = 2^(E-8))
        mx = max(mx, e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 32; ++q) mx = max(mx, sh[q]);
        *Eb = mx;
    }
}

// A8 [n_pad][s][MP]: digits of X_ij 2^(G_i - Eb), 2^G_i = 256 eaF[i]; rows >= n zero; warp per row
__global__ void __launch_bounds__(256)
oz_digits_ext_kernel(const double* __restrict__ X, int64_t ld, int m, int MP, int64_t n, int64_t n_pad, int s, const double* __restrict__ eaF,
                     const int* __restrict__ Eb, int8_t* __restrict__ A8) {
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (i >= n_pad) return;
    int8_t* out = A8 + i * static_cast<int64_t>(s) * MP;
    if (i >= n) {
        const int4 z = make_int4(0, 0, 0, 0);
        for (int b = lane * 16; b < s * MP; b += 512) *reinterpret_cast<int4*>(out + b) = z;
        return;
    }
    const int G = ((__double2hiint(eaF[i]) >> 20) & 0x7ff) - 1023 + 8;
    const double sc = ldexp(1.0, 8 * s + G - *Eb);
    const unsigned long long bias = 0x0080808080808080ull >> (8 * (8 - s));
    const double* row = X + i * ld;
    for (int j = lane * 4; j < MP; j += 128) {
        const double4 v = *reinterpret_cast<const double4*>(row + j);
        const double x[4] = {v.x, v.y, v.z, v.w};
        unsigned long long I[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) I[q] = oz_digit_bytes(j + q < m ? __double2ll_rn(x[q] * sc) : 0, bias);
        oz_store_digits4(I, s, out + j, MP);
    }
}

__global__ void oz_fill_scale_kernel(const int* __restrict__ Eb, int n, double* __restrict__ sr) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) sr[j] = ldexp(1.0, *Eb - 16);                     // 2^Eb 256^-2
}

int64_t oz_moment_workspace_bytes(int MP, int cols, int64_t rows) {
    const int64_t nch = oz_padded_rows(rows) / OZG_CH;
    return al256(ozmma_partial_doubles(MP, cols, 0, static_cast<int>(nch), 0) * 8) + al256(static_cast<int64_t>(MP) * 8) + 256;
}

// R [MP][ldr] (columns < cols) (+)= X' F.  X [rows][ld] (m valid columns) with its row scales eaX already known (from its own
// row digits), F8 / eaF: row digits of F [rows_pad][s][K128].  A8: scratch [rows_pad][s][MP].  ws: oz_moment_workspace_bytes.
int ozaki_moment_gemm(const double* X, int64_t ld, int m, int MP, int64_t rows, const double* eaX, const int8_t* F8, const double* eaF,
                      int K128, int cols, int s, int8_t* A8, int accumulate, double* R, int64_t ldr, void* ws, cudaStream_t st,
                      int64_t* launches) {
    const int64_t np = oz_padded_rows(rows);
    const int nch = static_cast<int>(np / OZG_CH);
    unsigned char* p = static_cast<unsigned char*>(ws);
    double* partial = reinterpret_cast<double*>(p);
    double* sr = reinterpret_cast<double*>(p + al256(ozmma_partial_doubles(MP, cols, 0, nch, 0) * 8));
    int* Eb = reinterpret_cast<int*>(p + al256(ozmma_partial_doubles(MP, cols, 0, nch, 0) * 8) + al256(static_cast<int64_t>(MP) * 8));
    oz_pair_exponent_kernel<<<1, 1024, 0, st>>>(eaX, eaF, rows, Eb);
    oz_digits_ext_kernel<<<static_cast<unsigned>(ceil_div(np, 8)), 256, 0, st>>>(X, ld, m, MP, rows, np, s, eaF, Eb, A8);
    oz_fill_scale_kernel<<<static_cast<unsigned>(ceil_div(MP, 256)), 256, 0, st>>>(Eb, MP, sr);
    GPZ_KERNEL_CHECK();
    *launches += 3;
    // addressed {row index, digit, k = i within chunk, chunk}
    const int64_t strA[3] = {MP, static_cast<int64_t>(s) * MP, static_cast<int64_t>(OZG_CH) * s * MP};
    const int64_t strB[3] = {K128, static_cast<int64_t>(s) * K128, static_cast<int64_t>(OZG_CH) * s * K128};
    return ozmma_gemm_nt(A8, strA, MP, F8, strB, cols, s, s + 1, OZG_CH, nch, 0, 1, partial, sr, nullptr, 1.0, accumulate, R, ldr, 0, st, launches);
}

// PHI <- exp(PHI) in place for columns < m (0 beyond), fused row dots out_q[i] = sum_j PHI_ij vec_q[j]; warp per row
__global__ void __launch_bounds__(256)
exp_rows_inplace_kernel(double* __restrict__ Phi, int64_t ld, int m, int MP, int64_t rows, DotSpec dots) {
    __shared__ double exp_sm[32];
    exp_tab_stage(exp_sm);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (i >= rows) return;
    double* row = Phi + i * ld;
    double s0 = 0.0, s1 = 0.0;
    for (int jb = 2 * lane; jb < MP; jb += 256) {                 // 4 double2 per lane and step: one grouped exp (exp_tab_vec)
        double ex[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = jb + 64 * q;
            const double2 v = j < MP ? *reinterpret_cast<const double2*>(row + j) : make_double2(0.0, 0.0);
            ex[2 * q] = v.x;
            ex[2 * q + 1] = v.y;
        }
        exp_tab_vec(ex, exp_sm);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = jb + 64 * q;
            if (j >= MP) continue;
            double2 v;
            v.x = (j < m) ? ex[2 * q] : 0.0;
            v.y = (j + 1 < m) ? ex[2 * q + 1] : 0.0;
            *reinterpret_cast<double2*>(row + j) = v;
            if (dots.n > 0) {
                const double2 a = __ldg(reinterpret_cast<const double2*>(dots.vec[0] + j));
                s0 = fma(v.x, a.x, fma(v.y, a.y, s0));
            }
            if (dots.n > 1) {
                const double2 b = __ldg(reinterpret_cast<const double2*>(dots.vec[1] + j));
                s1 = fma(v.x, b.x, fma(v.y, b.y, s1));
            }
        }
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0) {
        if (dots.n > 0) dots.out[0][i] = s0;
        if (dots.n > 1) dots.out[1][i] = s1;
    }
}

int exp_rows_inplace(double* Phi, int64_t ld, int m, int MP, int64_t rows, const DotSpec& dots, cudaStream_t st, int64_t* launches) {
    if (rows <= 0) return GPZ_OK;
    exp_rows_inplace_kernel<<<static_cast<unsigned>(ceil_div(rows, 8)), 256, 0, st>>>(Phi, ld, m, MP, rows, dots);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

// ---- PHI = exp(F W) through the int8 tensor cores (ozmma_phi) ---------------------------------------------------------
// digits of the monomial row features F [rows][ldf] (q valid columns, dataset constants: done once) -> FD8 [rows_pad][s][128]
int ozaki_feature_digits(const double* F, int64_t ldf, int q, int64_t rows, int s, int8_t* FD8, double* eaF, int* flag, cudaStream_t st,
                         int64_t* launches) {
    if (q > 128 || ldf % 4 != 0) {
        set_error("ozaki_feature_digits: feature width %d (ld %lld) not supported", q, static_cast<long long>(ldf));
        return GPZ_ERR_USAGE;
    }
    const int64_t np = oz_padded_rows(rows);
    launch_oz_digits(F, ldf, static_cast<int>(ldf < 128 ? ldf : 128), q, 128, rows, np, s, nullptr, nullptr, 0, FD8, nullptr, eaF, flag, st);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return GPZ_OK;
}

// coefficient columns W [kq][MP] (row k, basis j) -> WD8 [j][u][128] (K = k contiguous, zero padded), scales ebW[j]
__global__ void __launch_bounds__(128)
oz_wdigits_kernel(const double* __restrict__ W, int kq, int MP, int m, int s, int8_t* __restrict__ WD8, double* __restrict__ ebW,
                  int* __restrict__ flag) {
    __shared__ double sh[4];
    __shared__ int Esh;
    const int j = blockIdx.x, k = threadIdx.x;
    const double v = (k < kq && j < m) ? W[static_cast<int64_t>(k) * MP + j] : 0.0;
    if (!(fabs(v) <= 1.7e308) && flag != nullptr) atomicExch(flag, 1);
    double mx = fabs(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((k & 31) == 0) sh[k >> 5] = mx;
    __syncthreads();
    if (k == 0) {
        mx = fmax(fmax(sh[0], sh[1]), fmax(sh[2], sh[3]));
        const int E = oz_exponent(mx);
        Esh = E;
        ebW[j] = ldexp(1.0, E - 8);
    }
    __syncthreads();
    long long I = __double2ll_rn(v * ldexp(1.0, 8 * s - Esh));
    int8_t* out = WD8 + static_cast<int64_t>(j) * s * 128 + k;
    for (int u = s - 1; u >= 0; --u) out[u * 128] = static_cast<int8_t>(oz_digit(I));
}

int ozaki_phi(const int8_t* FD8, const double* eaF, const double* W, int kq, int MP, int m, int s, int64_t rows, int8_t* WD8, double* ebW,
              double* Phi, int ndot, const double* vec0, const double* vec1, double* part0, double* part1, int64_t part_ld,
              const double* ycol, int* flag, cudaStream_t st, int64_t* launches) {
    if (kq > 128) {
        set_error("ozaki_phi: K = %d > 128", kq);
        return GPZ_ERR_USAGE;
    }
    oz_wdigits_kernel<<<MP, 128, 0, st>>>(W, kq, MP, m, s, WD8, ebW, flag);
    GPZ_KERNEL_CHECK();
    ++*launches;
    return ozmma_phi(FD8, eaF, WD8, ebW, kq, MP, m, s, rows, Phi, ndot, vec0, vec1, part0, part1, part_ld, ycol, st, launches);
}

}  // namespace gpz
