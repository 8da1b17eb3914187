// Shared device/host helpers for libgpz_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

namespace gpz {

constexpr int TILE = 128;          // CTA tile edge of the big DMMA GEMMs; m is padded to a multiple of it
constexpr int KSTEP = 16;          // K depth of one pipeline stage
constexpr int LDT = TILE + 4;      // smem row stride (doubles) of a [k][128] operand tile: 132 % 16 == 4
constexpr int LDK = KSTEP + 4;     // smem row stride of a [128][k] operand tile: 20 % 16 == 4
                                   // (both make the DMMA.8x8x4 fragment loads bank-conflict free)

void set_error(const char* fmt, ...);

#define GPZ_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            gpz::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call,             \
                           cudaGetErrorString(e__));                                        \
            return GPZ_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define GPZ_KERNEL_CHECK()                                                                  \
    do {                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) {                                                           \
            gpz::set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__,             \
                           cudaGetErrorString(e__));                                        \
            return GPZ_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

// cudaFuncSetAttribute (opt-in shared memory) is per device: `need()` is true the first time a call site runs on a device.
// The library is single-caller, so no locking.
struct PerDeviceOnce {
    unsigned long long mask[2] = {0ull, 0ull};
    bool need() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return true;
        unsigned long long& m = mask[dev >> 6];
        const unsigned long long bit = 1ull << (dev & 63);
        if (m & bit) return false;
        m |= bit;
        return true;
    }
};

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// fp64 tensor-core primitive.  On sm_100a every f64 mma.sync shape lowers to DMMA.8x8x4, so this is
// the hardware atom: D(8x8) += A(8x4) * B(4x8).  Lane l: g = l>>2, t = l&3.
//   a = A[g][t]        b = B[t][g]        c0,c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte async copy global -> shared; src_bytes = 0 zero-fills (used for row/col edges)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem))), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// ---- exp for the PHI-build epilogue (one exp per (row, basis): 1e9 per evaluation at the headline size; the kernel is bound by
// the fp64 pipe, which DMMA and DFMA share, and libm's exp is ~22 fp64 operations).  Table-driven: x = (32 e + j) ln2/32 + r,
// |r| <= ln2/64, exp(x) = 2^e * T[j] * (1 + r + r^2/2 + .. + r^6/720) -- 11 fp64 operations, truncation error 3e-18, about one
// ulp overall (T[j] and the final FMA round once each).  T = 2^(j/32), correctly rounded, staged in shared memory by the
// caller (32 doubles; two entries per bank, so a lookup is at most a 2-way conflict).  Outside |x| <= 708 (underflow into
// subnormals, overflow, NaN) it defers to libm.
__constant__ static double c_exp2_32[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0,
    0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0,
    0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0, 0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0,
    0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

__device__ __forceinline__ void exp_tab_stage(double* tab_sm /* 32 doubles */) {
    if (threadIdx.x < 32) tab_sm[threadIdx.x] = c_exp2_32[threadIdx.x];
}

__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ tab_sm) {
    if (!(fabs(x) <= 708.0)) return exp(x);
    const double MAGIC = 6755399441055744.0;                                  // 1.5 * 2^52: the sum's low word is rint(.) as an int
    const double t = fma(x, 46.16624130844683, MAGIC);                        // 32 / ln 2
    const int n = __double2loint(t);
    const double nd = t - MAGIC;
    double r = fma(nd, -0x1.62e42fefa0000p-6, x);                             // ln2/32, leading 37 bits: n * hi is exact
    r = fma(nd, -0x1.cf79abc9e3b3ap-45, r);
    double q = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    q = fma(q, r, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    const double T = tab_sm[n & 31];
    const double v = fma(T, q * r, T);                                        // in [0.98, 2.1)
    return __hiloint2double(__double2hiint(v) + ((n >> 5) << 20), __double2loint(v));
}

static __device__ __noinline__ double exp_tab_cold(double x, const double* tab_sm) { return exp_tab(x, tab_sm); }

// x[q] <- exp(x[q]) for N values at once, bit-identical to exp_tab per element.  exp_tab's range check is a branch per element, so
// inside an unrolled epilogue every element's ~12-deep fp64 chain became its own basic block and ran serially (profiles/r02z_aux.md:
// "wait" and math-pipe stalls, 6 cycles per instruction).  Here the check is one predicate for the whole group; the N chains are
// straight-line code that the scheduler interleaves, and the per-element path only runs when some |x| > 708.
template <int N>
__device__ __forceinline__ void exp_tab_vec(double (&x)[N], const double* __restrict__ tab_sm) {
    bool slow = false;
#pragma unroll
    for (int q = 0; q < N; ++q) slow |= !(fabs(x[q]) <= 708.0);
    if (slow) {                                           // cold: one out-of-line copy instead of N inlined libm bodies per call site
#pragma unroll
        for (int q = 0; q < N; ++q) x[q] = exp_tab_cold(x[q], tab_sm);
        return;
    }
    const double MAGIC = 6755399441055744.0;
    int n[N];
    double r[N], p[N];
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const double t = fma(x[q], 46.16624130844683, MAGIC);
        n[q] = __double2loint(t);
        const double nd = t - MAGIC;
        r[q] = fma(nd, -0x1.cf79abc9e3b3ap-45, fma(nd, -0x1.62e42fefa0000p-6, x[q]));
    }
#pragma unroll
    for (int q = 0; q < N; ++q) p[q] = fma(r[q], 1.0 / 720.0, 1.0 / 120.0);
#pragma unroll
    for (int q = 0; q < N; ++q) p[q] = fma(p[q], r[q], 1.0 / 24.0);
#pragma unroll
    for (int q = 0; q < N; ++q) p[q] = fma(p[q], r[q], 1.0 / 6.0);
#pragma unroll
    for (int q = 0; q < N; ++q) p[q] = fma(p[q], r[q], 0.5);
#pragma unroll
    for (int q = 0; q < N; ++q) p[q] = fma(p[q], r[q], 1.0);
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const double T = tab_sm[n[q] & 31];
        const double v = fma(T, p[q] * r[q], T);
        x[q] = __hiloint2double(__double2hiint(v) + ((n[q] >> 5) << 20), __double2loint(v));
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block reduction (fixed order): result valid in thread 0
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sh /* >= NT/32 doubles */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) r += sh[i];
    }
    return r;
}
#endif

}  // namespace gpz
