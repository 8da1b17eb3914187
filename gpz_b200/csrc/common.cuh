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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
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
