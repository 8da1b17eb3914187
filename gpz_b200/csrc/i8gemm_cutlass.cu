// int8 x int8 -> int32 GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8, TMEM accumulators, TMA operand
// staging, 2-CTA clusters) instantiated from the CUTLASS/CuTe sm100 templates vendored in this image
// (flashinfer/data/cutlass, CUTLASS 4.5).  It is the engine of the error-free fp64 GEMM in ozaki.cu: tcgen05 has no
// fp64 kind, but int8 products with int32 accumulation are exact.
//   D[M x N] (row-major, ldd) = A[M x K] (row-major int8, lda) * B[K x N] (given as N x K row-major int8, ldb)
#include <cstdint>

#include "internal.cuh"

#if defined(GPZ_HAVE_CUTLASS)
#include "cute/tensor.hpp"
#include "cutlass/cutlass.h"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/util/packed_stride.hpp"

namespace gpz {
namespace {
using namespace cute;
using ElementA = int8_t;
using LayoutA = cutlass::layout::RowMajor;
using ElementB = int8_t;
using LayoutB = cutlass::layout::ColumnMajor;
using ElementC = int32_t;
using LayoutC = cutlass::layout::RowMajor;
using MmaTileShape = Shape<_256, _128, _128>;       // 2-CTA MMA: 256 x 128 output tile per CTA pair, K step 128 bytes
using ClusterShape = Shape<_2, _1, _1>;
using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTileShape, ClusterShape,
    cutlass::epilogue::collective::EpilogueTileAuto, int32_t, int32_t, ElementC, LayoutC, 4, ElementC, LayoutC, 4,
    cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;
using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutA, 16, ElementB, LayoutB, 16, int32_t, MmaTileShape,
    ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;
using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
}  // namespace

bool i8gemm_available() { return true; }

int i8gemm_tn(const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb, int32_t* D, int64_t ldd, int M, int N, int K,
              void* workspace, size_t ws_bytes, cudaStream_t st) {
    return i8gemm_tn_batched(A, lda, 0, B, ldb, 0, D, ldd, 0, M, N, K, 1, workspace, ws_bytes, st);
}

int i8gemm_tn_batched(const int8_t* A, int64_t lda, int64_t bsA, const int8_t* B, int64_t ldb, int64_t bsB, int32_t* D,
                      int64_t ldd, int64_t bsD, int M, int N, int K, int batches, void* workspace, size_t ws_bytes,
                      cudaStream_t st) {
    using StrideA = typename Gemm::GemmKernel::StrideA;
    using StrideB = typename Gemm::GemmKernel::StrideB;
    using StrideC = typename Gemm::GemmKernel::StrideC;
    StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, cute::make_shape(M, K, batches));
    StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, cute::make_shape(N, K, batches));
    StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, cute::make_shape(M, N, batches));
    get<0>(sa) = lda;
    get<0>(sb) = ldb;
    get<0>(sc) = ldd;
    get<2>(sa) = bsA;
    get<2>(sb) = bsB;
    get<2>(sc) = bsD;
    typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, batches}, {A, sa, B, sb}, {{1, 0}, D, sc, D, sc}};
    Gemm gemm;
    if (gemm.can_implement(args) != cutlass::Status::kSuccess) {
        set_error("i8gemm: CUTLASS cannot implement M=%d N=%d K=%d", M, N, K);
        return GPZ_ERR_USAGE;
    }
    if (Gemm::get_workspace_size(args) > ws_bytes) {
        set_error("i8gemm: workspace too small");
        return GPZ_ERR_USAGE;
    }
    if (gemm.initialize(args, workspace, st) != cutlass::Status::kSuccess || gemm.run(st) != cutlass::Status::kSuccess) {
        set_error("i8gemm: CUTLASS launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return GPZ_ERR_CUDA;
    }
    return GPZ_OK;
}
}  // namespace gpz
#else
namespace gpz {
bool i8gemm_available() { return false; }
int i8gemm_tn(const int8_t*, int64_t, const int8_t*, int64_t, int32_t*, int64_t, int, int, int, void*, size_t, cudaStream_t) {
    set_error("built without the CUTLASS headers: the tcgen05 int8 GEMM is unavailable");
    return GPZ_ERR_USAGE;
}
int i8gemm_tn_batched(const int8_t*, int64_t, int64_t, const int8_t*, int64_t, int64_t, int32_t*, int64_t, int64_t, int, int, int,
                      int, void*, size_t, cudaStream_t) {
    set_error("built without the CUTLASS headers: the tcgen05 int8 GEMM is unavailable");
    return GPZ_ERR_USAGE;
}
}  // namespace gpz
#endif
