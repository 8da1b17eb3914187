#include <cstdint>
#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/util/packed_stride.hpp"
using namespace cute;
using ElementA = int8_t;  using LayoutA = cutlass::layout::RowMajor;    constexpr int AlignA = 16;
using ElementB = int8_t;  using LayoutB = cutlass::layout::ColumnMajor; constexpr int AlignB = 16;
using ElementC = int32_t; using LayoutC = cutlass::layout::RowMajor;    constexpr int AlignC = 4;
using ElementAcc = int32_t;
using ArchTag = cutlass::arch::Sm100;
using OpClass = cutlass::arch::OpClassTensorOp;
using MmaTileShape = Shape<_256, _128, _128>;
using ClusterShape = Shape<_2, _1, _1>;
using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    ArchTag, OpClass, MmaTileShape, ClusterShape, cutlass::epilogue::collective::EpilogueTileAuto,
    ElementAcc, int32_t, ElementC, LayoutC, AlignC, ElementC, LayoutC, AlignC,
    cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;
using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    ArchTag, OpClass, ElementA, LayoutA, AlignA, ElementB, LayoutB, AlignB, ElementAcc, MmaTileShape, ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;
using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;

extern "C" int i8gemm(const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb, int32_t* D, int64_t ldd, int M, int N, int K,
                      void* workspace, size_t ws_bytes, cudaStream_t st) {
    using StrideA = typename Gemm::GemmKernel::StrideA; using StrideB = typename Gemm::GemmKernel::StrideB;
    using StrideC = typename Gemm::GemmKernel::StrideC; using StrideD = typename Gemm::GemmKernel::StrideD;
    StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, cute::make_shape(M, K, 1));
    StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, cute::make_shape(N, K, 1));
    StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, cute::make_shape(M, N, 1));
    StrideD sd = sc;
    get<0>(sa) = lda; get<0>(sb) = ldb; get<0>(sc) = ldd; get<0>(sd) = ldd;
    typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, 1}, {A, sa, B, sb}, {{1, 0}, D, sc, D, sd}};
    Gemm gemm;
    if (gemm.can_implement(args) != cutlass::Status::kSuccess) return 1;
    if (Gemm::get_workspace_size(args) > ws_bytes) return 2;
    if (gemm.initialize(args, workspace, st) != cutlass::Status::kSuccess) return 3;
    if (gemm.run(st) != cutlass::Status::kSuccess) return 4;
    return 0;
}
