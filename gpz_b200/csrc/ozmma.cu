// Hand-written tcgen05 kernel for the error-free int8-digit (Ozaki) fp64 GEMMs of the NLML evaluation:
//     T = PHI * iSigma          (GPz/GPz.m:69,72)   epilogue: nu_i = sum_j PHI_ij T_ij, H = rw_i PHI .* T, spare column -> PHI*w
//     S = PHI' diag(w) PHI      (GPz/GPz.m:63-65)   epilogue: fp64 partial tiles, reduced in fixed order by ozmma_gram_reduce
//
// Both operands are given as s signed 8-bit digits (two's-complement digit extraction, ozaki.cu):
//     a = sa * sum_t A_t 256^-(t-1),   b = sb * sum_u B_u 256^-(u-1),   A_t, B_u in [-128, 127].
// The product is sum over levels e = t+u of 256^-(e-2) * (A_t . B_u); every digit product is an EXACT int8 x int8 -> int32
// tensor-core GEMM (tcgen05.mma kind::i8, accumulators in TMEM).  Unlike a chain of library GEMMs (one int32 output per
// level, summed by a separate kernel) this kernel keeps the whole level loop on chip:
//   * one CTA PAIR (cta_group::2, UMMA 256 x 128 x 32) owns a 256 x 128 output tile; each CTA holds 128 rows;
//   * warp 0 (one lane) streams the digit tiles with TMA (4-D tensor maps {k, digit, row, chunk} or, for operands stored with the
//     row index contiguous (MN-major), {row, digit, k, chunk}; 128B / 64B swizzle) through an 8-stage mbarrier pipeline; warp 1 of the leader CTA (one lane) issues the MMAs of all digit pairs of a level into
//     one TMEM accumulator; TMEM holds two accumulators so level e-1 is multiplied while level e is folded;
//   * 16 epilogue warps read the finished level with tcgen05.ld and fold it into an fp64 running sum held in REGISTERS
//     (128 x 128 doubles per CTA = 32 per thread), smallest level first, so the int32 level results never touch HBM;
//   * after the last level of the tile the same warps apply the GEMM-specific epilogue directly from registers.
// Work units are (group of K chunks, tile), chunk-group major, dealt round-robin to the CTA pairs: at any time the pairs
// work on neighbouring tiles of the same K window, so every digit byte is fetched from HBM once and re-streamed from L2
// (a contiguous per-pair partition measured 21 GB of DRAM reads for 3 GB of operands).  The assignment is static, so
// results are bit-reproducible.  K per chunk is bounded by the caller so that no int32 accumulator can overflow.
#include <cuda.h>

#include "internal.cuh"

namespace gpz {
namespace {

constexpr int OM_STAGES = 6;
constexpr int OM_A_BYTES = 128 * 128;                 // A digit tile: 128 rows x 128 K-bytes per CTA
constexpr int OM_B_BYTES = 128 * 128;                 // B digit tile of a two-level step: 128 rows per CTA (a one-level step fills half of it)
constexpr int OM_STAGE_BYTES = OM_A_BYTES + OM_B_BYTES;
constexpr int OM_BAR_OFF = OM_STAGES * OM_STAGE_BYTES;
// barrier block (8 bytes each): full[8] @0, empty[8] @64, tfull[4] @128, tempty[4] @160, tmem slot @192
constexpr int OM_FULL = 0, OM_EMPTY = 64, OM_TFULL = 128, OM_TEMPTY = 160, OM_SLOT = 192;
constexpr int OM_SMEM_BYTES = OM_BAR_OFF + 256 + 6144 + 256 + 1024;   // barriers + tmem slot, row sums, exp table, + slack for the 1024-byte alignment
constexpr int OM_THREADS = 576;                       // warp 0: TMA, warp 1: MMA, warps 2..17: epilogue (4 lane quarters x 4 column groups)
constexpr uint32_t OM_TMEM_COLS = 512;                // two regions of 256 columns = two levels each: one region multiplied, one folded
// instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D = S32, A = B = signed 8 bit, both K-major,
// N = 128 or 256 (>>3 at bit 17), M = 256 (>>4 at bit 24)
constexpr uint32_t OM_IDESC0 = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 4) << 24);
constexpr uint32_t OM_IDESC_N128 = OM_IDESC0 | ((128u >> 3) << 17), OM_IDESC_N256 = OM_IDESC0 | ((256u >> 3) << 17);

struct OzmmaArgs {
    int s, emin, emax;        // digits per operand; levels emax, emax-1, .., emin are accumulated (weight 256^-(e-emin))
    int kblocks;              // K bytes per chunk / 128
    int tiles_n;              // 128-column tiles
    int ntiles;               // number of (256-row, 128-column) tiles in the list
    int nchunks;              // K chunks in total
    int gchunks;              // K chunks per work unit (folded one after the other into the fp64 registers)
    int ngroups;              // ceil(nchunks / gchunks); work unit u = group * ntiles + tile
    int lower;                // tile list = tiles touching the lower triangle (Gram); else all tiles_m x tiles_n
    int mode;                 // 0: fp64 partial tiles, 1: T-GEMM epilogue, 2: PHI = exp(.) epilogue, 3: C = scaled product, stored directly
    int mn_major;             // operands stored [k][row] (row index contiguous) instead of [row][k]
    int lgroup;               // levels multiplied together: 2 (default, shared operand tiles) or 1 (one level at a time)
    int prefetch;             // mode 1: prefetch the epilogue's PHI block into L2 (option "ozaki_prefetch", default 0)
    int kmma;                 // 32-byte K steps multiplied per 128-byte stage (4; fewer when the operands' K extent is < 97: PHI build)
    int nint;                 // the first nint levels of a unit (emax, emax-1, ..) are folded EXACTLY in int64; the host picks the
                              // largest count whose sum cannot overflow, and 0 when a unit folds more than one K chunk
    uint64_t hintA, hintB;    // L2 eviction policy of the operand loads
    // mode 1
    const double* ea;         // [rows] row scales (including 256^-1 .. see ozaki.cu)
    const double* eb;         // [cols] column scales
    const double* Phi;        // [rows][ld]
    int64_t ld, rows;
    const double* rw;         // [rows] or null
    double* H;                // [rows][ld] or null
    int accumulate;
    double* nupart;           // [tiles_n][nu_ld]
    int64_t nu_ld;
    int aug_col;              // -1: none
    double* pred;             // [rows]
    // mode 2: PHI_ij = exp(ea_i eb_j sum) for j < m, spare column m <- ycol, fused row dots with vec0 / vec1
    int m;
    const double* ycol;
    int ndot;
    const double* vec0;
    const double* vec1;
    double* part0;            // [tiles_n][nu_ld]
    double* part1;
    // mode 0
    double* partial;          // [ngroups*ntiles][256*128]
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one lane of the (converged) warp; the surrounding code stays warp-uniform so that barrier addresses, descriptors and
// coordinates live in uniform registers (a single-lane loop costs a R2UR waterfall per tcgen05 / TMA instruction)
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ uint32_t mapa_cta(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// bounded wait: a protocol bug traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 20000000000LL) __trap();     // ~10 s
    }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// L2 eviction policies (the encodings createpolicy.fractional.L2::evict_{normal,first,last} produces; cute TMA::CacheHintSm90)
constexpr uint64_t OM_EVICT_NORMAL = 0x1000000000000000ull, OM_EVICT_FIRST = 0x12F0000000000000ull, OM_EVICT_LAST = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t smem_dst, uint32_t bar_cluster, int c0, int c1, int c2,
                                            int c3, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, "
        "%6}], [%2], %7;"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(hint)
        : "memory");
}
// streaming epilogue traffic (PHI in, H / partials out) must not push the operand digits out of L2
__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(OM_EVICT_FIRST));
    return v;
}
// (not volatile / no memory clobber: the compiler may batch the 32 row loads of an epilogue; nothing in the kernel re-reads
// what these stores write)
__device__ __forceinline__ double ld_stream1(const double* p) {
    double v;
    asm("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(OM_EVICT_FIRST));
    return v;
}
__device__ __forceinline__ void st_stream1(double* p, double x) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(x), "l"(OM_EVICT_FIRST));
}
__device__ __forceinline__ void st_stream2(double* p, double x, double y) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(x), "d"(y), "l"(OM_EVICT_FIRST) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], 256 x 128 x 32 int8, both CTAs of the pair
__device__ __forceinline__ void umma_i8_2cta(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                             uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %5, {%7, %7, %7, %7, %7, %7, %7, %7}, p;\n\t}"
        ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// arrive (once) on the mbarrier at the same smem offset in both CTAs when all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// x * 2^e for x = 0 or a normal double, by exponent arithmetic on the integer pipe (results below 2^-1022 flush to zero):
// while tcgen05.mma is running, fp64 instructions of the same SM issue ~20x slower than nominal (profiles/r02e: DFMA / DMUL
// "math pipe throttle" was 60-70 % of all stall samples at 3.5 % fp64-pipe utilisation), so the epilogue avoids them where it can
__device__ __forceinline__ double scale_pow2(double x, int e) {
    const int hi = __double2hiint(x);
    const int ex = ((hi >> 20) & 0x7ff);
    const int nx = ex + e;
    if (ex == 0 || nx <= 0) return 0.0;
    if (nx >= 2047) return __hiloint2double((hi & 0x80000000) | 0x7ff00000, 0);
    return __hiloint2double(hi + (e << 20), __double2loint(x));
}
// exponent of a power of two (the row / column scales the digit kernels of ozaki.cu produce are ldexp(1.0, .))
__device__ __forceinline__ int pow2_exponent(double x) { return ((__double2hiint(x) >> 20) & 0x7ff) - 1023; }

__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lo_bits) { return ((smem_addr >> 4) & 0x3FFFu) | lo_bits; }

__device__ __forceinline__ void decode_tile(const OzmmaArgs& a, int tile, int& mt, int& nt) {
    if (!a.lower) {
        mt = tile / a.tiles_n;
        nt = tile - mt * a.tiles_n;
        return;
    }
    int acc = 0;
    for (mt = 0;; ++mt) {
        const int cnt = min(a.tiles_n, 2 * (mt + 1));
        if (tile < acc + cnt) {
            nt = tile - acc;
            return;
        }
        acc += cnt;
    }
}

// Schedule of one (work unit, K chunk): the levels are taken two at a time from the top, {emax, emax-1}, {emax-2, emax-3}, ..
// (a last single level when their number is odd).  Inside a group, for every 128-byte K block, the A digits t are swept in
// ascending order: A_t meets B_{eh-t} (level eh) and B_{el-t} (level el = eh-1).  Both products are ONE tcgen05.mma with
// N = 256: CTA 0 of the pair holds the 128 rows of B_{eh-t}, CTA 1 those of B_{el-t} (with cta_group::2 each CTA supplies half
// of the N extent), and the two levels' accumulators are the two 128-column halves of one 256-column TMEM region.  Against one
// level at a time (N = 128, every digit pair fetching its own A tile) this halves the A traffic and the number of MMA, TMA and
// barrier instructions per digit pair: that schedule sat on the L2 -> SM bandwidth (profiles/r01h: 57-66 % of the L2 peak at
// 52 % tensor-pipe activity), and its single issuing thread was busy 60 % of the time (profiles/r02c).  Steps where only one of
// the two levels has a partner digit (the ends of a sweep, the last single level) run as N = 128 into their half.
// Requires emax <= s + 1 (every caller: products below 256^-s of row scale x column scale are dropped), so that a group's
// first step is a two-level step and initialises both halves; lgroup = 1 selects one level at a time (all steps N = 128).
template <int MNMAJOR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(OM_THREADS, 1)
ozmma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapB2,
             const OzmmaArgs a) {
    extern __shared__ uint8_t om_smem_raw[];
    const uint32_t raw = smem_u32(om_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gbase = om_smem_raw + (base - raw);
    const uint32_t bar0 = base + OM_BAR_OFF;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + OM_BAR_OFF + OM_SLOT);
    double* rowsum_sm = reinterpret_cast<double*>(gbase + OM_BAR_OFF + 256);      // [2][3][128]
    double* exp_sm = rowsum_sm + 768;                                              // [32] table of exp_tab (PHI epilogue)
    if (a.mode == 2) exp_tab_stage(exp_sm);                                        // visible after the cluster barrier below
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();

    cluster_sync_all();                                   // both CTAs resident before the pair-wide TMEM allocation
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < OM_STAGES; ++i) {
            mbar_init(bar0 + OM_FULL + 8 * i, 2);         // one arrive per CTA's producer (+ the TMA bytes of both CTAs)
            mbar_init(bar0 + OM_EMPTY + 8 * i, 1);        // one MMA commit
        }
        for (int b = 0; b < 4; ++b) {
            mbar_init(bar0 + OM_TFULL + 8 * b, 1);        // one MMA commit
            mbar_init(bar0 + OM_TEMPTY + 8 * b, 32);      // 16 epilogue warps x 2 CTAs
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc2(smem_u32(const_cast<uint32_t*>(tmem_slot)), OM_TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int U = a.ntiles * a.ngroups;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer: one elected lane runs the whole loop
        if (elect_one()) {
            uint32_t stage = 0, ph = 0;
            const uint32_t full_leader0 = mapa_cta(bar0 + OM_FULL, 0);
            for (int u = pair; u < U; u += npairs) {
                const int group = u / a.ntiles, tile = u - group * a.ntiles;
                int mt, nt;
                decode_tile(a, tile, mt, nt);
                const int rowA = mt * 256 + static_cast<int>(rank) * 128, rowB = nt * 128, rowBh = rowB + static_cast<int>(rank) * 64;
                const int c1 = min(a.nchunks, (group + 1) * a.gchunks);
                for (int chunk = group * a.gchunks; chunk < c1; ++chunk)
                    for (int eh = a.emax; eh >= a.emin; eh -= a.lgroup) {
                        const int el = eh - 1;
                        const bool hasl = a.lgroup == 2 && el >= a.emin;
                        const int t0 = max(1, (hasl ? el : eh) - a.s), t1 = min(a.s, eh - 1);
                        for (int kb = 0; kb < a.kblocks; ++kb)
                            for (int t = t0; t <= t1; ++t) {
                                const int uh = eh - t, ul = el - t;
                                const bool vh = uh <= a.s, vl = hasl && ul >= 1;
                                const bool dual = vh && vl;
                                mbar_wait(bar0 + OM_EMPTY + 8 * stage, ph ^ 1u);
                                const uint32_t fl = full_leader0 + 8 * stage;
                                const uint32_t sA = base + stage * OM_STAGE_BYTES, sB = sA + OM_A_BYTES;
                                if (rank == 0) mbar_arrive_expect_tx(bar0 + OM_FULL + 8 * stage, dual ? 2 * OM_STAGE_BYTES : 2 * OM_A_BYTES + OM_B_BYTES);
                                else mbar_arrive_cluster(fl);
                                if (MNMAJOR) tma_load_4d(&mapA, sA, fl, rowA, t - 1, kb * 128, chunk, a.hintA);
                                else tma_load_4d(&mapA, sA, fl, kb * 128, t - 1, rowA, chunk, a.hintA);
                                if (dual) {                 // this CTA's 128 rows of the N = 256 operand: CTA 0 level eh, CTA 1 level el
                                    const int dg = (rank == 0 ? uh : ul) - 1;
                                    if (MNMAJOR) tma_load_4d(&mapB2, sB, fl, rowB, dg, kb * 128, chunk, a.hintB);
                                    else tma_load_4d(&mapB2, sB, fl, kb * 128, dg, rowB, chunk, a.hintB);
                                } else {                    // this CTA's 64 rows of the N = 128 operand
                                    const int dg = (vh ? uh : ul) - 1;
                                    if (MNMAJOR) tma_load_4d(&mapB, sB, fl, rowBh, dg, kb * 128, chunk, a.hintB);
                                    else tma_load_4d(&mapB, sB, fl, kb * 128, dg, rowBh, chunk, a.hintB);
                                }
                                if (++stage == OM_STAGES) {
                                    stage = 0;
                                    ph ^= 1u;
                                }
                            }
                    }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer: one elected lane of the leader CTA
        if (rank == 0 && elect_one()) {
            // smem matrix descriptors (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1
            // <<46 | layout <<61; the values follow deep_gemm's make_umma_desc).  K-major, 128B swizzle: rows of 128 K-bytes, 8-row
            // groups 1024 B apart (SBO), LBO unused (1), 32 K-bytes = +32 B.  MN-major: the tile is [128 k-rows][W bytes of the row
            // index], W = 128 (A and the two-level B, SWIZZLE_128B = 2) or 64 (one-level B, SWIZZLE_64B = 4: each CTA of the pair
            // holds 64 of the 128 N columns); 8-k-row groups 8 W bytes apart (SBO), LBO = stride between W-wide atoms (one atom
            // here), 32 K-rows = +32 W bytes.
            constexpr uint32_t VER = 1u << 14;
            constexpr uint32_t HI_A = (1024u >> 4) | VER | (2u << 29);
            constexpr uint32_t HI_B = MNMAJOR ? ((512u >> 4) | VER | (4u << 29)) : HI_A;
            constexpr uint32_t LO_A = MNMAJOR ? (((128u * 128u) >> 4) << 16) : (1u << 16);
            constexpr uint32_t LO_B = MNMAJOR ? (((128u * 64u) >> 4) << 16) : (1u << 16);
            constexpr uint32_t KA = MNMAJOR ? ((32u * 128u) >> 4) : (32u >> 4), KB = MNMAJOR ? ((32u * 64u) >> 4) : (32u >> 4);
            constexpr uint32_t TR = MNMAJOR ? ((1u << 15) | (1u << 16)) : 0u;
            constexpr uint32_t IDESC1 = OM_IDESC_N128 | TR, IDESC2 = OM_IDESC_N256 | TR;
            constexpr uint32_t STG = OM_STAGE_BYTES >> 4;
            const uint32_t a0 = (base >> 4) & 0x3FFFu;
            uint32_t G = 0, ubits = 0, stage = 0, ph = 0, aoff = a0, fullbar = bar0 + OM_FULL, emptybar = bar0 + OM_EMPTY;
            for (int u = pair; u < U; u += npairs) {
                const int group = u / a.ntiles;
                const int nch = min(a.nchunks, (group + 1) * a.gchunks) - group * a.gchunks;
                for (int ch = 0; ch < nch; ++ch)
                    for (int eh = a.emax; eh >= a.emin; eh -= a.lgroup, ++G) {
                        const int el = eh - 1;
                        const bool hasl = a.lgroup == 2 && el >= a.emin;
                        const int t0 = max(1, (hasl ? el : eh) - a.s), t1 = min(a.s, eh - 1);
                        // accumulator b = 2 (G & 1) + half; ubits bit b = parity of the number of times b has been used (a
                        // single-level group leaves its region's second half unused, so the halves need their own counts)
                        const uint32_t reg = G & 1u, bh = 2u * reg, bl = bh + 1u;
                        mbar_wait(bar0 + OM_TEMPTY + 8 * bh, ((ubits >> bh) & 1u) ^ 1u);      // the epilogue has drained it
                        if (hasl) mbar_wait(bar0 + OM_TEMPTY + 8 * bl, ((ubits >> bl) & 1u) ^ 1u);
                        tc_fence_after();
                        const uint32_t dreg = tmem_base + reg * 256u;
                        uint32_t acch = 0, accl = 0;
                        for (int kb = 0; kb < a.kblocks; ++kb)
                            for (int t = t0; t <= t1; ++t) {
                                const bool vh = eh - t <= a.s, vl = hasl && el - t >= 1;
                                mbar_wait(fullbar, ph);                             // both CTAs' tiles have landed
                                tc_fence_after();
                                const uint32_t alo = aoff | LO_A;
                                if (vh && vl) {
                                    const uint32_t blo = (aoff + (OM_A_BYTES >> 4)) | LO_A;    // 128 rows per CTA: A's layout
                                    umma_i8_2cta(dreg, alo, HI_A, blo, HI_A, IDESC2, acch);     // 4 x 32 K-bytes of the stage
                                    if (a.kmma > 1) umma_i8_2cta(dreg, alo + KA, HI_A, blo + KA, HI_A, IDESC2, 1u);
                                    if (a.kmma > 2) umma_i8_2cta(dreg, alo + 2 * KA, HI_A, blo + 2 * KA, HI_A, IDESC2, 1u);
                                    if (a.kmma > 3) umma_i8_2cta(dreg, alo + 3 * KA, HI_A, blo + 3 * KA, HI_A, IDESC2, 1u);
                                    acch = accl = 1u;
                                } else {
                                    const uint32_t blo = (aoff + (OM_A_BYTES >> 4)) | LO_B;
                                    const uint32_t d1 = vh ? dreg : dreg + 128u, acc1 = vh ? acch : accl;
                                    umma_i8_2cta(d1, alo, HI_A, blo, HI_B, IDESC1, acc1);
                                    if (a.kmma > 1) umma_i8_2cta(d1, alo + KA, HI_A, blo + KB, HI_B, IDESC1, 1u);
                                    if (a.kmma > 2) umma_i8_2cta(d1, alo + 2 * KA, HI_A, blo + 2 * KB, HI_B, IDESC1, 1u);
                                    if (a.kmma > 3) umma_i8_2cta(d1, alo + 3 * KA, HI_A, blo + 3 * KB, HI_B, IDESC1, 1u);
                                    if (vh) acch = 1u;
                                    else accl = 1u;
                                }
                                umma_commit_pair(emptybar);                         // frees the smem stage in both CTAs
                                ++stage;
                                aoff += STG;
                                fullbar += 8;
                                emptybar += 8;
                                if (stage == OM_STAGES) {
                                    stage = 0;
                                    ph ^= 1u;
                                    aoff = a0;
                                    fullbar = bar0 + OM_FULL;
                                    emptybar = bar0 + OM_EMPTY;
                                }
                            }
                        umma_commit_pair(bar0 + OM_TFULL + 8 * bh);                 // levels finished -> epilogue warps
                        ubits ^= 1u << bh;
                        if (hasl) {
                            umma_commit_pair(bar0 + OM_TFULL + 8 * bl);
                            ubits ^= 1u << bl;
                        }
                    }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue warps: fold levels in fp64 registers
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int cg = (warp - 2) >> 2;         // which 32 of the 128 accumulator columns
        const int rloc = q * 32 + lane;         // row inside this CTA's 128 rows
        double st[32];
        uint32_t G = 0, half = 0, ebits = 0;      // ebits: per accumulator, parity of its number of uses (as in the MMA issuer)
        const uint32_t tempty_leader0 = mapa_cta(bar0 + OM_TEMPTY, 0);
        for (int u = pair; u < U; u += npairs) {
            __syncwarp();
            const int group = u / a.ntiles, tile = u - group * a.ntiles;
#pragma unroll
            for (int c = 0; c < 32; ++c) st[c] = 0.0;
            const int nlev = (min(a.nchunks, (group + 1) * a.gchunks) - group * a.gchunks) * (a.emax - a.emin + 1);
            int mt, nt;
            decode_tile(a, tile, mt, nt);
            const int lv_prefetch = nlev > 5 ? nlev - 5 : 0;
            for (int lv = 0, e = a.emax; lv < nlev; ++lv, e = (e == a.emin ? a.emax : e - 1)) {
                if (a.mode == 1 && a.prefetch && lv == lv_prefetch) {
                    // the T-GEMM epilogue multiplies by this warp's 32 x 32 block of PHI: pull it into L2 while the last levels are
                    // still being multiplied (r02e capture: those loads, straight from HBM, were half of all stall samples and
                    // kept the epilogue warps -- and through the TMEM hand-shake the tensor pipe -- waiting)
                    const int64_t gr = static_cast<int64_t>(mt) * 256 + static_cast<int64_t>(rank) * 128 + q * 32 + lane;
                    if (gr < a.rows) {
                        const double* pp = a.Phi + gr * a.ld + nt * 128 + cg * 32;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + 16));
                    }
                }
                // level e is half `half` of group G (the MMA issuer's numbering): TMEM region G & 1, columns half * 128
                const uint32_t buf = 2u * (G & 1u) + half;
                mbar_wait(bar0 + OM_TFULL + 8 * buf, (ebits >> buf) & 1u);
                ebits ^= 1u << buf;
                tc_fence_after();
                if (a.lgroup == 2 && half == 0u && e > a.emin) half = 1u;
                else {
                    half = 0u;
                    ++G;
                }
                // running sum in units of the LOWEST level's (emax) least significant digit product: level e weighs 256^(emax-e).
                // The first a.nint levels are added exactly as 64-bit integers (st[] holds the integer's bits), the others in fp64.
                const int li = a.emax - e;
                const bool ilev = li < a.nint;
                if (li == a.nint && li > 0) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) st[c] = static_cast<double>(__double_as_longlong(st[c]));
                }
                const double wgt = __longlong_as_double(static_cast<long long>(1023 + 8 * li) << 52);          // 256^(emax-e)
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 128u + static_cast<uint32_t>(cg * 32);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t v[16];
                    tmem_ld16(taddr + hh * 16, v);
                    tmem_ld_wait();
                    if (ilev) {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            st[hh * 16 + c] = __longlong_as_double(__double_as_longlong(st[hh * 16 + c]) +
                                                                   (static_cast<long long>(static_cast<int>(v[c])) << (8 * li)));
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c) st[hh * 16 + c] = fma(static_cast<double>(static_cast<int>(v[c])), wgt, st[hh * 16 + c]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader0 + 8 * buf);
            }
            if (a.nint >= a.emax - a.emin + 1) {                    // every level was an integer level
#pragma unroll
                for (int c = 0; c < 32; ++c) st[c] = static_cast<double>(__double_as_longlong(st[c]));
            }
            const int efold = -8 * (a.emax - a.emin);               // back to units of the top level (emin): applied with the scales
            // Transpose the warp's 32 x 32 block in registers (5 butterfly stages of shuffles): before, lane = row and st[c] =
            // column c; after, lane = column and st[r] = row r.  Every global access below is then one 256-byte row segment
            // per warp instruction instead of 32 rows x 16 bytes.
#pragma unroll
            for (int K = 16; K > 0; K >>= 1) {
                const bool up = (lane & K) != 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i & K) continue;
                    const double give = up ? st[i] : st[i | K];
                    const double got = __shfl_xor_sync(0xffffffffu, give, K);
                    if (up) st[i] = got;
                    else st[i | K] = got;
                }
            }
            const int64_t gi0 = static_cast<int64_t>(mt) * 256 + static_cast<int64_t>(rank) * 128 + q * 32;   // first row of this warp
            const int col = nt * 128 + cg * 32 + lane;                                                          // this lane's column
            if (a.mode == 0) {
                double* out = a.partial + static_cast<int64_t>(u) * (256 * 128) + static_cast<int64_t>(static_cast<int>(rank) * 128 + q * 32) * 128 +
                              cg * 32 + lane;
#pragma unroll
                for (int r = 0; r < 32; ++r) st_stream1(out + r * 128, scale_pow2(st[r], efold));
                continue;
            }
            if (a.mode == 3) {                   // plain product, one K chunk per unit: out[row][col] = sum * ea_row * eb_col
                const bool colok = col < a.m;
                const int exb = colok ? pow2_exponent(a.eb[col]) + efold : 0;
                const int exa_l = (gi0 + lane < a.rows) ? pow2_exponent(a.ea[gi0 + lane]) : 0;
                double* op = a.H + gi0 * a.ld + col;
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const int ex = __shfl_sync(0xffffffffu, exa_l, r) + exb;
                    if (colok && gi0 + r < a.rows) st_stream1(op + r * a.ld, scale_pow2(st[r], ex));
                }
                continue;
            }
            // st[r] is overwritten by this lane's contribution to the row sum of row r (T-GEMM: PHI .* T; PHI build: PHI * vec0);
            // the PHI build's optional second dot is reduced first, from the PHI values still in st
            double second = 0.0;
            const int ndots = a.mode == 2 ? a.ndot : 1;
            if (a.mode == 2) {
                const int exb = pow2_exponent(a.eb[col]) + efold;
                const double v0 = a.ndot > 0 ? a.vec0[col] : 0.0;
                double* pp = a.H != nullptr ? a.H + gi0 * a.ld + col : nullptr;
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    double p = 0.0;
                    if (gi0 + r < a.rows) {
                        if (col < a.m) p = exp_tab(scale_pow2(st[r], pow2_exponent(a.ea[gi0 + r]) + exb), exp_sm);
                        else if (a.ycol != nullptr && col == a.m) p = a.ycol[gi0 + r];
                        if (pp != nullptr) pp[r * a.ld] = p;
                    }
                    st[r] = p;
                }
                if (a.ndot > 1) {
                    const double v1 = a.vec1[col];
                    double red[16];
                    {
                        const bool up = (lane & 16) != 0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const double give = (up ? st[i] : st[i + 16]) * v1, keep = (up ? st[i + 16] : st[i]) * v1;
                            red[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
                        }
                    }
#pragma unroll
                    for (int K = 8; K > 0; K >>= 1) {
                        const bool up = (lane & K) != 0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (i >= K) continue;
                            const double give = up ? red[i] : red[i + K], keep = up ? red[i + K] : red[i];
                            red[i] = keep + __shfl_xor_sync(0xffffffffu, give, K);
                        }
                    }
                    second = red[0];
                }
#pragma unroll
                for (int r = 0; r < 32; ++r) st[r] *= v0;
            } else {
                const int exb = pow2_exponent(a.eb[col]) + efold;                          // column scale (a power of two) + fold units
                const double* ph = a.Phi + gi0 * a.ld + col;
                double* hp = a.H != nullptr ? a.H + gi0 * a.ld + col : nullptr;
                const bool rowok = gi0 + lane < a.rows;
                const int exa_l = rowok ? pow2_exponent(a.ea[gi0 + lane]) : 0;            // row scale (a power of two) / weight of row gi0 + lane
                const double rw_l = (rowok && a.rw != nullptr) ? a.rw[gi0 + lane] : 1.0;
#pragma unroll
                for (int r0 = 0; r0 < 32; r0 += 8) {                                       // 8 rows at a time: loads first
                    double phv[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) phv[r] = (gi0 + r0 + r < a.rows) ? ld_stream1(ph + (r0 + r) * a.ld) : 0.0;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int rr = r0 + r;
                        const double t = scale_pow2(st[rr], __shfl_sync(0xffffffffu, exa_l, rr) + exb);
                        const double wr = __shfl_sync(0xffffffffu, rw_l, rr);
                        double h = phv[r] * t;
                        if (gi0 + rr < a.rows) {
                            if (col == a.aug_col) {
                                a.pred[gi0 + rr] = t;
                                h = 0.0;
                            }
                            if (hp != nullptr) {
                                double o = wr * h;
                                if (a.accumulate) o += ld_stream1(hp + rr * a.ld);
                                st_stream1(hp + rr * a.ld, o);
                            }
                        } else {
                            h = 0.0;
                        }
                        st[rr] = h;
                    }
                }
            }
            // row sums over the 32 columns of the warp: butterfly multi-reduction, the total of row r ends in lane r (fixed order)
#pragma unroll
            for (int K = 16; K > 0; K >>= 1) {
                const bool up = (lane & K) != 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i >= K) continue;
                    const double give = up ? st[i] : st[i + K], keep = up ? st[i + K] : st[i];
                    st[i] = keep + __shfl_xor_sync(0xffffffffu, give, K);
                }
            }
            // the four column groups of a row live in warps w, w+4, w+8, w+12: combine through smem (fixed order)
            const int64_t gi = gi0 + lane;
            if (cg > 0) {
                rowsum_sm[(cg - 1) * 128 + rloc] = st[0];
                if (ndots > 1) rowsum_sm[384 + (cg - 1) * 128 + rloc] = second;
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (cg == 0 && gi < a.rows) {
                double* o0 = a.mode == 2 ? a.part0 : a.nupart;
                if (ndots > 0) o0[static_cast<int64_t>(nt) * a.nu_ld + gi] = ((st[0] + rowsum_sm[rloc]) + rowsum_sm[128 + rloc]) + rowsum_sm[256 + rloc];
                if (ndots > 1)
                    a.part1[static_cast<int64_t>(nt) * a.nu_ld + gi] = ((second + rowsum_sm[384 + rloc]) + rowsum_sm[512 + rloc]) + rowsum_sm[640 + rloc];
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc2(tmem_base, OM_TMEM_COLS);
}

// out[r][c] (+)= scale * scale_r[r] * scale_c[c] * sum over the chunk groups of the tile's partials, in group order (fixed).
// lower != 0: only c <= r is produced, and mirrored to out[c][r].
__global__ void __launch_bounds__(256)
ozmma_reduce_kernel(const double* __restrict__ partial, int ngroups, int ntiles, int tiles_n, int lower, int R, int C,
                    const double* __restrict__ scale_r, const double* __restrict__ scale_c, double scale, int accumulate,
                    double* __restrict__ out, int64_t ldo) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= C || r >= R) return;
    if (lower && c > r) return;
    const int mt = r >> 8, nt = c >> 7;
    int tile;
    if (!lower) tile = mt * tiles_n + nt;
    else {
        int acc = 0;
        for (int q = 0; q < mt; ++q) acc += min(tiles_n, 2 * (q + 1));
        tile = acc + nt;
    }
    const double* src = partial + static_cast<int64_t>(tile) * (256 * 128) + static_cast<int64_t>(r & 255) * 128 + (c & 127);
    const int64_t gstride = static_cast<int64_t>(ntiles) * (256 * 128);
    double sum = 0.0;
    int g = 0;
    for (; g + 4 <= ngroups; g += 4) {
        const double v0 = src[g * gstride], v1 = src[(g + 1) * gstride], v2 = src[(g + 2) * gstride], v3 = src[(g + 3) * gstride];
        sum += v0;
        sum += v1;
        sum += v2;
        sum += v3;
    }
    for (; g < ngroups; ++g) sum += src[g * gstride];
    double v = sum * scale;
    if (scale_r != nullptr) v *= scale_r[r];
    if (scale_c != nullptr) v *= scale_c[c];
    const int64_t o = static_cast<int64_t>(r) * ldo + c;
    out[o] = (accumulate ? out[o] : 0.0) + v;
    if (lower && c != r) {
        const int64_t oT = static_cast<int64_t>(c) * ldo + r;
        out[oT] = (accumulate ? out[oT] : 0.0) + v;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int g_lgroup = 2, g_prefetch = 0, g_ifold = 1;       // prefetch: +1.5 % speed but +7 GB of DRAM reads per T-GEMM (r02p capture): off

// number of lowest levels (emax, emax-1, ..) whose exact integer sum sum_i L_i 256^i fits 62 bits, |L_e| <= pairs(e) K 2^14
int int_levels(int s, int emin, int emax, int kchunk, int gchunks) {
    if (!g_ifold || gchunks != 1) return 0;
    double cum = 0.0;
    int n = 0;
    for (int e = emax; e >= emin; --e, ++n) {
        const int li = emax - e;
        const int pairs = (s < e - 1 ? s : e - 1) - (1 > e - s ? 1 : e - s) + 1;
        const double b = static_cast<double>(pairs) * kchunk * 16384.0 * ldexp(1.0, 8 * li);
        if (8 * li > 32 || cum + b >= ldexp(1.0, 62)) break;
        cum += b;
    }
    return n;
}

void set_operand_layout(OzmmaArgs& a, int mn_major) {
    a.mn_major = mn_major;
    a.lgroup = g_lgroup;
    a.prefetch = g_prefetch;
    a.kmma = 4;
    a.hintA = a.hintB = OM_EVICT_NORMAL;
}

// int8 digits addressed by 4 coordinates (inner first); strides in bytes (multiples of 16) of coordinates 1..3
int make_map(CUtensorMap* map, const int8_t* ptr, const int64_t dims[4], const int64_t strides[3], int box_rows, int mn_major = 0) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("ozmma: cuTensorMapEncodeTiled is not available from this driver");
        return GPZ_ERR_CUDA;
    }
    cuuint64_t gd[4] = {static_cast<cuuint64_t>(dims[0]), static_cast<cuuint64_t>(dims[1]), static_cast<cuuint64_t>(dims[2]),
                        static_cast<cuuint64_t>(dims[3])};
    cuuint64_t gs[3] = {static_cast<cuuint64_t>(strides[0]), static_cast<cuuint64_t>(strides[1]), static_cast<cuuint64_t>(strides[2])};
    // K-major: box = 128 K-bytes x box_rows rows; MN-major: box = box_rows row-bytes x 128 k-rows
    cuuint32_t box[4] = {128, 1, static_cast<cuuint32_t>(box_rows), 1};
    if (mn_major) {
        box[0] = static_cast<cuuint32_t>(box_rows);
        box[2] = 128;
    }
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<int8_t*>(ptr), gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          (mn_major && box_rows == 64) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("ozmma: cuTensorMapEncodeTiled failed (%d) dims %lld %lld %lld %lld strides %lld %lld %lld", static_cast<int>(r),
                  static_cast<long long>(dims[0]), static_cast<long long>(dims[1]), static_cast<long long>(dims[2]),
                  static_cast<long long>(dims[3]), static_cast<long long>(strides[0]), static_cast<long long>(strides[1]),
                  static_cast<long long>(strides[2]));
        return GPZ_ERR_CUDA;
    }
    return GPZ_OK;
}

int g_pairs_dev[128];   // per device: CTA pairs that can be co-resident (one CTA per SM); 0 = not queried yet

int resident_pairs() {
    int dev_id = 0;
    if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 128) dev_id = 0;
    if (g_pairs_dev[dev_id] > 0) return g_pairs_dev[dev_id];
    if (cudaFuncSetAttribute(ozmma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, OM_SMEM_BYTES) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(ozmma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, OM_SMEM_BYTES) != cudaSuccess) return -1;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(sms / 2 * 2));
    cfg.blockDim = dim3(OM_THREADS);
    cfg.dynamicSmemBytes = OM_SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, ozmma_kernel<0>, &cfg) != cudaSuccess || ncl <= 0) {
        cudaGetLastError();
        ncl = sms / 2;
    }
    if (ncl > sms / 2) ncl = sms / 2;
    g_pairs_dev[dev_id] = ncl;
    return ncl;
}

int count_tiles(int tiles_m, int tiles_n, int lower) {
    if (!lower) return tiles_m * tiles_n;
    int t = 0;
    for (int mt = 0; mt < tiles_m; ++mt) t += (tiles_n < 2 * (mt + 1)) ? tiles_n : 2 * (mt + 1);
    return t;
}

int launch(const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mB2, const OzmmaArgs& a, int npairs, cudaStream_t st) {
    if (a.mn_major) ozmma_kernel<1><<<dim3(static_cast<unsigned>(2 * npairs)), dim3(OM_THREADS), OM_SMEM_BYTES, st>>>(mA, mB, mB2, a);
    else ozmma_kernel<0><<<dim3(static_cast<unsigned>(2 * npairs)), dim3(OM_THREADS), OM_SMEM_BYTES, st>>>(mA, mB, mB2, a);
    GPZ_KERNEL_CHECK();
    return GPZ_OK;
}

}  // namespace

bool ozmma_available() { return encode_fn() != nullptr; }

// 2-D fp64 tensor map {cols (contiguous), rows} with row stride ld doubles and a box of box_cols x box_rows, no swizzle;
// out: 128 bytes, 64-byte aligned (a CUtensorMap).  Used by the persistent PHI kernel (gemm.cu) for its row-feature tiles.
int tensor_map_2d_f64(void* out, const double* base, int64_t cols, int64_t rows, int64_t ld, int box_cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return GPZ_ERR_CUDA;
    }
    cuuint64_t gd[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gs[1] = {static_cast<cuuint64_t>(ld) * 8};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = fn(static_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gd, gs, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (fp64 2-D) failed (%d): cols %lld rows %lld ld %lld box %d x %d", static_cast<int>(r),
                  static_cast<long long>(cols), static_cast<long long>(rows), static_cast<long long>(ld), box_cols, box_rows);
        return GPZ_ERR_CUDA;
    }
    return GPZ_OK;
}

void ozmma_set_level_group(int g) { g_lgroup = g == 1 ? 1 : 2; }
void ozmma_set_prefetch(int on) { g_prefetch = on != 0; }
void ozmma_set_int_fold(int on) { g_ifold = on != 0; }

int ozmma_pairs() { return resident_pairs(); }

// chunks per work unit: as many as keep >= ~12 rounds of units per CTA pair (load balance), at most 64
static int group_chunks(int ntiles, int nchunks, int np) {
    int g = static_cast<int>(static_cast<int64_t>(ntiles) * nchunks / (12LL * np));
    if (g < 1) g = 1;
    if (g > 64) g = 64;
    if (g > nchunks) g = nchunks;
    return g;
}

int64_t ozmma_partial_doubles(int rowsA, int rowsB, int lower, int nchunks, int pairs_limit) {
    int np = resident_pairs();
    if (np <= 0) np = 1;
    if (pairs_limit > 0 && pairs_limit < np) np = pairs_limit;
    const int nt = count_tiles((rowsA + 255) / 256, (rowsB + 127) / 128, lower);
    const int g = group_chunks(nt, nchunks, np);
    return static_cast<int64_t>(nt) * ceil_div(nchunks, g) * 256 * 128;
}

// Generic form (also the self-test entry): out[rowsA][rowsB] (+)= scale * sr[r] * sc[c] * sum_chunks sum_e 256^-(e-2) sum_{t+u=e} A_t B_u'
// mn_major == 0: A, B are int8 digits addressed {k, digit, row, chunk} with byte strides strX = {digit, row, chunk};
// mn_major == 1: addressed {row, digit, k, chunk} (the row index is the contiguous one), strides = {digit, k, chunk}.
int ozmma_gemm_nt(const int8_t* A, const int64_t strA[3], int rowsA, const int8_t* B, const int64_t strB[3], int rowsB, int s, int emax,
                  int kchunk, int nchunks, int lower, int mn_major, double* partial, const double* sr, const double* sc, double scale,
                  int accumulate, double* out, int64_t ldo, int pairs_limit, cudaStream_t st, int64_t* launches) {
    if (s < 1 || s > 8 || kchunk % 128 != 0 || kchunk <= 0 || nchunks <= 0 || emax < 2 || emax > 2 * s) {
        set_error("ozmma_gemm_nt: bad arguments (s=%d kchunk=%d nchunks=%d emax=%d)", s, kchunk, nchunks, emax);
        return GPZ_ERR_USAGE;
    }
    int maxpairs = 0;
    for (int e = 2; e <= emax; ++e) {
        const int np = (s < e - 1 ? s : e - 1) - (1 > e - s ? 1 : e - s) + 1;
        if (np > maxpairs) maxpairs = np;
    }
    if (static_cast<int64_t>(maxpairs) * kchunk * 16384 >= 2147483648LL) {
        set_error("ozmma_gemm_nt: K chunk %d too long for exact int32 accumulation of %d digit pairs", kchunk, maxpairs);
        return GPZ_ERR_USAGE;
    }
    int np = resident_pairs();                                           // also sets the dynamic smem attribute
    if (np <= 0) {
        set_error("ozmma: cannot configure the tcgen05 kernel: %s", cudaGetErrorString(cudaGetLastError()));
        return GPZ_ERR_CUDA;
    }
    if (pairs_limit > 0 && pairs_limit < np) np = pairs_limit;
    CUtensorMap mA, mB, mB2;
    const int64_t dA[4] = {kchunk, s, rowsA, nchunks}, dB[4] = {kchunk, s, rowsB, nchunks};
    const int64_t tA[4] = {rowsA, s, kchunk, nchunks}, tB[4] = {rowsB, s, kchunk, nchunks};
    int rc;
    if ((rc = make_map(&mA, A, mn_major ? tA : dA, strA, 128, mn_major))) return rc;
    if ((rc = make_map(&mB, B, mn_major ? tB : dB, strB, 64, mn_major))) return rc;
    if ((rc = make_map(&mB2, B, mn_major ? tB : dB, strB, 128, mn_major))) return rc;
    OzmmaArgs a = {};
    set_operand_layout(a, mn_major);
    a.s = s;
    a.emin = 2;
    a.emax = emax;
    a.kblocks = kchunk / 128;
    a.tiles_n = (rowsB + 127) / 128;
    a.ntiles = count_tiles((rowsA + 255) / 256, a.tiles_n, lower);
    a.nchunks = nchunks;
    a.gchunks = group_chunks(a.ntiles, nchunks, np);
    a.nint = int_levels(s, 2, emax, kchunk, a.gchunks);
    a.ngroups = static_cast<int>(ceil_div(nchunks, a.gchunks));
    a.lower = lower;
    a.mode = 0;
    a.partial = partial;
    if (emax > s + 1) a.lgroup = 1;                // a group's first step must be a two-level step (see ozmma_kernel)
    if ((rc = launch(mA, mB, mB2, a, np, st))) return rc;
    dim3 g(static_cast<unsigned>(ceil_div(rowsB, 256)), static_cast<unsigned>(rowsA));
    ozmma_reduce_kernel<<<g, 256, 0, st>>>(partial, a.ngroups, a.ntiles, a.tiles_n, lower, rowsA, rowsB, sr, sc, scale, accumulate, out,
                                           ldo);
    GPZ_KERNEL_CHECK();
    if (launches) *launches += 2;
    return GPZ_OK;
}

// T-GEMM with the fused epilogue over `rows` rows.  A8: [rows][s][MP] digits of PHI (row scales ea), B8: [s][MP][MP] digits of
// iSigma columns, B8[j][u][l] = digit u of iSigma[l][j] (column scales eb).  nupart: [MP/128][nu_ld].
int ozmma_tgemm(const int8_t* A8, const int8_t* B8, int MP, int s, int emax, int64_t rows, const double* ea, const double* eb,
                const double* Phi, int64_t ld, const double* rw, double* H, int accumulate, double* nupart, int64_t nu_ld, int aug_col,
                double* pred, cudaStream_t st, int64_t* launches) {
    if (s < 1 || s > 8 || MP % 128 != 0 || static_cast<int64_t>(s) * MP * 16384 >= 2147483648LL) {
        set_error("ozmma_tgemm: unsupported s=%d MP=%d", s, MP);
        return GPZ_ERR_USAGE;
    }
    const int np = resident_pairs();
    if (np <= 0) {
        set_error("ozmma: cannot configure the tcgen05 kernel: %s", cudaGetErrorString(cudaGetLastError()));
        return GPZ_ERR_CUDA;
    }
    CUtensorMap mA, mB, mB2;
    const int64_t dA[4] = {MP, s, rows, 1}, sA[3] = {MP, static_cast<int64_t>(s) * MP, round_up(rows * s * MP, 16)};
    const int64_t dB[4] = {MP, s, MP, 1}, sB[3] = {MP, static_cast<int64_t>(s) * MP, static_cast<int64_t>(s) * MP * MP};
    int rc;
    if ((rc = make_map(&mA, A8, dA, sA, 128))) return rc;
    if ((rc = make_map(&mB, B8, dB, sB, 64))) return rc;
    if ((rc = make_map(&mB2, B8, dB, sB, 128))) return rc;
    OzmmaArgs a = {};
    set_operand_layout(a, 0);
    a.hintB = OM_EVICT_LAST;          // the iSigma digits are read by every tile
    a.s = s;
    a.emin = 2;
    a.emax = emax;
    a.kblocks = MP / 128;
    a.tiles_n = MP / 128;
    a.ntiles = static_cast<int>(ceil_div(rows, 256)) * a.tiles_n;
    a.nchunks = 1;
    a.gchunks = 1;
    a.nint = int_levels(s, 2, emax, MP, 1);
    a.ngroups = 1;
    a.lower = 0;
    a.mode = 1;
    a.ea = ea;
    a.eb = eb;
    a.Phi = Phi;
    a.ld = ld;
    a.rows = rows;
    a.rw = rw;
    a.H = H;
    a.accumulate = accumulate;
    a.nupart = nupart;
    a.nu_ld = nu_ld;
    a.aug_col = aug_col;
    a.pred = pred;
    if ((rc = launch(mA, mB, mB2, a, np, st))) return rc;
    if (launches) ++*launches;
    return GPZ_OK;
}

// C[rows][ldc] (columns < cols) = A B' with A8 [rows][s][K128] (row scales ea), B8 [>= cols][s][K128] (row scales eb), all
// scales powers of two; K128 a multiple of 128.  One launch, results stored directly (no partial tiles).
int ozmma_gemm_rows(const int8_t* A8, const double* ea, int64_t rows, const int8_t* B8, const double* eb, int cols, int K128, int s,
                    double* C, int64_t ldc, cudaStream_t st, int64_t* launches) {
    if (s < 1 || s > 8 || K128 % 128 != 0 || K128 <= 0 || cols < 1 || static_cast<int64_t>(s) * K128 * 16384 >= 2147483648LL ||
        rows >= (1LL << 31) - 256) {
        set_error("ozmma_gemm_rows: unsupported s=%d K=%d cols=%d", s, K128, cols);
        return GPZ_ERR_USAGE;
    }
    const int np = resident_pairs();
    if (np <= 0) {
        set_error("ozmma: cannot configure the tcgen05 kernel: %s", cudaGetErrorString(cudaGetLastError()));
        return GPZ_ERR_CUDA;
    }
    const int colsp = (cols + 127) / 128 * 128;
    CUtensorMap mA, mB, mB2;
    const int64_t dA[4] = {K128, s, rows, 1}, sA[3] = {K128, static_cast<int64_t>(s) * K128, round_up(rows * s * K128, 16)};
    const int64_t dB[4] = {K128, s, colsp, 1}, sB[3] = {K128, static_cast<int64_t>(s) * K128, static_cast<int64_t>(s) * K128 * colsp};
    int rc;
    if ((rc = make_map(&mA, A8, dA, sA, 128))) return rc;
    if ((rc = make_map(&mB, B8, dB, sB, 64))) return rc;
    if ((rc = make_map(&mB2, B8, dB, sB, 128))) return rc;
    OzmmaArgs a = {};
    set_operand_layout(a, 0);
    a.hintB = OM_EVICT_LAST;
    a.s = s;
    a.emin = 2;
    a.emax = s + 1;
    a.kblocks = K128 / 128;
    a.tiles_n = colsp / 128;
    a.ntiles = static_cast<int>(ceil_div(rows, 256)) * a.tiles_n;
    a.nchunks = 1;
    a.gchunks = 1;
    a.ngroups = 1;
    a.nint = int_levels(s, 2, s + 1, K128, 1);
    a.lower = 0;
    a.mode = 3;
    a.ea = ea;
    a.eb = eb;
    a.H = C;
    a.ld = ldc;
    a.rows = rows;
    a.m = cols;
    if ((rc = launch(mA, mB, mB2, a, np, st))) return rc;
    if (launches) ++*launches;
    return GPZ_OK;
}

// PHI = exp(F W) (GPz/getPHI.m:60-113 through the monomial expansion of the quadratic forms): FD8 [rows][s][128] digits of
// the row features (row scales eaF), WD8 [MP][s][128] digits of the per-basis coefficient columns (scales ebW), K = kq <= 128.
// Fused: spare column m <- ycol, up to two row dots sum_j PHI_ij vec_q[j] as partials [MP/128][part_ld].
int ozmma_phi(const int8_t* FD8, const double* eaF, const int8_t* WD8, const double* ebW, int kq, int MP, int m, int s, int64_t rows,
              double* Phi, int ndot, const double* vec0, const double* vec1, double* part0, double* part1, int64_t part_ld,
              const double* ycol, cudaStream_t st, int64_t* launches) {
    if (s < 1 || s > 8 || MP % 128 != 0 || kq < 1 || kq > 128 || ndot < 0 || ndot > 2) {
        set_error("ozmma_phi: unsupported s=%d MP=%d kq=%d", s, MP, kq);
        return GPZ_ERR_USAGE;
    }
    const int np = resident_pairs();
    if (np <= 0) {
        set_error("ozmma: cannot configure the tcgen05 kernel: %s", cudaGetErrorString(cudaGetLastError()));
        return GPZ_ERR_CUDA;
    }
    CUtensorMap mA, mB, mB2;
    const int64_t dA[4] = {128, s, rows, 1}, sA[3] = {128, static_cast<int64_t>(s) * 128, round_up(rows * s * 128, 16)};
    const int64_t dB[4] = {128, s, MP, 1}, sB[3] = {128, static_cast<int64_t>(s) * 128, static_cast<int64_t>(s) * 128 * MP};
    int rc;
    if ((rc = make_map(&mA, FD8, dA, sA, 128))) return rc;
    if ((rc = make_map(&mB, WD8, dB, sB, 64))) return rc;
    if ((rc = make_map(&mB2, WD8, dB, sB, 128))) return rc;
    OzmmaArgs a = {};
    set_operand_layout(a, 0);
    a.hintB = OM_EVICT_LAST;
    a.s = s;
    a.emin = 2;
    a.emax = s + 1;
    a.kblocks = 1;
    a.tiles_n = MP / 128;
    a.ntiles = static_cast<int>(ceil_div(rows, 256)) * a.tiles_n;
    a.nchunks = 1;
    a.gchunks = 1;
    a.nint = int_levels(s, 2, s + 1, 128, 1);
    a.kmma = (kq + 31) / 32;
    a.ngroups = 1;
    a.lower = 0;
    a.mode = 2;
    a.ea = eaF;
    a.eb = ebW;
    a.H = Phi;
    a.ld = MP;
    a.rows = rows;
    a.m = m;
    a.ycol = ycol;
    a.ndot = ndot;
    a.vec0 = vec0;
    a.vec1 = vec1;
    a.part0 = part0;
    a.part1 = part1;
    a.nu_ld = part_ld;
    if ((rc = launch(mA, mB, mB2, a, np, st))) return rc;
    if (launches) ++*launches;
    return GPZ_OK;
}

}  // namespace gpz
