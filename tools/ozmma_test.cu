// Stand-alone check + timing of the hand-written tcgen05 digit GEMM (gpz_b200/csrc/ozmma.cu) against an int64 CPU reference.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tools/ozmma_test.cu -Igpz_b200/csrc -Lgpz_b200 -lgpz_b200
//          -Xlinker -rpath -Xlinker '$ORIGIN/../gpz_b200' -o build/ozmma_test
//   run:   build/ozmma_test            (small exact checks, then a T-GEMM sized timing)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "internal.cuh"

using namespace gpz;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

static uint32_t rng_state = 12345u;
static inline uint32_t rnd() {
    rng_state = rng_state * 1664525u + 1013904223u;
    return rng_state >> 8;
}

// digits laid out [chunk][row][digit][k]  (mn: [chunk][k][digit][row])
static int run_case(int rowsA, int rowsB, int s, int emax, int kchunk, int nchunks, int lower, int pairs_limit, int mn = 0) {
    const int64_t szA = static_cast<int64_t>(nchunks) * rowsA * s * kchunk, szB = static_cast<int64_t>(nchunks) * rowsB * s * kchunk;
    std::vector<int8_t> hA(szA), hB(szB);
    for (auto& v : hA) v = static_cast<int8_t>(static_cast<int>(rnd() % 256) - 128);
    for (auto& v : hB) v = static_cast<int8_t>(static_cast<int>(rnd() % 256) - 128);
    if (lower) hB = hA;      // Gram-like: symmetric result
    int8_t *dA, *dB;
    CK(cudaMalloc(&dA, szA));
    CK(cudaMalloc(&dB, szB));
    CK(cudaMemcpy(dA, hA.data(), szA, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), szB, cudaMemcpyHostToDevice));
    const int64_t np = ozmma_partial_doubles(rowsA, rowsB, lower, nchunks, pairs_limit);
    double *dP, *dO;
    CK(cudaMalloc(&dP, np * 8));
    CK(cudaMalloc(&dO, static_cast<int64_t>(rowsA) * rowsB * 8));
    CK(cudaMemset(dO, 0xff, static_cast<int64_t>(rowsA) * rowsB * 8));
    int64_t strA[3] = {kchunk, static_cast<int64_t>(s) * kchunk, static_cast<int64_t>(rowsA) * s * kchunk};
    int64_t strB[3] = {kchunk, static_cast<int64_t>(s) * kchunk, static_cast<int64_t>(rowsB) * s * kchunk};
    if (mn) {
        strA[0] = rowsA; strA[1] = static_cast<int64_t>(s) * rowsA;
        strB[0] = rowsB; strB[1] = static_cast<int64_t>(s) * rowsB;
    }
    int64_t launches = 0;
    int rc = ozmma_gemm_nt(dA, strA, rowsA, dB, strB, rowsB, s, emax, kchunk, nchunks, lower, mn, dP, nullptr, nullptr, 1.0, 0, dO, rowsB,
                           pairs_limit, 0, &launches);
    if (rc) {
        printf("ozmma_gemm_nt failed rc=%d: %s\n", rc, gpz_last_error());
        return 1;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("kernel failed: %s\n", cudaGetErrorString(e));
        return 1;
    }
    std::vector<double> out(static_cast<int64_t>(rowsA) * rowsB);
    CK(cudaMemcpy(out.data(), dO, out.size() * 8, cudaMemcpyDeviceToHost));
    // reference on a sample of entries
    int bad = 0, checked = 0;
    double maxrel = 0.0;
    for (int trial = 0; trial < 4000; ++trial) {
        int r = rnd() % rowsA, c = rnd() % rowsB;
        if (trial < 8) { r = (trial & 1) ? rowsA - 1 : 0; c = (trial & 2) ? rowsB - 1 : 0; }
        if (lower && c > r) { const int t = r; r = c; c = t; }
        double ref = 0.0;
        for (int ch = 0; ch < nchunks; ++ch)
            for (int e = emax; e >= 2; --e) {
                long long acc = 0;
                for (int t = 1; t <= s; ++t) {
                    const int u = e - t;
                    if (u < 1 || u > s) continue;
                    if (!mn) {
                        const int8_t* pa = hA.data() + ((static_cast<int64_t>(ch) * rowsA + r) * s + (t - 1)) * kchunk;
                        const int8_t* pb = hB.data() + ((static_cast<int64_t>(ch) * rowsB + c) * s + (u - 1)) * kchunk;
                        for (int k = 0; k < kchunk; ++k) acc += static_cast<int>(pa[k]) * static_cast<int>(pb[k]);
                    } else {
                        for (int k = 0; k < kchunk; ++k)
                            acc += static_cast<int>(hA[((static_cast<int64_t>(ch) * kchunk + k) * s + (t - 1)) * rowsA + r]) *
                                   static_cast<int>(hB[((static_cast<int64_t>(ch) * kchunk + k) * s + (u - 1)) * rowsB + c]);
                    }
                }
                ref += static_cast<double>(acc) * ldexp(1.0, -8 * (e - 2));
            }
        const double got = out[static_cast<int64_t>(r) * rowsB + c];
        const double rel = fabs(got - ref) / (fabs(ref) + 1e-300);
        if (!(rel < 1e-13)) {
            if (bad < 10) printf("  mismatch at (%d,%d): got %.17g ref %.17g\n", r, c, got, ref);
            ++bad;
        }
        if (rel > maxrel) maxrel = rel;
        ++checked;
        if (lower) {
            const double gotT = out[static_cast<int64_t>(c) * rowsB + r];
            if (gotT != got) { if (bad < 10) printf("  mirror mismatch at (%d,%d)\n", r, c); ++bad; }
        }
    }
    printf("case rowsA=%d rowsB=%d s=%d emax=%d kchunk=%d nchunks=%d lower=%d pairs=%d mn=%d: %d/%d bad, max rel %.2e\n", rowsA, rowsB, s,
           emax, kchunk, nchunks, lower, pairs_limit, mn, bad, checked, maxrel);
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dP);
    cudaFree(dO);
    return bad != 0;
}

__global__ void fill_hash(int8_t* p, int64_t n, uint32_t seed, int mode) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        uint32_t h = static_cast<uint32_t>(i) * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = mode ? static_cast<int8_t>(h & 255) : static_cast<int8_t>(0x11);
    }
}

static void time_tgemm(int64_t rows, int MP, int s, int random) {
    const int64_t szA = rows * s * MP, szB = static_cast<int64_t>(MP) * s * MP;
    int8_t *dA, *dB;
    CK(cudaMalloc(&dA, szA));
    CK(cudaMalloc(&dB, szB));
    fill_hash<<<1184, 256>>>(dA, szA, 1u, random);
    fill_hash<<<1184, 256>>>(dB, szB, 7u, random);
    double *ea, *eb, *Phi, *H, *nup, *pred;
    CK(cudaMalloc(&ea, rows * 8));
    CK(cudaMalloc(&eb, MP * 8));
    CK(cudaMalloc(&Phi, rows * MP * 8));
    CK(cudaMalloc(&H, rows * MP * 8));
    CK(cudaMalloc(&nup, 2 * (MP / 128) * rows * 8));
    CK(cudaMalloc(&pred, rows * 8));
    CK(cudaMemset(ea, 0, rows * 8));
    CK(cudaMemset(eb, 0, MP * 8));
    CK(cudaMemset(Phi, 0, rows * MP * 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int64_t launches = 0;
    for (int rep = 0; rep < 4; ++rep) {
        if (rep == 1) cudaEventRecord(e0, 0);
        int rc = ozmma_tgemm(dA, dB, MP, s, s + 1, rows, ea, eb, Phi, MP, nullptr, H, 0, nup, rows, MP - 1, pred, 0, &launches);
        if (rc) {
            printf("ozmma_tgemm failed: %s\n", gpz_last_error());
            return;
        }
    }
    cudaEventRecord(e1, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("timing kernel failed: %s\n", cudaGetErrorString(e));
        return;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 3;
    const double ops = 2.0 * rows * MP * static_cast<double>(MP) * (s * (s + 1) / 2);
    printf("tgemm random=%d rows=%lld MP=%d s=%d: %.3f ms  -> %.1f TOP/s int8, fp64-equivalent %.1f TFLOP/s\n", random, static_cast<long long>(rows), MP, s, ms,
           ops / ms * 1e-9, 2.0 * rows * MP * static_cast<double>(MP) / ms * 1e-9);
}

static void time_gram(int64_t nrows, int MP, int s, int kchunk, int mn = 0) {
    const int nch = static_cast<int>((nrows + kchunk - 1) / kchunk);
    const int64_t sz = static_cast<int64_t>(nch) * MP * s * kchunk;
    int8_t *dA, *dB;
    CK(cudaMalloc(&dA, sz));
    CK(cudaMalloc(&dB, sz));
    fill_hash<<<1184, 256>>>(dA, sz, 3u, 1);
    fill_hash<<<1184, 256>>>(dB, sz, 5u, 1);
    const int64_t np = ozmma_partial_doubles(MP, MP, 1, nch, 0);
    double *dP, *dO;
    CK(cudaMalloc(&dP, np * 8));
    CK(cudaMalloc(&dO, static_cast<int64_t>(MP) * MP * 8));
    int64_t str[3] = {kchunk, static_cast<int64_t>(s) * kchunk, static_cast<int64_t>(MP) * s * kchunk};
    if (mn) { str[0] = MP; str[1] = static_cast<int64_t>(s) * MP; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int64_t launches = 0;
    for (int rep = 0; rep < 3; ++rep) {
        if (rep == 1) cudaEventRecord(e0, 0);
        int rc = ozmma_gemm_nt(dA, str, MP, dB, str, MP, s, s + 1, kchunk, nch, 1, mn, dP, nullptr, nullptr, 1.0, 0, dO, MP, 0, 0, &launches);
        if (rc) {
            printf("gram failed: %s\n", gpz_last_error());
            return;
        }
    }
    cudaEventRecord(e1, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("gram kernel failed: %s\n", cudaGetErrorString(e));
        return;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 2;
    printf("gram mn=%d rows=%lld MP=%d s=%d kchunk=%d (partials %.2f GB): %.3f ms -> fp64-equivalent %.1f TFLOP/s (2 n m^2)\n",
           mn, static_cast<long long>(nrows), MP, s, kchunk, np * 8e-9, ms, 2.0 * nrows * MP * static_cast<double>(MP) / ms * 1e-9);
    cudaFree(dA); cudaFree(dB); cudaFree(dP); cudaFree(dO);
}

int main(int argc, char** argv) {
    printf("ozmma available: %d, resident pairs: %d\n", static_cast<int>(ozmma_available()), ozmma_pairs());
    int fail = 0;
    fail |= run_case(256, 128, 1, 2, 128, 1, 0, 1);          // one tile, one digit pair, one k-block
    if (!fail) fail |= run_case(256, 128, 1, 2, 512, 1, 0, 1);   // 4 k-blocks
    if (!fail) fail |= run_case(256, 128, 2, 3, 256, 1, 0, 1);   // 2 levels
    if (!fail) fail |= run_case(512, 256, 3, 4, 256, 3, 0, 2);   // several tiles, chunks, pipeline wrap
    if (!fail) fail |= run_case(500, 200, 3, 4, 256, 2, 0, 0);   // ragged rows (TMA zero fill), all pairs
    if (!fail) fail |= run_case(1024, 1024, 7, 8, 1024, 1, 0, 0);
    if (!fail) fail |= run_case(1024, 1024, 3, 4, 2048, 5, 1, 0);   // Gram-like: lower tiles
    if (!fail) fail |= run_case(512, 512, 2, 3, 256, 300, 1, 0);    // many chunks: grouped units + ordered reduction
    if (!fail) fail |= run_case(256, 128, 1, 2, 128, 1, 0, 1, 1);       // MN-major operands: one tile, one k-block
    if (!fail) fail |= run_case(256, 128, 2, 3, 512, 2, 0, 1, 1);
    if (!fail) fail |= run_case(1024, 1024, 3, 4, 1024, 7, 1, 0, 1);    // Gram layout
    if (!fail) fail |= run_case(1024, 1024, 7, 8, 1024, 3, 1, 0, 1);
    if (fail) {
        printf("FAILED\n");
        return 1;
    }
    printf("all exact checks passed\n");
    if (argc > 1 && strcmp(argv[1], "time") == 0) {
        time_tgemm(131072, 1024, 7, 0);
        time_tgemm(131072, 1024, 7, 1);
        time_tgemm(1000192, 1024, 7, 1);
        time_gram(1000000, 1024, 7, 1024);
        time_gram(1000000, 1024, 7, 1024, 1);
        time_gram(1000000, 1024, 6, 1024, 1);
        time_gram(1000000, 512, 7, 1024, 1);
    }
    return 0;
}
