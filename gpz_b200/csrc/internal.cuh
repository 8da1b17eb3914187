// Internal declarations shared by the translation units of libgpz_b200.
#pragma once
#include <vector>

#include "../../include/gpz_b200.h"
#include "common.cuh"

namespace gpz {

enum Mode { GL = 0, VL = 1, GD = 2, VD = 3, GC = 4, VC = 5 };
__host__ __device__ inline bool mode_is_cov(int mode) { return mode >= GC; }

// theta-derived per-basis parameter arrays (device), rebuilt by prep_params at every call.
// All [..][MP] arrays are zero-padded in the basis index j >= m.
struct Params {
    int d, dp, k, m, MP, mode, het;
    int64_t g_dim, p;
    // offsets into theta
    int64_t oG, oA, oB, oV, oT;
    double* Pt;     // [d][MP]      P(j,a)
    double* Gt;     // [d][MP]      diag modes: Gamma(j,a)
    double* Ct;     // [d][MP]      diag modes: Gamma(j,a)*P(j,a);  cov modes: (Gamma_j p_j)_b
    double* Gam;    // [d][dp][MP]  cov modes: Gamma_j(b,a) at ((b*dp+a)*MP + j), zero for a>=d
    double* Aj;     // [d][d][MP]   cov modes: (Gamma_j' Gamma_j)(a,b)  at ((a*d+b)*MP+j)
    double* Sj;     // [d][d][MP]   cov modes: inverse of Aj (only built when Psi is present)
    double* lndS;   // [MP]         cov modes: ln det Sigma_j = -ln det Aj
    double* alpha;  // [k][MP]
    double* v;      // [k][MP]
    double* tau;    // [k][MP]
    double* bk;     // [k]
    // fast path (no Psi, no NaN): PHI = exp(F W) on the fp64 tensor pipe
    int q, KQ, QP;  // feature count, K extent (multiple of 16), row stride of F (multiple of 32)
    double* Wc;     // [npat][KQ][MP]  coefficients of the monomial features (zero rows/columns in the padding)
    double* xshift; // [d]          constant subtracted from X at upload (and from P here): only x - p matters
    // missing-input patterns of the covariance modes (getPHI.m:43-54,76; GPz.m:151-159); npat = 1 and all
    // dims observed when the data have no NaN
    int npat;
    unsigned char* obs;   // [npat][d]          1 = dim observed in this pattern
    double* Mg;     // [npat][d*d][MP]  (Sigma_j(o,o))^-1 embedded in d x d (zero on missing rows/cols)
    double* Gg;     // [npat][d*d][MP]  iSigma(u,u)^-1 iSigma(u,o) embedded (row e in u, col b in o)
    double* lndM;   // [npat][MP]       ln det of the marginal precision = -ln det Sigma_j(o,o)
};

struct RowData {          // one resident row set (training or validation rows of this rank)
    int64_t n = 0;
    double* X = nullptr;      // [d][n]  (column-major n x d)
    double* Y = nullptr;      // [k][n]
    double* omega = nullptr;  // [n]
    double* Psi = nullptr;    // diag: [d][n]; cov: [n][d*d] (MATLAB d x d x n)
    int has_nan = 0;
    double* F = nullptr;      // [n][QP] monomial row features (fast path only)
    const double* ycol = nullptr;   // set (to Y) when PHI's spare column m should carry y (see api.cu "aug")
    // int8 PHI build (ozaki.cu ozaki_phi): digits of F (dataset constants) and per-call scratch for the digits of W
    int8_t* FD8 = nullptr;
    double* eaF = nullptr;
    int8_t* WD8 = nullptr;
    double* ebW = nullptr;
    int phi_digits = 0;
    int* flag = nullptr;
    // GC + Psi fast path (gcpsi.cu): per-evaluation row features [gc_chunk][KQ], basis tables W [KQ][MP], G [MP][KQ]
    double *gcF = nullptr, *gcW = nullptr, *gcG = nullptr;
    const double* gc_ones = nullptr;
    int64_t gc_chunk = 0;
    // ... with its K = 1 + d + d(d+1)/2 GEMMs on the int8 tensor cores (d > 16): digit buffers (aliases of the context's
    // digit workspaces, free at those points of the evaluation) and the digits of the basis tables
    int gc_digits = 0;            // 0: fp64 DMMA GEMMs
    int8_t* gcA8 = nullptr;       // [rows_pad][s][K128] row digits of the left operand (features, then dPHI)
    double* gcEa = nullptr;       // [rows_pad]
    int8_t* gcWD8 = nullptr;      // [MP][s][K128]   digits of the columns of gcW
    double* gcEbW = nullptr;      // [MP]
    int8_t* gcGD8 = nullptr;      // [KN][s][MP]     digits of the columns of gcG
    double* gcEbG = nullptr;      // [KN]
    int8_t* gcF8 = nullptr;       // [rows_pad][s][K128] row digits of the features, for the moment GEMM dPHI' F
    double* gcEaF = nullptr;      // [rows_pad]
    int8_t* gcX8 = nullptr;       // [rows_pad][s][MP]   digits of dPHI 2^(G_i - Eb) (alias of the second PHI digit buffer)
    void* gcMws = nullptr;        // oz_moment_workspace_bytes
    // covariance modes with missing inputs: rows are stored sorted by NaN pattern (NaN entries zero-filled),
    // group g = rows [g_r0[g], g_r1[g]) with pattern g_pat[g]; perm[sorted position] = position in selection order
    std::vector<int64_t> g_r0, g_r1, perm;
    std::vector<int> g_pat;
};

struct DotSpec {          // up to 2 fused row-dots  out_q[i] = sum_j PHI_ij vec_q[j]
    int n;
    const double* vec[2];
    double* out[2];
};

// ---- phi.cu
int prep_params(const double* d_theta, const Params& P, int need_sigma, cudaStream_t st, int64_t* launches);
// dot_scratch: >= 2 * (MP/128) * (r1-r0) doubles (row-dot partials of the tensor-core path)
int phi_build(const Params& P, const RowData& R, int64_t r0, int64_t r1, double* Phi /*[rows][MP] at row r0 -> index 0*/,
              const DotSpec& dots, double* dot_scratch, cudaStream_t st, int64_t* launches);
// N = PHI .* exp(lnN - lnPHI): the normalised basis densities, getPHI's 4th output (getPHI.m:77,87,98,105,114)
int phi_to_density(const Params& P, const RowData& R, int64_t r0, int64_t r1, const double* Phi, double* N, cudaStream_t st,
                   int64_t* launches);
int rowdot(const double* Phi, int64_t ld, int m, int64_t n, const DotSpec& dots, cudaStream_t st, int64_t* launches);
int dxy_device(const double* X, int64_t n, const double* Y, int m, int d, double* D, cudaStream_t st);
int dxy_colmean_device(const double* X, int64_t n, const double* Y, int m, int d, double* part, double* mean, cudaStream_t st);
int64_t dxy_colmean_chunks(int64_t n);
int transpose_out(const double* src_rowmajor, int64_t ld, int64_t n, int m, double* dst_colmajor, cudaStream_t st);

// ---- gemm.cu
extern int g_gemm_warps;
extern int g_phi_persist;
extern int g_prep_block;
extern int g_moment_warps;
int gram_nsplit(int MP, int sm_count);
int gram_syrk_main(const double* Phi, int64_t ld, int MP, const double* wgt, int64_t row0, int64_t row1, int nsplit,
                   double* partial, int accumulate, cudaStream_t st, int64_t* launches);
int gram_syrk_finish(const double* partial, int nsplit, int MP, int reduce, double* S, cudaStream_t st, int64_t* launches);
int atb_general(const double* A, int64_t lda, int MP, const double* B, int64_t ldb, int QP, const double* wgt,
                int64_t row0, int64_t row1, int nsplit, double* partial, int accumulate, int reduce, double* R,
                cudaStream_t st, int64_t* launches);
int tgemm(const double* Phi, int64_t ld, const double* Sinv, int MP, int m, int64_t n, const double* rw, double* H,
          int accumulate, double* nupart, int64_t nu_ld, double* pred_aug, cudaStream_t st, int64_t* launches);
int phi_gemm(const double* F, int64_t ldf, int kq, int kvalid, const double* W, int MP, int m, int64_t n, double* Phi, int ndot,
             const double* vec0, const double* vec1, double* part0, double* part1, int64_t part_ld, const double* ycol,
             cudaStream_t st, int64_t* launches);
int gemm_rows(const double* A, int64_t lda, int K, const double* B, int N, int64_t n, double* C, cudaStream_t st, int64_t* launches);
int atb_dphi(const double* Phi, const double* H, int64_t ld, int MP, const double* F, int QP, int q, const double* cw,
             const double* dbeta, const double* w, const double* v, int64_t row0, int64_t row1, int nsplit, double* partial,
             double* colp, int accumulate, int col_accumulate, int reduce, double* R, cudaStream_t st, int64_t* launches);
int sgemm(int M, int N, int K, double alpha, const double* A, int64_t sAi, int64_t sAk, const double* B, int64_t sBk,
          int64_t sBj, double beta, double* C, int64_t ldc, int lower_only, cudaStream_t st, int64_t* launches);

int sgemm_batched(int M, int N, int K, double alpha, const double* A, int64_t sAi, int64_t sAk, int64_t bsA, const double* B,
                  int64_t sBk, int64_t sBj, int64_t bsB, double beta, double* C, int64_t ldc, int64_t bsC, int batches, int mlim,
                  int mstep, int k_follows_m, cudaStream_t st, int64_t* launches);

// ---- solve.cu
struct SolveWs {
    double* W = nullptr;      // [MP][MP] L^{-1}
    double* Linv = nullptr;   // [nb][64*64] inverses of the diagonal blocks
    double* tmp = nullptr;    // [64][MP]
    int* flag = nullptr;      // device int: !=0 -> non-positive pivot
    // look-ahead factorisation (solve.cu): L in its own matrix, panel GEMMs on a side stream, events between the two chains
    static constexpr int MAXBLK = 64;
    double* Lbuf = nullptr;   // [MP][MP]
    cudaStream_t side = nullptr;
    cudaEvent_t evP[MAXBLK] = {};
    cudaEvent_t evT[MAXBLK] = {};
    // incremental inverse (solve_lookahead = 2): a third chain builds L^-1 and Sinv = W'W block row by block row behind the factorisation
    cudaStream_t inv = nullptr, acc = nullptr;   // block rows of W; their rank-64 updates of Sinv
    cudaEvent_t evL[MAXBLK] = {};   // panel solve k done: column panel k of L is complete
    cudaEvent_t evW[MAXBLK] = {};   // block row k of W done
    cudaEvent_t evF = nullptr, evI = nullptr;   // fork / join of the third chain
};
extern int g_solve_lookahead;
int solve_ws_alloc(SolveWs& ws, int MP);
void solve_ws_free(SolveWs& ws);
// S (MP x MP row-major, lower triangle read, overwritten by L) -> Sinv (full symmetric), *d_logdet
int spd_inverse(double* S, int m, int MP, double* Sinv, double* d_logdet, SolveWs& ws, cudaStream_t st,
                int64_t* launches);

// reference semantics of inv_logdet.m (SVD pseudo-inverse with truncation) for the exported gpz_inv_logdet: one-sided Jacobi
int svd_pinv_logdet(double* G, int m, int MP, double* V, double* Xi, double* h_logdet, int* d_counter, double* d_vec, cudaStream_t st,
                    int64_t* launches);
int chol_diag_minmax(const SolveWs& ws, const double* S, int m, int MP, double* d_out2, cudaStream_t st);

// ---- backproj.cu
int build_features(const Params& P, const double* X, int64_t n, int64_t r0, int64_t r1, double* F, cudaStream_t st,
                   int64_t* launches);
int feature_count(const Params& P);
// moments of one pattern group -> dP / per-basis dGamma (in `full`), accumulated when accumulate != 0
int moments_to_grad(const Params& P, int pat, const double* Rm /*[MP][QP]*/, int QP, double* dP /*m*d*/, double* full,
                    int accumulate, cudaStream_t st, int64_t* launches);
int mode_reduce(const Params& P, const double* full, double* dG, cudaStream_t st, int64_t* launches);
int backproj_diag_generic(const Params& P, const RowData& R, int64_t r0, int64_t r1, const double* dPhi, int64_t ld,
                          double* partial, int nslab, int accumulate, cudaStream_t st, int64_t* launches);
int backproj_diag_generic_finish(const Params& P, const double* partial, int nslab, double* dP, double* dG,
                                 double* scratch, cudaStream_t st, int64_t* launches);
int backproj_cov_psi(const Params& P, const RowData& R, int64_t r0, int64_t r1, const double* dPhi, int64_t ld,
                     double* partial, int nslab, int accumulate, cudaStream_t st, int64_t* launches);
int backproj_cov_psi_finish(const Params& P, const RowData& R, const double* partial, int nslab, double* dP, double* dG, double* scratch,
                            cudaStream_t st, int64_t* launches);
int64_t backproj_partial_doubles(const Params& P, int nslab, int has_psi, int has_nan);

// ---- gcpsi.cu: GC + Psi through one d x d factorisation per row and GEMMs
int gc_feature_width(int d);
int gc_features(const Params& P, const RowData& R, int64_t r0, int64_t r1, cudaStream_t st, int64_t* launches);
int64_t gc_backproj_ws_doubles(const Params& P, int64_t chunk_rows, int nsplit, int sm_count);
int gc_backproj(const Params& P, const RowData& R, int64_t r0, int64_t r1, const double* dPhi, int64_t ld, double* ws, int nsplit,
                int sm_count, int accumulate, int last, cudaStream_t st, int64_t* launches);
int gc_backproj_finish(const Params& P, const RowData& R, double* ws, int nsplit, int sm_count, double* dP, double* dG, cudaStream_t st,
                       int64_t* launches);

// ---- predict.cu
int predict_noisy_diag(const Params& P, const RowData& R, const double* w /*[k][MP]*/, const double* Sinv /*[k][MP][MP]*/,
                       const double* ElnS /*[k][n] = b + PHI v*/, const double* mu /*[k][n]*/, double* nu, double* beta_i,
                       double* gamma, cudaStream_t st, int64_t* launches);

int predict_noisy_cov(const Params& P, const RowData& R, const double* w, const double* Sinv, const double* ElnS,
                      const double* mu, double* nu, double* beta_i, double* gamma, cudaStream_t st, int64_t* launches);
int predict_missing_diag(const Params& P, const double* X, const double* Psi, int64_t n, const unsigned char* ob,
                         const double* prior, const double* w, const double* Sinv, double* mu, double* nu, double* beta_i,
                         double* gamma, double* Phi, cudaStream_t st, int64_t* launches);

int predict_missing_cov(const Params& P, const double* X, const double* Psi, int64_t n, const unsigned char* ob, const double* prior,
                        const double* w, const double* Sinv, double* mu, double* nu, double* beta_i, double* gamma, double* Phi,
                        cudaStream_t st, int64_t* launches);

// ---- ozaki.cu: fp64 operands -> base-256 digits; the two n x m x m products through ozmma.cu
int64_t oz_padded_rows(int64_t rows);
int64_t oz_digit_bytes(int MP, int s, int64_t rows);
int64_t oz_workspace_bytes(int MP, int s);
int64_t oz_gram_workspace_bytes(int MP, int64_t rows);
int ozaki_digits(const double* Phi, int64_t ld, int MP, int m, int64_t rows, int s, const double* wgt, const double* d_scal, int aug,
                 int8_t* D8, int8_t* F8, double* ea, int* flag, cudaStream_t st, int64_t* launches);
int ozaki_row_digits(const double* A, int64_t lda, int cols, int K128, int64_t rows, int s, int8_t* A8, double* ea, int* flag,
                     cudaStream_t st, int64_t* launches);
int ozaki_transpose_digits(const double* src, int64_t lds, int R, int C, int Rpad, int Cpad, int s, int8_t* out, double* scale, int* flag,
                           cudaStream_t st, int64_t* launches);
int64_t oz_moment_workspace_bytes(int MP, int cols, int64_t rows);
int ozaki_moment_gemm(const double* X, int64_t ld, int m, int MP, int64_t rows, const double* eaX, const int8_t* F8, const double* eaF,
                      int K128, int cols, int s, int8_t* A8, int accumulate, double* R, int64_t ldr, void* ws, cudaStream_t st,
                      int64_t* launches);
int exp_rows_inplace(double* Phi, int64_t ld, int m, int MP, int64_t rows, const DotSpec& dots, cudaStream_t st, int64_t* launches);
int ozaki_feature_digits(const double* F, int64_t ldf, int q, int64_t rows, int s, int8_t* FD8, double* eaF, int* flag, cudaStream_t st,
                         int64_t* launches);
int ozaki_phi(const int8_t* FD8, const double* eaF, const double* W, int kq, int MP, int m, int s, int64_t rows, int8_t* WD8, double* ebW,
              double* Phi, int ndot, const double* vec0, const double* vec1, double* part0, double* part1, int64_t part_ld,
              const double* ycol, int* flag, cudaStream_t st, int64_t* launches);
// S (+)= PHI' diag(wgt) PHI over `rows` rows through the int8 tensor cores (digits from ozaki_digits)
int ozaki_gram(const int8_t* F8, const int8_t* D8, int MP, int m, int64_t rows, int s, int gs, const double* d_scal, int aug,
               int accumulate, double* S, void* ws, const int* flag, cudaStream_t st, int64_t* launches);
int ozaki_tgemm(const double* Phi, int64_t ld, const int8_t* D8, const double* ea, const double* Sinv, int MP, int m, int64_t n, int s,
                const double* rw, double* H, int accumulate, double* nupart, int64_t nu_ld, const double* waug, double* pred, void* ws,
                cudaStream_t st, cudaEvent_t tev0, cudaEvent_t tev1, int64_t* launches);

}  // namespace gpz

namespace gpz {
// ---- ozmma.cu: hand-written tcgen05 (cta_group::2, TMA, TMEM) digit-level GEMM with on-chip level folding
bool ozmma_available();
int tensor_map_2d_f64(void* out /*128 B, 64-byte aligned*/, const double* base, int64_t cols, int64_t rows, int64_t ld, int box_cols, int box_rows);
void ozmma_set_prefetch(int on);
void ozmma_set_int_fold(int on);     // 1 (default): lowest levels folded exactly in int64 where a unit is one level sweep
void ozmma_set_level_group(int g);   // 2 (default): two levels share their operand tiles; 1: one level at a time (A/B measurements)
int ozmma_pairs();
int64_t ozmma_partial_doubles(int rowsA, int rowsB, int lower, int nchunks, int pairs_limit);
int ozmma_gemm_nt(const int8_t* A, const int64_t strA[3], int rowsA, const int8_t* B, const int64_t strB[3], int rowsB, int s, int emax,
                  int kchunk, int nchunks, int lower, int mn_major, double* partial, const double* sr, const double* sc, double scale,
                  int accumulate, double* out, int64_t ldo, int pairs_limit, cudaStream_t st, int64_t* launches);
int ozmma_gemm_rows(const int8_t* A8, const double* ea, int64_t rows, const int8_t* B8, const double* eb, int cols, int K128, int s,
                    double* C, int64_t ldc, cudaStream_t st, int64_t* launches);
int ozmma_phi(const int8_t* FD8, const double* eaF, const int8_t* WD8, const double* ebW, int kq, int MP, int m, int s, int64_t rows,
              double* Phi, int ndot, const double* vec0, const double* vec1, double* part0, double* part1, int64_t part_ld,
              const double* ycol, cudaStream_t st, int64_t* launches);
int ozmma_tgemm(const int8_t* A8, const int8_t* B8, int MP, int s, int emax, int64_t rows, const double* ea, const double* eb,
                const double* Phi, int64_t ld, const double* rw, double* H, int accumulate, double* nupart, int64_t nu_ld, int aug_col,
                double* pred, cudaStream_t st, int64_t* launches);
// train.cu: minFunc's L-BFGS + Wolfe line search + GPz/callBack.m, device-resident (see gpz_train in gpz_b200.h)
int lbfgs_train(int64_t p, gpz_objective_dev fn, void* fn_user, const gpz_train_options* opt, double* theta,
                double* best_theta, double* best_valid, gpz_train_callback cb, void* user, gpz_train_result* res,
                cudaStream_t st, int64_t* launches);

}  // namespace gpz
