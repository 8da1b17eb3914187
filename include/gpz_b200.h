/*
 * gpz_b200.h -- C ABI of libgpz_b200.so: the B200 (sm_100a) implementation of GPz's
 * marginal-likelihood objective / gradient / fit / predict hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI for this path (it is
 * pure MATLAB); the plug point is the closure
 *     f = @(params) GPz(params,model,X,Y,Psi,omega,training,validation)      GPz/train.m:40, GPz/init.m:89
 * and the only FFI precedent in the tree is the MEX gateway convention of
 *     minFunc_2012/minFunc/mex/lbfgsProdC.c:7-44   (mexFunction, mxGetPr in, mxCreateDoubleMatrix out).
 * Each entry point below names the reference function (file:line) it replaces.  The MEX gateway a
 * maintainer adds on the MATLAB side is matlab/gpz_b200_mex.cpp.
 *
 * Conventions
 *   - all matrices are fp64, COLUMN-MAJOR (MATLAB layout), passed as plain host pointers unless
 *     the function name ends in _dev; the library never writes to its inputs and keeps no host
 *     pointer after a call returns (the dataset is copied to the device in gpz_create);
 *   - X is already z-scored, Y centred and Psi normalised by fixPsi (GPz/train.m:30-38);
 *   - NaN in X means "missing" (GPz/getPHI.m:43-54);
 *   - return value: 0 = ok.  Numerical trouble (a non-positive Cholesky pivot of SIGMA) is NOT an
 *     error: the call returns 0 and the outputs are NaN, which minFunc's isLegal tests handle
 *     (minFunc_2012/minFunc/WolfeLineSearch.m:53).  Non-zero = usage / CUDA / NCCL error; the text is
 *     in gpz_last_error().
 *   - single caller: one host thread per context (MATLAB calls MEX serially).
 *   - there is NO CPU fallback: every entry point needs a CUDA device of compute capability 10.x.
 */
#ifndef GPZ_B200_H
#define GPZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpz_ctx gpz_ctx;

/* model struct fields that parameterise the kernels (GPz/init.m:16-20,86) */
typedef struct gpz_model {
    int32_t d;                /* input dimension                                  */
    int32_t k;                /* number of outputs                                */
    int32_t m;                /* number of basis functions                        */
    char    method[4];        /* "GL","VL","GD","VD","GC","VC" (NUL padded)       */
    int32_t heteroscedastic;  /* 0/1                                              */
} gpz_model;

/* status codes */
enum {
    GPZ_OK = 0,
    GPZ_ERR_USAGE = 1,        /* bad argument / unsupported combination           */
    GPZ_ERR_CUDA = 2,
    GPZ_ERR_NCCL = 3,
    GPZ_ERR_NODEVICE = 4
};

const char* gpz_last_error(void);
int gpz_version(void);

/* length of theta for a model: [P(:); Gamma(:); lnAlpha(:); b(:); v(:); lnTau(:)]  GPz/init.m:87,97 */
int64_t gpz_theta_len(const gpz_model* model);
int64_t gpz_g_dim(const gpz_model* model);

/* ---- dataset context: replaces the closure capture at GPz/train.m:40 ---------------------------
 * X n_all x d, Y n_all x k, omega n_all (NULL = ones, GPz.m:20-22), training/validation logical
 * n_all (NULL training = all rows, GPz.m:16-18; NULL validation = none, GPz.m:239).
 * Psi: NULL, or n_all x d (methods ?L/?D) or d x d x n_all (methods ?C), as GPz/fixPsi.m returns it.
 * The rows selected by the masks are gathered once and stay resident in HBM.
 * device: CUDA ordinal to use.                                                                   */
int gpz_create(gpz_ctx** out, const gpz_model* model, int64_t n_all,
               const double* X, const double* Y, const double* Psi, const double* omega,
               const uint8_t* training, const uint8_t* validation, int device);
void gpz_destroy(gpz_ctx* ctx);

/* ---- multi-GPU: one process per GPU, rows sharded, NCCL sum-allreduce inside eval/fit ------------
 * Every rank calls gpz_create on ITS row shard, then gpz_comm_init with the same 128-byte id that
 * rank 0 obtained from gpz_comm_unique_id (distribute it with any host-side channel).            */
int gpz_comm_unique_id(char id[128]);
int gpz_comm_init(gpz_ctx* ctx, int rank, int world, const char id[128]);

/* ---- multi-GPU from ONE process / one caller thread (the reference's caller is a single interpreter thread holding
 * the closure of GPz/train.m:40; SURVEY.md 8b "threading", 8e "backend"): the rows selected by the masks are split into
 * ngpus contiguous blocks, one context per device, one NCCL communicator over all of them; every call below runs on all
 * devices (one internal worker thread each) and returns rank 0's -- identical -- results.  devices: ngpus CUDA ordinals,
 * NULL = 0 .. ngpus-1.  Arguments otherwise as gpz_create / gpz_eval / gpz_fit / gpz_get_prior / gpz_train.          */
typedef struct gpz_multi gpz_multi;
int gpz_create_multi(gpz_multi** out, const gpz_model* model, int64_t n_all,
                     const double* X, const double* Y, const double* Psi, const double* omega,
                     const uint8_t* training, const uint8_t* validation, int ngpus, const int* devices);
void gpz_destroy_multi(gpz_multi* mc);
int gpz_multi_devices(const gpz_multi* mc);
gpz_ctx* gpz_multi_ctx(gpz_multi* mc, int rank);            /* the per-device context (timing, launch counts) */
int gpz_multi_eval(gpz_multi* mc, const double* theta, double* nlogML, double* grad, double stats[4]);
int gpz_multi_fit(gpz_multi* mc, const double* theta, double* nlogML_k, double* w, double* iSigma_w);
int gpz_multi_get_prior(gpz_multi* mc, const double* theta, double* prior);
int gpz_multi_set_option(gpz_multi* mc, const char* name, double value);

/* ---- [nlogML,grad] = GPz(theta,...) with nargout<=2   (GPz/GPz.m:1-263, full path) ---------------
 * stats[4] = trainRMSE, trainLL, validRMSE, validLL  (the globals of GPz/GPz.m:3-7,236-259;
 * valid* are NaN without a validation mask).                                                     */
int gpz_eval(gpz_ctx* ctx, const double* theta, double* nlogML, double* grad, double stats[4]);

/* same, device-resident: d_theta (p doubles) and d_out (p+5 doubles: nlogML, grad[p], stats[4]) are
 * device pointers; the work is enqueued on gpz_stream(ctx) and NOT synchronised.                 */
int gpz_eval_dev(gpz_ctx* ctx, const double* d_theta, double* d_out);

/* ---- [~,~,w,iSigma_w] = GPz(theta,...) with nargout>2  (GPz/GPz.m:84-87 early exit) --------------
 * nlogML_k: 1 x k UN-normalised values as the reference returns on this exit (may be NULL);
 * w m x k; iSigma_w m x m x k.                                                                   */
int gpz_fit(gpz_ctx* ctx, const double* theta, double* nlogML_k, double* w, double* iSigma_w);

/* ---- [PHI,~,lnBeta_i,N] = getPHI(X,Psi,theta,model,selection) on the context's rows --------------
 * (GPz/getPHI.m:1-127).  which: 0 = training rows, 1 = validation rows.  PHI n x m, lnBeta_i n x k,
 * N n x m (normalised densities, getPHI.m:114); any of them may be NULL.                         */
int gpz_phi(gpz_ctx* ctx, const double* theta, int which, double* PHI, double* lnBeta_i, double* N);

/* ---- prior = getPrior(X,Psi,theta,model,training)  (GPz/getPrior.m:1-22, called at train.m:59,74) ---
 * EM for the mixture weights of the bases on the training rows; prior: m doubles.                */
int gpz_get_prior(gpz_ctx* ctx, const double* theta, double* prior);
int64_t gpz_rows(const gpz_ctx* ctx, int which);

/* ---- predict core (GPz/predict.m:45-73 grouping + dispatch; predictDiag.m:58-295, predictCov.m:53-336) ---
 * Xz n x d z-scored rows (NaN = missing); Psi NULL, n x d (methods ?L/?D) or d x d x n (methods ?C), fixPsi-normalised.
 * Rows are grouped by NaN pattern (predict.m:45-52).  Complete rows: predictFull / predictNoisy.  Rows with NaN:
 * predictMissing / predictNoisyMissing, which need priors (m doubles, model.best.priors; may be NULL otherwise).
 * All four branches exist for the diagonal and for the covariance methods.
 * theta/w/iSigma_w as stored in model.best / model.last.
 * Outputs n x k each: mu (WITHOUT muY), nu, beta_i, gamma; PHI n x m (may be NULL).
 * sigma = nu + beta_i + gamma and mu += muY stay on the host (predict.m:72-73).                  */
int gpz_predict(const gpz_model* model, const double* theta, const double* w, const double* iSigma_w,
                int64_t n, const double* Xz, const double* Psi, const double* priors,
                double* mu, double* nu, double* beta_i, double* gamma, double* PHI, int device);

/* ---- [Xi,logdet] = inv_logdet(X)  (GPz/inv_logdet.m:1-15) for SPD X (blocked Cholesky) ----------- */
int gpz_inv_logdet(int32_t m, const double* X, double* Xi, double* logdet, int device);

/* ---- D = Dxy(X,Y)  (GPz/Dxy.m:1-10): X n x d, Y m x d -> D n x m -------------------------------- */
int gpz_dxy(int64_t n, int32_t m, int32_t d, const double* X, const double* Y, double* D, int device);
/* mean(Dxy(X,Y)) -- the 1 x m column means init.m:62 takes for its length-scale heuristic -- without forming the
 * n x m matrix (8 GB at the headline size).  Same per-entry formula and fma order as gpz_dxy.                   */
int gpz_dxy_colmean(int64_t n, int32_t m, int32_t d, const double* X, const double* Y, double* mean, int device);

/* ---- C = A * B'  in fp64 through the int8 tensor cores (the engine of the Gram and T-GEMM steps of GPz.m:63-72,
 * exposed for testing and reuse): A is M x K, B is N x K, C is M x N, all ROW-major with leading dimensions lda, ldb, ldc.
 * Every row of A and B is scaled by a power of two and cut into `digits` (3..7) balanced base-256 digits; the digit
 * products are exact int8 tcgen05 GEMMs, digit pairs below 256^-digits of (row scale x row scale) are dropped.
 * digits = 7 is more accurate than an fp64 GEMM (error <= ~K 2^-56 max|A_i| max|B_j|).  K <= 16384.              */
int gpz_dgemm_nt(int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                 int64_t ldc, int32_t digits, int device);

/* ---- theta = minFunc(f,theta,options) as GPz/train.m:38-48 configures it, with GPz/callBack.m -----
 * Replaces, for this one configuration, minFunc_2012/minFunc/minFunc.m:258-1170 (method 'lbfgs'),
 * WolfeLineSearch.m:32-263, ArmijoBacktrack.m:32-143, polyinterp.m:41-58, lbfgsAdd.m / lbfgsProd.m (and the MEX
 * lbfgsAddC.c / lbfgsProdC.c) and the best-theta / early-stopping logic of GPz/callBack.m:21-48.
 * theta, the gradients, the search direction and the (S,Y) history stay on the device; per evaluation only
 * nine scalars cross PCIe.  Defaults are minFunc's (minFunc_processInputOptions.m:62-67,117-146).             */
typedef struct gpz_train_options {
    int32_t max_iter;         /* train.m 'maxIter' (200)                                                   */
    int32_t training_only;    /* 1: no validation set (callBack.m:21-24); 0: best theta by validLL; -1: auto
                                 from this context's validation rows (sharded callers must set it)         */
    double  max_attempts;     /* train.m 'maxAttempts' (inf): stop after this many non-improving iterations */
    int32_t corrections;      /* 100 */
    int32_t max_ls;           /* 25 line-search evaluations (minFunc.m:1066)                               */
    double  opt_tol;          /* 1e-5 */
    double  prog_tol;         /* 1e-9 */
    double  c1, c2;           /* 1e-4, 0.9 */
    double  max_fun_evals;    /* inf (train.m:45)                                                          */
} gpz_train_options;
typedef struct gpz_train_iter {   /* what callBack.m prints per iteration                                  */
    int32_t iter, fun_evals;
    int32_t improved;         /* this iterate became best_theta                                            */
    int32_t attempts;         /* -1 while the reference's `attempts` global is still unset                 */
    double  f, t, gtd, opt_cond;
    double  stats[4];         /* trainRMSE, trainLL, validRMSE, validLL of the LAST evaluation (the globals) */
} gpz_train_iter;
typedef int (*gpz_train_callback)(void* user, const gpz_train_iter* it);    /* non-zero return stops the run */
typedef struct gpz_train_result {
    int32_t iterations, fun_evals;
    int32_t exitflag;         /* minFunc's: 1 optTol, 2 progTol family, 0 limits, -1 stopped by callback/attempts,
                                 -3 illegal direction                                                      */
    int32_t reason;           /* index for gpz_train_reason()                                              */
    int32_t attempts, skipped_pairs;
    double  f, opt_cond, best_valid;
    double  ms_total;         /* host wall time of the call                                                */
    double  ms_eval;          /* device time inside the objective evaluations (CUDA events)                */
} gpz_train_result;
void gpz_train_default_options(gpz_train_options* o);
const char* gpz_train_reason(int reason);
/* theta [p]: in = start (model.last.theta), out = final iterate.  best_theta [p] / best_valid: in = model.best.theta /
 * model.best.LL (NaN = empty), out = updated as callBack.m does.  cb may be NULL.                           */
int gpz_train(gpz_ctx* ctx, const gpz_train_options* opt, double* theta, double* best_theta, double* best_valid,
              gpz_train_callback cb, void* user, gpz_train_result* res);
int gpz_multi_train(gpz_multi* mc, const gpz_train_options* opt, double* theta, double* best_theta, double* best_valid,
                    gpz_train_callback cb, void* user, gpz_train_result* res);   /* opt->training_only must be 0 or 1 */
/* the same optimiser on a caller-supplied objective (host callback that fills f and g for a DEVICE x): exposed so the
 * optimiser can be tested on analytic functions.  fn(user, d_x, d_out) must enqueue on `stream` (a cudaStream_t) and
 * leave d_out = [f, g[p], stats[4]].                                                                        */
typedef int (*gpz_objective_dev)(void* user, const double* d_x, double* d_out, void* stream);
int gpz_minimize_dev(int64_t p, gpz_objective_dev fn, void* fn_user, const gpz_train_options* opt, double* theta,
                     double* best_theta, double* best_valid, gpz_train_callback cb, void* user, gpz_train_result* res,
                     int device);

/* ---- plumbing / measurement ------------------------------------------------------------------- */
void* gpz_stream(gpz_ctx* ctx);                 /* cudaStream_t the context enqueues on            */
int   gpz_sync(gpz_ctx* ctx);
int64_t gpz_launch_count(const gpz_ctx* ctx);   /* kernels launched by this context so far         */
int64_t gpz_graph_replays(const gpz_ctx* ctx);  /* evaluations replayed as one CUDA graph (small, launch-bound problems:
                                                   the evaluation is captured the second time it runs on the same
                                                   (theta, out) buffers -- always the case for gpz_eval and gpz_train) */
/* device time (ms, CUDA events on the context stream) of the phases of the LAST eval:
 * [0] phi build  [1] row weights + Gram + PHI'Wy (+allreduce #1)  [2] solve  [3] PHI w + T-GEMM + row gradients
 * [4] dPHI + back-projection + validation + finish (+allreduce #2)  [5] total
 * [6] the Gram step alone  [7] the T-GEMM step alone (fp64 DMMA kernel, or column digits + the tcgen05 digit GEMM)
 * [8] the tcgen05 digit-GEMM launch of T = PHI iSigma (ms)  [9] int8 operations that launch executed
 * [10] base-256 digits in use (0 = fp64 DMMA path)  [11] 1 if the Gram also runs on the int8 tensor cores
 * When the last evaluation was a CUDA-graph replay, [5] is the time of that replay and the phases are those of the
 * last evaluation that ran as plain launches.                                                                  */
int gpz_last_timing(gpz_ctx* ctx, double ms[12]);
/* device time (ms) of single kernels of the last evaluation that ran as plain launches, first row chunk; -1 = not run:
 * [0] PHI build (the PHI kernel + the ordered sum of its row-dot partials)   [1] digit extraction of PHI
 * [2] Gram digit GEMM + its fixed-order reduce   [3] moment GEMM of the back-projection (dPHI' F)             */
int gpz_kernel_timing(gpz_ctx* ctx, double ms[4]);
int gpz_set_option(gpz_ctx* ctx, const char* name, double value);

#ifdef __cplusplus
}
#endif
#endif /* GPZ_B200_H */
