// gpz_b200_mex.cpp -- MEX gateway between MATLAB and libgpz_b200.so (C ABI in include/gpz_b200.h).
//
// Build (on a machine with MATLAB; there is no mex.h in the build container, so this file is shipped
// as source and only syntax-checked against tests/mex_stub/mex.h):
//     mex -R2017b -I../include gpz_b200_mex.cpp -L../gpz_b200 -lgpz_b200
// Convention follows the reference's own MEX files (minFunc_2012/minFunc/mex/lbfgsProdC.c:7-44):
// plain mexFunction, mxGetPr in, mxCreateDoubleMatrix out, mexErrMsgIdAndTxt on misuse.  Inputs are never
// written (unlike lbfgsAddC.c:30-33).  Usage from MATLAB (see GPz.m / getPHI.m / predict_core.m here):
//     h = gpz_b200_mex('create', model, X, Y, Psi, omega, training, validation)    -> uint64 handle
//     [f, g, stats] = gpz_b200_mex('eval', h, theta)
//     [nl, w, iSigma_w] = gpz_b200_mex('fit', h, theta)
//     [PHI, lnBeta_i, N] = gpz_b200_mex('phi', h, theta, which, model)
//     prior = gpz_b200_mex('get_prior', h, theta, model)
//     [theta, best_theta, best_valid, info] = gpz_b200_mex('train', h, theta, best_theta, best_valid, maxIter,
//                                                          maxAttempts, trainingOnly, display)
//         info = [iterations funEvals exitflag reason attempts skippedPairs f optCond msTotal msEval]
//     [mu, nu, beta_i, gamma, PHI] = gpz_b200_mex('predict', model, theta, w, iSigma_w, Xz, Psi, priors)
//     [Xi, logdet] = gpz_b200_mex('inv_logdet', X)
//     D = gpz_b200_mex('dxy', X, Y)
//     mD = gpz_b200_mex('dxy_colmean', X, Y)                  % mean(Dxy(X,Y)) without the n x m matrix (init.m:62)
//     gpz_b200_mex('destroy', h)
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "gpz_b200.h"
#include "mex.h"

namespace {

std::map<uint64_t, gpz_ctx*> g_ctx;
uint64_t g_next = 1;

void fail(const char* what) { mexErrMsgIdAndTxt("gpz_b200:error", "%s: %s", what, gpz_last_error()); }

void destroy_all() {
    for (auto& kv : g_ctx) gpz_destroy(kv.second);
    g_ctx.clear();
}

gpz_model read_model(const mxArray* s) {
    if (!mxIsStruct(s)) mexErrMsgIdAndTxt("gpz_b200:usage", "model must be a struct");
    gpz_model m;
    std::memset(&m, 0, sizeof(m));
    m.d = static_cast<int32_t>(mxGetScalar(mxGetField(s, 0, "d")));
    m.k = static_cast<int32_t>(mxGetScalar(mxGetField(s, 0, "k")));
    m.m = static_cast<int32_t>(mxGetScalar(mxGetField(s, 0, "m")));
    m.heteroscedastic = mxGetScalar(mxGetField(s, 0, "heteroscedastic")) != 0;
    char buf[8] = {0};
    mxGetString(mxGetField(s, 0, "method"), buf, sizeof(buf));
    m.method[0] = buf[0];
    m.method[1] = buf[1];
    return m;
}

const double* opt_double(const mxArray* a) { return (a == nullptr || mxIsEmpty(a)) ? nullptr : mxGetPr(a); }

// MATLAB logical / double mask -> uint8 vector (empty -> NULL = "all rows" / "no validation")
const uint8_t* mask(const mxArray* a, std::vector<uint8_t>& store, size_t n) {
    if (a == nullptr || mxIsEmpty(a)) return nullptr;
    store.resize(n);
    if (mxIsLogical(a)) {
        const mxLogical* p = mxGetLogicals(a);
        for (size_t i = 0; i < n; ++i) store[i] = p[i] ? 1 : 0;
    } else {
        const double* p = mxGetPr(a);
        for (size_t i = 0; i < n; ++i) store[i] = p[i] != 0.0;
    }
    return store.data();
}

gpz_ctx* lookup(const mxArray* h) {
    const uint64_t id = static_cast<uint64_t>(mxGetScalar(h));
    auto it = g_ctx.find(id);
    if (it == g_ctx.end()) mexErrMsgIdAndTxt("gpz_b200:usage", "invalid context handle");
    return it->second;
}

// the table GPz/callBack.m:14-34 prints, one row per iteration
struct TrainPrint {
    bool display, training_only;
};
int train_row(void* user, const gpz_train_iter* it) {
    const TrainPrint* tp = static_cast<const TrainPrint*>(user);
    if (!tp->display) return 0;
    if (it->iter == 1)
        mexPrintf(tp->training_only ? "\tIter\tlogML/n\t\tTrain RMSE\tTrain MLL\n"
                                    : "\tIter\tlogML/n\t\tTrain RMSE\tTrain MLL\tValid RMSE\tValid MLL\n");
    if (tp->training_only)
        mexPrintf("\t%d\t%1.5e\t%1.5e\t %1.5e\n", it->iter, -it->f, it->stats[0], it->stats[1]);
    else
        mexPrintf(it->improved ? "\t%d\t%1.5e\t%1.5e\t%1.5e\t%1.5e\t[%1.5e]\n" : "\t%d\t%1.5e\t%1.5e\t%1.5e\t%1.5e\t %1.5e\n",
                  it->iter, -it->f, it->stats[0], it->stats[1], it->stats[2], it->stats[3]);
    mexEvalString("drawnow;");
    return 0;
}

}  // namespace

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("gpz_b200:usage", "first argument must be a command string");
    char cmd[32];
    mxGetString(prhs[0], cmd, sizeof(cmd));
    const std::string c(cmd);
    static bool locked = false;
    if (!locked) {
        mexLock();
        mexAtExit(destroy_all);
        locked = true;
    }
    if (c == "create") {
        if (nrhs < 4) mexErrMsgIdAndTxt("gpz_b200:usage", "create(model,X,Y,Psi,omega,training,validation)");
        gpz_model m = read_model(prhs[1]);
        const size_t n = mxGetM(prhs[2]);
        std::vector<uint8_t> tr, va;
        gpz_ctx* ctx = nullptr;
        const int rc = gpz_create(&ctx, &m, static_cast<int64_t>(n), mxGetPr(prhs[2]), mxGetPr(prhs[3]),
                                  nrhs > 4 ? opt_double(prhs[4]) : nullptr, nrhs > 5 ? opt_double(prhs[5]) : nullptr,
                                  nrhs > 6 ? mask(prhs[6], tr, n) : nullptr, nrhs > 7 ? mask(prhs[7], va, n) : nullptr, 0);
        if (rc) fail("gpz_create");
        const uint64_t id = g_next++;
        g_ctx[id] = ctx;
        plhs[0] = mxCreateDoubleScalar(static_cast<double>(id));
    } else if (c == "destroy") {
        const uint64_t id = static_cast<uint64_t>(mxGetScalar(prhs[1]));
        auto it = g_ctx.find(id);
        if (it != g_ctx.end()) {
            gpz_destroy(it->second);
            g_ctx.erase(it);
        }
    } else if (c == "eval") {
        gpz_ctx* ctx = lookup(prhs[1]);
        const size_t p = mxGetNumberOfElements(prhs[2]);
        plhs[0] = mxCreateDoubleMatrix(1, 1, mxREAL);
        mxArray* g = mxCreateDoubleMatrix(p, 1, mxREAL);
        mxArray* st = mxCreateDoubleMatrix(4, 1, mxREAL);
        if (gpz_eval(ctx, mxGetPr(prhs[2]), mxGetPr(plhs[0]), mxGetPr(g), mxGetPr(st))) fail("gpz_eval");
        if (nlhs > 1) plhs[1] = g; else mxDestroyArray(g);
        if (nlhs > 2) plhs[2] = st; else mxDestroyArray(st);
    } else if (c == "fit") {
        gpz_ctx* ctx = lookup(prhs[1]);
        const mxArray* model = prhs[3];
        gpz_model m = read_model(model);
        plhs[0] = mxCreateDoubleMatrix(1, m.k, mxREAL);
        mxArray* w = mxCreateDoubleMatrix(m.m, m.k, mxREAL);
        const mwSize dims[3] = {static_cast<mwSize>(m.m), static_cast<mwSize>(m.m), static_cast<mwSize>(m.k)};
        mxArray* iS = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        if (gpz_fit(ctx, mxGetPr(prhs[2]), mxGetPr(plhs[0]), mxGetPr(w), mxGetPr(iS))) fail("gpz_fit");
        if (nlhs > 1) plhs[1] = w; else mxDestroyArray(w);
        if (nlhs > 2) plhs[2] = iS; else mxDestroyArray(iS);
    } else if (c == "phi") {
        gpz_ctx* ctx = lookup(prhs[1]);
        const int which = nrhs > 3 ? static_cast<int>(mxGetScalar(prhs[3])) : 0;
        gpz_model m = read_model(prhs[4]);
        const int64_t n = gpz_rows(ctx, which);
        plhs[0] = mxCreateDoubleMatrix(n, m.m, mxREAL);
        mxArray* lb = mxCreateDoubleMatrix(n, m.k, mxREAL);
        mxArray* N = nlhs > 2 ? mxCreateDoubleMatrix(n, m.m, mxREAL) : nullptr;
        if (gpz_phi(ctx, mxGetPr(prhs[2]), which, mxGetPr(plhs[0]), mxGetPr(lb), N ? mxGetPr(N) : nullptr)) fail("gpz_phi");
        if (nlhs > 1) plhs[1] = lb; else mxDestroyArray(lb);
        if (nlhs > 2) plhs[2] = N;
    } else if (c == "get_prior") {
        gpz_ctx* ctx = lookup(prhs[1]);
        gpz_model m = read_model(prhs[3]);
        plhs[0] = mxCreateDoubleMatrix(1, m.m, mxREAL);
        if (gpz_get_prior(ctx, mxGetPr(prhs[2]), mxGetPr(plhs[0]))) fail("gpz_get_prior");
    } else if (c == "train") {
        if (nrhs < 8) mexErrMsgIdAndTxt("gpz_b200:usage", "train(h,theta,best_theta,best_valid,maxIter,maxAttempts,trainingOnly[,display])");
        gpz_ctx* ctx = lookup(prhs[1]);
        const size_t p = mxGetNumberOfElements(prhs[2]);
        if (mxGetNumberOfElements(prhs[3]) != p) mexErrMsgIdAndTxt("gpz_b200:usage", "theta and best_theta differ in length");
        plhs[0] = mxCreateDoubleMatrix(p, 1, mxREAL);                 // inputs are never written: work on copies
        mxArray* best = mxCreateDoubleMatrix(p, 1, mxREAL);
        std::memcpy(mxGetPr(plhs[0]), mxGetPr(prhs[2]), sizeof(double) * p);
        std::memcpy(mxGetPr(best), mxGetPr(prhs[3]), sizeof(double) * p);
        double bv = mxIsEmpty(prhs[4]) ? NAN : mxGetScalar(prhs[4]);  // isempty(best_valid), callBack.m:26
        gpz_train_options o;
        gpz_train_default_options(&o);
        o.max_iter = static_cast<int32_t>(mxGetScalar(prhs[5]));
        o.max_attempts = mxGetScalar(prhs[6]);
        o.training_only = mxGetScalar(prhs[7]) != 0;
        TrainPrint tp{nrhs > 8 && mxGetScalar(prhs[8]) != 0, o.training_only != 0};
        gpz_train_result r;
        std::memset(&r, 0, sizeof(r));
        if (gpz_train(ctx, &o, mxGetPr(plhs[0]), mxGetPr(best), &bv, train_row, &tp, &r)) fail("gpz_train");
        if (tp.display) mexPrintf("%s\n", r.reason == 7 ? "No improvment after maximum number of attempts" : gpz_train_reason(r.reason));
        if (nlhs > 1) plhs[1] = best; else mxDestroyArray(best);
        if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(bv);
        if (nlhs > 3) {
            plhs[3] = mxCreateDoubleMatrix(1, 10, mxREAL);
            double* q = mxGetPr(plhs[3]);
            q[0] = r.iterations, q[1] = r.fun_evals, q[2] = r.exitflag, q[3] = r.reason, q[4] = r.attempts;
            q[5] = r.skipped_pairs, q[6] = r.f, q[7] = r.opt_cond, q[8] = r.ms_total, q[9] = r.ms_eval;
        }
    } else if (c == "predict") {
        gpz_model m = read_model(prhs[1]);
        const size_t n = mxGetM(prhs[5]);
        mxArray* out[5];
        for (int i = 0; i < 4; ++i) out[i] = mxCreateDoubleMatrix(n, m.k, mxREAL);
        out[4] = mxCreateDoubleMatrix(n, m.m, mxREAL);
        if (gpz_predict(&m, mxGetPr(prhs[2]), mxGetPr(prhs[3]), mxGetPr(prhs[4]), static_cast<int64_t>(n), mxGetPr(prhs[5]),
                        nrhs > 6 ? opt_double(prhs[6]) : nullptr, nrhs > 7 ? opt_double(prhs[7]) : nullptr, mxGetPr(out[0]),
                        mxGetPr(out[1]), mxGetPr(out[2]), mxGetPr(out[3]), mxGetPr(out[4]), 0))
            fail("gpz_predict");
        for (int i = 0; i < 5; ++i) {
            if (i < nlhs || i == 0) plhs[i] = out[i]; else mxDestroyArray(out[i]);
        }
    } else if (c == "inv_logdet") {
        const size_t m = mxGetM(prhs[1]);
        plhs[0] = mxCreateDoubleMatrix(m, m, mxREAL);
        double ld = 0.0;
        if (gpz_inv_logdet(static_cast<int32_t>(m), mxGetPr(prhs[1]), mxGetPr(plhs[0]), &ld, 0)) fail("gpz_inv_logdet");
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(ld);
    } else if (c == "dxy") {
        const size_t n = mxGetM(prhs[1]), d = mxGetN(prhs[1]), m = mxGetM(prhs[2]);
        plhs[0] = mxCreateDoubleMatrix(n, m, mxREAL);
        if (gpz_dxy(static_cast<int64_t>(n), static_cast<int32_t>(m), static_cast<int32_t>(d), mxGetPr(prhs[1]), mxGetPr(prhs[2]),
                    mxGetPr(plhs[0]), 0))
            fail("gpz_dxy");
    } else if (c == "dxy_colmean") {                                  // mean(Dxy(X,Y)), init.m:62
        const size_t n = mxGetM(prhs[1]), d = mxGetN(prhs[1]), m = mxGetM(prhs[2]);
        plhs[0] = mxCreateDoubleMatrix(1, m, mxREAL);
        if (gpz_dxy_colmean(static_cast<int64_t>(n), static_cast<int32_t>(m), static_cast<int32_t>(d), mxGetPr(prhs[1]),
                            mxGetPr(prhs[2]), mxGetPr(plhs[0]), 0))
            fail("gpz_dxy_colmean");
    } else {
        mexErrMsgIdAndTxt("gpz_b200:usage", "unknown command '%s'", cmd);
    }
}
