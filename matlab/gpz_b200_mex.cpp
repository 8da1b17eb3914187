// gpz_b200_mex.cpp -- MEX gateway between MATLAB and libgpz_b200.so (C ABI in include/gpz_b200.h).
//
// Build (on a machine with MATLAB):
//     mex -R2017b -I../include gpz_b200_mex.cpp -L../gpz_b200 -lgpz_b200
// There is no MATLAB in the build container: there the file is compiled against tests/mex_stub/mex.h and linked with the
// minimal libmx mock tests/mex_stub/mex_mock.cpp, and tests/test_mex_gateway.py drives every command through mexFunction.
// Convention follows the reference's own MEX files (minFunc_2012/minFunc/mex/lbfgsProdC.c:7-44):
// plain mexFunction, mxGetPr in, mxCreateDoubleMatrix out, mexErrMsgIdAndTxt on misuse.  Inputs are never
// written (unlike lbfgsAddC.c:30-33).  Usage from MATLAB (see GPz.m / getPHI.m / predict_core.m here):
//     h = gpz_b200_mex('create', model, X, Y, Psi, omega, training, validation[, ngpus])    -> handle
//         ngpus > 1: the rows are split over that many GPUs inside the library (gpz_create_multi), still one caller thread
//     [f, g, stats] = gpz_b200_mex('eval', h, theta)
//     [nl, w, iSigma_w] = gpz_b200_mex('fit', h, theta[, model])
//     [PHI, lnBeta_i, N] = gpz_b200_mex('phi', h, theta[, which[, model]])
//     prior = gpz_b200_mex('get_prior', h, theta[, model])
//   (the handle remembers the model it was created with: theta lengths are checked against it, outputs are sized from it,
//    and an optional model argument must agree with it -- a mismatch raises gpz_b200:usage instead of corrupting the heap)
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

struct Entry {
    gpz_ctx* ctx;           // one device ...
    gpz_multi* mc;          // ... or several, driven from this one thread (create(..., ngpus) with ngpus > 1)
    gpz_model model;        // the model the context was created for: every output is sized from THIS, never from an argument
};
std::map<uint64_t, Entry> g_ctx;
uint64_t g_next = 1;

void fail(const char* what) { mexErrMsgIdAndTxt("gpz_b200:error", "%s: %s", what, gpz_last_error()); }
void usage(const char* what) { mexErrMsgIdAndTxt("gpz_b200:usage", "%s", what); }

void destroy_entry(Entry& e) {
    if (e.ctx) gpz_destroy(e.ctx);
    if (e.mc) gpz_destroy_multi(e.mc);
    e.ctx = nullptr;
    e.mc = nullptr;
}

void destroy_all() {
    for (auto& kv : g_ctx) destroy_entry(kv.second);
    g_ctx.clear();
}

const mxArray* field(const mxArray* s, const char* name) {
    const mxArray* f = mxGetField(s, 0, name);
    if (f == nullptr || mxIsEmpty(f)) mexErrMsgIdAndTxt("gpz_b200:usage", "model.%s is missing", name);
    return f;
}

gpz_model read_model(const mxArray* s) {
    if (s == nullptr || !mxIsStruct(s)) usage("model must be a struct");
    gpz_model m;
    std::memset(&m, 0, sizeof(m));
    m.d = static_cast<int32_t>(mxGetScalar(field(s, "d")));
    m.k = static_cast<int32_t>(mxGetScalar(field(s, "k")));
    m.m = static_cast<int32_t>(mxGetScalar(field(s, "m")));
    m.heteroscedastic = mxGetScalar(field(s, "heteroscedastic")) != 0;
    char buf[8] = {0};
    mxGetString(field(s, "method"), buf, sizeof(buf));
    m.method[0] = buf[0];
    m.method[1] = buf[1];
    if (gpz_theta_len(&m) < 0) fail("model");
    return m;
}

bool same_model(const gpz_model& a, const gpz_model& b) {
    return a.d == b.d && a.k == b.k && a.m == b.m && a.method[0] == b.method[0] && a.method[1] == b.method[1] &&
           (a.heteroscedastic != 0) == (b.heteroscedastic != 0);
}

void need(int nrhs, int n, const char* sig) {
    if (nrhs < n) mexErrMsgIdAndTxt("gpz_b200:usage", "usage: %s", sig);
}

// theta must have exactly the length the context's model defines: the library reads that many doubles
void check_theta(const mxArray* th, const gpz_model& m, const char* what) {
    const int64_t p = gpz_theta_len(&m);
    if (th == nullptr || mxGetPr(th) == nullptr || static_cast<int64_t>(mxGetNumberOfElements(th)) != p)
        mexErrMsgIdAndTxt("gpz_b200:usage", "%s: theta has %lld elements, the model needs %lld", what,
                          static_cast<long long>(th ? mxGetNumberOfElements(th) : 0), static_cast<long long>(p));
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

Entry& lookup(const mxArray* h) {
    if (h == nullptr || mxIsEmpty(h)) usage("invalid context handle");
    const uint64_t id = static_cast<uint64_t>(mxGetScalar(h));
    auto it = g_ctx.find(id);
    if (it == g_ctx.end()) usage("invalid context handle");
    return it->second;
}

// an optional model argument must describe the context it is used with (a stale handle with another model would make the
// caller mis-size its own arrays)
void check_model_arg(int nrhs, const mxArray* prhs[], int pos, const gpz_model& m) {
    if (nrhs > pos && prhs[pos] != nullptr && !mxIsEmpty(prhs[pos]) && !same_model(read_model(prhs[pos]), m))
        usage("the model argument differs from the model this context was created with");
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
        need(nrhs, 4, "h = gpz_b200_mex('create',model,X,Y,Psi,omega,training,validation)");
        gpz_model m = read_model(prhs[1]);
        const size_t n = mxGetM(prhs[2]);
        if (mxGetPr(prhs[2]) == nullptr || mxGetN(prhs[2]) != static_cast<size_t>(m.d)) usage("create: X must be n x model.d (double)");
        if (mxGetPr(prhs[3]) == nullptr || mxGetNumberOfElements(prhs[3]) != n * static_cast<size_t>(m.k)) usage("create: Y must be n x model.k (double)");
        const double* psi = nrhs > 4 ? opt_double(prhs[4]) : nullptr;
        if (psi != nullptr) {
            const size_t want = (m.method[1] == 'C') ? n * m.d * m.d : n * static_cast<size_t>(m.d);
            if (mxGetNumberOfElements(prhs[4]) != want) usage("create: Psi must be n x d (?L/?D) or d x d x n (?C), as fixPsi returns it");
        }
        const double* om = nrhs > 5 ? opt_double(prhs[5]) : nullptr;
        if (om != nullptr && mxGetNumberOfElements(prhs[5]) != n) usage("create: omega must have n elements");
        for (int q = 6; q <= 7; ++q)
            if (nrhs > q && prhs[q] != nullptr && !mxIsEmpty(prhs[q]) && mxGetNumberOfElements(prhs[q]) != n)
                usage("create: training / validation masks must have n elements");
        std::vector<uint8_t> tr, va;
        const int ngpus = (nrhs > 8 && !mxIsEmpty(prhs[8])) ? static_cast<int>(mxGetScalar(prhs[8])) : 1;
        if (ngpus < 1) usage("create: ngpus must be >= 1");
        Entry e{nullptr, nullptr, m};
        const uint8_t* trp = nrhs > 6 ? mask(prhs[6], tr, n) : nullptr;
        const uint8_t* vap = nrhs > 7 ? mask(prhs[7], va, n) : nullptr;
        if (ngpus == 1) {
            if (gpz_create(&e.ctx, &m, static_cast<int64_t>(n), mxGetPr(prhs[2]), mxGetPr(prhs[3]), psi, om, trp, vap, 0)) fail("gpz_create");
        } else {            // one MATLAB thread, ngpus devices: rows split and NCCL set up inside the library
            if (gpz_create_multi(&e.mc, &m, static_cast<int64_t>(n), mxGetPr(prhs[2]), mxGetPr(prhs[3]), psi, om, trp, vap, ngpus, nullptr))
                fail("gpz_create_multi");
        }
        const uint64_t id = g_next++;
        g_ctx[id] = e;
        plhs[0] = mxCreateDoubleScalar(static_cast<double>(id));
    } else if (c == "destroy") {
        need(nrhs, 2, "gpz_b200_mex('destroy',h)");
        if (mxIsEmpty(prhs[1])) return;
        const uint64_t id = static_cast<uint64_t>(mxGetScalar(prhs[1]));
        auto it = g_ctx.find(id);
        if (it != g_ctx.end()) {
            destroy_entry(it->second);
            g_ctx.erase(it);
        }
    } else if (c == "eval") {
        need(nrhs, 3, "[f,g,stats] = gpz_b200_mex('eval',h,theta)");
        Entry& e = lookup(prhs[1]);
        check_theta(prhs[2], e.model, "eval");
        const int64_t p = gpz_theta_len(&e.model);
        plhs[0] = mxCreateDoubleMatrix(1, 1, mxREAL);
        mxArray* g = mxCreateDoubleMatrix(p, 1, mxREAL);
        mxArray* st = mxCreateDoubleMatrix(4, 1, mxREAL);
        if (e.mc ? gpz_multi_eval(e.mc, mxGetPr(prhs[2]), mxGetPr(plhs[0]), mxGetPr(g), mxGetPr(st))
                 : gpz_eval(e.ctx, mxGetPr(prhs[2]), mxGetPr(plhs[0]), mxGetPr(g), mxGetPr(st)))
            fail("gpz_eval");
        if (nlhs > 1) plhs[1] = g; else mxDestroyArray(g);
        if (nlhs > 2) plhs[2] = st; else mxDestroyArray(st);
    } else if (c == "fit") {
        need(nrhs, 3, "[nl,w,iSigma_w] = gpz_b200_mex('fit',h,theta[,model])");
        Entry& e = lookup(prhs[1]);
        check_theta(prhs[2], e.model, "fit");
        check_model_arg(nrhs, prhs, 3, e.model);
        const gpz_model& m = e.model;
        plhs[0] = mxCreateDoubleMatrix(1, m.k, mxREAL);
        mxArray* w = mxCreateDoubleMatrix(m.m, m.k, mxREAL);
        const mwSize dims[3] = {static_cast<mwSize>(m.m), static_cast<mwSize>(m.m), static_cast<mwSize>(m.k)};
        mxArray* iS = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        if (e.mc ? gpz_multi_fit(e.mc, mxGetPr(prhs[2]), mxGetPr(plhs[0]), mxGetPr(w), mxGetPr(iS))
                 : gpz_fit(e.ctx, mxGetPr(prhs[2]), mxGetPr(plhs[0]), mxGetPr(w), mxGetPr(iS)))
            fail("gpz_fit");
        if (nlhs > 1) plhs[1] = w; else mxDestroyArray(w);
        if (nlhs > 2) plhs[2] = iS; else mxDestroyArray(iS);
    } else if (c == "phi") {
        need(nrhs, 3, "[PHI,lnBeta_i,N] = gpz_b200_mex('phi',h,theta[,which[,model]])");
        Entry& e = lookup(prhs[1]);
        if (e.mc) usage("phi: not available on a multi-GPU handle (the rows live on several devices); create a one-GPU handle");
        check_theta(prhs[2], e.model, "phi");
        const int which = (nrhs > 3 && !mxIsEmpty(prhs[3])) ? static_cast<int>(mxGetScalar(prhs[3])) : 0;
        if (which != 0 && which != 1) usage("phi: which must be 0 (training rows) or 1 (validation rows)");
        check_model_arg(nrhs, prhs, 4, e.model);
        const gpz_model& m = e.model;
        const int64_t n = gpz_rows(e.ctx, which);
        plhs[0] = mxCreateDoubleMatrix(n, m.m, mxREAL);
        mxArray* lb = mxCreateDoubleMatrix(n, m.k, mxREAL);
        mxArray* N = nlhs > 2 ? mxCreateDoubleMatrix(n, m.m, mxREAL) : nullptr;
        if (gpz_phi(e.ctx, mxGetPr(prhs[2]), which, mxGetPr(plhs[0]), mxGetPr(lb), N ? mxGetPr(N) : nullptr)) fail("gpz_phi");
        if (nlhs > 1) plhs[1] = lb; else mxDestroyArray(lb);
        if (nlhs > 2) plhs[2] = N;
    } else if (c == "get_prior") {
        need(nrhs, 3, "prior = gpz_b200_mex('get_prior',h,theta[,model])");
        Entry& e = lookup(prhs[1]);
        check_theta(prhs[2], e.model, "get_prior");
        check_model_arg(nrhs, prhs, 3, e.model);
        plhs[0] = mxCreateDoubleMatrix(1, e.model.m, mxREAL);
        if (e.mc ? gpz_multi_get_prior(e.mc, mxGetPr(prhs[2]), mxGetPr(plhs[0])) : gpz_get_prior(e.ctx, mxGetPr(prhs[2]), mxGetPr(plhs[0])))
            fail("gpz_get_prior");
    } else if (c == "train") {
        if (nrhs < 8) mexErrMsgIdAndTxt("gpz_b200:usage", "train(h,theta,best_theta,best_valid,maxIter,maxAttempts,trainingOnly[,display])");
        Entry& e = lookup(prhs[1]);
        check_theta(prhs[2], e.model, "train");
        check_theta(prhs[3], e.model, "train (best_theta)");
        const size_t p = mxGetNumberOfElements(prhs[2]);
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
        if (e.mc ? gpz_multi_train(e.mc, &o, mxGetPr(plhs[0]), mxGetPr(best), &bv, train_row, &tp, &r)
                 : gpz_train(e.ctx, &o, mxGetPr(plhs[0]), mxGetPr(best), &bv, train_row, &tp, &r))
            fail("gpz_train");
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
        need(nrhs, 6, "[mu,nu,beta_i,gamma,PHI] = gpz_b200_mex('predict',model,theta,w,iSigma_w,Xz,Psi,priors)");
        gpz_model m = read_model(prhs[1]);
        check_theta(prhs[2], m, "predict");
        const size_t n = mxGetM(prhs[5]);
        const size_t mk = static_cast<size_t>(m.m) * m.k;
        if (mxGetPr(prhs[3]) == nullptr || mxGetNumberOfElements(prhs[3]) != mk) usage("predict: w must be m x k");
        if (mxGetPr(prhs[4]) == nullptr || mxGetNumberOfElements(prhs[4]) != mk * m.m) usage("predict: iSigma_w must be m x m x k");
        if (n > 0 && (mxGetPr(prhs[5]) == nullptr || mxGetN(prhs[5]) != static_cast<size_t>(m.d))) usage("predict: X must be n x model.d");
        if (nrhs > 6 && !mxIsEmpty(prhs[6])) {
            const size_t want = (m.method[1] == 'C') ? n * m.d * m.d : n * static_cast<size_t>(m.d);
            if (mxGetNumberOfElements(prhs[6]) != want) usage("predict: Psi must be n x d (?L/?D) or d x d x n (?C)");
        }
        if (nrhs > 7 && !mxIsEmpty(prhs[7]) && mxGetNumberOfElements(prhs[7]) != static_cast<size_t>(m.m)) usage("predict: priors must have m elements");
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
        need(nrhs, 2, "[Xi,logdet] = gpz_b200_mex('inv_logdet',X)");
        const size_t m = mxGetM(prhs[1]);
        if (mxGetPr(prhs[1]) == nullptr || mxGetN(prhs[1]) != m) usage("inv_logdet: X must be square (double)");
        plhs[0] = mxCreateDoubleMatrix(m, m, mxREAL);
        double ld = 0.0;
        if (gpz_inv_logdet(static_cast<int32_t>(m), mxGetPr(prhs[1]), mxGetPr(plhs[0]), &ld, 0)) fail("gpz_inv_logdet");
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(ld);
    } else if (c == "dxy") {
        need(nrhs, 3, "D = gpz_b200_mex('dxy',X,Y)");
        const size_t n = mxGetM(prhs[1]), d = mxGetN(prhs[1]), m = mxGetM(prhs[2]);
        if (mxGetN(prhs[2]) != d) usage("dxy: X and Y must have the same number of columns");
        plhs[0] = mxCreateDoubleMatrix(n, m, mxREAL);
        if (gpz_dxy(static_cast<int64_t>(n), static_cast<int32_t>(m), static_cast<int32_t>(d), mxGetPr(prhs[1]), mxGetPr(prhs[2]),
                    mxGetPr(plhs[0]), 0))
            fail("gpz_dxy");
    } else if (c == "dxy_colmean") {                                  // mean(Dxy(X,Y)), init.m:62
        need(nrhs, 3, "mD = gpz_b200_mex('dxy_colmean',X,Y)");
        const size_t n = mxGetM(prhs[1]), d = mxGetN(prhs[1]), m = mxGetM(prhs[2]);
        if (mxGetN(prhs[2]) != d) usage("dxy_colmean: X and Y must have the same number of columns");
        plhs[0] = mxCreateDoubleMatrix(1, m, mxREAL);
        if (gpz_dxy_colmean(static_cast<int64_t>(n), static_cast<int32_t>(m), static_cast<int32_t>(d), mxGetPr(prhs[1]),
                            mxGetPr(prhs[2]), mxGetPr(plhs[0]), 0))
            fail("gpz_dxy_colmean");
    } else {
        mexErrMsgIdAndTxt("gpz_b200:usage", "unknown command '%s'", cmd);
    }
}
