// Single-process multi-GPU boundary (SURVEY.md 8b "threading", 8e "backend"): the reference's caller is ONE interpreter thread
// holding one closure f = @(params) GPz(params,model,X,Y,Psi,omega,training,validation) (GPz/train.m:40), so a drop-in must be
// able to drive N GPUs from that one thread.  gpz_create_multi splits the selected rows into N contiguous blocks, builds one
// ordinary gpz_ctx per device and joins them into one NCCL communicator; every call is then executed by the caller's thread
// (rank 0) and N-1 persistent worker threads (one per further device), each running the unchanged per-rank entry point -- kernels, the two allreduces of an
// evaluation, the replicated m x m solve -- on its own device and stream.  All ranks finish with identical results
// (fixed-order reductions + allreduce), rank 0's are returned.  All calls block, the library stays single-caller.
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "internal.cuh"

using namespace gpz;

namespace {

enum Cmd { CMD_NONE = 0, CMD_INIT, CMD_EVAL, CMD_FIT, CMD_PRIOR, CMD_TRAIN, CMD_OPTION, CMD_QUIT };

struct Job {
    int cmd = CMD_NONE;
    const double* theta = nullptr;
    double* f = nullptr;
    double* grad = nullptr;
    double* stats = nullptr;
    double* nl = nullptr;
    double* w = nullptr;
    double* iS = nullptr;
    double* prior = nullptr;
    const gpz_train_options* topt = nullptr;
    double* best_theta = nullptr;
    double* best_valid = nullptr;
    gpz_train_callback cb = nullptr;
    void* cb_user = nullptr;
    gpz_train_result* res = nullptr;
    const char* opt_name = nullptr;
    double opt_value = 0.0;
    char id[128];
};

}  // namespace

struct gpz_multi {
    int n = 0;
    gpz_model model{};
    int64_t p = 0;
    std::vector<gpz_ctx*> ctx;
    std::vector<int> dev;
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    uint64_t epoch = 0;
    int pending = 0;
    Job job;
    std::vector<int> rc;
    std::vector<std::string> err;
    // scratch outputs of ranks > 0 (every rank computes the same values; only rank 0's are handed out)
    std::vector<std::vector<double>> g_scr, th_scr, bt_scr, big_scr;
};

namespace {

// one rank's share of a job; called on a worker thread for ranks > 0 and on the CALLER's thread for rank 0, so that
// callbacks (the callBack.m table a MEX gateway prints with mexPrintf) never run on a foreign thread
int exec_job(gpz_multi* M, int r, const Job& j) {
    int rc = GPZ_OK;
    const int64_t p = M->p;
    const int m = M->model.m, k = M->model.k;
    switch (j.cmd) {
        case CMD_INIT:
            rc = gpz_comm_init(M->ctx[r], r, M->n, j.id);
            break;
        case CMD_EVAL: {
            double f = 0.0, st[4];
            double* g = r == 0 ? j.grad : M->g_scr[r].data();
            rc = gpz_eval(M->ctx[r], j.theta, r == 0 ? j.f : &f, g, r == 0 ? j.stats : st);
            break;
        }
        case CMD_FIT: {
            if (r == 0) rc = gpz_fit(M->ctx[r], j.theta, j.nl, j.w, j.iS);
            else {
                std::vector<double>& b = M->big_scr[r];
                b.resize(static_cast<size_t>(k) + static_cast<size_t>(m) * k + static_cast<size_t>(m) * m * k);
                rc = gpz_fit(M->ctx[r], j.theta, b.data(), b.data() + k, b.data() + k + static_cast<size_t>(m) * k);
            }
            break;
        }
        case CMD_PRIOR: {
            std::vector<double> pr(static_cast<size_t>(m));
            rc = gpz_get_prior(M->ctx[r], j.theta, r == 0 ? j.prior : pr.data());
            break;
        }
        case CMD_TRAIN: {
            // the optimiser is replicated: same objective values on every rank -> same decisions
            double bv = *j.best_valid;
            gpz_train_result res;
            std::memset(&res, 0, sizeof(res));
            if (r == 0) rc = gpz_train(M->ctx[r], j.topt, const_cast<double*>(j.theta), j.best_theta, j.best_valid, j.cb, j.cb_user, j.res);
            else {
                M->th_scr[r].assign(j.theta, j.theta + p);
                M->bt_scr[r].assign(j.best_theta, j.best_theta + p);
                rc = gpz_train(M->ctx[r], j.topt, M->th_scr[r].data(), M->bt_scr[r].data(), &bv, nullptr, nullptr, &res);
            }
            break;
        }
        case CMD_OPTION:
            rc = gpz_set_option(M->ctx[r], j.opt_name, j.opt_value);
            break;
        default:
            break;
    }
    return rc;
}

void worker(gpz_multi* M, int r) {
    uint64_t seen = 0;
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(M->mu);
            M->cv_go.wait(lk, [&] { return M->epoch != seen; });
            seen = M->epoch;
            j = M->job;
        }
        const int rc = exec_job(M, r, j);
        {
            std::lock_guard<std::mutex> lk(M->mu);
            M->rc[r] = rc;
            if (rc) M->err[r] = gpz_last_error();
            if (--M->pending == 0) M->cv_done.notify_all();
        }
        if (j.cmd == CMD_QUIT) return;
    }
}

int run(gpz_multi* M, const Job& j) {
    {
        std::lock_guard<std::mutex> lk(M->mu);
        M->job = j;
        M->pending = M->n - 1;
        ++M->epoch;
    }
    M->cv_go.notify_all();
    // rank 0 on this (the caller's) thread
    M->rc[0] = exec_job(M, 0, j);
    if (M->rc[0]) M->err[0] = gpz_last_error();
    {
        std::unique_lock<std::mutex> lk(M->mu);
        M->cv_done.wait(lk, [&] { return M->pending == 0; });
    }
    for (int r = 0; r < M->n; ++r)
        if (M->rc[r]) {
            set_error("device %d (rank %d of %d): %s", M->dev[r], r, M->n, M->err[r].c_str());
            return M->rc[r];
        }
    return GPZ_OK;
}

}  // namespace

extern "C" {

int gpz_create_multi(gpz_multi** out, const gpz_model* model, int64_t n_all, const double* X, const double* Y, const double* Psi,
                     const double* omega, const uint8_t* training, const uint8_t* validation, int ngpus, const int* devices) {
    if (!out || !model || !X || n_all < 1 || ngpus < 1 || ngpus > 64) {
        set_error("gpz_create_multi: bad arguments (1 <= ngpus <= 64)");
        return GPZ_ERR_USAGE;
    }
    *out = nullptr;
    const int64_t p = gpz_theta_len(model);
    if (p < 0) return GPZ_ERR_USAGE;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
        set_error("no CUDA device available; libgpz_b200 has no CPU fallback");
        return GPZ_ERR_NODEVICE;
    }
    for (int r = 0; r < ngpus; ++r) {
        const int d = devices ? devices[r] : r;
        if (d < 0 || d >= count) {
            set_error("gpz_create_multi: device %d requested, %d visible", d, count);
            return GPZ_ERR_USAGE;
        }
    }
    // contiguous blocks of the SELECTED rows per rank: rank r keeps training rows [nt r / N, nt (r+1) / N) in
    // selection order, validation rows likewise
    int64_t nt = 0, nv = 0;
    for (int64_t i = 0; i < n_all; ++i) {
        nt += training ? (training[i] != 0) : 1;
        nv += validation ? (validation[i] != 0) : 0;
    }
    if (nt < ngpus) {
        set_error("gpz_create_multi: %lld training rows cannot be split over %d GPUs", static_cast<long long>(nt), ngpus);
        return GPZ_ERR_USAGE;
    }
    gpz_multi* M = new gpz_multi;
    M->n = ngpus;
    M->model = *model;
    M->p = p;
    M->ctx.assign(ngpus, nullptr);
    M->dev.resize(ngpus);
    M->rc.assign(ngpus, 0);
    M->err.assign(ngpus, std::string());
    M->g_scr.resize(ngpus);
    M->th_scr.resize(ngpus);
    M->bt_scr.resize(ngpus);
    M->big_scr.resize(ngpus);
    std::vector<uint8_t> tr(static_cast<size_t>(n_all)), va(static_cast<size_t>(n_all));
    int rc = GPZ_OK;
    for (int r = 0; r < ngpus && !rc; ++r) {
        M->dev[r] = devices ? devices[r] : r;
        const int64_t t0 = nt * r / ngpus, t1 = nt * (r + 1) / ngpus, v0 = nv * r / ngpus, v1 = nv * (r + 1) / ngpus;
        int64_t it = 0, iv = 0;
        for (int64_t i = 0; i < n_all; ++i) {
            const bool st = training ? (training[i] != 0) : true, sv = validation ? (validation[i] != 0) : false;
            tr[i] = st && it >= t0 && it < t1;
            va[i] = sv && iv >= v0 && iv < v1;
            it += st;
            iv += sv;
        }
        rc = gpz_create(&M->ctx[r], model, n_all, X, Y, Psi, omega, tr.data(), nv > 0 ? va.data() : nullptr, M->dev[r]);
        if (r > 0) M->g_scr[r].resize(static_cast<size_t>(p));
    }
    if (rc) {
        const std::string e = gpz_last_error();
        for (gpz_ctx* c : M->ctx) gpz_destroy(c);
        delete M;
        set_error("%s", e.c_str());
        return rc;
    }
    for (int r = 1; r < ngpus; ++r) M->th.emplace_back(worker, M, r);       // rank 0 runs on the caller's thread
    Job j;
    j.cmd = CMD_INIT;
    std::memset(j.id, 0, sizeof(j.id));
    if (ngpus > 1 && (rc = gpz_comm_unique_id(j.id))) {
        const std::string e = gpz_last_error();
        gpz_destroy_multi(M);
        set_error("%s", e.c_str());
        return rc;
    }
    if ((rc = run(M, j))) {
        const std::string e = gpz_last_error();
        gpz_destroy_multi(M);
        set_error("%s", e.c_str());
        return rc;
    }
    *out = M;
    return GPZ_OK;
}

void gpz_destroy_multi(gpz_multi* M) {
    if (!M) return;
    if (!M->th.empty()) {
        Job j;
        j.cmd = CMD_QUIT;
        run(M, j);
        for (std::thread& t : M->th) t.join();
        M->th.clear();
    }
    for (gpz_ctx* c : M->ctx) gpz_destroy(c);
    delete M;
}

int gpz_multi_devices(const gpz_multi* M) { return M ? M->n : -1; }
gpz_ctx* gpz_multi_ctx(gpz_multi* M, int rank) { return (M && rank >= 0 && rank < M->n) ? M->ctx[rank] : nullptr; }

int gpz_multi_eval(gpz_multi* M, const double* theta, double* nlogML, double* grad, double stats[4]) {
    if (!M || !theta || !nlogML || !grad || !stats) {
        set_error("gpz_multi_eval: NULL argument");
        return GPZ_ERR_USAGE;
    }
    Job j;
    j.cmd = CMD_EVAL;
    j.theta = theta;
    j.f = nlogML;
    j.grad = grad;
    j.stats = stats;
    return run(M, j);
}

int gpz_multi_fit(gpz_multi* M, const double* theta, double* nlogML_k, double* w, double* iSigma_w) {
    if (!M || !theta || !w || !iSigma_w) {
        set_error("gpz_multi_fit: NULL argument");
        return GPZ_ERR_USAGE;
    }
    Job j;
    j.cmd = CMD_FIT;
    j.theta = theta;
    j.nl = nlogML_k;
    j.w = w;
    j.iS = iSigma_w;
    return run(M, j);
}

int gpz_multi_get_prior(gpz_multi* M, const double* theta, double* prior) {
    if (!M || !theta || !prior) {
        set_error("gpz_multi_get_prior: NULL argument");
        return GPZ_ERR_USAGE;
    }
    Job j;
    j.cmd = CMD_PRIOR;
    j.theta = theta;
    j.prior = prior;
    return run(M, j);
}

int gpz_multi_train(gpz_multi* M, const gpz_train_options* opt, double* theta, double* best_theta, double* best_valid,
                    gpz_train_callback cb, void* user, gpz_train_result* res) {
    if (!M || !opt || !theta || !best_theta || !best_valid || !res) {
        set_error("gpz_multi_train: NULL argument");
        return GPZ_ERR_USAGE;
    }
    if (opt->training_only < 0) {
        set_error("gpz_multi_train: set training_only explicitly (a rank's own validation rows do not tell)");
        return GPZ_ERR_USAGE;
    }
    Job j;
    j.cmd = CMD_TRAIN;
    j.theta = theta;
    j.topt = opt;
    j.best_theta = best_theta;
    j.best_valid = best_valid;
    j.cb = cb;
    j.cb_user = user;
    j.res = res;
    return run(M, j);
}

int gpz_multi_set_option(gpz_multi* M, const char* name, double value) {
    if (!M || !name) return GPZ_ERR_USAGE;
    Job j;
    j.cmd = CMD_OPTION;
    j.opt_name = name;
    j.opt_value = value;
    return run(M, j);
}

}  // extern "C"
