// train.cu -- the optimiser GPz trains with, device-resident.
//
// Reference behaviour: GPz/train.m:38-48 runs M. Schmidt's minFunc with method 'lbfgs' and GPz/callBack.m as the
// output function.  On that configuration minFunc is: L-BFGS with a circular store of 100 (s,y) pairs
// (minFunc.m:560-576, lbfgsAdd.m, lbfgsProd.m / mex/lbfgsProdC.c), a bracketing Wolfe line search with cubic
// interpolation (WolfeLineSearch.m:32-263, polyinterp.m:41-58) that falls back to Armijo backtracking when it
// steps into a non-finite region (ArmijoBacktrack.m:32-143), and the stopping rules of minFunc.m:1094-1152.
//
// B200 design: theta, the gradients of the bracket points, the direction and the (S,Y) history never leave
// HBM.  The reference's two-loop recursion streams S and Y four times through ONE host core, strictly one
// vector after the other (O(p x corrections) = 350 MB per iteration at the headline size -- as long as our
// whole objective evaluation on 8 GPUs).  Here the recursion runs in COEFFICIENT space: the direction is a
// linear combination of {s_i}, {y_i} and g, so with the Gram matrix B of those 2c+1 vectors the two loops
// only touch B (host, O(c^2) flops).  Per iteration the device makes two full-bandwidth passes over the
// history: one that adds the rows of B for the new s, y and g (rows kernel), one that forms the direction
// from the coefficients (combine kernel).  Everything is summed in a fixed order: runs are reproducible and
// every rank of a sharded run takes bit-identical decisions (f and g are identical after the all-reduce).
#include <math.h>
#include <string.h>

#include <chrono>
#include <limits>
#include <vector>

#include "internal.cuh"

namespace gpz {
namespace {

constexpr int RT = 256;        // threads of the vector kernels
constexpr int RMAXB = 148;     // blocks of a reduction over one p-vector
constexpr int RSEG = 8192;     // elements per CTA of the rows kernel

template <int NT>
__device__ __forceinline__ double block_max(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) r = fmax(r, sh[i]);
    }
    return r;
}

enum { OP_DIR = 0, OP_EVAL = 1, OP_PAIR = 2 };

struct RedArgs {
    const double* a;
    const double* b;
    const double* c;
    double* o1;
    double* o2;
    double t;
    int64_t p;
    double* partial;          // [gridDim.x][4]
    unsigned int* ticket;
    double* res;              // [9]
    const double* out;        // OP_EVAL: evaluation buffer [f, g[p], stats[4]]
};

// One pass over a p-vector pair, four reductions, finished by the last block in block order.
//   OP_DIR : a = g, b = d            -> res = { g'd, sum|g|, max|d|, d non-finite }
//   OP_EVAL: a = g_new, b = d, o1 = copy of g_new -> res = { g_new'd, 0, max|g_new|, g_new non-finite, f, stats[4] }
//   OP_PAIR: a = g_new, b = g_old, c = d: o1 = s = t d, o2 = y = g_new - g_old -> res = { y's, y'y }
template <int OP>
__global__ void __launch_bounds__(RT) reduce_kernel(const RedArgs A) {
    __shared__ double sh[RT / 32];
    __shared__ int last;
    const int64_t chunk = (A.p + gridDim.x - 1) / gridDim.x;
    const int64_t i0 = chunk * blockIdx.x, i1 = min(A.p, i0 + chunk);
    double s0 = 0.0, s1 = 0.0, m2 = 0.0, m3 = 0.0;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += RT) {
        if (OP == OP_DIR) {
            const double g = A.a[i], d = A.b[i];
            s0 += g * d;
            s1 += fabs(g);
            m2 = fmax(m2, fabs(d));
            if (!isfinite(d)) m3 = 1.0;
        } else if (OP == OP_EVAL) {
            const double g = A.a[i], d = A.b[i];
            A.o1[i] = g;                          // keep this point's gradient (the objective always writes one fixed buffer)
            s0 += g * d;
            m2 = fmax(m2, fabs(g));
            if (!isfinite(g)) m3 = 1.0;
        } else {
            const double y = A.a[i] - A.b[i], s = A.t * A.c[i];
            A.o1[i] = s;
            A.o2[i] = y;
            s0 += y * s;
            s1 += y * y;
        }
    }
    const double r0 = block_sum<RT>(s0, sh);
    const double r1 = block_sum<RT>(s1, sh);
    const double r2 = block_max<RT>(m2, sh);
    const double r3 = block_max<RT>(m3, sh);
    if (threadIdx.x == 0) {
        double* q = A.partial + 4 * blockIdx.x;
        q[0] = r0;
        q[1] = r1;
        q[2] = r2;
        q[3] = r3;
        __threadfence();
        last = atomicAdd(A.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) {
            t0 += __ldcg(A.partial + 4 * b);
            t1 += __ldcg(A.partial + 4 * b + 1);
            t2 = fmax(t2, __ldcg(A.partial + 4 * b + 2));
            t3 = fmax(t3, __ldcg(A.partial + 4 * b + 3));
        }
        A.res[0] = t0;
        A.res[1] = t1;
        A.res[2] = t2;
        A.res[3] = t3;
        if (OP == OP_EVAL) {
            A.res[4] = A.out[0];
            for (int q = 0; q < 4; ++q) A.res[5 + q] = A.out[1 + A.p + q];
        }
        *A.ticket = 0u;
    }
}

__global__ void __launch_bounds__(RT) axpy_kernel(double* xt, const double* x, const double* __restrict__ d, double t,
                                                  int64_t p) {      // xt may alias x
    const int64_t i = static_cast<int64_t>(blockIdx.x) * RT + threadIdx.x;
    if (i < p) xt[i] = x[i] + t * d[i];
}

__global__ void __launch_bounds__(RT) neg_kernel(double* __restrict__ d, const double* __restrict__ g, int64_t p) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * RT + threadIdx.x;
    if (i < p) d[i] = -g[i];
}

// Rows of the Gram matrix of the basis {S(:,0..nf-1), Y(:,0..nf-1), g} against (v0, v1, v2) = (new s, new y, g):
// block (seg, j) sums one segment of basis vector j; the finish kernel adds the segments in order.
__global__ void __launch_bounds__(RT) rows_kernel(const double* __restrict__ S, const double* __restrict__ Y,
                                                  const double* __restrict__ g, int nf, int64_t p, int64_t ldp,
                                                  const double* __restrict__ v0, const double* __restrict__ v1,
                                                  const double* __restrict__ v2, double* __restrict__ part) {
    __shared__ double sh[RT / 32];
    const int j = blockIdx.y;
    const double* b = j < nf ? S + ldp * j : (j < 2 * nf ? Y + ldp * (j - nf) : g);
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * RSEG, i1 = min(p, i0 + RSEG);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += RT) {
        const double a = b[i];
        s0 += a * v0[i];
        s1 += a * v1[i];
        s2 += a * v2[i];
    }
    const double r0 = block_sum<RT>(s0, sh);
    const double r1 = block_sum<RT>(s1, sh);
    const double r2 = block_sum<RT>(s2, sh);
    if (threadIdx.x == 0) {
        double* q = part + 3 * (static_cast<int64_t>(j) * gridDim.x + blockIdx.x);
        q[0] = r0;
        q[1] = r1;
        q[2] = r2;
    }
}

__global__ void rows_finish_kernel(const double* __restrict__ part, int nvec, int nseg, double* __restrict__ out) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nvec * 3) return;
    const int j = id / 3, r = id - 3 * j;
    double s = 0.0;
    for (int q = 0; q < nseg; ++q) s += part[3 * (static_cast<int64_t>(j) * nseg + q) + r];
    out[id] = s;
}

// d = delta_g g + sum_j (delta_s[j] S(:,j) + delta_y[j] Y(:,j)); delta = [delta_s[nf], delta_y[nf], delta_g]
__global__ void __launch_bounds__(RT) combine_kernel(const double* __restrict__ S, const double* __restrict__ Y,
                                                     const double* __restrict__ g, int nf, int64_t p, int64_t ldp,
                                                     const double* __restrict__ delta, double* __restrict__ d) {
    extern __shared__ double sd[];
    for (int q = threadIdx.x; q < 2 * nf + 1; q += RT) sd[q] = delta[q];
    __syncthreads();
    const int64_t i = static_cast<int64_t>(blockIdx.x) * RT + threadIdx.x;
    if (i >= p) return;
    double acc = sd[2 * nf] * g[i];
    for (int j = 0; j < nf; ++j) {
        acc += sd[j] * S[ldp * j + i];
        acc += sd[nf + j] * Y[ldp * j + i];
    }
    d[i] = acc;
}

// polyinterp.m:41-58 (two points, values and slopes known); fmin/fmax drop a NaN operand like MATLAB's min/max
double cubic2(double x0, double f0, double g0, double x1, double f1, double g1, double lo, double hi) {
    double xa = x0, fa = f0, ga = g0, xb = x1, fb = f1, gb = g1;
    if (!(x0 <= x1)) {
        xa = x1, fa = f1, ga = g1;
        xb = x0, fb = f0, gb = g0;
    }
    const double d1 = ga + gb - 3.0 * (fa - fb) / (xa - xb);
    const double rad = d1 * d1 - ga * gb;
    if (rad < 0.0) return (hi + lo) / 2.0;
    const double d2 = sqrt(rad);
    const double t = xb - (xb - xa) * ((gb + d2 - d1) / (gb - ga + 2.0 * d2));
    return fmin(fmax(t, lo), hi);
}
double cubic2(double x0, double f0, double g0, double x1, double f1, double g1) {
    return cubic2(x0, f0, g0, x1, f1, g1, fmin(x0, x1), fmax(x0, x1));
}
// polyinterp.m:60-111 for [0 f0 g0; t f1 ?]: quadratic fit, lowest value among bounds, abscissae and the root
double quad2(double f0, double g0, double t, double f1, double lo, double hi) {
    const double a = (f1 - f0 - g0 * t) / (t * t), b = g0, c = f0;
    double cand[5] = {lo, hi, 0.0, t, 0.0};
    int nc = 4;
    if (isfinite(2.0 * a) && isfinite(b) && a != 0.0) cand[nc++] = -b / (2.0 * a);
    double best = (lo + hi) / 2.0, fbest = std::numeric_limits<double>::infinity();
    for (int q = 0; q < nc; ++q) {
        const double x = cand[q];
        if (x >= lo && x <= hi) {
            const double fx = (a * x + b) * x + c;
            if (fx < fbest) best = x, fbest = fx;
        }
    }
    return best;
}

struct Pt {                // one evaluated point of the line search
    double t = 0.0, f = 0.0, gtd = 0.0, optc = 0.0;
    int buf = -1;          // which evaluation buffer holds its gradient
    bool legal_g = true;
};

struct Trainer {
    int64_t p = 0;
    gpz_objective_dev fn = nullptr;
    void* fn_user = nullptr;
    gpz_train_options o{};
    cudaStream_t st = nullptr;
    int64_t* launches = nullptr;
    int64_t own_launches = 0;
    // device
    double *x = nullptr, *xt = nullptr, *d = nullptr, *s_tmp = nullptr, *y_tmp = nullptr, *best = nullptr;
    double* out[4] = {nullptr, nullptr, nullptr, nullptr};     // gradients of the line-search points ([0] unused, g at [1..p])
    double* out_eval = nullptr;    // the objective always reads xt and writes here: one fixed pointer pair, so a context can
                                   // replay its evaluation as a CUDA graph (api.cu eval_device)
    double *S = nullptr, *Y = nullptr, *partial = nullptr, *res = nullptr, *rows_part = nullptr, *rows_out = nullptr, *delta = nullptr;
    unsigned int* ticket = nullptr;
    // pinned host
    double *h_res = nullptr, *h_rows = nullptr, *h_delta = nullptr;
    std::vector<void*> dev_allocs, host_allocs;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // L-BFGS state
    int cols = 1;                  // allocated history columns
    int64_t ldp = 0;
    int start = 1, end = 0;        // 1-based, lbfgsAdd.m
    int nf = 0;                    // filled slots
    double Hdiag = 1.0;
    std::vector<double> B, YS, al, dl;
    int nb = 0;                    // edge of B: 2 cols + 1
    int cur = 0;                   // buffer of the gradient at x
    double last_stats[4] = {NAN, NAN, NAN, NAN};
    double ms_eval = 0.0;
    int fun_evals = 0;
    int skipped = 0;

    ~Trainer() {
        for (void* q : dev_allocs) cudaFree(q);
        for (void* q : host_allocs) cudaFreeHost(q);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
    void count(int n = 1) {
        own_launches += n;
        if (launches) *launches += n;
    }
    int dalloc(double** q, int64_t n) {
        GPZ_CUDA(cudaMalloc(q, sizeof(double) * static_cast<size_t>(n > 0 ? n : 1)));
        dev_allocs.push_back(*q);
        return GPZ_OK;
    }
    int halloc(double** q, int64_t n) {
        GPZ_CUDA(cudaMallocHost(q, sizeof(double) * static_cast<size_t>(n > 0 ? n : 1)));
        host_allocs.push_back(*q);
        return GPZ_OK;
    }
    int blocks() const { return static_cast<int>(ceil_div(p, RT)); }
    int rblocks() const { return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(RMAXB, ceil_div(p, 4 * RT)))); }
    double* grad(int b) const { return out[b] + 1; }

    int setup() {
        int rc;
        // one pair is staged per iteration: a run shorter than the history needs fewer columns (the circular logic
        // keeps using o.corrections, it just never wraps)
        cols = std::max(1, std::min(o.corrections, o.max_iter));
        ldp = round_up(p, 32);
        if ((rc = dalloc(&x, p)) || (rc = dalloc(&xt, p)) || (rc = dalloc(&d, p)) || (rc = dalloc(&s_tmp, p)) ||
            (rc = dalloc(&y_tmp, p)) || (rc = dalloc(&best, p)))
            return rc;
        for (int b = 0; b < 4; ++b)
            if ((rc = dalloc(&out[b], p + 5))) return rc;
        if ((rc = dalloc(&out_eval, p + 5))) return rc;
        if ((rc = dalloc(&S, ldp * cols)) || (rc = dalloc(&Y, ldp * cols))) return rc;
        nb = 2 * cols + 1;
        const int nseg = static_cast<int>(ceil_div(p, RSEG));
        if ((rc = dalloc(&partial, 4 * RMAXB)) || (rc = dalloc(&res, 16)) || (rc = dalloc(&rows_part, 3ll * nb * nseg)) ||
            (rc = dalloc(&rows_out, 3ll * nb)) || (rc = dalloc(&delta, nb)))
            return rc;
        double* tk = nullptr;
        if ((rc = dalloc(&tk, 1))) return rc;
        ticket = reinterpret_cast<unsigned int*>(tk);
        GPZ_CUDA(cudaMemsetAsync(ticket, 0, sizeof(double), st));
        GPZ_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * p, st));       // the first evaluation forms x + 0 * d
        if ((rc = halloc(&h_res, 16)) || (rc = halloc(&h_rows, 3ll * nb)) || (rc = halloc(&h_delta, nb))) return rc;
        GPZ_CUDA(cudaEventCreate(&ev0));
        GPZ_CUDA(cudaEventCreate(&ev1));
        B.assign(static_cast<size_t>(nb) * nb, 0.0);
        YS.assign(cols, 0.0);
        al.assign(cols, 0.0);
        dl.assign(nb, 0.0);
        return GPZ_OK;
    }

    // f, g at x + t d (at_x: at x itself); the gradient lands in buffer `buf`
    int eval(double t, int buf, bool at_x, Pt* pt) {
        axpy_kernel<<<blocks(), RT, 0, st>>>(xt, x, d, at_x ? 0.0 : t, p);
        GPZ_KERNEL_CHECK();
        count();
        GPZ_CUDA(cudaEventRecord(ev0, st));
        const int rc = fn(fn_user, xt, out_eval, st);
        if (rc) return rc;
        GPZ_CUDA(cudaEventRecord(ev1, st));
        RedArgs A{};
        A.a = out_eval + 1;
        A.b = at_x ? out_eval + 1 : d;
        A.o1 = grad(buf);
        A.p = p;
        A.partial = partial;
        A.ticket = ticket;
        A.res = res;
        A.out = out_eval;
        reduce_kernel<OP_EVAL><<<rblocks(), RT, 0, st>>>(A);
        GPZ_KERNEL_CHECK();
        count();
        GPZ_CUDA(cudaMemcpyAsync(h_res, res, sizeof(double) * 9, cudaMemcpyDeviceToHost, st));
        GPZ_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        GPZ_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        ms_eval += ms;
        ++fun_evals;
        pt->t = t;
        pt->f = h_res[4];
        pt->gtd = h_res[0];
        pt->optc = h_res[2];
        pt->legal_g = h_res[3] == 0.0;
        pt->buf = buf;
        memcpy(last_stats, h_res + 5, sizeof(double) * 4);
        return GPZ_OK;
    }

    int free_buf(int a, int b = -1, int c = -1) const {
        for (int q = 0; q < 4; ++q)
            if (q != a && q != b && q != c) return q;
        return -1;
    }
    static bool legal(const Pt& q) { return isfinite(q.f) && q.legal_g; }

    // ArmijoBacktrack.m:32-143 (LS_interp 2, LS_multi 0), entered from the Wolfe search only
    int armijo(double t, double f, double gtd, double nrm_d, Pt* acc) {
        int rc;
        Pt nw;
        if ((rc = eval(t, free_buf(cur), false, &nw))) return rc;
        while (!isfinite(nw.f) || nw.f > f + o.c1 * t * gtd) {
            const double temp = t;
            if (!isfinite(nw.f))
                t = 0.5 * t;
            else if (!nw.legal_g)
                t = quad2(f, gtd, t, nw.f, 0.0, t);
            else
                t = cubic2(0.0, f, gtd, t, nw.f, nw.gtd, 0.0, t);
            if (t < temp * 1e-3)
                t = temp * 1e-3;
            else if (t > temp * 0.6)
                t = temp * 0.6;
            if ((rc = eval(t, free_buf(cur), false, &nw))) return rc;
            if (fabs(t) * nrm_d <= o.prog_tol) {         // max(abs(t*d)) <= progTol: give up, stay at x
                acc->t = 0.0;
                acc->f = f;
                acc->buf = cur;
                return GPZ_OK;
            }
        }
        *acc = nw;
        return GPZ_OK;
    }

    // WolfeLineSearch.m:32-263 (LS_interp 2)
    int wolfe(double t, double f, double gtd, double nrm_d, Pt* acc) {
        int rc;
        Pt p0;
        p0.t = 0.0, p0.f = f, p0.gtd = gtd, p0.buf = cur;
        Pt nw, prev = p0, br[2];
        int nbr = 0, ls = 0;
        bool done = false;
        if ((rc = eval(t, free_buf(cur), false, &nw))) return rc;
        while (ls < o.max_ls) {
            if (!legal(nw)) return armijo((nw.t + prev.t) / 2.0, f, gtd, nrm_d, acc);
            if (nw.f > f + o.c1 * nw.t * gtd || (ls > 1 && nw.f >= prev.f)) {
                br[0] = prev, br[1] = nw, nbr = 2;
                break;
            } else if (fabs(nw.gtd) <= -o.c2 * gtd) {
                br[0] = nw, nbr = 1, done = true;
                break;
            } else if (nw.gtd >= 0.0) {
                br[0] = prev, br[1] = nw, nbr = 2;
                break;
            }
            const double temp = prev.t;
            const double min_step = nw.t + 0.01 * (nw.t - temp), max_step = nw.t * 10.0;
            const double tn = cubic2(temp, prev.f, prev.gtd, nw.t, nw.f, nw.gtd, min_step, max_step);
            prev = nw;
            if ((rc = eval(tn, free_buf(cur, prev.buf), false, &nw))) return rc;
            ++ls;
        }
        if (ls == o.max_ls) br[0] = p0, br[1] = nw, nbr = 2;
        bool insuf = false;
        while (!done && ls < o.max_ls) {
            const int lo = (br[1].f < br[0].f || (isnan(br[0].f) && !isnan(br[1].f))) ? 1 : 0, hi = 1 - lo;
            const double f_lo = br[lo].f;
            double tn;
            if (!legal(br[0]) || !legal(br[1]))
                tn = (br[0].t + br[1].t) / 2.0;
            else
                tn = cubic2(br[0].t, br[0].f, br[0].gtd, br[1].t, br[1].f, br[1].gtd);
            const double bmax = fmax(br[0].t, br[1].t), bmin = fmin(br[0].t, br[1].t);
            if (fmin(bmax - tn, tn - bmin) / (bmax - bmin) < 0.1) {
                if (insuf || tn >= bmax || tn <= bmin) {
                    tn = fabs(tn - bmax) < fabs(tn - bmin) ? bmax - 0.1 * (bmax - bmin) : bmin + 0.1 * (bmax - bmin);
                    insuf = false;
                } else {
                    insuf = true;
                }
            } else {
                insuf = false;
            }
            if ((rc = eval(tn, free_buf(cur, br[0].buf, br[1].buf), false, &nw))) return rc;
            ++ls;
            const bool arm = nw.f < f + o.c1 * nw.t * gtd;
            if (!arm || nw.f >= f_lo) {
                br[hi] = nw;
            } else {
                if (fabs(nw.gtd) <= -o.c2 * gtd)
                    done = true;
                else if (nw.gtd * (br[hi].t - br[lo].t) >= 0.0)
                    br[hi] = br[lo];
                br[lo] = nw;
            }
            if (!done && fabs(br[0].t - br[1].t) * nrm_d < o.prog_tol) break;
        }
        int lo = 0;
        if (nbr == 2 && (br[1].f < br[0].f || (isnan(br[0].f) && !isnan(br[1].f)))) lo = 1;
        *acc = br[lo];
        return GPZ_OK;
    }

    // lbfgsAdd.m:2-30 on the staged pair (s_tmp, y_tmp), then the rows of B for the new s, y and the current g
    int add_pair_and_rows(double ys, double yy, bool have_pair) {
        int slot = -1;
        if (have_pair) {
            if (ys > 1e-10) {
                if (end < o.corrections) {
                    ++end;
                    if (start != 1) start = start == o.corrections ? 1 : start + 1;
                } else {
                    start = std::min(2, o.corrections);
                    end = 1;
                }
                slot = end - 1;
                if (slot >= cols) {
                    set_error("gpz_train: history slot %d outside the %d allocated columns", slot, cols);
                    return GPZ_ERR_USAGE;
                }
                GPZ_CUDA(cudaMemcpyAsync(S + ldp * slot, s_tmp, sizeof(double) * p, cudaMemcpyDeviceToDevice, st));
                GPZ_CUDA(cudaMemcpyAsync(Y + ldp * slot, y_tmp, sizeof(double) * p, cudaMemcpyDeviceToDevice, st));
                YS[slot] = ys;
                Hdiag = ys / yy;
                nf = std::max(nf, end);
            } else {
                ++skipped;
            }
        }
        const int nvec = 2 * nf + 1, nseg = static_cast<int>(ceil_div(p, RSEG));
        const double* g = grad(cur);
        rows_kernel<<<dim3(nseg, nvec), RT, 0, st>>>(S, Y, g, nf, p, ldp, slot >= 0 ? S + ldp * slot : g,
                                                     slot >= 0 ? Y + ldp * slot : g, g, rows_part);
        GPZ_KERNEL_CHECK();
        rows_finish_kernel<<<static_cast<int>(ceil_div(3 * nvec, 128)), 128, 0, st>>>(rows_part, nvec, nseg, rows_out);
        GPZ_KERNEL_CHECK();
        count(2);
        GPZ_CUDA(cudaMemcpyAsync(h_rows, rows_out, sizeof(double) * 3 * nvec, cudaMemcpyDeviceToHost, st));
        GPZ_CUDA(cudaStreamSynchronize(st));
        // basis index: s_i -> i, y_i -> cols + i, g -> 2 cols
        auto idx = [&](int j) { return j < nf ? j : (j < 2 * nf ? cols + (j - nf) : 2 * cols); };
        for (int j = 0; j < nvec; ++j) {
            const int bj = idx(j);
            if (slot >= 0) {
                B[static_cast<size_t>(slot) * nb + bj] = B[static_cast<size_t>(bj) * nb + slot] = h_rows[3 * j];
                B[static_cast<size_t>(cols + slot) * nb + bj] = B[static_cast<size_t>(bj) * nb + cols + slot] = h_rows[3 * j + 1];
            }
            B[static_cast<size_t>(2 * cols) * nb + bj] = B[static_cast<size_t>(bj) * nb + 2 * cols] = h_rows[3 * j + 2];
        }
        return GPZ_OK;
    }

    // lbfgsProd.m:9-32 in coefficient space, then d = basis * coefficients on the device
    int direction() {
        std::vector<int> ind;
        if (start == 1) {
            for (int i = 0; i < end; ++i) ind.push_back(i);
        } else {
            for (int i = start - 1; i < o.corrections; ++i) ind.push_back(i);
            for (int i = 0; i < end; ++i) ind.push_back(i);
        }
        const int G = 2 * cols;
        std::fill(dl.begin(), dl.end(), 0.0);
        dl[G] = -1.0;
        auto dotB = [&](int row) {
            const double* br = B.data() + static_cast<size_t>(row) * nb;
            double s = br[G] * dl[G];
            for (int j = 0; j < nf; ++j) s += br[j] * dl[j] + br[cols + j] * dl[cols + j];
            return s;
        };
        for (int q = static_cast<int>(ind.size()) - 1; q >= 0; --q) {
            const int i = ind[q];
            al[i] = dotB(i) / YS[i];
            dl[cols + i] -= al[i];
        }
        for (int j = 0; j < nf; ++j) dl[j] *= Hdiag, dl[cols + j] *= Hdiag;
        dl[G] *= Hdiag;
        for (int i : ind) {
            const double be = dotB(cols + i) / YS[i];
            dl[i] += al[i] - be;
        }
        for (int j = 0; j < nf; ++j) h_delta[j] = dl[j], h_delta[nf + j] = dl[cols + j];
        h_delta[2 * nf] = dl[G];
        GPZ_CUDA(cudaMemcpyAsync(delta, h_delta, sizeof(double) * (2 * nf + 1), cudaMemcpyHostToDevice, st));
        combine_kernel<<<blocks(), RT, sizeof(double) * (2 * nf + 1), st>>>(S, Y, grad(cur), nf, p, ldp, delta, d);
        GPZ_KERNEL_CHECK();
        count();
        return GPZ_OK;
    }

    int run(double* theta, double* best_theta, double* best_valid, gpz_train_callback cb, void* user, gpz_train_result* R) {
        int rc;
        const auto t_begin = std::chrono::steady_clock::now();
        GPZ_CUDA(cudaMemcpyAsync(x, theta, sizeof(double) * p, cudaMemcpyHostToDevice, st));
        GPZ_CUDA(cudaMemcpyAsync(best, best_theta, sizeof(double) * p, cudaMemcpyHostToDevice, st));
        double bv = *best_valid;
        int attempts = -1;                                   // the reference's global starts out empty (train.m:3)
        Pt at;
        cur = 0;
        if ((rc = eval(0.0, cur, true, &at))) return rc;     // minFunc.m:313-314
        double f = at.f, opt_cond = at.optc, t = 1.0, gtd = 0.0, nrm_d = 0.0;
        int exitflag = 0, reason = 0, i = 0;
        if (opt_cond <= o.opt_tol) {                         // minFunc.m:350-362
            exitflag = 1, reason = 9;
        } else {
            for (i = 1; i <= o.max_iter; ++i) {
                if (i == 1) {                                // minFunc.m:562-569
                    neg_kernel<<<blocks(), RT, 0, st>>>(d, grad(cur), p);
                    GPZ_KERNEL_CHECK();
                    count();
                } else {
                    if ((rc = direction())) return rc;       // minFunc.m:571-576
                }
                RedArgs A{};
                A.a = grad(cur), A.b = d, A.p = p, A.partial = partial, A.ticket = ticket, A.res = res;
                reduce_kernel<OP_DIR><<<rblocks(), RT, 0, st>>>(A);
                GPZ_KERNEL_CHECK();
                count();
                GPZ_CUDA(cudaMemcpyAsync(h_res, res, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
                GPZ_CUDA(cudaStreamSynchronize(st));
                gtd = h_res[0];
                nrm_d = h_res[2];
                if (h_res[3] != 0.0) {                       // minFunc.m:963-967
                    exitflag = -3, reason = 8;
                    break;
                }
                if (gtd > -o.prog_tol) {                     // minFunc.m:975-979
                    exitflag = 2, reason = 2;
                    break;
                }
                t = i == 1 ? fmin(1.0, 1.0 / h_res[1]) : 1.0;    // minFunc.m:982-991 (LS_init 0)
                const double f_old = f;
                Pt acc;
                if ((rc = wolfe(t, f, gtd, nrm_d, &acc))) return rc;     // minFunc.m:1061-1068
                t = acc.t;
                f = acc.f;
                if (acc.buf != cur) {
                    // the pair lbfgsAdd will see at the next iteration: y = g - g_old, s = t d (minFunc.m:571)
                    RedArgs Pp{};
                    Pp.a = grad(acc.buf), Pp.b = grad(cur), Pp.c = d, Pp.o1 = s_tmp, Pp.o2 = y_tmp, Pp.t = t, Pp.p = p;
                    Pp.partial = partial, Pp.ticket = ticket, Pp.res = res;
                    reduce_kernel<OP_PAIR><<<rblocks(), RT, 0, st>>>(Pp);
                    GPZ_KERNEL_CHECK();
                    axpy_kernel<<<blocks(), RT, 0, st>>>(x, x, d, t, p);     // x = x + t d
                    GPZ_KERNEL_CHECK();
                    count(2);
                    GPZ_CUDA(cudaMemcpyAsync(h_res, res, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
                    GPZ_CUDA(cudaStreamSynchronize(st));
                    opt_cond = acc.optc;
                    cur = acc.buf;
                    if ((rc = add_pair_and_rows(h_res[0], h_res[1], true))) return rc;
                } else {
                    // failed backtracking: t = 0, same point; y = 0 so lbfgsAdd skips the pair
                    if ((rc = add_pair_and_rows(0.0, 0.0, true))) return rc;
                }
                // callBack.m:20-35,48
                gpz_train_iter it{};
                it.iter = i, it.fun_evals = fun_evals, it.f = f, it.t = t, it.gtd = gtd, it.opt_cond = opt_cond;
                memcpy(it.stats, last_stats, sizeof(it.stats));
                it.improved = 1;
                if (o.training_only) {
                    bv = last_stats[1];
                } else if (isnan(bv) || last_stats[3] >= bv) {
                    bv = last_stats[3];
                    attempts = 0;
                } else {
                    it.improved = 0;
                    if (attempts >= 0) ++attempts;
                }
                if (it.improved) GPZ_CUDA(cudaMemcpyAsync(best, x, sizeof(double) * p, cudaMemcpyDeviceToDevice, st));
                it.attempts = attempts;
                bool stop = attempts >= 0 && static_cast<double>(attempts) == o.max_attempts;
                if (cb && cb(user, &it)) stop = true;
                if (stop) {                                   // minFunc.m:1109-1116
                    exitflag = -1, reason = 7;
                    break;
                }
                if (opt_cond <= o.opt_tol) {                  // minFunc.m:1119-1123
                    exitflag = 1, reason = 1;
                    break;
                }
                if (fabs(t) * nrm_d <= o.prog_tol) {          // minFunc.m:1127-1131
                    exitflag = 2, reason = 3;
                    break;
                }
                if (fabs(f - f_old) < o.prog_tol) {           // minFunc.m:1134-1138
                    exitflag = 2, reason = 4;
                    break;
                }
                if (static_cast<double>(fun_evals) >= o.max_fun_evals) {     // minFunc.m:1142-1146
                    exitflag = 0, reason = 5;
                    break;
                }
                if (i == o.max_iter) {                        // minFunc.m:1148-1152
                    exitflag = 0, reason = 6;
                    break;
                }
            }
        }
        GPZ_CUDA(cudaMemcpyAsync(theta, x, sizeof(double) * p, cudaMemcpyDeviceToHost, st));
        GPZ_CUDA(cudaMemcpyAsync(best_theta, best, sizeof(double) * p, cudaMemcpyDeviceToHost, st));
        GPZ_CUDA(cudaStreamSynchronize(st));
        *best_valid = bv;
        if (R) {
            R->iterations = std::min(i, o.max_iter);
            R->fun_evals = fun_evals;
            R->exitflag = exitflag;
            R->reason = reason;
            R->attempts = attempts;
            R->skipped_pairs = skipped;
            R->f = f;
            R->opt_cond = opt_cond;
            R->best_valid = bv;
            R->ms_eval = ms_eval;
            R->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        }
        return GPZ_OK;
    }
};

}  // namespace

int lbfgs_train(int64_t p, gpz_objective_dev fn, void* fn_user, const gpz_train_options* opt, double* theta,
                double* best_theta, double* best_valid, gpz_train_callback cb, void* user, gpz_train_result* res,
                cudaStream_t st, int64_t* launches) {
    if (p < 1 || !fn || !theta || !best_theta || !best_valid) {
        set_error("gpz_train: NULL argument or empty theta");
        return GPZ_ERR_USAGE;
    }
    Trainer T;
    T.p = p;
    T.fn = fn;
    T.fn_user = fn_user;
    if (opt)
        T.o = *opt;
    else
        gpz_train_default_options(&T.o);
    if (T.o.max_iter < 1 || T.o.corrections < 1 || T.o.max_ls < 1 || T.o.training_only < 0) {
        set_error("gpz_train: bad options (max_iter %d, corrections %d, max_ls %d, training_only %d)", T.o.max_iter,
                  T.o.corrections, T.o.max_ls, T.o.training_only);
        return GPZ_ERR_USAGE;
    }
    T.st = st;
    T.launches = launches;
    int rc;
    if ((rc = T.setup())) return rc;
    return T.run(theta, best_theta, best_valid, cb, user, res);
}

}  // namespace gpz

extern "C" {

void gpz_train_default_options(gpz_train_options* o) {
    if (!o) return;
    o->max_iter = 200;
    o->training_only = -1;
    o->max_attempts = std::numeric_limits<double>::infinity();
    o->corrections = 100;
    o->max_ls = 25;
    o->opt_tol = 1e-5;
    o->prog_tol = 1e-9;
    o->c1 = 1e-4;
    o->c2 = 0.9;
    o->max_fun_evals = std::numeric_limits<double>::infinity();
}

const char* gpz_train_reason(int reason) {
    static const char* msg[10] = {"",
                                  "Optimality Condition below optTol",
                                  "Directional Derivative below progTol",
                                  "Step Size below progTol",
                                  "Function Value changing by less than progTol",
                                  "Reached Maximum Number of Function Evaluations",
                                  "Reached Maximum Number of Iterations",
                                  "Stopped by output function",
                                  "Step direction is illegal",
                                  "Optimality Condition below optTol (initial point)"};
    return reason >= 0 && reason < 10 ? msg[reason] : "";
}

int gpz_minimize_dev(int64_t p, gpz_objective_dev fn, void* fn_user, const gpz_train_options* opt, double* theta,
                     double* best_theta, double* best_valid, gpz_train_callback cb, void* user, gpz_train_result* res,
                     int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        gpz::set_error("gpz_minimize_dev: no CUDA device");
        return GPZ_ERR_NODEVICE;
    }
    GPZ_CUDA(cudaSetDevice(device));
    gpz_train_options o;
    if (opt)
        o = *opt;
    else
        gpz_train_default_options(&o);
    if (o.training_only < 0) o.training_only = 1;
    cudaStream_t st = nullptr;
    GPZ_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const int rc = gpz::lbfgs_train(p, fn, fn_user, &o, theta, best_theta, best_valid, cb, user, res, st, nullptr);
    cudaStreamDestroy(st);
    return rc;
}

}  // extern "C"
