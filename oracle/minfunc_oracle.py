"""minfunc_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

numpy restatement of the optimiser GPz trains with: M. Schmidt's minFunc (vendored in the reference as
minFunc_2012/), restricted to the configuration GPz/train.m:42-48 selects --

    options.method = 'lbfgs', maxIter = <user>, MaxFunEvals = inf, outputFcn = callBack, everything else default
    (minFunc_processInputOptions.m:62-67,117-146): Wolfe bracketing line search (LS_type 1) with cubic
    interpolation (LS_interp 2, LS_multi 0), unit initial step after the first iteration (LS_init 0),
    c1 = 1e-4, c2 = 0.9, 100 corrections, optTol = 1e-5, progTol = 1e-9, Fref = 1, no damping,

-- plus the training loop's callback (GPz/callBack.m) that tracks the best theta and stops early.

PARITY UNPINNED: the reference ships no recorded optimiser trajectories and there is no MATLAB/Octave in this
image, so this file is checked only through properties (tests/test_minfunc_oracle.py: the Wolfe conditions at
every accepted step, the secant equation of the stored pairs, convergence on convex and Rosenbrock problems,
agreement of the circular two-loop product with a dense BFGS inverse).

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import math

import numpy as np

INF = float("inf")


def is_legal(v) -> bool:
    """isLegal.m:1 (real, no NaN, no Inf)."""
    return bool(np.all(np.isfinite(np.asarray(v, dtype=np.float64))))


def _mmax(a, b):
    """MATLAB max(a,b): a NaN operand is ignored."""
    if math.isnan(a):
        return b
    if math.isnan(b):
        return a
    return max(a, b)


def _mmin(a, b):
    if math.isnan(a):
        return b
    if math.isnan(b):
        return a
    return min(a, b)


def polyinterp_cubic2(x0, f0, g0, x1, f1, g1, lo=None, hi=None):
    """polyinterp.m:41-58 -- minimiser of the cubic through two points with values and slopes, clamped to
    [lo, hi] (defaults: the two abscissae, polyinterp.m:26-35); bisection when the discriminant is negative."""
    xmin, xmax = min(x0, x1), max(x0, x1)
    lo = xmin if lo is None else lo
    hi = xmax if hi is None else hi
    if x0 <= x1:                       # [minVal minPos] = min(points(:,1)): first of equal abscissae
        xa, fa, ga, xb, fb, gb = x0, f0, g0, x1, f1, g1
    else:
        xa, fa, ga, xb, fb, gb = x1, f1, g1, x0, f0, g0
    with np.errstate(all="ignore"):
        d1 = np.float64(ga) + gb - 3.0 * (np.float64(fa) - fb) / (np.float64(xa) - xb)
        rad = d1 * d1 - np.float64(ga) * gb
        if rad < 0:                    # sqrt is complex -> ~isreal(d2)
            return (hi + lo) / 2.0
        d2 = np.sqrt(rad)
        t = xb - (xb - xa) * ((gb + d2 - d1) / (gb - ga + 2.0 * d2))
    return float(_mmin(_mmax(float(t), lo), hi))


def polyinterp_quad(f0, g0, t, f1, lo, hi):
    """polyinterp.m:60-111 for points [0 f0 g0; t f1 <unknown>] (order 2): fit a x^2 + b x + c, test the
    bounds, the abscissae and the stationary point, keep the lowest polynomial value inside [lo, hi]."""
    with np.errstate(all="ignore"):
        a = (np.float64(f1) - f0 - np.float64(g0) * t) / (np.float64(t) * t)
        b, c = np.float64(g0), np.float64(f0)
        cand = [lo, hi, 0.0, t]
        if np.isfinite(2.0 * a) and np.isfinite(b) and a != 0.0:
            cand.append(float(-b / (2.0 * a)))
        best, fbest = (lo + hi) / 2.0, INF
        for x in cand:
            if lo <= x <= hi:
                fx = (a * x + b) * x + c
                if fx < fbest:
                    best, fbest = x, float(fx)
    return float(best)


def armijo_backtrack(fun, x, t, d, f, fr, g, gtd, c1=1e-4, prog_tol=1e-9):
    """ArmijoBacktrack.m:32-143 with LS_interp = 2, LS_multi = 0.  Returns t, f_new, g_new, evals."""
    f_new, g_new = fun(x + t * d)
    evals = 1
    while (not is_legal(f_new)) or f_new > fr + c1 * t * gtd:
        temp = t
        if not is_legal(f_new):
            t = 0.5 * t                                                         # :43-48
        elif not is_legal(g_new):
            t = polyinterp_quad(f, gtd, t, f_new, 0.0, t)                       # :49-56
        else:
            t = polyinterp_cubic2(0.0, f, gtd, t, f_new, float(g_new @ d), 0.0, t)   # :67-72
        if t < temp * 1e-3:                                                     # :91-102
            t = temp * 1e-3
        elif t > temp * 0.6:
            t = temp * 0.6
        f_new, g_new = fun(x + t * d)
        evals += 1
        if np.max(np.abs(t * d)) <= prog_tol:                                   # :121-128
            return 0.0, f, g, evals
    return t, f_new, g_new, evals


def wolfe_line_search(fun, x, t, d, f, g, gtd, c1=1e-4, c2=0.9, max_ls=25, prog_tol=1e-9):
    """WolfeLineSearch.m:32-263 with LS_interp = 2.  Returns t, f_new, g_new, evals."""
    f_new, g_new = fun(x + t * d)
    evals = 1
    gtd_new = float(g_new @ d)
    ls_iter, t_prev, f_prev, g_prev, gtd_prev = 0, 0.0, f, g, gtd
    nrm_d = float(np.max(np.abs(d)))
    done = False
    br_t = br_f = br_g = None
    while ls_iter < max_ls:                                                     # bracketing phase :51-126
        if not is_legal(f_new) or not is_legal(g_new):                          # :54-73
            t = (t + t_prev) / 2.0
            t, f_new, g_new, ev = armijo_backtrack(fun, x, t, d, f, f, g, gtd, c1, prog_tol)
            return t, f_new, g_new, evals + ev
        if f_new > f + c1 * t * gtd or (ls_iter > 1 and f_new >= f_prev):       # :76-80
            br_t, br_f, br_g = [t_prev, t], [f_prev, f_new], [g_prev, g_new]
            break
        elif abs(gtd_new) <= -c2 * gtd:                                         # :81-86
            br_t, br_f, br_g = [t], [f_new], [g_new]
            done = True
            break
        elif gtd_new >= 0:                                                      # :87-92
            br_t, br_f, br_g = [t_prev, t], [f_prev, f_new], [g_prev, g_new]
            break
        temp = t_prev                                                           # :93-110
        t_prev = t
        min_step = t + 0.01 * (t - temp)
        max_step = t * 10
        t = polyinterp_cubic2(temp, f_prev, gtd_prev, t, f_new, gtd_new, min_step, max_step)
        f_prev, g_prev, gtd_prev = f_new, g_new, gtd_new
        f_new, g_new = fun(x + t * d)
        evals += 1
        gtd_new = float(g_new @ d)
        ls_iter += 1
    if ls_iter == max_ls:                                                       # :128-132
        br_t, br_f, br_g = [0.0, t], [f, f_new], [g, g_new]

    insuf = False
    while not done and ls_iter < max_ls:                                        # zoom phase :142-243
        lo = 0 if not (br_f[1] < br_f[0]) else 1                                # [f_LO LOpos] = min(bracketFval)
        if math.isnan(br_f[0]) and not math.isnan(br_f[1]):
            lo = 1
        f_lo = br_f[lo]
        hi = 1 - lo
        if not is_legal(br_f) or not is_legal(br_g[0]) or not is_legal(br_g[1]):
            t = (br_t[0] + br_t[1]) / 2.0                                       # :149-153
        else:
            t = polyinterp_cubic2(br_t[0], br_f[0], float(br_g[0] @ d), br_t[1], br_f[1], float(br_g[1] @ d))
        bmax, bmin = max(br_t), min(br_t)
        with np.errstate(all="ignore"):
            ratio = np.float64(min(bmax - t, t - bmin)) / np.float64(bmax - bmin)
        if ratio < 0.1:                                                         # :174-196
            if insuf or t >= bmax or t <= bmin:
                if abs(t - bmax) < abs(t - bmin):
                    t = bmax - 0.1 * (bmax - bmin)
                else:
                    t = bmin + 0.1 * (bmax - bmin)
                insuf = False
            else:
                insuf = True
        else:
            insuf = False
        f_new, g_new = fun(x + t * d)                                           # :199-206
        evals += 1
        gtd_new = float(g_new @ d)
        ls_iter += 1
        armijo = f_new < f + c1 * t * gtd
        if not armijo or f_new >= f_lo:                                         # :209-214
            br_t[hi], br_f[hi], br_g[hi] = t, f_new, g_new
        else:
            if abs(gtd_new) <= -c2 * gtd:                                       # :216-218
                done = True
            elif gtd_new * (br_t[hi] - br_t[lo]) >= 0:                          # :219-223
                br_t[hi], br_f[hi], br_g[hi] = br_t[lo], br_f[lo], br_g[lo]
            br_t[lo], br_f[lo], br_g[lo] = t, f_new, g_new                      # :234-238
        if not done and abs(br_t[0] - br_t[1]) * nrm_d < prog_tol:              # :241-246
            break

    lo = 0                                                                      # :257-261
    if len(br_f) == 2 and (br_f[1] < br_f[0] or (math.isnan(br_f[0]) and not math.isnan(br_f[1]))):
        lo = 1
    return br_t[lo], br_f[lo], br_g[lo], evals


class LbfgsMemory:
    """The circular (S, Y, YS, start, end, Hdiag) store of minFunc.m:560-576."""

    def __init__(self, p, corrections=100):
        self.S = np.zeros((p, corrections))
        self.Y = np.zeros((p, corrections))
        self.YS = np.zeros(corrections)
        self.start, self.end, self.Hdiag = 1, 0, 1.0          # 1-based like the reference

    def add(self, y, s):
        """lbfgsAdd.m:2-30.  Returns True when the pair was skipped (curvature y's <= 1e-10)."""
        ys = float(y @ s)
        cor = self.S.shape[1]
        if not ys > 1e-10:
            return True
        if self.end < cor:
            self.end += 1
            if self.start != 1:
                self.start = 1 if self.start == cor else self.start + 1
        else:
            self.start = min(2, cor)
            self.end = 1
        self.S[:, self.end - 1] = s
        self.Y[:, self.end - 1] = y
        self.YS[self.end - 1] = ys
        self.Hdiag = ys / float(y @ y)
        return False

    def order(self):
        """lbfgsProd.m:9-15: column indices (0-based here), oldest first."""
        cor = self.S.shape[1]
        if self.start == 1:
            return list(range(0, self.end))
        return list(range(self.start - 1, cor)) + list(range(0, self.end))

    def prod(self, g):
        """lbfgsProd.m:19-32 / lbfgsProdC.c:47-88: d = -H g by the two-loop recursion."""
        ind = self.order()
        al = {}
        d = -g.copy()
        for i in reversed(ind):
            al[i] = float(self.S[:, i] @ d) / self.YS[i]
            d = d - al[i] * self.Y[:, i]
        d = self.Hdiag * d
        for i in ind:
            be = float(self.Y[:, i] @ d) / self.YS[i]
            d = d + self.S[:, i] * (al[i] - be)
        return d


def minfunc_lbfgs(fun, x0, max_iter=500, output_fcn=None, corrections=100, opt_tol=1e-5, prog_tol=1e-9,
                  c1=1e-4, c2=0.9, max_fun_evals=INF, trace=None):
    """minFunc.m:258-1170 restricted to method LBFGS with the options of the module docstring.

    fun(x) -> (f, g).  output_fcn(x, kind, i, fun_evals, f, t, gtd, g, d, opt_cond) -> stop (kind is 'init',
    'iter' or 'done').  Returns x, f, exitflag, dict(iterations, funcCount, message, firstorderopt).
    trace, when a list, receives one dict per iteration (t, f, gtd, optCond, ls_evals, skipped)."""
    x = np.array(x0, dtype=np.float64).copy()
    p = x.size
    f, g = fun(x)                                                               # :313-314
    g = np.asarray(g, dtype=np.float64)
    fun_evals = 1
    opt_cond = float(np.max(np.abs(g))) if p else 0.0
    info = dict(iterations=0, funcCount=1, message="", firstorderopt=opt_cond)
    if opt_cond <= opt_tol:                                                     # :350-362
        info["message"] = "Optimality Condition below optTol"
        return x, f, 1, info
    if output_fcn is not None and output_fcn(x, "init", 0, fun_evals, f, None, None, g, None, opt_cond):   # :365-378
        info["message"] = "Stopped by output function"
        return x, f, -1, info
    mem = LbfgsMemory(p, corrections)
    t, d, gtd = 1.0, np.zeros(p), 0.0
    g_old = g
    exitflag, msg, i = 0, "", 0
    for i in range(1, int(max_iter) + 1):
        skipped = False
        if i == 1:                                                              # :562-569
            d = -g
        else:                                                                   # :571-576
            skipped = mem.add(g - g_old, t * d)
            d = mem.prod(g)
        g_old = g
        if not is_legal(d):                                                     # :963-967 (the reference pauses; we stop)
            exitflag, msg = -3, "Step direction is illegal"
            break
        gtd = float(g @ d)                                                      # :972
        if gtd > -prog_tol:                                                     # :975-979
            exitflag, msg = 2, "Directional Derivative below progTol"
            break
        if i == 1:                                                              # :982-987
            t = min(1.0, 1.0 / float(np.sum(np.abs(g))))
        else:                                                                   # LS_init == 0, :989-991
            t = 1.0
        f_old = f
        t, f, g, ls_evals = wolfe_line_search(fun, x, t, d, f, g, gtd, c1, c2, 25, prog_tol)      # :1061-1068
        fun_evals += ls_evals
        x = x + t * d
        opt_cond = float(np.max(np.abs(g)))                                     # :1094
        if trace is not None:
            trace.append(dict(i=i, t=t, f=f, gtd=gtd, optCond=opt_cond, ls_evals=ls_evals, skipped=skipped))
        if output_fcn is not None and output_fcn(x, "iter", i, fun_evals, f, t, gtd, g, d, opt_cond):      # :1109-1116
            exitflag, msg = -1, "Stopped by output function"
            break
        if opt_cond <= opt_tol:                                                 # :1119-1123
            exitflag, msg = 1, "Optimality Condition below optTol"
            break
        if float(np.max(np.abs(t * d))) <= prog_tol:                            # :1127-1131
            exitflag, msg = 2, "Step Size below progTol"
            break
        if abs(f - f_old) < prog_tol:                                           # :1134-1138
            exitflag, msg = 2, "Function Value changing by less than progTol"
            break
        if fun_evals >= max_fun_evals:                                          # :1142-1146
            exitflag, msg = 0, "Reached Maximum Number of Function Evaluations"
            break
        if i == max_iter:                                                       # :1148-1152
            exitflag, msg = 0, "Reached Maximum Number of Iterations"
            break
    if output_fcn is not None:                                                  # :1165-1167
        output_fcn(x, "done", i, fun_evals, f, t, gtd, g, d, float(np.max(np.abs(g))))
    info.update(iterations=i, funcCount=fun_evals, message=msg, firstorderopt=float(np.max(np.abs(g))))
    return x, f, exitflag, info


def train_loop(fun_stats, theta0, best_theta, best_valid, max_iter=200, max_attempts=INF, training_only=True,
               log=None, **kw):
    """The optimisation part of GPz/train.m:5-48 with GPz/callBack.m as the output function.

    fun_stats(theta) -> (f, g, stats) with stats = (trainRMSE, trainLL, validRMSE, validLL): the values the
    reference leaves in globals on EVERY evaluation (GPz.m:3-7,236-259); the callback therefore reads those of the
    LAST evaluation of the line search, which need not be the accepted point (callBack.m:22-33).
    Returns theta_last, best_theta, best_valid, exitflag, info."""
    state = dict(stats=None, best_theta=np.array(best_theta, dtype=np.float64).copy(), best_valid=best_valid, attempts=None)

    def fun(th):
        f, g, st = fun_stats(th)
        state["stats"] = st
        return f, g

    def cb(x, kind, i, fun_evals, f, t, gtd, g, d, opt_cond):
        if kind == "iter":
            tr_rmse, tr_ll, va_rmse, va_ll = state["stats"]
            improved = True
            if training_only:                                                   # callBack.m:21-24
                state["best_valid"], state["best_theta"] = tr_ll, x.copy()
            elif state["best_valid"] is None or va_ll >= state["best_valid"]:   # callBack.m:26-30
                state["best_valid"], state["best_theta"], state["attempts"] = va_ll, x.copy(), 0
            else:                                                               # callBack.m:31-33 ([]+1 stays [])
                improved = False
                if state["attempts"] is not None:
                    state["attempts"] += 1
            if log is not None:
                log.append(dict(i=i, f=f, stats=tuple(state["stats"]), improved=improved, fun_evals=fun_evals, t=t))
        return state["attempts"] is not None and state["attempts"] == max_attempts      # callBack.m:48

    x, f, flag, info = minfunc_lbfgs(fun, theta0, max_iter=max_iter, output_fcn=cb, **kw)
    return x, state["best_theta"], state["best_valid"], flag, info
