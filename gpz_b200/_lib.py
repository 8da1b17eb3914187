"""ctypes binding of libgpz_b200.so (the C ABI declared in include/gpz_b200.h).

This is the Python stand-in for the MEX gateway (matlab/gpz_b200_mex.cpp): the same entry points, the
same column-major fp64 buffers.  There is no CPU fallback: if the shared library is missing or no
sm_100 device is present every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPZ_B200_LIB") or os.path.join(_HERE, "libgpz_b200.so")

EXPORTS = [
    "gpz_last_error", "gpz_version", "gpz_theta_len", "gpz_g_dim", "gpz_create", "gpz_destroy",
    "gpz_comm_unique_id", "gpz_comm_init", "gpz_eval", "gpz_eval_dev", "gpz_fit", "gpz_phi", "gpz_get_prior", "gpz_rows",
    "gpz_predict", "gpz_inv_logdet", "gpz_dxy", "gpz_stream", "gpz_sync", "gpz_launch_count", "gpz_graph_replays",
    "gpz_last_timing", "gpz_kernel_timing", "gpz_set_option", "gpz_dgemm_nt",
    "gpz_dxy_colmean", "gpz_train", "gpz_minimize_dev", "gpz_train_default_options", "gpz_train_reason",
    "gpz_create_multi", "gpz_destroy_multi", "gpz_multi_devices", "gpz_multi_ctx", "gpz_multi_eval", "gpz_multi_fit",
    "gpz_multi_get_prior", "gpz_multi_set_option", "gpz_multi_train",
]


class GpzModel(C.Structure):
    _fields_ = [("d", C.c_int32), ("k", C.c_int32), ("m", C.c_int32), ("method", C.c_char * 4),
                ("heteroscedastic", C.c_int32)]


class TrainOptions(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("training_only", C.c_int32), ("max_attempts", C.c_double),
                ("corrections", C.c_int32), ("max_ls", C.c_int32), ("opt_tol", C.c_double), ("prog_tol", C.c_double),
                ("c1", C.c_double), ("c2", C.c_double), ("max_fun_evals", C.c_double)]


class TrainIter(C.Structure):
    _fields_ = [("iter", C.c_int32), ("fun_evals", C.c_int32), ("improved", C.c_int32), ("attempts", C.c_int32),
                ("f", C.c_double), ("t", C.c_double), ("gtd", C.c_double), ("opt_cond", C.c_double),
                ("stats", C.c_double * 4)]


class TrainResult(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("fun_evals", C.c_int32), ("exitflag", C.c_int32), ("reason", C.c_int32),
                ("attempts", C.c_int32), ("skipped_pairs", C.c_int32), ("f", C.c_double), ("opt_cond", C.c_double),
                ("best_valid", C.c_double), ("ms_total", C.c_double), ("ms_eval", C.c_double)]


TRAIN_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(TrainIter))
OBJECTIVE_DEV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)


class GpzError(RuntimeError):
    pass


_lib = None
_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


def load():
    """Load the shared library (raises GpzError if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpzError(f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()); "
                       "gpz_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.gpz_last_error.restype = C.c_char_p
    lib.gpz_version.restype = C.c_int
    lib.gpz_theta_len.restype = C.c_int64
    lib.gpz_theta_len.argtypes = [C.POINTER(GpzModel)]
    lib.gpz_g_dim.restype = C.c_int64
    lib.gpz_g_dim.argtypes = [C.POINTER(GpzModel)]
    lib.gpz_create.restype = C.c_int
    lib.gpz_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(GpzModel), C.c_int64, _dp, _dp, _dp, _dp, _u8p, _u8p, C.c_int]
    lib.gpz_destroy.restype = None
    lib.gpz_destroy.argtypes = [C.c_void_p]
    lib.gpz_comm_unique_id.restype = C.c_int
    lib.gpz_comm_unique_id.argtypes = [C.c_char_p]
    lib.gpz_comm_init.restype = C.c_int
    lib.gpz_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    lib.gpz_eval.restype = C.c_int
    lib.gpz_eval.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.gpz_eval_dev.restype = C.c_int
    lib.gpz_eval_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gpz_fit.restype = C.c_int
    lib.gpz_fit.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.gpz_phi.restype = C.c_int
    lib.gpz_phi.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp, _dp]
    lib.gpz_get_prior.restype = C.c_int
    lib.gpz_get_prior.argtypes = [C.c_void_p, _dp, _dp]
    lib.gpz_rows.restype = C.c_int64
    lib.gpz_rows.argtypes = [C.c_void_p, C.c_int]
    lib.gpz_predict.restype = C.c_int
    lib.gpz_predict.argtypes = [C.POINTER(GpzModel), _dp, _dp, _dp, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int]
    lib.gpz_inv_logdet.restype = C.c_int
    lib.gpz_inv_logdet.argtypes = [C.c_int32, _dp, _dp, _dp, C.c_int]
    lib.gpz_dxy.restype = C.c_int
    lib.gpz_dxy.argtypes = [C.c_int64, C.c_int32, C.c_int32, _dp, _dp, _dp, C.c_int]
    lib.gpz_dxy_colmean.restype = C.c_int
    lib.gpz_dxy_colmean.argtypes = [C.c_int64, C.c_int32, C.c_int32, _dp, _dp, _dp, C.c_int]
    lib.gpz_stream.restype = C.c_void_p
    lib.gpz_stream.argtypes = [C.c_void_p]
    lib.gpz_sync.restype = C.c_int
    lib.gpz_sync.argtypes = [C.c_void_p]
    lib.gpz_launch_count.restype = C.c_int64
    lib.gpz_launch_count.argtypes = [C.c_void_p]
    lib.gpz_graph_replays.restype = C.c_int64
    lib.gpz_graph_replays.argtypes = [C.c_void_p]
    lib.gpz_dgemm_nt.restype = C.c_int
    lib.gpz_dgemm_nt.argtypes = [C.c_int64, C.c_int64, C.c_int64, _dp, C.c_int64, _dp, C.c_int64, _dp, C.c_int64, C.c_int32, C.c_int]
    lib.gpz_last_timing.restype = C.c_int
    lib.gpz_last_timing.argtypes = [C.c_void_p, _dp]
    lib.gpz_kernel_timing.restype = C.c_int
    lib.gpz_kernel_timing.argtypes = [C.c_void_p, _dp]
    lib.gpz_set_option.restype = C.c_int
    lib.gpz_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    lib.gpz_train_default_options.restype = None
    lib.gpz_train_default_options.argtypes = [C.POINTER(TrainOptions)]
    lib.gpz_train_reason.restype = C.c_char_p
    lib.gpz_train_reason.argtypes = [C.c_int]
    lib.gpz_train.restype = C.c_int
    lib.gpz_train.argtypes = [C.c_void_p, C.POINTER(TrainOptions), _dp, _dp, _dp, TRAIN_CALLBACK, C.c_void_p,
                              C.POINTER(TrainResult)]
    lib.gpz_minimize_dev.restype = C.c_int
    lib.gpz_minimize_dev.argtypes = [C.c_int64, OBJECTIVE_DEV, C.c_void_p, C.POINTER(TrainOptions), _dp, _dp, _dp,
                                     TRAIN_CALLBACK, C.c_void_p, C.POINTER(TrainResult), C.c_int]
    lib.gpz_create_multi.restype = C.c_int
    lib.gpz_create_multi.argtypes = [C.POINTER(C.c_void_p), C.POINTER(GpzModel), C.c_int64, _dp, _dp, _dp, _dp, _u8p, _u8p, C.c_int,
                                     C.POINTER(C.c_int)]
    lib.gpz_destroy_multi.restype = None
    lib.gpz_destroy_multi.argtypes = [C.c_void_p]
    lib.gpz_multi_devices.restype = C.c_int
    lib.gpz_multi_devices.argtypes = [C.c_void_p]
    lib.gpz_multi_ctx.restype = C.c_void_p
    lib.gpz_multi_ctx.argtypes = [C.c_void_p, C.c_int]
    lib.gpz_multi_eval.restype = C.c_int
    lib.gpz_multi_eval.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.gpz_multi_fit.restype = C.c_int
    lib.gpz_multi_fit.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.gpz_multi_get_prior.restype = C.c_int
    lib.gpz_multi_get_prior.argtypes = [C.c_void_p, _dp, _dp]
    lib.gpz_multi_set_option.restype = C.c_int
    lib.gpz_multi_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    lib.gpz_multi_train.restype = C.c_int
    lib.gpz_multi_train.argtypes = [C.c_void_p, C.POINTER(TrainOptions), _dp, _dp, _dp, TRAIN_CALLBACK, C.c_void_p,
                                    C.POINTER(TrainResult)]
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise GpzError(f"libgpz_b200 error {rc}: {load().gpz_last_error().decode()}")


def f64(a, shape=None):
    """Column-major contiguous fp64 copy/view (MATLAB layout)."""
    a = np.asarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape, order="F")
    return np.asfortranarray(a)


def ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def make_model(d: int, k: int, m: int, method: str, heteroscedastic: bool) -> GpzModel:
    mm = GpzModel()
    mm.d, mm.k, mm.m = int(d), int(k), int(m)
    mm.method = method.encode()
    mm.heteroscedastic = 1 if heteroscedastic else 0
    return mm


class Context:
    """Device-resident dataset + workspaces: the closure `f = @(params) GPz(params,model,X,Y,Psi,omega,
    training,validation)` of GPz/train.m:40 as an object.  Inputs are already normalised."""

    def __init__(self, model: GpzModel, X, Y, Psi=None, omega=None, training=None, validation=None, device=0):
        lib = load()
        X = f64(X)
        n_all, d = X.shape
        assert d == model.d
        Y = f64(Y).reshape(n_all, -1, order="F")
        assert Y.shape[1] == model.k
        self.model = model
        self.p = int(lib.gpz_theta_len(C.byref(model)))
        psi = None
        if Psi is not None:
            psi = f64(Psi)
            if model.method[1:2] == b"C":
                assert psi.shape == (d, d, n_all), psi.shape
            else:
                assert psi.shape == (n_all, d), psi.shape
        om = None if omega is None else f64(omega).reshape(-1)
        tr = None if training is None else np.ascontiguousarray(np.asarray(training).reshape(-1) != 0, dtype=np.uint8)
        va = None if validation is None else np.ascontiguousarray(np.asarray(validation).reshape(-1) != 0, dtype=np.uint8)
        h = C.c_void_p()
        check(lib.gpz_create(C.byref(h), C.byref(model), n_all, ptr(X), ptr(Y), ptr(psi), ptr(om),
                             None if tr is None else tr.ctypes.data_as(_u8p),
                             None if va is None else va.ctypes.data_as(_u8p), int(device)))
        self._h = h
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gpz_destroy(self._h)
            self._h = None

    __del__ = close

    def set_option(self, name: str, value: float):
        check(self._lib.gpz_set_option(self._h, name.encode(), float(value)))

    def comm_init(self, rank: int, world: int, uid: bytes):
        check(self._lib.gpz_comm_init(self._h, rank, world, uid))

    def eval(self, theta):
        """[nlogML, grad] = GPz(theta, ...) plus the four global statistics (GPz.m:236-259)."""
        th = f64(theta).reshape(-1)
        assert th.size == self.p, (th.size, self.p)
        f = C.c_double()
        g = np.empty(self.p)
        st = np.empty(4)
        check(self._lib.gpz_eval(self._h, ptr(th), C.byref(f), ptr(g), ptr(st)))
        return f.value, g, dict(trainRMSE=st[0], trainLL=st[1], validRMSE=st[2], validLL=st[3])

    def eval_dev(self, d_theta_ptr: int, d_out_ptr: int):
        check(self._lib.gpz_eval_dev(self._h, C.c_void_p(d_theta_ptr), C.c_void_p(d_out_ptr)))

    def fit(self, theta, want_nlogML=True):
        """[nlogML(1xk, un-normalised), ~, w, iSigma_w] = GPz(theta, ...) (GPz.m:84-87)."""
        th = f64(theta).reshape(-1)
        m, k = self.model.m, self.model.k
        nl = np.empty(k)
        w = np.empty((m, k), order="F")
        iS = np.empty((m, m, k), order="F")
        check(self._lib.gpz_fit(self._h, ptr(th), ptr(nl) if want_nlogML else None, ptr(w), ptr(iS)))
        return (nl.reshape(1, k) if want_nlogML else None), w, iS

    def rows(self, which=0):
        return int(self._lib.gpz_rows(self._h, which))

    def phi(self, theta, which=0, want_phi=True, want_N=False):
        th = f64(theta).reshape(-1)
        n = self.rows(which)
        PHI = np.empty((n, self.model.m), order="F") if want_phi else None
        lnb = np.empty((n, self.model.k), order="F")
        N = np.empty((n, self.model.m), order="F") if want_N else None
        check(self._lib.gpz_phi(self._h, ptr(th), which, ptr(PHI), ptr(lnb), ptr(N)))
        return (PHI, lnb, N) if want_N else (PHI, lnb)

    def get_prior(self, theta):
        """prior = getPrior(X,Psi,theta,model,training) (GPz/getPrior.m)."""
        th = f64(theta).reshape(-1)
        pr = np.empty(self.model.m)
        check(self._lib.gpz_get_prior(self._h, ptr(th), ptr(pr)))
        return pr

    def train(self, theta, best_theta, best_valid, callback=None, **options):
        """theta = minFunc(f, theta, options) + callBack.m on the device (gpz_train).  Returns theta_last, best_theta,
        best_valid, result dict.  callback(dict) -> truthy stops the run."""
        th = f64(theta).reshape(-1).copy()
        bt = f64(best_theta).reshape(-1).copy()
        assert th.size == self.p and bt.size == self.p
        return _run_train(lambda o, cb, bv, res: self._lib.gpz_train(self._h, C.byref(o), ptr(th), ptr(bt), C.byref(bv), cb,
                                                                     None, C.byref(res)), th, bt, best_valid, callback, options)

    def stream(self) -> int:
        return int(self._lib.gpz_stream(self._h) or 0)

    def sync(self):
        check(self._lib.gpz_sync(self._h))

    def launch_count(self) -> int:
        return int(self._lib.gpz_launch_count(self._h))

    def graph_replays(self) -> int:
        return int(self._lib.gpz_graph_replays(self._h))

    def last_timing(self):
        ms = np.empty(12)
        check(self._lib.gpz_last_timing(self._h, ptr(ms)))
        return dict(phi=ms[0], gram=ms[1], solve=ms[2], tgemm=ms[3], backproj=ms[4], total=ms[5],
                    gram_kernel=ms[6], tgemm_kernel=ms[7], i8_gemms_ms=ms[8], i8_gemms_ops=ms[9],
                    int8_slices=int(ms[10]), int8_gram=int(ms[11]))

    def kernel_timing(self):
        ms = np.empty(4)
        check(self._lib.gpz_kernel_timing(self._h, ptr(ms)))
        return dict(phi_build=ms[0], digits=ms[1], gram_gemm=ms[2], moment_gemm=ms[3])


class MultiContext:
    """One caller thread driving several GPUs (gpz_create_multi): the same closure object as Context, rows split over
    `ngpus` devices inside the library, NCCL allreduces inside every call."""

    def __init__(self, model: GpzModel, X, Y, Psi=None, omega=None, training=None, validation=None, ngpus=2, devices=None):
        lib = load()
        X = f64(X)
        n_all, d = X.shape
        assert d == model.d
        Y = f64(Y).reshape(n_all, -1, order="F")
        self.model = model
        self.p = int(lib.gpz_theta_len(C.byref(model)))
        psi = None if Psi is None else f64(Psi)
        om = None if omega is None else f64(omega).reshape(-1)
        tr = None if training is None else np.ascontiguousarray(np.asarray(training).reshape(-1) != 0, dtype=np.uint8)
        va = None if validation is None else np.ascontiguousarray(np.asarray(validation).reshape(-1) != 0, dtype=np.uint8)
        dv = None if devices is None else (C.c_int * int(ngpus))(*[int(x) for x in devices])
        h = C.c_void_p()
        check(lib.gpz_create_multi(C.byref(h), C.byref(model), n_all, ptr(X), ptr(Y), ptr(psi), ptr(om),
                                   None if tr is None else tr.ctypes.data_as(_u8p),
                                   None if va is None else va.ctypes.data_as(_u8p), int(ngpus), dv))
        self._h = h
        self._lib = lib
        self.ngpus = int(ngpus)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gpz_destroy_multi(self._h)
            self._h = None

    __del__ = close

    def set_option(self, name: str, value: float):
        check(self._lib.gpz_multi_set_option(self._h, name.encode(), float(value)))

    def eval(self, theta):
        th = f64(theta).reshape(-1)
        assert th.size == self.p, (th.size, self.p)
        f = C.c_double()
        g = np.empty(self.p)
        st = np.empty(4)
        check(self._lib.gpz_multi_eval(self._h, ptr(th), C.cast(C.byref(f), _dp), ptr(g), ptr(st)))
        return f.value, g, dict(trainRMSE=st[0], trainLL=st[1], validRMSE=st[2], validLL=st[3])

    def fit(self, theta):
        th = f64(theta).reshape(-1)
        m, k = self.model.m, self.model.k
        nl = np.empty(k)
        w = np.empty((m, k), order="F")
        iS = np.empty((m, m, k), order="F")
        check(self._lib.gpz_multi_fit(self._h, ptr(th), ptr(nl), ptr(w), ptr(iS)))
        return nl, w, iS

    def get_prior(self, theta):
        th = f64(theta).reshape(-1)
        pr = np.empty(self.model.m)
        check(self._lib.gpz_multi_get_prior(self._h, ptr(th), ptr(pr)))
        return pr

    def train(self, theta, best_theta, best_valid, callback=None, **options):
        th = f64(theta).reshape(-1).copy()
        bt = f64(best_theta).reshape(-1).copy()
        assert th.size == self.p and bt.size == self.p
        return _run_train(lambda o, cb, bv, res: self._lib.gpz_multi_train(self._h, C.byref(o), ptr(th), ptr(bt), C.cast(C.byref(bv), _dp),
                                                                           cb, None, C.byref(res)),
                          th, bt, best_valid, callback, options)

    def rank_timing(self, rank=0):
        """gpz_last_timing of one device's context (all ranks block on the same two allreduces)."""
        ms = np.empty(12)
        check(self._lib.gpz_last_timing(self._lib.gpz_multi_ctx(self._h, rank), ptr(ms)))
        return dict(phi=ms[0], gram=ms[1], solve=ms[2], tgemm=ms[3], backproj=ms[4], total=ms[5])


def train_options(**kw) -> TrainOptions:
    o = TrainOptions()
    load().gpz_train_default_options(C.byref(o))
    for name, value in kw.items():
        if not hasattr(o, name):
            raise GpzError(f"unknown train option {name!r}")
        setattr(o, name, value)
    return o


def _run_train(call, th, bt, best_valid, callback, options):
    lib = load()
    o = train_options(**options)
    raised = []

    def _cb(_user, it):
        if callback is None:
            return 0
        i = it.contents
        try:
            return 1 if callback(dict(iter=i.iter, fun_evals=i.fun_evals, improved=bool(i.improved), attempts=i.attempts,
                                      f=i.f, t=i.t, gtd=i.gtd, opt_cond=i.opt_cond, trainRMSE=i.stats[0],
                                      trainLL=i.stats[1], validRMSE=i.stats[2], validLL=i.stats[3])) else 0
        except BaseException as e:        # never unwind through the C frames
            raised.append(e)
            return 1

    cb = TRAIN_CALLBACK(_cb)
    bv = C.c_double(np.nan if best_valid is None else float(best_valid))
    res = TrainResult()
    check(call(o, cb, bv, res))
    if raised:
        raise raised[0]
    info = {name: getattr(res, name) for name, _ in TrainResult._fields_}
    info["message"] = lib.gpz_train_reason(res.reason).decode()
    return th, bt, bv.value, info


def minimize_dev(p, objective, theta, best_theta=None, best_valid=None, callback=None, device=0, **options):
    """gpz_minimize_dev: the optimiser on a caller-supplied objective(d_x_ptr, d_out_ptr, stream_ptr) -> rc that fills
    the DEVICE buffer d_out = [f, g[p], stats[4]] (used by the tests to run it on analytic functions)."""
    lib = load()
    th = f64(theta).reshape(-1).copy()
    bt = th.copy() if best_theta is None else f64(best_theta).reshape(-1).copy()
    fn = OBJECTIVE_DEV(lambda _u, dx, do, st: int(objective(dx, do, st) or 0))
    return _run_train(lambda o, cb, bv, res: lib.gpz_minimize_dev(int(p), fn, None, C.byref(o), ptr(th), ptr(bt), C.byref(bv),
                                                                  cb, None, C.byref(res), int(device)),
                      th, bt, best_valid, callback, options)


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(load().gpz_comm_unique_id(buf))
    return buf.raw


def predict_core(model: GpzModel, theta, w, iSigma_w, Xz, Psi=None, want_phi=False, device=0, priors=None):
    lib = load()
    Xz = f64(Xz)
    n = Xz.shape[0]
    k, m = model.k, model.m
    th = f64(theta).reshape(-1)
    w = f64(w).reshape(m, k, order="F")
    iS = f64(iSigma_w).reshape(m, m, k, order="F")
    psi = None if Psi is None else f64(Psi)
    pri = None if priors is None else f64(priors).reshape(-1)
    mu, nu, be, ga = (np.empty((n, k), order="F") for _ in range(4))
    PHI = np.empty((n, m), order="F") if want_phi else None
    check(lib.gpz_predict(C.byref(model), ptr(th), ptr(w), ptr(iS), n, ptr(Xz), ptr(psi), ptr(pri), ptr(mu), ptr(nu), ptr(be),
                          ptr(ga), ptr(PHI), int(device)))
    return mu, nu, be, ga, PHI


def inv_logdet(A, device=0):
    A = f64(A)
    m = A.shape[0]
    Xi = np.empty((m, m), order="F")
    ld = C.c_double()
    check(load().gpz_inv_logdet(m, ptr(A), ptr(Xi), C.byref(ld), int(device)))
    return Xi, ld.value


def dxy(X, Y, device=0):
    X, Y = f64(X), f64(Y)
    n, d = X.shape
    m = Y.shape[0]
    D = np.empty((n, m), order="F")
    check(load().gpz_dxy(n, m, d, ptr(X), ptr(Y), ptr(D), int(device)))
    return D


def dxy_colmean(X, Y, device=0):
    """mean(Dxy(X,Y)) over the rows of X (init.m:62), reduced on the device."""
    X, Y = f64(X), f64(Y)
    n, d = X.shape
    m = Y.shape[0]
    out = np.empty(m)
    check(load().gpz_dxy_colmean(n, m, d, ptr(X), ptr(Y), ptr(out), int(device)))
    return out


def dgemm_nt(A, B, digits=7, device=0):
    """C = A @ B.T in fp64 through the int8 tensor cores (gpz_dgemm_nt): A (M x K), B (N x K) C-contiguous."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    M, K = A.shape
    N, K2 = B.shape
    assert K == K2
    Cm = np.empty((M, N), dtype=np.float64)
    check(load().gpz_dgemm_nt(M, N, K, ptr(A), K, ptr(B), K, ptr(Cm), N, int(digits), int(device)))
    return Cm
